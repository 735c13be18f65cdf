"""CPU check of the kernels' algebra (csrc/gtf_math.cuh compiled for the host) against the golden
fixtures produced by the unmodified reference and against the oracle's pure helpers.
Tolerance: 1e-9 relative (BASELINE.json north_star)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import golden_util as gu
import oracle_lib as ol

CSRC = os.path.join(gu.REPO, "gnn-track-finding_b200", "csrc")
GEOM = np.array([0.3, 0.4, 0.6, 550.0])
dp = ctypes.POINTER(ctypes.c_double)


def P(a):
    return a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def hm():
    so = os.path.join(CSRC, "libgtf_hostmath.so")
    src = [os.path.join(CSRC, f) for f in ("gtf_hostmath.cpp", "gtf_math.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-x", "c++", "-o", so, src[0], "-lm"])
    L = ctypes.CDLL(so)
    L.gtfh_var_ms.restype = ctypes.c_double
    L.gtfh_var_ms.argtypes = [ctypes.c_double] * 7
    L.gtfh_pair_chi2.restype = ctypes.c_double
    L.gtfh_kl.restype = ctypes.c_double
    L.gtfh_seed_entry.argtypes = [dp, dp, ctypes.c_double, ctypes.c_double, dp, dp]
    L.gtfh_extrapolate.argtypes = [dp, dp, dp, ctypes.c_double, ctypes.c_double, dp, dp]
    L.gtfh_pair_chi2.argtypes = [dp] * 6
    L.gtfh_merge.argtypes = [dp, dp, dp]
    L.gtfh_kl.argtypes = [dp, dp]
    return L


def xyzr(hb, i):
    return np.array([hb["x"][i], hb["y"][i], hb["z"][i], hb["r"][i]])


@pytest.mark.parametrize("name", ["barrel25_deg6", "barrel40_eta1"])
def test_seed_entry_matches_reference(hm, name):
    fx = gu.load(name)
    hb = gu.stage_batch(fx, "seed")
    worst = 0.0
    for i in range(len(hb["x"])):
        s0, s1 = hb["in_off"][i], hb["in_off"][i + 1]
        d = s1 - s0
        for k in range(d):
            s = s0 + k
            other = hb["in_src"][s0 + d - 1 - k]       # quirk 5: tau of the mirrored neighbour
            dz, dr = hb["z"][other] - hb["z"][i], hb["r"][other] - hb["r"][i]
            tau = dz / dr
            sr, sz = (0.4, 0.6) if abs(hb["z"][i]) < 550 else (0.6, 0.4)
            srn, szn = (0.4, 0.6) if abs(hb["z"][other]) < 550 else (0.6, 0.4)
            vt = (sz ** 2 + szn ** 2) / dr ** 2 + (dz / dr ** 2) ** 2 * (sr ** 2 + srn ** 2)
            out = np.zeros(8)
            hm.gtfh_seed_entry(P(xyzr(hb, i)), P(xyzr(hb, hb["in_src"][s])), tau, vt * vt, P(GEOM), P(out))
            want = np.array([hb["tse_" + f][s] for f in ("a", "b", "c", "tau", "p00", "p01", "p11", "p22")])
            worst = max(worst, gu.rel_err(out, want))
    assert worst <= gu.RTOL, worst


@pytest.mark.parametrize("name", ["barrel25_deg6", "barrel40_eta1"])
def test_extrapolate_matches_reference(hm, name):
    """Stage E on the fixture's post-extraction state: per-source var_ms prefix in successor order
    (quirk 2), gate decisions bit-exact, updated states / likelihood within 1e-9."""
    fx = gu.load(name)
    hb = gu.stage_batch(fx, "x1")
    want = gu.stage_batch(fx, "e2")
    ex = gu.edge_exists(hb)
    n_sent = n_pass = 0
    worst = 0.0
    for u in range(len(hb["x"])):
        if not (hb["alive"][u] and hb["has_merged"][u] and hb["sub_state"][hb["sub"][u]] == 0):
            continue
        p11 = hb["m_p11"][u]
        for o in range(hb["out_off"][u], hb["out_off"][u + 1]):
            s = hb["out_slot"][o]
            v = hb["slot_dst"][s]
            if not (ex[s] and hb["active"][s] == 1):
                continue
            vms = hm.gtfh_var_ms(hb["m_a"][u], hb["m_b"][u], hb["x"][v], hb["r"][v] - hb["r"][u],
                                 hb["z"][v] - hb["z"][u], hb["z"][u], 550.0)
            p11 = p11 + vms
            m7 = np.array([hb["m_a"][u], hb["m_b"][u], hb["m_c"][u], hb["m_p00"][u], hb["m_p01"][u], p11, hb["m_p22"][u]])
            out = np.zeros(11)
            hm.gtfh_extrapolate(P(xyzr(hb, u)), P(xyzr(hb, v)), P(m7), vms, 2.0, P(GEOM), P(out))
            n_sent += 1
            passed = bool(out[2])
            assert passed == bool(want["uts_present"][s]), (u, v, out[0])
            if passed:
                n_pass += 1
                w = np.array([want["uts_" + f][s] for f in ("lik", "a", "b", "c", "tau", "p00", "p01", "p11", "p22")])
                worst = max(worst, gu.rel_err(out[[1, 3, 4, 5, 6, 7, 8, 9, 10]], w))
        # the accumulated value persists on the node (extrapolate_merged_states.py:127-128)
        assert gu.rel_err(np.array([p11]), np.array([want["m_p11"][u]])) <= gu.RTOL
    assert n_sent > 100 and n_pass > 20
    assert worst <= gu.RTOL, worst


def test_pair_merge_kl_match_oracle(hm):
    """gtf_pair_chi2 / gtf_merge / gtf_kl (block-covariance closed forms) vs the oracle's literal 3x3
    general-inverse restatement of clustering.py:11-105, on real seeded states."""
    fx = gu.load("barrel40_eta1")
    hb = gu.stage_batch(fx, "seed")
    L = ol.lib()
    rng = np.random.default_rng(0)
    worst = {"chi2": 0.0, "merge": 0.0, "kl": 0.0}
    n = 0
    for i in range(len(hb["x"])):
        s0, s1 = hb["in_off"][i], hb["in_off"][i + 1]
        if s1 - s0 < 3:
            continue
        a, b = rng.choice(np.arange(s0, s1), 2, replace=False)

        def st(s):
            return np.array([hb["tse_" + f][s] for f in ("a", "b", "c", "tau", "p00", "p01", "p11", "p22")])

        def full(s8, joint):
            m = np.array([s8[0], s8[1], s8[3] if joint else s8[2]])
            c = np.array([s8[4], s8[5], 0, s8[5], s8[6], 0, 0, 0, s8[7]], dtype=np.float64)
            return m, c
        sa, sb = st(a), st(b)
        ma, ca = full(sa, True)
        mb, cb = full(sb, True)
        node, na, nb = xyzr(hb, i), xyzr(hb, hb["in_src"][a]), xyzr(hb, hb["in_src"][b])
        want = L.gtfo_mahalanobis(P(ma), P(ca), P(mb), P(cb), P(node), P(na), P(nb), 0.4, 0.6, 550.0)
        got = hm.gtfh_pair_chi2(P(sa), P(sb), P(node), P(na), P(nb), P(GEOM))
        worst["chi2"] = max(worst["chi2"], gu.rel_err([got], [want]))
        mm, mc = np.zeros(3), np.zeros(9)
        L.gtfo_merge_states(P(ma), P(ca), P(mb), P(cb), P(mm), P(mc))
        pm, pc = np.zeros(3), np.zeros(9)
        mpa, _ = full(sa, False)
        mpb, _ = full(sb, False)
        L.gtfo_merge_states(P(mpa), P(ca), P(mpb), P(cb), P(pm), P(pc))
        out = np.zeros(8)
        hm.gtfh_merge(P(sa), P(sb), P(out))
        w = np.array([mm[0], mm[1], pm[2], mm[2], mc[0], mc[1], mc[4], mc[8]])
        worst["merge"] = max(worst["merge"], gu.rel_err(out, w))
        wantkl = L.gtfo_kl_distance(P(ma), P(ca), P(mm), P(mc))
        gotkl = hm.gtfh_kl(P(sa), P(out))
        worst["kl"] = max(worst["kl"], gu.rel_err([gotkl], [wantkl]))
        n += 1
    assert n > 100
    assert all(v <= gu.RTOL for v in worst.values()), worst


@pytest.mark.parametrize("name,prev,stage,key,chi2,kl", [
    ("barrel25_deg6", "seed", "c1", "tse", 1.0, 2.0), ("barrel40_eta1", "seed", "c1", "tse", 1.0, 2.0),
    ("barrel25_deg6", "m2", "c3", "uts", 1000.0, 100.0), ("barrel40_eta1", "m2", "c3", "uts", 1000.0, 100.0)])
def test_information_form_clustering_matches_reference(hm, name, prev, stage, key, chi2, kl):
    """the kernels' greedy merge loop runs in information form (sum of inverse covariances); check the
    merged states and the set of un-absorbed components against the unmodified reference's cluster()"""
    hm.gtfh_cluster_node.argtypes = [dp, dp, ctypes.c_int, dp, dp, ctypes.c_double, ctypes.c_double, dp, dp, dp,
                                     ctypes.POINTER(ctypes.c_uint)]
    fx = gu.load(name)
    hb = gu.stage_batch(fx, prev)
    want = gu.stage_batch(fx, stage)
    order = gu.dict_order(hb, key)
    ex = gu.edge_exists(hb)
    n_checked = 0
    worst = 0.0
    for i in range(len(hb["x"])):
        if not gu.inplay_nodes(hb)[i] or (key == "uts" and not hb["has_uts"][i]):
            continue
        sl = order[i]
        n = len(sl)
        if n < 3 or n > 15:
            continue
        S = np.array([[hb["%s_%s" % (key, f)][s] for f in ("a", "b", "c", "tau", "p00", "p01", "p11", "p22")] for s in sl])
        pr = np.array([hb[key + "_prior"][s] for s in sl])
        nb = np.array([xyzr(hb, hb["in_src"][s]) for s in sl])
        m8, mp, rem = np.zeros(8), ctypes.c_double(0), ctypes.c_uint(0)
        ok = hm.gtfh_cluster_node(P(np.ascontiguousarray(S)), P(pr), n, P(xyzr(hb, i)), P(np.ascontiguousarray(nb)), chi2, kl,
                                  P(GEOM), P(m8), ctypes.byref(mp), ctypes.byref(rem))
        newly = bool(ok)
        if key == "tse":
            assert newly == bool(want["has_merged"][i])
        if not newly:
            continue
        w = np.array([want["m_a"][i], want["m_b"][i], want["m_c"][i], want["m_p00"][i], want["m_p01"][i], want["m_p11"][i],
                      want["m_p22"][i], want["m_prior"][i]])
        got = np.array([m8[0], m8[1], m8[2], m8[4], m8[5], m8[6], m8[7], mp.value])
        worst = max(worst, gu.rel_err(got, w))
        for k, s in enumerate(sl):      # un-absorbed components are switched off, absorbed ones keep their flag
            if (rem.value >> k) & 1 and ex[s]:
                assert want["active"][s] == 0
        n_checked += 1
    assert n_checked > (20 if key == "tse" else -1), n_checked
    print(name, stage, "nodes", n_checked, "worst rel err %.3g" % worst)
    assert worst <= gu.RTOL, worst


def test_general_kl_known_answer_from_reference_csv(hm):
    """the kernel algebra for general 3x3 covariances (gtf_kl_general, closed-form cofactor inverse) reproduces the
    reference's shipped golden vector (1_events_training_data.csv, 7,574 KL values) as a sorted multiset"""
    import ctypes
    fx = np.load(gu.GOLDEN + "/kl_parabolic_known_answer.npz")
    dp = ctypes.POINTER(ctypes.c_double)
    hm.gtfh_kl_general.restype = ctypes.c_double
    hm.gtfh_kl_general.argtypes = [dp] * 4
    mean, cov, off = fx["mean"], fx["cov"].reshape(-1, 9), fx["off"]
    out = []
    for a, b in zip(off[:-1], off[1:]):
        for i in range(a, b):
            for j in range(a, i):
                mi, ci, mj, cj = (np.ascontiguousarray(x) for x in (mean[i], cov[i], mean[j], cov[j]))
                out.append(hm.gtfh_kl_general(mi.ctypes.data_as(dp), ci.ctypes.data_as(dp), mj.ctypes.data_as(dp),
                                              cj.ctypes.data_as(dp)))
    got = np.sort(np.array(out))
    assert len(got) == 7574
    assert gu.rel_err(got, fx["kl_sorted"]) <= 1e-9


def test_parabolic_seeding_algebra_and_oracle_vs_reference(hm):
    """learn_KL_parabolic_model/.../utils.py:221-299: the kernel's closed-form inverse (gtf_seed_parabolic in gtf_math.cuh, host
    build) and the oracle's literal restatement (np.linalg.inv -> pivoted Gauss-Jordan) against the unmodified reference's
    outputs (tests/golden/parabolic_seed.npz): 1e-9 (measured 1e-14)."""
    fx = np.load(os.path.join(gu.REPO, "tests", "golden", "parabolic_seed.npz"))
    O = ol.lib()
    sig = [dp, dp, ctypes.c_double, ctypes.c_double, ctypes.c_double, dp, dp]
    hm.gtfh_seed_parabolic.argtypes = sig
    hm.gtfh_seed_parabolic.restype = None
    O.gtfo_seed_parabolic.argtypes = sig
    O.gtfo_seed_parabolic.restype = None
    for fn in (hm.gtfh_seed_parabolic, O.gtfo_seed_parabolic):
        for i in range(len(fx["node"])):
            n, b = np.ascontiguousarray(fx["node_xy"][i]), np.ascontiguousarray(fx["nbr_xy"][i])
            sv, cv = np.zeros(3), np.zeros(9)
            fn(P(n), P(b), 4.0, 0.1, 0.1, P(sv), P(cv))
            assert np.allclose(sv, fx["state"][i], rtol=1e-9, atol=1e-300)
            assert np.allclose(cv, fx["cov"][i].ravel(), rtol=1e-9, atol=0)
