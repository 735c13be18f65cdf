"""CPU: the oracle (oracle/gtf_oracle.c) against the golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  This is what pins the oracle; the GPU tests then compare against the
oracle and the same fixtures."""
import numpy as np
import pytest

import golden_util as gu
import oracle_lib as ol

ALL = ("alive", "active", "merged", "tse", "uts", "degree", "edge_w")
FIX = ["barrel25_deg6", "barrel40_eta1"]


def blank_seed(hb):
    hb = dict(hb)
    for f in list(hb):
        if f.startswith("tse_") and f != "tse_present":
            hb[f] = np.full_like(hb[f], np.nan)
    hb["tse_present"] = np.zeros_like(hb["tse_present"])
    hb["active"] = np.zeros_like(hb["active"])
    return hb


@pytest.mark.parametrize("name", FIX)
def test_seed(name):
    fx = gu.load(name)
    ob = ol.OracleBatch(blank_seed(gu.stage_batch(fx, "seed")))
    ob.seed()
    assert ob.err == 0
    assert gu.compare_states(ob.hb, gu.stage_batch(fx, "seed"), ("active", "tse", "degree")) == []


STEPS = [
    ("seed", "c1", lambda o: o.cluster(0, 1.0, 2.0)),
    ("x1", "e2", lambda o: o.extrapolate_stage(2.0)),
    ("x2", "m2", lambda o: o.remove_state_metadata()),
    ("m2", "c3", lambda o: o.cluster(1, 1000.0, 100.0)),
]


@pytest.mark.parametrize("name", FIX)
@pytest.mark.parametrize("step", range(len(STEPS)))
def test_stage(name, step):
    prev, stage, fn = STEPS[step]
    fx = gu.load(name)
    ob = ol.OracleBatch(gu.stage_batch(fx, prev))
    fn(ob)
    assert ob.err == 0
    assert gu.compare_states(ob.hb, gu.stage_batch(fx, stage), ALL) == []


@pytest.mark.parametrize("name", FIX)
@pytest.mark.parametrize("prev,stage", [("c1", "x1"), ("e2", "x2"), ("c3", "x3")])
def test_extract(name, prev, stage):
    fx = gu.load(name)
    ob = ol.OracleBatch(gu.stage_batch(fx, prev))
    n, acc, pxy, pzr = ob.extract()
    assert np.array_equal(acc, fx[stage + "/accepted"])
    assert gu.compare_states(ob.hb, gu.stage_batch(fx, stage), ("alive",)) == []
    lab = fx[stage + "/cand_label"]
    roots = sorted(set(lab[lab >= 0].tolist()))
    want = fx[stage + "/pvals"]
    assert n == len(roots) == len(want)
    if roots:
        got = np.array([[pxy[r], pzr[r]] for r in roots])
        assert gu.rel_err(np.sort(got[:, 0]), np.sort(want[:, 0])) <= 1e-8   # p = Q(k/2, chi2/2) amplifies chi2's 1e-10
        assert gu.rel_err(np.sort(got[:, 1]), np.sort(want[:, 1])) <= 1e-8


@pytest.mark.parametrize("name", FIX + ["barrel100_cfg1", "barrel60_deg16"])
def test_full_schedule_chained(name, rtol=1e-7):
    """run_gnn_trackml_mod.sh:71-148 schedule, chained from the seeds: every decision bit-exact"""
    fx = gu.load(name)
    ob = ol.OracleBatch(blank_seed(gu.stage_batch(fx, "seed")))
    ob.seed()
    W = ("alive", "active", "merged", "degree")

    def chk(stage):
        assert gu.compare_states(ob.hb, gu.stage_batch(fx, stage), W, rtol=rtol) == [], stage

    chk("seed")
    ob.cluster(0, 1.0, 2.0)
    chk("c1")
    assert np.array_equal(ob.extract()[1], fx["x1/accepted"])
    chk("x1")
    ob.extrapolate_stage(2.0)
    chk("e2")
    assert np.array_equal(ob.extract()[1], fx["x2/accepted"])
    chk("x2")
    ob.remove_state_metadata()
    chk("m2")
    ob.cluster(1, 1000.0, 100.0)
    chk("c3")
    assert np.array_equal(ob.extract()[1], fx["x3/accepted"])
    chk("x3")
    assert ob.err == 0


def test_chi2_sf_matches_scipy():
    from scipy.stats import chi2
    L = ol.lib()
    rng = np.random.default_rng(1)
    for k in range(1, 20):
        for x in np.concatenate([rng.uniform(0, 60, 20), [0.0, 1e-9, 200.0]]):
            want = chi2.sf(x, k)
            got = L.gtfo_chi2_sf(float(x), float(k))
            assert abs(got - want) <= 1e-12 * max(want, 1e-300) + 1e-300 or abs(got - want) / want < 1e-10


def test_cca_matches_scipy():
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    fx = gu.load("barrel40_eta1")
    for stage in ("c1", "e2", "c3"):
        ob = ol.OracleBatch(gu.stage_batch(fx, stage))
        lab = ob.cca()
        h = ob.hb
        ex = gu.edge_exists(h) & (h["active"] != 0)
        N = ob.N
        nc, comp = connected_components(coo_matrix((np.ones(int(ex.sum())), (h["in_src"][ex], h["slot_dst"][ex])), shape=(N, N)),
                                        directed=True, connection="weak")
        mn = np.full(nc, N)
        np.minimum.at(mn, comp, np.arange(N))
        want = np.where(h["alive"] > 0, mn[comp], -1)
        inplay = gu.inplay_nodes(h)
        assert np.array_equal(lab[inplay], want[inplay])


def test_kl_distance_known_answer_from_reference_csv():
    """The reference's one shipped golden vector (1_events_training_data.csv, 7,574 KL values):
    the oracle's KLDistance on the reference-seeded components reproduces it as a sorted multiset.  With a true
    matrix product inside the trace the values are off by orders of magnitude, so this pins the element-wise
    trace (clustering.py:93)."""
    import ctypes
    fx = np.load(gu.GOLDEN + "/kl_parabolic_known_answer.npz")
    L = ol.lib()
    dp = ctypes.POINTER(ctypes.c_double)
    mean, cov, off = fx["mean"], fx["cov"], fx["off"]
    out = []
    for a, b in zip(off[:-1], off[1:]):
        for i in range(a, b):
            for j in range(a, i):
                mi, ci, mj, cj = (np.ascontiguousarray(x) for x in (mean[i], cov[i], mean[j], cov[j]))
                out.append(L.gtfo_kl_distance(mi.ctypes.data_as(dp), ci.ctypes.data_as(dp), mj.ctypes.data_as(dp),
                                              cj.ctypes.data_as(dp)))
    got = np.sort(np.array(out))
    want = fx["kl_sorted"]
    assert len(got) == len(want) == 7574
    assert gu.rel_err(got, want) <= 1e-9
    # the discriminating power of the fixture: a proper matrix product in the trace does NOT reproduce it
    alt = []
    for a, b in zip(off[:-1], off[1:]):
        if b - a < 3:
            continue
        c = cov[a:b].reshape(-1, 3, 3)
        inv = np.linalg.inv(c)
        for i in range(b - a):
            for j in range(i):
                d = mean[a + i] - mean[a + j]
                alt.append(np.trace((c[i] - c[j]) @ (inv[j] - inv[i])) + d @ (inv[i] + inv[j]) @ d)
        if len(alt) > 500:
            break
    alt = np.array(alt)
    # every element-wise value is in the CSV; the matrix-product values are not
    idx = np.searchsorted(want, alt)
    idx = np.clip(idx, 1, len(want) - 1)
    nearest = np.minimum(np.abs(want[idx] - alt), np.abs(want[idx - 1] - alt))
    assert (nearest > 1e-6 * np.abs(alt)).mean() > 0.5


# ---------------------------------------------------------------------------------------------- round-2 pins
@pytest.mark.parametrize("name", FIX + ["barrel100_cfg1", "barrel1000_cfg2", "lut_barrel40"])
def test_emp_var_vs_reference(name):
    """`xy_edge_gradient_mean_var[1]` (np.var of the xy edge gradients, helper.py:446): the input of the LUT bin"""
    fx = gu.load(name)
    ob = ol.OracleBatch(blank_seed(gu.stage_batch(fx, "seed")))
    ob.hb["emp_var"][:] = np.nan
    ob.seed()
    assert gu.rel_err(ob.hb["emp_var"], fx["topo_emp_var"]) <= 1e-9
    bins = lambda v: np.clip(np.floor(v / 0.05), 0, 27)    # noqa: E731  the decision the value feeds: bit-exact
    assert np.array_equal(bins(ob.hb["emp_var"]), bins(fx["topo_emp_var"]))


@pytest.mark.parametrize("prev,stage,key,chi2,table", [("seed", "c1lut", 0, 1.0, "lut"), ("seed", "c1str", 0, 1.0, "lut_stress"),
                                                       ("m2", "c3lut", 1, 1000.0, "lut"), ("m2", "c3str", 1, 1000.0, "lut_stress")])
def test_lut_mode_vs_reference_wrapper(prev, stage, key, chi2, table):
    """LUT-threshold mode against the reference's own cluster() with KL_threshold swapped per node
    (tests/golden/ref_harness.NodeKLThreshold, SURVEY.md 8c last row): shipped table and a table that bites"""
    fx = gu.load("lut_barrel40")
    ob = ol.OracleBatch(gu.stage_batch(fx, prev))
    ob.cluster(key, chi2, 123.0, lut=fx[table])           # the scalar threshold must be ignored in LUT mode
    assert ob.err == 0
    # the stress table absorbs up to 13 components per node: one merged `a` ends at 2.9e-8, three orders below the field's
    # typical magnitude, and carries 3e-8 of cancellation error -- measured against that magnitude (still 1e-9)
    assert gu.compare_states(ob.hb, gu.stage_batch(fx, stage), ALL, rtol=gu.RTOL, chained=(table == "lut_stress")) == []
    if table == "lut_stress":
        assert int(fx[stage + "_scalar_diff"]) > 0         # the per-node threshold changes decisions on this event


def test_load_lut_parses_the_reference_format(tmp_path):
    """`bin kl_min kl_max` per line (learn_KL_linear_model/create_lut/plot_lut.py:10-17); the fixture holds the values of
    the shipped learn_KL_linear_model/output/empvar/empvar.lut"""
    import os
    from gtf_b200 import stages
    fx = gu.load("lut_barrel40")
    p = tmp_path / "empvar.lut"
    p.write_text("".join("%d 0 %d\n" % (b, int(v)) for b, v in enumerate(fx["lut"])))
    assert np.array_equal(stages.load_lut(str(p)), fx["lut"])
    shipped = "/root/reference/learn_KL_linear_model/output/empvar/empvar.lut"
    if os.path.exists(shipped):                            # build container only
        assert np.array_equal(stages.load_lut(shipped), fx["lut"])
    assert fx["lut"].shape == (28,) and fx["lut"][0] == 13 and fx["lut"][2] == 32 and fx["lut"][27] == 0


def test_tag_propagation_vs_reference_script():
    """tag_propagation/tag_propagation.py:64-164 run unmodified (ref_harness.run_tag_propagation): final tags, number of
    sweeps; directed successor-only rule, isolated hits, stop at flipped/work <= 0.1 (not a fixed point)"""
    fx = gu.load("tagprop_barrel30")
    hb = gu.stage_batch(fx, "seed")         # topology only: every hit alive
    hb["alive"][:] = 1
    ob = ol.OracleBatch(hb)
    n, tags = ob.tag_propagation(fx["tags0"], 0.1)
    assert n == int(fx["sweeps"]) and n > 1
    assert np.array_equal(tags, fx["tags"])
    assert int((fx["tags"] != fx["tags0"]).sum()) > 100


def test_cfg2_size_schedule_chained_vs_reference():
    """BASELINE configs[1] size (1000 tracks, 10k hits, 100k directed edges) through the reference's own schedule,
    chained from the seeds: every decision and candidate set bit-exact (COMPACT fixture, 222 s of reference time)"""
    test_full_schedule_chained("barrel1000_cfg2")


def test_shipped_event_schedule_chained_vs_reference():
    """the reference's SHIPPED TrackML-derived event (volumes 7-9: 30,387 hits, 73,230 directed edges, mean in-degree 2.4,
    thousands of tiny sub-graphs) through the reference's own schedule: every decision and candidate set bit-exact.
    Values: everything within 1e-9 except merged_cov[1,1] of 16 nodes (<= 4.1e-7): the reference's close-pair merge
    (extract_track_candidates.py:101-118) writes the pair's midpoint into `GNN_Measurement` of a SHALLOW graph copy, i.e. into
    the object the remaining graph shares, so those hits move for every later extrapolation even when the candidate is
    rejected; the hit record here is immutable (not reproduced; no decision of this event depends on it)"""
    test_full_schedule_chained("shipped_vol79", rtol=1e-6)
