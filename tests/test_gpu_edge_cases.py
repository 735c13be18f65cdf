"""GPU edge cases through the C-ABI, checked against the oracle: empty / ragged batches, isolated nodes, exact ties
and duplicate hits (np.where / np.nonzero semantics), degenerate geometry (inf / NaN like numpy), reference-side
error conditions, tile overflow."""
import numpy as np
import pytest

import golden_util as gu
import oracle_lib as ol
import gtf_b200
from gtf_b200 import synth, lib as L

pytestmark = pytest.mark.gpu
ALL = ("alive", "active", "merged", "tse", "uts", "degree", "edge_w")


def tiny_event(xyz, layers, pairs):
    xyz = np.asarray(xyz, float)
    ev = {"x": xyz[:, 0], "y": xyz[:, 1], "z": xyz[:, 2], "r": np.hypot(xyz[:, 0], xyz[:, 1]),
          "layer": np.asarray(layers, np.int32), "volume": np.full(len(xyz), 8, np.int32),
          "truth": np.arange(len(xyz), dtype=np.int64),
          "edge_a": np.array([p[0] for p in pairs], np.int32), "edge_b": np.array([p[1] for p in pairs], np.int32)}
    hb = synth.event_to_host(ev)
    hb.pop("truth")
    hb.pop("orig_id")
    return hb


def run_both(hb, raise_ref=False):
    ob = ol.OracleBatch(hb)
    ob.seed()
    ob.cluster(0, 1.0, 2.0)
    ob.extrapolate_stage(2.0)
    ob.cluster(1, 1000.0, 100.0)
    b = gtf_b200.EventBatch(hb, raise_ref_errors=raise_ref)
    b.seed()
    errs = 0
    errs |= b.cluster(0, 1.0, 2.0)["ref_errors"]
    errs |= b.iterate(max_iter=1, stop_when_converged=False)[0]["ref_errors"]
    return ob, b, errs


def test_empty_batch():
    hb = {k: np.zeros(0, v) for k, v in (("x", float), ("y", float), ("z", float), ("r", float), ("layer", np.int32),
                                           ("volume", np.int32), ("sub", np.int32), ("alive", np.uint8),
                                           ("in_src", np.int32), ("slot_dst", np.int32), ("out_slot", np.int32),
                                           ("rev_slot", np.int32), ("sub_state", np.uint8), ("sub_event", np.int32))}
    hb["in_off"] = np.zeros(1, np.int32)
    hb["out_off"] = np.zeros(1, np.int32)
    hb["sub_off"] = np.zeros(1, np.int32)
    b = gtf_b200.EventBatch(hb)
    b.seed()
    b.cluster(0, 1.0, 2.0)
    assert b.iterate(max_iter=2) is not None
    n, acc, _, _ = b.extract()
    assert n == 0 and len(b.candidates()) == 0 and len(b.CCA()) == 0


def test_isolated_nodes_and_ragged_subgraphs():
    # two isolated hits, one 2-hit graph, one star with 4 neighbours
    xyz = [(30, 1, 0), (70, -2, 5), (32, 0.5, 1), (72, 1.5, 3), (116, 2, 4), (32.5, -1, 0), (71, 0, 2), (115, 1, 5), (170, 3, 8)]
    hb = tiny_event(xyz, [2, 4, 2, 4, 6, 2, 4, 6, 8], [(2, 3), (6, 5), (6, 7), (6, 8), (6, 4)])
    assert len(hb["sub_off"]) - 1 == 4
    ob, b, errs = run_both(hb)
    assert errs == ob.err == 0
    assert gu.compare_states(b.download(), ob.hb, ALL, rtol=1e-7) == []
    assert np.array_equal(b.CCA(), ob.cca())


def test_duplicate_hits_zero_chi2_and_ties():
    """two neighbours at identical coordinates: their pairwise chi2 is exactly 0 (dropped by np.nonzero,
    clustering.py:119) and their chi2 to any third component ties exactly (np.where returns both, :122-123)"""
    centre = (116.0, 3.0, 10.0)
    nb = [(72.0, 1.0, 6.0), (72.0, 1.0, 6.0), (172.0, 5.5, 15.0), (72.5, 2.5, 6.2), (171.0, 4.0, 14.5)]
    xyz = [centre] + nb
    hb = tiny_event(xyz, [6, 4, 4, 8, 4, 8], [(0, k) for k in range(1, 6)])
    ob, b, errs = run_both(hb)
    assert errs == ob.err
    assert gu.compare_states(b.download(), ob.hb, ("active", "merged", "tse", "degree"), rtol=1e-7) == []


def test_all_components_identical_raises_like_reference():
    """every pairwise chi2 == 0 -> np.min([]) -> ValueError in the reference (clustering.py:120)"""
    xyz = [(116.0, 3.0, 10.0)] + [(72.0, 1.0, 6.0)] * 3
    hb = tiny_event(xyz, [6, 4, 4, 4], [(0, 1), (0, 2), (0, 3)])
    ob = ol.OracleBatch(hb)
    ob.seed()
    ob.cluster(0, 1.0, 2.0)
    assert ob.err & 1
    b = gtf_b200.EventBatch(hb)
    b.seed()
    with pytest.raises(ValueError):
        b.cluster(0, 1.0, 2.0)


def test_degenerate_geometry_matches_numpy_semantics():
    """a neighbour at the same radius (dr = 0 -> tau = +-inf, var(tau) = inf / NaN): non-finite values must appear in
    the same places and comparisons with NaN must behave like numpy (no cluster, gate fails).
    (A neighbour exactly on the node's local y axis makes the 3-point design matrix singular: the reference raises
    LinAlgError there, helper.py:384 -- undefined behaviour, not tested.)"""
    c = (100.0, 0.0, 5.0)
    r = 100.0
    same_r = (r * np.cos(0.3), r * np.sin(0.3), 9.0)
    xyz = [c, same_r, (60.0, 1.0, 3.0), (140.0, -1.0, 7.0), (101.0, 30.0, 6.0), (61.0, -2.0, 2.5)]
    hb = tiny_event(xyz, [6, 6, 4, 8, 6, 4], [(0, k) for k in range(1, 6)])
    ob, b, errs = run_both(hb)
    g = b.download()
    for f in ("tse_tau", "tse_p22", "tse_a"):
        assert np.array_equal(np.isnan(g[f]), np.isnan(ob.hb[f])), f
        assert np.array_equal(np.isinf(g[f]), np.isinf(ob.hb[f])), f
    assert np.array_equal(g["active"], ob.hb["active"])
    assert np.array_equal(g["has_merged"], ob.hb["has_merged"])
    assert errs == ob.err


def test_degree_beyond_tile_is_rejected():
    n = 800     # more in-slots than one tile holds
    ang = np.linspace(0, 0.5, n)
    xyz = [(100.0, 0.0, 0.0)] + [(60.0 * np.cos(a), 60.0 * np.sin(a), 1.0) for a in ang]
    hb = tiny_event(xyz, [6] + [4] * n, [(0, k) for k in range(1, n + 1)])
    with pytest.raises(L.GtfError, match="GTF_TILE_SLOTS"):
        gtf_b200.EventBatch(hb)


def test_wide_nodes_take_the_generic_path():
    """in-degree 33..200 (> one warp): generic shared-memory node program, same results as the oracle"""
    rng = np.random.default_rng(3)
    n = 120
    ang = rng.uniform(-0.05, 0.05, n)
    rr = np.where(np.arange(n) % 2 == 0, 72.0, 172.0)
    xyz = [(116.0, 0.0, 10.0)] + [(rr[k] * np.cos(ang[k]), rr[k] * np.sin(ang[k]), 10.0 * rr[k] / 116.0 + rng.normal(0, 0.3))
                                  for k in range(n)]
    hb = tiny_event(xyz, [6] + [4 if k % 2 == 0 else 8 for k in range(n)], [(0, k) for k in range(1, n + 1)])
    ob, b, errs = run_both(hb)
    assert errs == ob.err == 0
    assert gu.compare_states(b.download(), ob.hb, ALL, rtol=1e-7) == []


@pytest.mark.parametrize("seed", [11, 22, 33, 44])
def test_randomised_iteration_plans_vs_oracle(seed):
    """random event shape / in-degree (wide nodes reach k_big) and a random way of issuing five committed iterations
    (singly, as one burst, or mixed with uncommitted passes, partial downloads and the asynchronous call): same state
    as the oracle's five iterations (tools/stress_parity.py is the long form of this test)"""
    rng = np.random.default_rng(seed)
    n_ev, tracks = int(rng.integers(1, 4)), int(rng.choice([60, 150]))
    deg, eta = float(rng.choice([3.0, 10.0, 16.0, 28.0])), float(rng.choice([0.5, 1.0]))
    hbs = [synth.event_to_host(synth.barrel_event(tracks, seed=int(rng.integers(1, 10**6)), eta_max=eta, target_degree=deg), e)
           for e in range(n_ev)]
    hb = synth.concat_host_batches(hbs)
    hb.pop("truth")
    hb.pop("orig_id")
    ob = ol.OracleBatch(hb)
    ob.seed()
    ob.cluster(0, 1.0, 2.0)
    b = gtf_b200.EventBatch(hb, raise_ref_errors=False)
    b.seed()
    b.cluster(0, 1.0, 2.0)
    for _ in range(5):
        ob.extrapolate_stage(2.0)
        ob.cluster(1, 1000.0, 100.0)
    plan = rng.choice(["single", "burst", "mixed"])
    if plan == "single":
        for _ in range(5):
            b.iterate(max_iter=1, stop_when_converged=False)
    elif plan == "burst":
        b.iterate(max_iter=5, stop_when_converged=False)
    else:
        b.iterate(max_iter=2, stop_when_converged=False)
        b.iterate_dry()
        b.download(["uts_w", "active"])
        b.iterate(max_iter=1, stop_when_converged=False, want_stats=False)
        b.iterate_dry()
        b.iterate(max_iter=2, stop_when_converged=False)
    assert gu.compare_states(b.download(), ob.hb, ("alive", "active", "merged", "uts", "degree", "edge_w"), rtol=1e-7) == []
    assert np.array_equal(b.CCA(), ob.cca())


def test_cfg1_toy_event_nan_regime_vs_oracle():
    """BASELINE configs[0]: the 2-D toy event (z = r = 0 for every hit) through the 3-D path.  tau = dz / dr is 0 / 0, so every
    slope, its variance and every KL are NaN: numpy semantics make each comparison false -- no cluster forms, no merged state,
    no message -- and the device must land on exactly the oracle's state (NaN patterns included), without raising."""
    from gtf_b200 import synth
    ev = synth.toy_event(seed=3)
    hb = synth.event_to_host(ev)
    hb.pop("truth")
    hb.pop("orig_id")
    ob = ol.OracleBatch(hb)
    ob.seed()
    ob.cluster(0, 1.0, 2.0)
    b = gtf_b200.EventBatch(hb)
    b.seed()
    st = b.cluster(0, 1.0, 2.0)
    assert st["nodes_merged"] == 0 and st["ref_errors"] == 0
    what = ("active", "merged", "tse", "uts", "degree", "edge_w")
    assert gu.compare_states(b.download(), ob.hb, what) == []
    got = b.download(["tse_tau", "tse_p22", "tse_a", "tse_present"])
    assert np.isnan(got["tse_tau"][got["tse_present"] > 0]).all() and np.isfinite(got["tse_a"][got["tse_present"] > 0]).all()
    for _ in range(2):
        ob.extrapolate_stage(2.0)
        ob.cluster(1, 1000.0, 100.0)
        s = b.iterate(max_iter=1, stop_when_converged=False)[0]
        assert s["edges_sent"] == 0 and s["active_edges"] == len(hb["in_src"]) and s["active_changed"] == 0
        assert gu.compare_states(b.download(), ob.hb, what, rtol=1e-7) == []
    assert np.array_equal(b.CCA(), ob.cca())
