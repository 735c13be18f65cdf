"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C-ABI
(libgtf_b200.so via gtf_b200.EventBatch).

Gates (BASELINE.json north_star): activation bitmap, dict membership / order, merged-state existence,
candidate node sets: bit-exact.  States / covariances / weights / likelihoods / p-values: 1e-9 relative
per stage on identical inputs (chained runs accumulate through ill-conditioned 2x2 inverses: 1e-7)."""
import numpy as np
import pytest

import golden_util as gu
import oracle_lib as ol
import gtf_b200
from gtf_b200 import synth

pytestmark = pytest.mark.gpu

ALL = ("alive", "active", "merged", "tse", "uts", "degree", "edge_w")
FIX = ["barrel25_deg6", "barrel40_eta1"]


def gpu_batch(hb):
    return gtf_b200.EventBatch(hb)


def blank_seed(hb):
    hb = dict(hb)
    for f in list(hb):
        if f.startswith("tse_") and f != "tse_present":
            hb[f] = np.full_like(hb[f], np.nan)
    hb["tse_present"] = np.zeros_like(hb["tse_present"])
    hb["active"] = np.zeros_like(hb["active"])
    return hb


def state_of(b):
    hb = b.download()
    return hb


@pytest.mark.parametrize("name", FIX)
def test_seed_vs_reference(name):
    fx = gu.load(name)
    b = gpu_batch(blank_seed(gu.stage_batch(fx, "seed")))
    b.seed()
    assert gu.compare_states(state_of(b), gu.stage_batch(fx, "seed"), ("active", "tse", "degree")) == []


STEPS = [
    ("seed", "c1", lambda b: b.cluster(0, 1.0, 2.0), ALL),
    ("x1", "e2", lambda b: b.extrapolate_stage(2.0), ALL),
    ("x2", "m2", lambda b: b.remove_state_metadata(), ALL),
    ("m2", "c3", lambda b: b.cluster(1, 1000.0, 100.0), ALL),
]


@pytest.mark.parametrize("name", FIX)
@pytest.mark.parametrize("step", range(len(STEPS)))
def test_stage_vs_reference(name, step):
    """one reference stage on the reference's own previous state"""
    prev, stage, fn, what = STEPS[step]
    fx = gu.load(name)
    b = gpu_batch(gu.stage_batch(fx, prev))
    fn(b)
    assert gu.compare_states(state_of(b), gu.stage_batch(fx, stage), what) == []


@pytest.mark.parametrize("name", FIX)
def test_message_passing_then_reweight_separately(name):
    """the un-fused entry points compose to the same result as extrapolate_stage"""
    fx = gu.load(name)
    b = gpu_batch(gu.stage_batch(fx, "x1"))
    b.message_passing(2.0)
    for _ in range(2):
        b.compute_prior_probabilities("updated_track_states")
        b.reweight("updated_track_states")
    b.query_node_degree_in_edges()
    assert gu.compare_states(state_of(b), gu.stage_batch(fx, "e2"), ALL) == []


@pytest.mark.parametrize("name", FIX)
@pytest.mark.parametrize("prev,stage", [("c1", "x1"), ("e2", "x2"), ("c3", "x3")])
def test_extract_vs_reference(name, prev, stage):
    fx = gu.load(name)
    b = gpu_batch(gu.stage_batch(fx, prev))
    n, acc, pxy, pzr = b.extract()
    assert np.array_equal(acc, fx[stage + "/accepted"])
    assert gu.compare_states(state_of(b), gu.stage_batch(fx, stage), ("alive",)) == []
    lab = fx[stage + "/cand_label"]
    roots = sorted(set(lab[lab >= 0].tolist()))
    assert n == len(roots)
    want = fx[stage + "/pvals"]
    if len(roots):
        got = np.array([[pxy[r], pzr[r]] for r in roots])
        assert gu.rel_err(np.sort(got[:, 0]), np.sort(want[:, 0])) <= 1e-8
        assert gu.rel_err(np.sort(got[:, 1]), np.sort(want[:, 1])) <= 1e-8


@pytest.mark.parametrize("name", FIX + ["barrel100_cfg1", "barrel60_deg16"])
def test_full_schedule_vs_reference(name):
    """run_gnn_trackml_mod.sh schedule end to end on the GPU: decisions bit-exact at every stage"""
    fx = gu.load(name)
    b = gpu_batch(blank_seed(gu.stage_batch(fx, "seed")))
    W = ("alive", "active", "merged", "degree")
    b.seed()

    def chk(stage, rtol=1e-7):
        assert gu.compare_states(state_of(b), gu.stage_batch(fx, stage), W, rtol=rtol) == [], stage

    chk("seed")
    b.cluster("track_state_estimates", 1.0, 2.0)
    chk("c1")
    n, acc, _, _ = b.extract()
    assert np.array_equal(acc, fx["x1/accepted"])
    chk("x1")
    b.extrapolate_stage(2.0)
    chk("e2")
    n, acc, _, _ = b.extract()
    assert np.array_equal(acc, fx["x2/accepted"])
    chk("x2")
    b.remove_state_metadata()
    chk("m2")
    b.cluster("updated_track_states", 1000.0, 100.0)
    chk("c3")
    n, acc, _, _ = b.extract()
    assert np.array_equal(acc, fx["x3/accepted"])
    chk("x3")
    rows = b.candidates()
    total = (fx["x1/accepted"] | fx["x2/accepted"] | fx["x3/accepted"]).sum()
    assert len(rows) == total


def synth_batch(n_events, n_tracks, seed0, **kw):
    hbs = [synth.event_to_host(synth.barrel_event(n_tracks, seed=seed0 + i, **kw), i) for i in range(n_events)]
    hb = synth.concat_host_batches(hbs)
    hb.pop("truth")
    hb.pop("orig_id")
    return hb


@pytest.mark.parametrize("n_events,n_tracks,eta", [(3, 200, 0.5), (1, 1000, 1.0)])
def test_fused_iterate_vs_oracle(n_events, n_tracks, eta):
    """seed -> cluster -> 3 fused iterations on synthetic events (cfg2 shape) vs the oracle chain
    [message_passing, (prior, reweight) x2, cluster(updated states)]"""
    hb = synth_batch(n_events, n_tracks, 2000, eta_max=eta)
    ob = ol.OracleBatch(hb)
    ob.seed()
    ob.cluster(0, 1.0, 2.0)
    b = gpu_batch(hb)
    b.seed()
    b.cluster(0, 1.0, 2.0)
    assert gu.compare_states(state_of(b), ob.hb, ALL) == []
    for it in range(3):
        ob.extrapolate_stage(2.0)
        ob.cluster(1, 1000.0, 100.0)
        st = b.iterate(max_iter=1, stop_when_converged=False)
        bad = gu.compare_states(state_of(b), ob.hb, ALL, rtol=1e-7)
        assert bad == [], (it, bad)
        assert st[0]["active_edges"] == int((ob.hb["active"][gu.edge_exists(ob.hb)] == 1).sum())
    lab = b.CCA()
    assert np.array_equal(lab, ob.cca())


def same_state(a, c, rtol=1e-9):
    """integer / flag fields identical, value fields within rtol (two separately compiled kernels may contract their
    multiply-adds differently); returns the offending fields"""
    bad = []
    for f in a:
        x, y = np.asarray(a[f]), np.asarray(c[f])
        if x.dtype.kind != "f":
            if not np.array_equal(x, y):
                bad.append(f)
            continue
        nx, ny = np.isnan(x), np.isnan(y)
        if not np.array_equal(nx, ny):
            bad.append(f + " (NaN pattern)")
            continue
        xv = x[~nx]
        d = np.abs(xv - y[~ny])
        if d.size:
            ref = np.maximum(np.abs(xv), np.median(np.abs(xv)))      # (components crossing zero: the field's own magnitude)
            ref = np.where(ref > 0, ref, 1.0)
            if not (d <= rtol * ref).all():
                bad.append("%s (max rel %.3g)" % (f, float((d / ref).max())))
    return bad


def test_fused_send_execute_kernel_equals_the_two_kernel_path(monkeypatch):
    """GTF_FUSED_SX=1: k_send + k_exec as the one warp-specialised kernel k_sx (scanning / executing warp groups, message ring
    in shared memory).  Same iterations on the same event through both paths: every flag / order / counter identical, values
    to 1e-9 (measured: 1e-11); and the fused path against the oracle.  A tile whose every out-edge carries a message (first iteration: all
    edges active) fills the ring to its bound."""
    hb = synth_batch(3, 400, 2050, eta_max=1.0)
    ob = ol.OracleBatch(hb)
    ob.seed()
    ob.cluster(0, 1.0, 2.0)
    ref = gpu_batch(hb)
    monkeypatch.setenv("GTF_FUSED_SX", "1")
    fus = gpu_batch(hb)
    monkeypatch.delenv("GTF_FUSED_SX")
    for b in (ref, fus):
        b.seed()
        b.cluster(0, 1.0, 2.0)
    assert fus.iteration_launches() == 0
    for it in range(4):
        ob.extrapolate_stage(2.0)
        ob.cluster(1, 1000.0, 100.0)
        s_ref = ref.iterate(max_iter=1, stop_when_converged=False)[0]
        l0 = fus.iteration_launches()
        s_fus = fus.iterate(max_iter=1, stop_when_converged=False)[0]
        assert fus.iteration_launches() - l0 == 9       # k_begin, k_sx, k_node2, k_hv x 4, k_big, k_iter_end
        assert s_ref == s_fus, (it, s_ref, s_fus)
        c = state_of(fus)
        assert same_state(state_of(ref), c) == [], it
        bad = gu.compare_states(c, ob.hb, ALL, rtol=1e-7)
        assert bad == [], (it, bad)
    # uncommitted passes and a longer committed loop through the fused kernel
    d1 = fus.iterate_dry(want_stats=True)
    assert d1 == ref.iterate_dry(want_stats=True)
    assert fus.iterate(max_iter=6)[-1] == ref.iterate(max_iter=6)[-1]
    assert same_state(state_of(ref), state_of(fus)) == []


def test_iterate_dry_is_idempotent_and_matches_commit():
    hb = synth_batch(2, 300, 2100)
    b = gpu_batch(hb)
    b.seed()
    b.cluster(0, 1.0, 2.0)
    s1 = b.iterate_dry(want_stats=True)
    before = state_of(b)
    s2 = b.iterate_dry(want_stats=True)
    assert s1 == s2
    after = state_of(b)
    for f in ("active", "has_merged", "m_a", "m_p11"):
        assert np.array_equal(before[f], after[f], equal_nan=True)
    s3 = b.iterate(max_iter=1, stop_when_converged=False)[0]
    assert s3 == s1


@pytest.mark.parametrize("stop,graph", [(False, "1"), (True, "1"), (False, "0")])
def test_device_side_loop_equals_single_iterations(stop, graph, monkeypatch):
    """gtf_iterate(max_iter = 8) queues its iterations on the device (k_iter_end files the counters and raises the stop flag;
    from the second iteration on the few out-edges still active are sent by k_send_sparse from the compacted lists) -- against
    the same iterations issued one call at a time (one read-back each, tiled k_send throughout) and against the oracle:
    same counters per iteration, same number of iterations, identical flags and orders, values to 1e-12."""
    hb = synth_batch(3, 400, 2150, eta_max=1.0)
    one = gpu_batch(hb)
    monkeypatch.setenv("GTF_GRAPH", graph)           # 0: plain launches instead of graph replays (read when the batch is created)
    loop = gpu_batch(hb)
    monkeypatch.delenv("GTF_GRAPH")
    ob = ol.OracleBatch(hb)
    ob.seed()
    ob.cluster(0, 1.0, 2.0)
    for b in (one, loop):
        b.seed()
        b.cluster(0, 1.0, 2.0)
    singles = []
    for it in range(8):
        singles.append(one.iterate(max_iter=1, stop_when_converged=False)[0])
        if stop and singles[-1]["active_changed"] == 0:
            break
    st = loop.iterate(max_iter=8, stop_when_converged=stop)
    assert st == singles
    assert same_state(state_of(one), state_of(loop), rtol=1e-12) == []
    for _ in range(len(st)):
        ob.extrapolate_stage(2.0)
        ob.cluster(1, 1000.0, 100.0)
    assert gu.compare_states(state_of(loop), ob.hb, ALL, rtol=1e-7) == []
    # the loop left no flag behind: a single iteration, an uncommitted pass and a second loop behave as before
    assert loop.iterate(max_iter=1, stop_when_converged=False) == one.iterate(max_iter=1, stop_when_converged=False)
    assert loop.iterate_dry(want_stats=True) == one.iterate_dry(want_stats=True)
    assert loop.iterate(max_iter=4, stop_when_converged=False) == [one.iterate(max_iter=1, stop_when_converged=False)[0] for _ in range(4)]
    assert same_state(state_of(one), state_of(loop), rtol=1e-12) == []


def test_tag_propagation_vs_oracle():
    hb = synth_batch(2, 200, 2200)
    ob = ol.OracleBatch(hb)
    ob.seed()
    b = gpu_batch(hb)
    b.seed()
    tags0 = np.arange(len(hb["x"]), dtype=np.int32)
    n_o, t_o = ob.tag_propagation(tags0)
    n_g, t_g = b.tag_propagation(tags0)
    assert n_o == n_g and np.array_equal(t_o, t_g)


def test_converged_loop_and_candidates_vs_oracle():
    hb = synth_batch(2, 150, 2300)
    ob = ol.OracleBatch(hb)
    ob.seed()
    ob.cluster(0, 1.0, 2.0)
    b = gpu_batch(hb)
    b.seed()
    b.cluster(0, 1.0, 2.0)
    stats = b.iterate(max_iter=10, stop_when_converged=True)
    for _ in range(len(stats)):
        ob.extrapolate_stage(2.0)
        ob.cluster(1, 1000.0, 100.0)
    assert gu.compare_states(state_of(b), ob.hb, ("active", "merged", "degree"), rtol=1e-7) == []
    n_o, acc_o, pxy_o, pzr_o = ob.extract()
    n_g, acc_g, pxy_g, pzr_g = b.extract()
    assert n_o == n_g and np.array_equal(acc_o, acc_g)
    assert gu.rel_err(pxy_g, pxy_o) <= 1e-7 and gu.rel_err(pzr_g, pzr_o) <= 1e-7


def test_iterate_after_extraction_vs_oracle():
    """extraction removes the accepted candidates' nodes: the packed iteration afterwards must treat their edges as
    non-existing (existing-edge bitmap, one-node sub-graphs, fragments) exactly like the reference's graph surgery"""
    hb = synth_batch(2, 200, 2350, eta_max=1.0)
    ob = ol.OracleBatch(hb)
    ob.seed()
    ob.cluster(0, 1.0, 2.0)
    b = gpu_batch(hb)
    b.seed()
    b.cluster(0, 1.0, 2.0)
    # nodes that lose every neighbour make the reference itself raise (1/len({}), helper.py:90) AFTER it has mutated the
    # graph; both sides record the condition and carry on, and the states must still agree
    b.raise_ref_errors = False
    removed = 0
    for rnd in range(3):
        for _ in range(2):
            ob.extrapolate_stage(2.0)
            ob.cluster(1, 1000.0, 100.0)
        b.iterate(max_iter=2, stop_when_converged=False)
        n_o, acc_o, _, _ = ob.extract()
        n_g, acc_g, _, _ = b.extract()
        assert n_o == n_g and np.array_equal(acc_o, acc_g), rnd
        removed += int(acc_g.sum())
        ob.remove_state_metadata()      # pops the dict entries of removed neighbours, like the reference's schedule
        b.remove_state_metadata()       # (run_gnn_trackml_mod.sh:131-139); without it the reference raises KeyError
        assert gu.compare_states(state_of(b), ob.hb, ("alive", "active", "merged", "uts", "degree"), rtol=1e-7) == [], rnd
    assert removed > 0
    ob.extrapolate_stage(2.0)
    ob.cluster(1, 1000.0, 100.0)
    b.iterate(max_iter=1, stop_when_converged=False)
    assert gu.compare_states(state_of(b), ob.hb, ("alive", "active", "merged", "uts", "degree"), rtol=1e-7) == []
    assert np.array_equal(b.CCA(), ob.cca())


def test_stage_wrappers_on_networkx_graphs():
    """the reference-signature wrappers (gtf_b200.stages) on nx.DiGraph lists give the same graphs as the
    device-resident EventBatch path"""
    from gtf_b200 import nxio, stages
    fx = gu.load("barrel25_deg6")
    hb = gu.stage_batch(fx, "x1")
    graphs = nxio.host_to_graphs(hb, orig_id=fx["topo_orig_id"], truth=fx["topo_truth"])
    stages.message_passing(graphs, 2.0, 0.3, 0.4, 0.6, 550.0)
    for _ in range(2):
        stages.compute_prior_probabilities(graphs, "updated_track_states")
        stages.reweight(graphs, "updated_track_states")
    got = nxio.graphs_to_host(graphs)
    b = gpu_batch(hb)
    b.message_passing(2.0)
    for _ in range(2):
        b.compute_prior_probabilities("updated_track_states")
        b.reweight("updated_track_states")
    want = nxio.graphs_to_host(nxio.host_to_graphs(b.download(), orig_id=fx["topo_orig_id"], truth=fx["topo_truth"]))
    for k in want:
        if k == "degree":
            continue
        assert np.array_equal(got[k], want[k], equal_nan=True), k
    parts = stages.CCA(graphs[0].copy())
    assert sum(p.number_of_nodes() for p in parts) == graphs[0].number_of_nodes()


def test_lut_threshold_mode_vs_oracle():
    """per-node KL threshold from the emp_var bin (the hook the reference never wired in)"""
    hb = synth_batch(1, 300, 2400)
    lut = np.where(np.arange(28) % 2 == 1, 1e9, 1e-9)       # alternately absorb everything / nothing
    ob = ol.OracleBatch(hb)
    ob.seed()
    ob.cluster(0, 1.0, 2.0, lut=lut)
    b = gpu_batch(hb)
    b.seed()
    b.cluster(0, 1.0, 2.0, KL_lut=lut)
    assert gu.compare_states(state_of(b), ob.hb, ALL) == []
    ob2 = ol.OracleBatch(hb)
    ob2.seed()
    ob2.cluster(0, 1.0, 2.0)
    assert not np.array_equal(ob2.hb["active"], ob.hb["active"])     # the LUT really changes decisions


def test_driver_schedules_run_and_agree_with_oracle():
    from gtf_b200 import driver
    hb = synth_batch(2, 120, 2500, eta_max=1.0)
    b = gpu_batch(hb)
    res = driver.reference_schedule(b)
    ob = ol.OracleBatch(hb)
    ob.seed()
    ob.cluster(0, 1.0, 2.0)
    a1 = ob.extract()[1]
    ob.extrapolate_stage(2.0)
    a2 = ob.extract()[1]
    ob.remove_state_metadata()
    ob.cluster(1, 1000.0, 100.0)
    a3 = ob.extract()[1]
    for (n, acc), want in zip(res, (a1, a2, a3)):
        assert np.array_equal(acc, want)
    rows = b.candidates()
    assert len(rows) == int((a1 | a2 | a3).sum())
    assert set(rows[:, 0].tolist()) <= {0, 1}


def test_packed_layout_stays_in_sync_with_the_fields():
    """the iteration runs on packed records / bitmaps; uploads, downloads and per-stage calls in between must see
    (and be seen by) the same state: iterate -> per-stage call -> upload -> iterate, against the oracle"""
    hb = synth_batch(2, 300, 2600)
    ob = ol.OracleBatch(hb)
    ob.seed()
    ob.cluster(0, 1.0, 2.0)
    b = gpu_batch(hb)
    b.seed()
    b.cluster(0, 1.0, 2.0)

    def oracle_iter():
        ob.extrapolate_stage(2.0)
        ob.cluster(1, 1000.0, 100.0)

    oracle_iter()
    b.iterate(max_iter=1, stop_when_converged=False)
    # a per-stage call on the fields right after a packed iteration
    ob.remove_state_metadata()
    b.remove_state_metadata()
    assert gu.compare_states(state_of(b), ob.hb, ALL, rtol=1e-7) == []
    oracle_iter()
    b.iterate(max_iter=1, stop_when_converged=False)
    # partial downloads (one group at a time), then an upload that edits the activation flags and one weight array
    act = b.download(["active"])["active"].copy()
    w = b.download(["uts_w"])["uts_w"].copy()
    kill = np.flatnonzero(act == 1)[::7]
    act[kill] = 0
    ob.hb["active"][kill] = 0
    b.upload({"active": act, "uts_w": w})
    ma = b.download(["m_a"])["m_a"].copy()
    b.upload({"m_a": ma})
    oracle_iter()
    b.iterate(max_iter=1, stop_when_converged=False)
    assert gu.compare_states(state_of(b), ob.hb, ALL, rtol=1e-7) == []
    assert np.array_equal(b.CCA(), ob.cca())


def test_copies_of_an_event_inside_a_large_batch_evolve_identically():
    """size-independent property at batch scale: events are independent, so the copies of one generated event inside a
    tiled batch (the benchmark's construction) must end in BIT-identical states whatever tile / CTA / list position
    they land on, and one copy must match the oracle run on that event alone"""
    import bench
    n_ev, distinct, tracks = 48, 3, 400
    hb = bench.build_batch(n_ev, tracks, 5000, distinct)
    b = gpu_batch(hb)
    b.seed()
    b.cluster(0, 1.0, 2.0)
    stats = b.iterate(max_iter=3, stop_when_converged=False)
    out = state_of(b)
    # node / slot range of every event (events are concatenated)
    first_sub = np.searchsorted(out["sub_event"], np.arange(n_ev + 1))
    node0 = out["sub_off"][first_sub]
    slot0 = out["in_off"][node0]
    node_fields = ["has_merged", "degree", "has_uts", "uts_next"] + ["m_" + k for k in ("a", "b", "c", "p00", "p01", "p11", "p22", "prior")]
    slot_fields = ["active", "uts_present", "uts_rank", "uts_side", "edge_w"] + \
                  ["uts_" + k for k in ("a", "b", "c", "tau", "p00", "p01", "p11", "p22", "lik", "prior", "w", "lrn")]
    for e in range(distinct, n_ev):
        r = e % distinct
        for f in node_fields:
            assert np.array_equal(out[f][node0[e]:node0[e + 1]], out[f][node0[r]:node0[r + 1]], equal_nan=True), (e, f)
        for f in slot_fields:
            assert np.array_equal(out[f][slot0[e]:slot0[e + 1]], out[f][slot0[r]:slot0[r + 1]], equal_nan=True), (e, f)
    for k in ("nodes_merged", "edges_sent", "edges_gated", "active_edges"):
        assert stats[-1][k] % (n_ev // distinct) == 0
    # and the first event against the oracle
    one = bench.build_batch(1, tracks, 5000, 1)
    ob = ol.OracleBatch(one)
    ob.seed()
    ob.cluster(0, 1.0, 2.0)
    for _ in range(3):
        ob.extrapolate_stage(2.0)
        ob.cluster(1, 1000.0, 100.0)
    n1, s1 = int(node0[1]), int(slot0[1])
    assert np.array_equal(out["active"][:s1], ob.hb["active"])
    assert np.array_equal(out["has_merged"][:n1], ob.hb["has_merged"])
    assert np.array_equal(out["uts_present"][:s1], ob.hb["uts_present"])
    m = ob.hb["has_merged"] > 0
    assert gu.rel_err(out["m_a"][:n1][m], ob.hb["m_a"][m], gu.field_floor(ob.hb["m_a"][m])) <= 1e-7


def test_gate_chi2_is_recorded_on_request():
    """uts_chi2 is a diagnostic (the reference appends it to a CSV): the fused iteration stores it only when asked"""
    hb = synth_batch(1, 300, 2650)
    ob = ol.OracleBatch(hb)
    ob.seed()
    ob.cluster(0, 1.0, 2.0)
    ob.message_passing(2.0)
    b = gpu_batch(hb)
    b.seed()
    b.cluster(0, 1.0, 2.0)
    b.iterate(max_iter=1, stop_when_converged=False, record_chi2=True)
    got, want = b.download(["uts_chi2"])["uts_chi2"], ob.hb["uts_chi2"]
    sent = want != 0
    assert sent.sum() > 1000
    assert gu.rel_err(got[sent], want[sent]) <= 1e-9


def test_cfg3_high_pileup_event_vs_oracle():
    """BASELINE configs[2]: one synthetic high-pileup event (10k tracks -> 100k hits, 1M directed edges):
    seed, cluster, two fused iterations, components -- every decision bit-exact against the oracle"""
    hb = synth_batch(1, 10000, 3300)
    assert len(hb["in_src"]) == 1_000_000
    ob = ol.OracleBatch(hb)
    ob.seed()
    ob.cluster(0, 1.0, 2.0)
    b = gpu_batch(hb)
    b.seed()
    b.cluster(0, 1.0, 2.0)
    # 1e-9, with components that cross zero measured against the field's typical magnitude (a seed parameter
    # a = h02 * m_B inherits the ABSOLUTE rounding of the rotated coordinate m_B, ~1e-13 mm)
    assert gu.compare_states(state_of(b), ob.hb, ALL, rtol=gu.RTOL, chained=True) == []
    for it in range(2):
        ob.extrapolate_stage(2.0)
        ob.cluster(1, 1000.0, 100.0)
        b.iterate(max_iter=1, stop_when_converged=False)
        assert gu.compare_states(state_of(b), ob.hb, ("active", "merged", "uts", "degree"), rtol=1e-7) == [], it
    assert np.array_equal(b.CCA(), ob.cca())


@pytest.mark.parametrize("deg", [1, 2, 4, 8])
def test_cfg5_degree_sweep_vs_oracle(deg):
    """BASELINE configs[4]: mixture components per node 1..8 (d < 3 exercises the clustering skip path)"""
    hb = synth_batch(1, 500, 3400 + deg, target_degree=float(deg))
    ob = ol.OracleBatch(hb)
    ob.seed()
    ob.cluster(0, 1.0, 2.0)
    ob.extrapolate_stage(2.0)
    ob.cluster(1, 1000.0, 100.0)
    b = gpu_batch(hb)
    b.seed()
    b.cluster(0, 1.0, 2.0)
    b.iterate(max_iter=1, stop_when_converged=False)
    assert gu.compare_states(state_of(b), ob.hb, ALL, rtol=1e-7) == []


def dict_orders_equal(a, b):
    return gu.dict_order(a) == gu.dict_order(b)


def test_pairwise_kl_kernel_reproduces_the_reference_golden_csv():
    """the reference's one shipped known-answer file (7,574 KL values of its LUT training-data generator)
    through the GPU kernel behind gtf_kl_pairs / stages.KLDistance_pairs: same values as a sorted multiset (1e-9), and
    pair for pair against the oracle"""
    import ctypes
    from gtf_b200 import stages
    fx = np.load(gu.GOLDEN + "/kl_parabolic_known_answer.npz")
    mean, cov, off = fx["mean"], fx["cov"].reshape(-1, 9), fx["off"]
    got = stages.KLDistance_pairs(mean, cov, off)
    assert len(got) == 7574
    assert gu.rel_err(np.sort(got), fx["kl_sorted"]) <= 1e-9
    L = ol.lib()
    dp = ctypes.POINTER(ctypes.c_double)
    want = []
    for a, b in zip(off[:-1], off[1:]):
        for i in range(a, b):
            for j in range(a, i):
                mi, ci, mj, cj = (np.ascontiguousarray(x) for x in (mean[i], cov[i], mean[j], cov[j]))
                want.append(L.gtfo_kl_distance(mi.ctypes.data_as(dp), ci.ctypes.data_as(dp), mj.ctypes.data_as(dp),
                                               cj.ctypes.data_as(dp)))
    assert gu.rel_err(got, np.array(want)) <= 1e-9
    assert len(stages.KLDistance_pairs(mean[:0], cov[:0], np.zeros(1, np.int32))) == 0
