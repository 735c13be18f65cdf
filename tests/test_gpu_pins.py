"""GPU parity against the round-2 reference pins (tests/golden/make_lut_tag_golden.py, make_golden.py barrel1000_cfg2):
LUT-threshold mode vs the reference's cluster() with a per-node KL_threshold, emp_var vs helper.py:446, tag propagation
vs the unmodified tag_propagation.py, a cfg2-size event through the reference's own schedule.  All through the C-ABI."""
import numpy as np
import pytest

import golden_util as gu
import gtf_b200
from test_gpu_parity import ALL, blank_seed, state_of

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["barrel25_deg6", "barrel40_eta1", "barrel1000_cfg2", "lut_barrel40"])
def test_emp_var_vs_reference(name):
    fx = gu.load(name)
    hb = blank_seed(gu.stage_batch(fx, "seed"))
    hb["emp_var"] = np.full_like(hb["emp_var"], np.nan)
    b = gtf_b200.EventBatch(hb)
    b.seed()
    got = b.download(["emp_var"])["emp_var"]
    assert gu.rel_err(got, fx["topo_emp_var"]) <= 1e-9
    bins = lambda v: np.clip(np.floor(v / 0.05), 0, 27)    # noqa: E731
    assert np.array_equal(bins(got), bins(fx["topo_emp_var"]))


@pytest.mark.parametrize("prev,stage,key,chi2,table", [("seed", "c1lut", 0, 1.0, "lut"), ("seed", "c1str", 0, 1.0, "lut_stress"),
                                                       ("m2", "c3lut", 1, 1000.0, "lut"), ("m2", "c3str", 1, 1000.0, "lut_stress")])
def test_lut_mode_vs_reference_wrapper(prev, stage, key, chi2, table):
    fx = gu.load("lut_barrel40")
    b = gtf_b200.EventBatch(gu.stage_batch(fx, prev))
    b.cluster(key, chi2, 123.0, KL_lut=fx[table])
    # stress table: see tests/test_oracle_golden.py (one merged component three orders below the field's magnitude)
    assert gu.compare_states(state_of(b), gu.stage_batch(fx, stage), ALL, rtol=gu.RTOL, chained=(table == "lut_stress")) == []


def test_lut_mode_in_the_fused_iteration_vs_reference_wrapper():
    """LUT mode inside gtf_iterate (k_hv / k_big on the packed layout) = the per-stage kernels (pinned above against the
    reference wrapper) applied in the same order, on the reference's own post-cluster state"""
    fx = gu.load("lut_barrel40")
    hb = gu.stage_batch(fx, "c1str")
    a = gtf_b200.EventBatch(hb)
    a.iterate(max_iter=2, stop_when_converged=False, KL_lut=fx["lut_stress"])
    c = gtf_b200.EventBatch(hb)
    for _ in range(2):
        c.extrapolate_stage(2.0)
        c.cluster(1, 1000.0, 100.0, KL_lut=fx["lut_stress"])
    assert gu.compare_states(state_of(a), state_of(c), ALL, rtol=1e-7) == []


def test_tag_propagation_vs_reference_script():
    fx = gu.load("tagprop_barrel30")
    hb = gu.stage_batch(fx, "seed")
    hb["alive"][:] = 1
    b = gtf_b200.EventBatch(hb)
    n, tags = b.tag_propagation(fx["tags0"], 0.1)
    assert n == int(fx["sweeps"])
    assert np.array_equal(tags, fx["tags"])


def test_cfg2_size_schedule_vs_reference():
    """BASELINE configs[1] size through run_gnn_trackml_mod.sh's schedule, chained from the seeds on the GPU, against the
    unmodified reference: every decision and candidate set bit-exact"""
    from test_gpu_parity import test_full_schedule_vs_reference
    test_full_schedule_vs_reference("barrel1000_cfg2")


def test_fragment_subgraphs_leave_play_vs_oracle():
    """sub-graphs left with 1..3 nodes after an extraction become fragments (extract_track_candidates.py:463-467) and are
    dropped from the list: no later stage may touch their nodes, count their edges or raise for them.  Small sparse events
    (61 sub-graphs) leave fragments and empty sub-graphs after every extraction; states (incl. the frozen rows of the
    out-of-play nodes) and the per-iteration counters are compared with the oracle."""
    import oracle_lib as ol
    from test_gpu_parity import synth_batch
    hb = synth_batch(6, 20, 3100, target_degree=4.0)
    ob = ol.OracleBatch(hb)
    b = gtf_b200.EventBatch(hb, raise_ref_errors=False)
    ob.seed()
    b.seed()
    ob.cluster(0, 1.0, 2.0)
    b.cluster(0, 1.0, 2.0)
    W = ("alive", "active", "merged", "uts", "degree", "edge_w")
    seen_fragment = False
    for rnd in range(3):
        n_o, acc_o, _, _ = ob.extract()
        n_g, acc_g, _, _ = b.extract()
        assert n_o == n_g and np.array_equal(acc_o, acc_g), rnd
        if rnd:
            ob.remove_state_metadata()
            b.remove_state_metadata()
        seen_fragment |= bool((ob.hb["sub_state"] == 1).any())
        for it in range(2):
            s0 = (ob.stats.edges_sent, ob.stats.edges_gated, ob.stats.edges_reweight_off)
            ob.extrapolate_stage(2.0)
            ob.cluster(1, 1000.0, 100.0)
            st = b.iterate(max_iter=1, stop_when_converged=False)[0]
            assert gu.compare_states(state_of(b), ob.hb, W, rtol=1e-7) == [], (rnd, it)
            assert st["edges_sent"] == ob.stats.edges_sent - s0[0], (rnd, it)
            assert st["edges_gated"] == ob.stats.edges_gated - s0[1], (rnd, it)
            assert st["edges_reweight_off"] == ob.stats.edges_reweight_off - s0[2], (rnd, it)
            ex = gu.edge_exists(ob.hb) & gu.inplay_nodes(ob.hb)[ob.hb["slot_dst"]]
            assert st["active_edges"] == int((ob.hb["active"][ex] == 1).sum()), (rnd, it)
            assert st["ref_errors"] == ob.err, (rnd, it)
    assert seen_fragment


def test_load_events_equals_full_upload_and_batch_reuse():
    """gtf_batch_load_events (hits + the two CSR orders only; slot_dst / rev_slot / initial state on the device) gives the
    same batch as uploading every array of gtf_fields.h, bit for bit, through the whole schedule; the same batch object then
    takes a second, smaller set of events (capacity reuse) without leaking state from the first"""
    from test_gpu_parity import synth_batch
    from gtf_b200 import driver

    def run(b):
        b.seed()
        b.cluster("track_state_estimates", 1.0, 2.0)
        st = b.iterate(max_iter=4, stop_when_converged=False)
        n = b.extract()[0]
        return st, n, b.download(), b.candidates()

    def same(x, y):
        assert x[0] == y[0] and x[1] == y[1]
        for k in x[2]:
            assert np.array_equal(x[2][k], y[2][k], equal_nan=True), k
        assert np.array_equal(x[3], y[3])

    hb1 = synth_batch(3, 150, 5100, eta_max=1.0)
    hb2 = synth_batch(2, 90, 5200)
    ref1, ref2 = run(gtf_b200.EventBatch(hb1)), run(gtf_b200.EventBatch(hb2))
    b = gtf_b200.EventBatch.with_capacity(len(hb1["x"]) + 7, len(hb1["in_src"]) + 11, len(hb1["sub_event"]) + 3)
    b.load_events(hb1)
    got = b.download(["slot_dst", "rev_slot", "alive", "sub_state", "uts_rank", "label"])
    assert np.array_equal(got["slot_dst"], hb1["slot_dst"]) and np.array_equal(got["rev_slot"], hb1["rev_slot"])
    assert got["alive"].all() and not got["sub_state"].any() and (got["uts_rank"] == -1).all() and (got["label"] == -1).all()
    same(run(b), ref1)
    t = b.candidates_device()
    import torch
    assert np.array_equal(torch.as_tensor(t, device="cuda").cpu().numpy(), ref1[3])
    b.load_events(hb2)
    same(run(b), ref2)
    b.load_events(hb1)
    same(run(b), ref1)
