"""GPU parity against the round-2 reference pins (tests/golden/make_lut_tag_golden.py, make_golden.py barrel1000_cfg2):
LUT-threshold mode vs the reference's cluster() with a per-node KL_threshold, emp_var vs helper.py:446, tag propagation
vs the unmodified tag_propagation.py, a cfg2-size event through the reference's own schedule.  All through the C-ABI."""
import os

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
    # the same SHAPE with other hits: the captured iteration graphs are replayed as they are (sizes and pointers are equal,
    # the tile tables and everything else they read live in device memory) -- no state of the previous events may show
    hb3 = {k: np.array(v, copy=True) for k, v in hb1.items()}
    rng = np.random.default_rng(5)
    for f in ("x", "y", "z"):
        hb3[f] = hb3[f] + rng.normal(size=len(hb3[f])) * 0.05
    hb3["r"] = np.sqrt(hb3["x"] ** 2 + hb3["y"] ** 2)
    ref3 = run(gtf_b200.EventBatch(hb3))
    assert not np.array_equal(ref3[2]["m_a"], ref1[2]["m_a"], equal_nan=True)
    b.load_events(hb3)
    same(run(b), ref3)
    b.load_events(hb1)
    same(run(b), ref1)


def test_boundary_flip_candidates_are_enumerated():
    """north_star: "any threshold-boundary flips enumerated" -- decisions taken within 1e-9 relative of their threshold are
    counted in gtf_stats.near_threshold and listed by gtf_batch_near_threshold.  None occurs on the reference fixtures with
    the schedule's thresholds; a gate cut set EXACTLY on one message's chi2 is reported by both the per-stage kernel and
    the packed iteration (and that message still passes: chi2 <= cut)."""
    from test_gpu_parity import STEPS
    for name in ("barrel25_deg6", "barrel40_eta1"):
        fx = gu.load(name)
        for prev, stage, fn, what in STEPS:
            b = gtf_b200.EventBatch(gu.stage_batch(fx, prev))
            fn(b)
            assert b.last_stats["near_threshold"] == 0 and b.near_threshold() == (0, []), (name, stage)
    fx = gu.load("barrel40_eta1")
    hb = gu.stage_batch(fx, "x1")
    b = gtf_b200.EventBatch(hb)
    b.message_passing(2.0)
    out = b.download(["uts_chi2", "uts_present", "active"])
    passing = np.sort(out["uts_chi2"][(out["uts_present"] > 0) & (out["uts_chi2"] <= 2.0)])
    cut = float(passing[len(passing) // 2])
    n_at_cut = int((out["uts_chi2"] == cut).sum())
    b2 = gtf_b200.EventBatch(hb)
    st = b2.message_passing(cut)
    n, recs = b2.near_threshold()
    assert st["near_threshold"] == n == n_at_cut >= 1
    assert all(k == 0 and v == cut and t == cut for k, _, v, t in recs)
    slots = sorted(i for _, i, _, _ in recs)
    assert slots == sorted(np.nonzero(out["uts_chi2"] == cut)[0].tolist())
    assert b2.download(["uts_present"])["uts_present"][slots].all()          # chi2 <= cut: the message passes
    b3 = gtf_b200.EventBatch(hb)
    st3 = b3.iterate(max_iter=1, stop_when_converged=False, chi2_cut=cut)[0]
    assert st3["near_threshold"] >= n_at_cut
    assert sorted(i for k, i, _, _ in b3.near_threshold()[1] if k == 0) == slots


def test_reference_helper_names_vs_reference():
    """the stand-alone helpers of clustering.py:11-124 and extrapolate_validate (extrapolate_merged_states.py:26) under the
    reference's names (gtf_b200.stages), on the GPU through the C-ABI, against values produced by the unmodified reference
    functions (tests/golden/make_helpers_golden.py)"""
    import networkx as nx
    from gtf_b200 import stages, nxio
    fx = gu.load("helpers")
    svs, covs, node, nbrs = fx["svs"], fx["covs"], fx["node"], fx["nbrs"]
    got = stages.calc_pairwise_distances_chi2(len(svs), svs, covs, node, nbrs, 0.4, 0.6, 550.0)
    assert gu.rel_err(got, fx["chi2_matrix"]) <= 1e-9 and np.array_equal(got == 0, fx["chi2_matrix"] == 0)
    assert abs(stages.mahalanobis_distance(svs[2], covs[2], svs[5], covs[5], node, nbrs[2], nbrs[5], 0.4, 0.6, 550.0)
               - float(fx["chi2_pair"])) <= 1e-9 * abs(float(fx["chi2_pair"]))
    sm, idx = stages.get_smallest_dist_idx(got)
    assert sm == got[np.nonzero(got)].min() and got[idx[0], idx[len(idx) // 2]] == sm
    gm, gc = fx["gm"], fx["gc"]
    mm, mc = stages.merge_states(gm[0], gc[0], gm[1], gc[1])
    assert gu.rel_err(mm, fx["merged_mean"]) <= 1e-9 and gu.rel_err(mc, fx["merged_cov"]) <= 1e-9
    assert abs(stages.KLDistance(gm[0], gc[0], gm[1], gc[1]) - float(fx["kl"])) <= 1e-9 * abs(float(fx["kl"]))
    d = stages.calc_dist_to_merged_state(len(gm) - 2, gm[2:], gc[2:], mm, mc)
    assert gu.rel_err(np.array(d), fx["kl_to_merged"]) <= 1e-9
    sm, k = stages.get_smallest_dist_idx(d)
    assert d[k] == sm == min(d)
    # extrapolate_validate, edge by edge
    n_pass = n_fail = 0
    for row in fx["edges"]:
        G = nx.DiGraph()
        G.add_node(0, GNN_Measurement=nxio.Measurement(*row[0:4]), truth_particle=1,
                   track_state_estimates={1: {"mixture_weight": 0.25}})
        G.add_node(1, GNN_Measurement=nxio.Measurement(*row[4:8]), truth_particle=1)
        G.add_edge(0, 1, activated=1)
        cov = row[11:20].reshape(3, 3).copy()
        out = stages.extrapolate_validate(G, 0, G.nodes[0], 1, G.nodes[1], row[33], row[8:11].copy(), cov, 0.3, 0.4, 0.6, 550.0)
        assert abs(cov[1, 1] - row[20]) <= 1e-9 * abs(row[20])               # merged_cov[1, 1] += var_ms, in place
        assert abs(out[1] - row[21]) <= 1e-9 * abs(row[21])
        if row[22]:
            n_pass += 1
            dct = out[0]
            got = list(dct["edge_state_vector"]) + [dct["joint_vector"][2]] + \
                [dct["joint_vector_covariance"][i, j] for i, j in ((0, 0), (0, 1), (1, 1), (2, 2))] + [dct["likelihood"]]
            assert gu.rel_err(np.array(got), row[23:32]) <= 1e-9
            assert dct["edge_covariance"] is dct["joint_vector_covariance"] and dct["mixture_weight"] == 0.25
            assert out[2:] == (0, 0, 1, 1) and G[0][1]["activated"] == 1
        else:
            n_fail += 1
            assert out[0] is None and out[2:] == (0, 1, 0, 0) and G[0][1]["activated"] == 0
    assert n_pass >= 10 and n_fail >= 10


def test_shipped_event_native_ingest_to_candidates_vs_reference():
    """SURVEY.md 8f row 3 + VERDICT r1 item 9: the reference's shipped TrackML-derived event (volumes 7-9: 30,387 hits,
    73,230 directed edges, mean in-degree 2.4, 27 % of the nodes in the 3..15 clustering window, thousands of 1-3 node
    sub-graphs) from the CSV content through the native ingest (reference graph / dict orders), the device-side batch
    load and the reference's own schedule on the GPU: every decision and every candidate set equals the unmodified
    reference's (tests/golden/make_shipped_golden.py)"""
    from test_ingest import shipped_event_from_fixture
    from test_gpu_parity import test_full_schedule_vs_reference
    from gtf_b200 import synth
    fx, ev = shipped_event_from_fixture()
    hb = synth.event_to_host(ev, 0, dict_order="pyset")
    b = gtf_b200.EventBatch.with_capacity(len(hb["x"]), len(hb["in_src"]), len(hb["sub_event"]))
    b.load_events(hb)
    W = ("alive", "active", "merged", "degree")
    b.seed()

    def chk(stage):   # rtol: merged_cov[1,1] of 16 nodes, see tests/test_oracle_golden.py (reference quirk 13)
        assert gu.compare_states(state_of(b), gu.stage_batch(fx, stage), W, rtol=1e-6) == [], stage

    chk("seed")
    b.cluster("track_state_estimates", 1.0, 2.0)
    chk("c1")
    assert np.array_equal(b.extract()[1], fx["x1/accepted"])
    chk("x1")
    b.extrapolate_stage(2.0)
    chk("e2")
    assert np.array_equal(b.extract()[1], fx["x2/accepted"])
    chk("x2")
    b.remove_state_metadata()
    chk("m2")
    b.cluster("updated_track_states", 1000.0, 100.0)
    chk("c3")
    assert np.array_equal(b.extract()[1], fx["x3/accepted"])
    chk("x3")
    rows = b.candidates()
    assert len(rows) == int((fx["x1/accepted"] | fx["x2/accepted"] | fx["x3/accepted"]).sum()) > 1000
    # and the packed iteration on the same event (k_send / k_exec / k_node2 / k_hv on a low-degree, fragmented graph)
    import oracle_lib as ol
    hb.pop("truth"), hb.pop("orig_id")
    ob = ol.OracleBatch(hb)
    ob.seed()
    ob.cluster(0, 1.0, 2.0)
    b.load_events(hb)
    b.seed()
    b.cluster("track_state_estimates", 1.0, 2.0)
    b.raise_ref_errors = False
    for _ in range(3):
        ob.extrapolate_stage(2.0)
        ob.cluster(1, 1000.0, 100.0)
    b.iterate(max_iter=3, stop_when_converged=False)
    assert gu.compare_states(state_of(b), ob.hb, ("active", "merged", "uts", "degree", "edge_w"), rtol=1e-7) == []


def test_seed_cluster_equals_seed_then_cluster():
    """gtf_seed_cluster (one pass of the packed node kernels over the freshly seeded dicts) = gtf_seed_all followed by
    gtf_cluster on the seeds, against the reference fixtures and bit for bit against the two-call form (scalar and LUT mode)"""
    for name in ("barrel25_deg6", "barrel40_eta1", "barrel1000_cfg2"):
        fx = gu.load(name)
        hb = blank_seed(gu.stage_batch(fx, "seed"))
        a = gtf_b200.EventBatch(hb)
        sa = a.seed_cluster(1.0, 2.0)
        what = ALL if name != "barrel1000_cfg2" else ("alive", "active", "merged", "degree")
        assert gu.compare_states(state_of(a), gu.stage_batch(fx, "c1"), what) == [], name
        c = gtf_b200.EventBatch(hb)
        c.seed()
        sc = c.cluster("track_state_estimates", 1.0, 2.0)
        ga, gc = state_of(a), state_of(c)
        for k in ga:
            assert np.array_equal(ga[k], gc[k], equal_nan=True), (name, k)
        assert sa == sc
        # and the packed iteration picks both up identically (after gtf_seed_cluster the dict-entry records still hold the
        # seed dict and only the presence bitmap is cleared; after the two-call form everything is re-packed)
        a2 = gtf_b200.EventBatch(hb)
        a2.seed_cluster(1.0, 2.0)
        ia = a2.iterate(max_iter=2, stop_when_converged=False)
        ic = c.iterate(max_iter=2, stop_when_converged=False)
        assert ia == ic
        ga, gc = state_of(a2), state_of(c)
        for k in ga:
            assert np.array_equal(ga[k], gc[k], equal_nan=True), (name, k, "after two iterations")
    fx = gu.load("lut_barrel40")
    a = gtf_b200.EventBatch(blank_seed(gu.stage_batch(fx, "seed")))
    a.seed_cluster(1.0, 123.0, KL_lut=fx["lut_stress"])
    assert gu.compare_states(state_of(a), gu.stage_batch(fx, "c1str"), ALL, rtol=gu.RTOL, chained=True) == []


def test_parabolic_seeding_of_the_lut_training_pipeline_vs_reference():
    """learn_KL_parabolic_model/src/generate_training_data/utils.py:221-299 (compute_track_state_estimates with rotate_track):
    the kernel behind gtf_seed_parabolic_pairs against the outputs of the UNMODIFIED reference function on a toy graph
    (tests/golden/make_parabolic_golden.py): state vectors and full 3x3 covariances to 1e-9 (measured 1e-14), and the
    reference-named graph wrapper: same dict keys in the same order, same xy_edge_gradient_mean_var."""
    import networkx as nx
    from gtf_b200 import stages
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "parabolic_seed.npz"))
    sv, cov = stages.seed_parabolic_pairs(fx["node_xy"], fx["nbr_xy"])
    assert np.abs(sv - fx["state"]).max() <= 1e-9 * np.abs(fx["state"]).max()
    assert (np.abs(cov - fx["cov"]) <= 1e-9 * np.abs(fx["cov"])).all()
    assert (sv[:, 2] == 0).all()

    class M:
        def __init__(self, x, y):
            self.x, self.y = x, y
    G = nx.DiGraph()
    xy = {}
    for n, p in zip(fx["node"], fx["node_xy"]):
        xy[int(n)] = p
    for n in range(int(fx["n_nodes"])):
        G.add_node(n, GNN_Measurement=M(*xy[n]))
    G.add_edges_from((int(u), int(v)) for u, v in fx["edges"])
    stages.compute_track_state_estimates_parabolic([G])
    i = 0
    for n, mv in zip(fx["grad_node"], fx["grad_mean_var"]):
        a = G.nodes[int(n)]
        assert np.allclose(a["xy_edge_gradient_mean_var"], mv, rtol=1e-12, atol=0)
        for k, e in a["track_state_estimates"].items():
            assert (int(fx["node"][i]), int(fx["nbr"][i])) == (int(n), int(k))
            assert np.allclose(e["edge_state_vector"], fx["state"][i], rtol=1e-9, atol=1e-300)
            assert np.allclose(e["edge_covariance"], fx["cov"][i], rtol=1e-9, atol=0)
            i += 1
    assert i == len(fx["node"])
