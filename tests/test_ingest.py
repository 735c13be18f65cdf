"""CPU: native CSV ingest (helper.py:524-545 / :465-520 semantics)."""
import os

import numpy as np
import pytest

import gtf_b200
from gtf_b200 import ingest, synth

REF_EVENT = "/root/reference/src/trackml_mod/event_network/minCurv_0.3_134/event_1_filtered_graph_"


def test_csv_roundtrip(tmp_path):
    ev = synth.barrel_event(30, seed=5)
    pre = str(tmp_path) + "/ev_"
    ingest.write_event_csv(ev, pre)
    back = ingest.load_event_csv(pre, 8, 8)
    for k in ("x", "y", "z", "r", "layer", "volume", "edge_a", "edge_b"):
        assert np.array_equal(back[k], ev[k]), k
    hb0, hb1 = synth.event_to_host(ev), synth.event_to_host(back)
    for k in hb0:
        if k != "truth":
            assert np.array_equal(hb0[k], hb1[k]), k
    # volume filter: nothing of volume 8 survives a [7, 7] selection except layer_id == 8000 (pandas `between`)
    none = ingest.load_event_csv(pre, 9, 9)
    assert len(none["x"]) == 0


@pytest.mark.skipif(not os.path.exists(REF_EVENT + "nodes.csv"), reason="reference data only in the build container")
def test_shipped_event_statistics():
    """the shipped TrackML-derived event: 55,701 nodes / 165,472 undirected edges in the file; volumes 7-9 give the
    30,387 nodes / 73,230 directed edges quoted in SURVEY.md 8d"""
    ev = ingest.load_event_csv(REF_EVENT, 7, 9)
    assert len(ev["x"]) == 30387
    hb = synth.event_to_host(ev)
    assert len(hb["in_src"]) == 73230
    deg = np.diff(hb["in_off"])
    assert abs(deg.mean() - 2.4) < 0.05
    full = ingest.load_event_csv(REF_EVENT, 0, 99)
    assert len(full["x"]) == 55701


def shipped_event_from_fixture():
    """the csv_* arrays of tests/golden/shipped_vol79.npz = what ingest.load_event_csv reads from the shipped files"""
    import golden_util as gu
    fx = gu.load("shipped_vol79")
    ev = {k: fx["csv_" + k] for k in ("x", "y", "z", "layer", "volume", "edge_a", "edge_b", "node_idx")}
    ev["r"] = ingest.radius_like_reference(ev["x"], ev["y"])
    ev["truth"] = np.full(len(ev["x"]), -1, np.int64)
    return fx, ev


def test_native_ingest_reproduces_the_reference_graph_of_the_shipped_event():
    """nodes.csv / edges.csv -> flat layout without pandas / networkx.  Against the graph the
    UNMODIFIED reference builds from the same files (helper.py:524-545 load_nodes_edges, :465-520 construct_graph,
    event_conversion.py:76-96; tests/golden/make_shipped_golden.py): node order, sub-graph split and order, successor order
    and the state-dict order (CPython set iteration of the neighbour ids, helper.py:280) are all identical."""
    fx, ev = shipped_event_from_fixture()
    hb = synth.event_to_host(ev, 0, dict_order="pyset")
    assert len(hb["x"]) == 30387 and len(hb["in_src"]) == 73230
    assert np.array_equal(ev["node_idx"][hb["orig_id"]], fx["topo_orig_id"])          # node order, sub-graphs concatenated
    for k in ("x", "y", "z", "r", "layer", "volume", "sub", "sub_off", "in_off", "in_src", "slot_dst", "out_off", "out_slot", "rev_slot"):
        assert np.array_equal(hb[k], fx["topo_" + k]), k
    # the default (insertion-order) layout is a different, equally valid labelling of the same graph: set iteration
    # re-orders nodes inside small sub-graphs and the neighbours inside the state dicts
    hb_ins = synth.event_to_host(ev, 0)
    assert not np.array_equal(hb_ins["orig_id"], hb["orig_id"]) and np.array_equal(np.sort(hb_ins["orig_id"]), np.sort(hb["orig_id"]))


@pytest.mark.skipif(not os.path.exists(REF_EVENT + "nodes.csv"), reason="reference data only in the build container")
def test_fixture_holds_the_shipped_files_content():
    fx, _ = shipped_event_from_fixture()
    ev = ingest.load_event_csv(REF_EVENT, 7, 9)
    for k in ("x", "y", "z", "layer", "volume", "edge_a", "edge_b", "node_idx"):
        assert np.array_equal(ev[k], fx["csv_" + k]), k


def test_cfg1_toy_event_shape_and_host_layout():
    """BASELINE configs[0]: the 2-D toy detector of toyMC_model/track_simulation_xy.py:36-160 (88 straight tracks x 10 hits, the
    collision point and every node on a long-dx edge removed), seeded: the event is deterministic, its edges join layers one or
    two apart, and the host layout built from it is a consistent pair of CSR orders."""
    from gtf_b200 import synth
    ev, ev2 = synth.toy_event(seed=3), synth.toy_event(seed=3)
    for k in ev:
        assert np.array_equal(ev[k], ev2[k])
    assert not np.array_equal(ev["y"], synth.toy_event(seed=4)["y"][:len(ev["y"])]) or len(ev["y"]) != len(synth.toy_event(seed=4)["y"])
    n, e, mean_deg, frac = synth.degree_stats(ev)
    assert 50 <= n <= 880 and e == 2 * len(ev["edge_a"]) and 1.0 < mean_deg < 10.0
    dl = np.abs(ev["layer"][ev["edge_a"]] - ev["layer"][ev["edge_b"]])
    assert dl.min() >= 1 and dl.max() <= 2 and (ev["edge_a"] != ev["edge_b"]).all()
    assert (ev["z"] == 0).all() and (ev["r"] == 0).all() and (ev["layer"] >= 1).all()        # the collision point is gone
    hb = synth.event_to_host(ev)
    N, E = len(hb["x"]), len(hb["in_src"])
    assert N == n and E == e
    assert hb["in_off"][0] == 0 and hb["in_off"][-1] == E and hb["out_off"][0] == 0 and hb["out_off"][-1] == E
    assert sorted(hb["out_slot"].tolist()) == list(range(E))                                   # every slot is somebody's out-edge
    rs = hb["rev_slot"]
    assert (rs >= 0).all() and np.array_equal(rs[rs], np.arange(E))                            # u->v and v->u cross-linked
    assert np.array_equal(hb["slot_dst"], np.repeat(np.arange(N), np.diff(hb["in_off"])))
    src_of_out = np.repeat(np.arange(N), np.diff(hb["out_off"]))
    assert np.array_equal(hb["in_src"][hb["out_slot"]], src_of_out)
