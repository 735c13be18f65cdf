"""CPU: networkx <-> flat layout round trip (the data half of the drop-in boundary)."""
import numpy as np

import golden_util as gu
import gtf_b200
from gtf_b200 import nxio, fields as F


def roundtrip(stage):
    fx = gu.load("barrel25_deg6")
    hb = gu.stage_batch(fx, stage)
    graphs = nxio.host_to_graphs(hb, orig_id=fx["topo_orig_id"], truth=fx["topo_truth"])
    back = nxio.graphs_to_host(graphs)
    return fx, hb, graphs, back


def test_roundtrip_preserves_state_and_orders():
    for stage in ("seed", "c1", "e2", "c3"):
        fx, hb, graphs, back = roundtrip(stage)
        alive = hb["alive"] > 0
        real = back["alive"] > 0            # removed neighbours still referenced by stale dict keys come back as ghost rows
        assert int(real.sum()) == int(alive.sum())
        # node order and identity
        assert np.array_equal(back["orig_id"][real], fx["topo_orig_id"][alive])
        # every live slot comes back with the same key, state and flags, in the same dict order
        node_of = {int(o): i for i, o in enumerate(fx["topo_orig_id"])}
        for s2 in range(len(back["in_src"])):
            dst = node_of[int(back["orig_id"][back["slot_dst"][s2]])]
            key = node_of[int(back["in_key"][s2])]
            cand = [s for s in range(hb["in_off"][dst], hb["in_off"][dst + 1]) if hb["in_src"][s] == key]
            assert len(cand) == 1
            s = cand[0]
            for f in ("tse_present", "uts_present"):
                assert back[f][s2] == hb[f][s]
            if hb["tse_present"][s]:
                for f in ("a", "b", "c", "tau", "p00", "p01", "p11", "p22", "prior", "w"):
                    assert np.array_equal(back["tse_" + f][s2], hb["tse_" + f][s], equal_nan=True)
            if hb["uts_present"][s]:
                for f in ("a", "b", "c", "tau", "p00", "p01", "p11", "p22", "prior", "w", "lik", "lrn"):
                    assert np.array_equal(back["uts_" + f][s2], hb["uts_" + f][s], equal_nan=True), f
            if hb["alive"][key]:
                assert back["active"][s2] == hb["active"][s]
        # merged state
        m = alive & (hb["has_merged"] > 0)
        assert np.array_equal(back["has_merged"][real], hb["has_merged"][alive])
        for f in ("m_a", "m_b", "m_c", "m_p00", "m_p01", "m_p11", "m_p22", "m_prior"):
            assert np.array_equal(back[f][back["has_merged"] > 0], hb[f][m])


def test_uts_dict_order_survives():
    fx, hb, graphs, back = roundtrip("e2")
    full = F.complete_host_batch({k: v for k, v in back.items() if k not in ("truth", "orig_id", "in_key")})
    # ranks come back as dict positions: order within each node must match the original ranks
    node_of = {int(o): i for i, o in enumerate(fx["topo_orig_id"])}
    for i2 in range(len(back["x"])):
        if not back["alive"][i2]:
            continue
        i = node_of[int(back["orig_id"][i2])]
        a = [(hb["uts_rank"][s], int(fx["topo_orig_id"][hb["in_src"][s]])) for s in range(hb["in_off"][i], hb["in_off"][i + 1])
             if hb["uts_present"][s]]
        b = [(full["uts_rank"][s], int(back["in_key"][s])) for s in range(full["in_off"][i2], full["in_off"][i2 + 1])
             if full["uts_present"][s]]
        assert [k for _, k in sorted(a)] == [k for _, k in sorted(b)]
