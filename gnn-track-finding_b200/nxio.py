"""networkx <-> flat structure-of-arrays ("host batch") conversion.

This is the data half of the drop-in boundary: the reference's
stages exchange `list[nx.DiGraph]` whose node attributes hold Python dicts of
numpy 3-vectors / 3x3 matrices (schema: helper.py:432-441, 497-508;
extrapolate_merged_states.py:375-385).  The B200 path keeps the same
information as flat arrays:

* nodes in *graph iteration order*, sub-graphs concatenated (order is data: the
  reference's results depend on it, SURVEY.md §7 "Order as data");
* in-CSR by destination: one *slot* per (neighbour -> node) entry, slots in the
  insertion order of the node's `track_state_estimates` dict (helper.py:375);
* out-CSR by source: the slot ids of a node's out-edges in `G.successors(u)` order
  (extrapolate_merged_states.py:430 iterates in this order);
* `updated_track_states` dict order is dynamic -> stored as `uts_rank` per slot.

Covariances are stored as the 4 numbers (p00, p01, p11, p22) the reference's
matrices actually carry (row/col 2 are zeroed, helper.py:423-425,
extrapolate_merged_states.py:363-365; p10 is p01 up to rounding).
"""
import numpy as np

F64_FIELDS_TSE = ("a", "b", "c", "tau", "p00", "p01", "p11", "p22", "prior", "w")
F64_FIELDS_UTS = ("a", "b", "c", "tau", "p00", "p01", "p11", "p22", "lik", "prior", "w", "lrn")
NODE_MERGED = ("m_a", "m_b", "m_c", "m_p00", "m_p01", "m_p11", "m_p22", "m_prior")


def empty_host_batch(N, E, S):
    hb = {
        "x": np.zeros(N), "y": np.zeros(N), "z": np.zeros(N), "r": np.zeros(N),
        "layer": np.zeros(N, np.int32), "volume": np.zeros(N, np.int32),
        "truth": np.zeros(N, np.int64), "orig_id": np.zeros(N, np.int64),
        "sub": np.zeros(N, np.int32), "alive": np.ones(N, np.uint8),
        "sub_off": np.zeros(S + 1, np.int32), "sub_event": np.zeros(S, np.int32),
        "in_off": np.zeros(N + 1, np.int32), "in_src": np.zeros(E, np.int32),
        "out_off": np.zeros(N + 1, np.int32), "out_slot": np.zeros(E, np.int32),
        "rev_slot": np.full(E, -1, np.int32), "slot_dst": np.zeros(E, np.int32),
        "active": np.zeros(E, np.uint8), "edge_w": np.full(E, np.nan),
        "tse_present": np.zeros(E, np.uint8),
        "uts_present": np.zeros(E, np.uint8), "uts_rank": np.full(E, -1, np.int32),
        "uts_side": np.zeros(E, np.int8),
        "has_merged": np.zeros(N, np.uint8), "degree": np.zeros(N, np.int32),
        "has_uts": np.zeros(N, np.uint8), "in_key": np.zeros(E, np.int64),
        "emp_var": np.full(N, np.nan),
    }
    for f in F64_FIELDS_TSE:
        hb["tse_" + f] = np.full(E, np.nan)
    for f in F64_FIELDS_UTS:
        hb["uts_" + f] = np.full(E, np.nan)
    for f in NODE_MERGED:
        hb[f] = np.full(N, np.nan)
    return hb


def _put_state(hb, prefix, s, ent):
    sv = ent["edge_state_vector"]
    jv = ent["joint_vector"]
    cov = ent["joint_vector_covariance"]
    hb[prefix + "a"][s] = sv[0]
    hb[prefix + "b"][s] = sv[1]
    hb[prefix + "c"][s] = sv[2]
    hb[prefix + "tau"][s] = jv[2]
    hb[prefix + "p00"][s] = cov[0, 0]
    hb[prefix + "p01"][s] = cov[0, 1]
    hb[prefix + "p11"][s] = cov[1, 1]
    hb[prefix + "p22"][s] = cov[2, 2]
    if "prior" in ent:
        hb[prefix + "prior"][s] = ent["prior"]
    if "mixture_weight" in ent:
        hb[prefix + "w"][s] = ent["mixture_weight"]


def graphs_to_host(graphs, events=None):
    """Flatten a list of reference-schema DiGraphs.

    Slot order at a node = key order of its `track_state_estimates` dict when present, else predecessor
    order.  A dict key whose node was already removed from the graph (extract_track_candidates.py:460-462
    leaves such stale entries until remove_state_metadata pops them) becomes a *ghost* row: alive = 0,
    coordinates taken from the entry's own 'xyzr', appended after the sub-graph's real nodes.
    """
    S = len(graphs)
    per_graph_nodes, per_graph_ghosts, slot_keys_all = [], [], []
    for gi, g in enumerate(graphs):
        nodes = list(g.nodes())
        per_graph_nodes.append(nodes)
        ghosts, seen = [], set()
        keys_g = []
        for n in nodes:
            attr = g.nodes[n]
            if "track_state_estimates" in attr:
                keys = list(attr["track_state_estimates"].keys())
                kset = set(keys)
                for p in g.predecessors(n):
                    if p not in kset:
                        keys.append(p)
                        kset.add(p)
                for p in attr.get("updated_track_states", {}).keys():
                    if p not in kset:
                        keys.append(p)
                        kset.add(p)
            else:
                keys = list(g.predecessors(n))
            keys_g.append(keys)
            for k in keys:
                if k not in g and k not in seen:
                    seen.add(k)
                    ent = attr.get("track_state_estimates", {}).get(k) or attr.get("updated_track_states", {}).get(k)
                    ghosts.append((k, ent["xyzr"]))
        per_graph_ghosts.append(ghosts)
        slot_keys_all.append(keys_g)
    N = sum(len(a) + len(b) for a, b in zip(per_graph_nodes, per_graph_ghosts))
    E = sum(len(k) for kg in slot_keys_all for k in kg)
    hb = empty_host_batch(N, E, S)
    index = {}
    pos = 0
    for gi in range(S):
        for n in per_graph_nodes[gi]:
            index[(gi, n)] = pos
            pos += 1
        for k, _ in per_graph_ghosts[gi]:
            index[(gi, k)] = pos
            pos += 1
    slot_of = {}
    s = 0
    i = 0
    for gi, g in enumerate(graphs):
        hb["sub_off"][gi] = i
        hb["sub_event"][gi] = 0 if events is None else events[gi]
        for ni, n in enumerate(per_graph_nodes[gi]):
            attr = g.nodes[n]
            if "GNN_Measurement" in attr:
                gm = attr["GNN_Measurement"]
                hb["x"][i], hb["y"][i], hb["z"][i], hb["r"][i] = gm.x, gm.y, gm.z, gm.r
            else:
                hb["x"][i], hb["y"][i], hb["z"][i], hb["r"][i] = attr["xyzr"]
            hb["layer"][i] = attr["in_volume_layer_id"]
            hb["volume"][i] = attr["volume_id"]
            hb["truth"][i] = attr["truth_particle"]
            hb["orig_id"][i] = n
            hb["sub"][i] = gi
            hb["degree"][i] = attr.get("degree", 0)
            if "xy_edge_gradient_mean_var" in attr:          # helper.py:446 (np.mean, np.var) of the xy edge gradients
                hb["emp_var"][i] = attr["xy_edge_gradient_mean_var"][1]
            hb["in_off"][i] = s
            tse = attr.get("track_state_estimates", {})
            uts = attr.get("updated_track_states", None)
            hb["has_uts"][i] = uts is not None
            uts_pos = {} if uts is None else {k: r for r, k in enumerate(uts.keys())}
            for k in slot_keys_all[gi][ni]:
                slot_of[(gi, k, n)] = s
                hb["in_src"][s] = index[(gi, k)]
                hb["slot_dst"][s] = i
                hb["in_key"][s] = k
                if g.has_edge(k, n):
                    ed = g[k][n]
                    hb["active"][s] = ed.get("activated", 0)
                    if "mixture_weight" in ed:
                        hb["edge_w"][s] = ed["mixture_weight"]
                if k in tse:
                    hb["tse_present"][s] = 1
                    _put_state(hb, "tse_", s, tse[k])
                if uts is not None and k in uts:
                    ent = uts[k]
                    hb["uts_present"][s] = 1
                    hb["uts_rank"][s] = uts_pos[k]
                    _put_state(hb, "uts_", s, ent)
                    hb["uts_lik"][s] = ent["likelihood"]
                    if "lr_layer_norm" in ent:
                        hb["uts_lrn"][s] = ent["lr_layer_norm"]
                    if "side" in ent:
                        hb["uts_side"][s] = 1 if ent["side"] == "left" else 2
                s += 1
            if "merged_state" in attr:
                hb["has_merged"][i] = 1
                ms, mc = attr["merged_state"], attr["merged_cov"]
                hb["m_a"][i], hb["m_b"][i], hb["m_c"][i] = ms[0], ms[1], ms[2]
                hb["m_p00"][i], hb["m_p01"][i] = mc[0, 0], mc[0, 1]
                hb["m_p11"][i], hb["m_p22"][i] = mc[1, 1], mc[2, 2]
                hb["m_prior"][i] = attr["merged_prior"]
            i += 1
        for k, xyzr in per_graph_ghosts[gi]:
            hb["x"][i], hb["y"][i], hb["z"][i], hb["r"][i] = xyzr
            hb["alive"][i] = 0
            hb["orig_id"][i] = k
            hb["sub"][i] = gi
            hb["truth"][i] = -1
            hb["in_off"][i] = s
            i += 1
    hb["sub_off"][S] = N
    hb["in_off"][N] = E
    # out-CSR: successor order for real nodes (extrapolate_merged_states.py:430), then the slots keyed by ghosts
    o = 0
    i = 0
    for gi, g in enumerate(graphs):
        ghost_out = {k: [] for k, _ in per_graph_ghosts[gi]}
        for ni, n in enumerate(per_graph_nodes[gi]):
            for k in slot_keys_all[gi][ni]:
                if k in ghost_out:
                    ghost_out[k].append(slot_of[(gi, k, n)])
        for n in per_graph_nodes[gi]:
            hb["out_off"][i] = o
            for v in g.successors(n):
                hb["out_slot"][o] = slot_of[(gi, n, v)]
                o += 1
            i += 1
        for k, _ in per_graph_ghosts[gi]:
            hb["out_off"][i] = o
            for sl in ghost_out[k]:
                hb["out_slot"][o] = sl
                o += 1
            i += 1
    hb["out_off"][N] = o
    assert o == E, (o, E)
    for (gi, k, n), sl in slot_of.items():
        hb["rev_slot"][sl] = slot_of.get((gi, n, k), -1)
    return hb


def events_to_graphs(ev, event_id=0):
    """Build the DiGraph the reference's `construct_graph` (helper.py:465-520) would build for a
    synthetic event dict (synth.py), then split it into weakly-connected sub-graphs exactly like
    event_conversion.py:76-84.  Needs the reference's GNN_Measurement class on sys.path."""
    import networkx as nx
    from GNN_Measurement import GNN_Measurement as gnn
    G = nx.DiGraph()
    for i in range(len(ev["x"])):
        x, y, z, r = float(ev["x"][i]), float(ev["y"][i]), float(ev["z"][i]), float(ev["r"][i])
        t = int(ev["truth"][i])
        gm = gnn.GNN_Measurement(x, y, z, r, truth_particle=t, n=i)
        G.add_node(i, GNN_Measurement=gm, xy=(x, y), zr=(z, r), xyzr=(x, y, z, r),
                   volume_id=int(ev["volume"][i]), in_volume_layer_id=int(ev["layer"][i]),
                   vivl_id=(int(ev["volume"][i]), int(ev["layer"][i])),
                   module_id=np.array([i]), truth_particle=t,
                   hit_dissociation={"hit_id": np.array([i]), "particle_id": [t]},
                   tags=[i])
    for a, b in zip(ev["edge_a"].tolist(), ev["edge_b"].tolist()):
        G.add_edge(a, b)
        G.add_edge(b, a)
    G = nx.DiGraph(G)
    return [G.subgraph(c).copy() for c in nx.weakly_connected_components(G)]


class Measurement(object):
    """Stand-in for the reference's GNN_Measurement (GNN_Measurement.py:1-9) when graphs are rebuilt from
    flat arrays without the reference package on the path."""

    def __init__(self, x, y, z, r, truth_particle=-1, n=None):
        self.x, self.y, self.z, self.r = x, y, z, r
        self.truth_particle = truth_particle
        self.node = n


def _cov3(p00, p01, p11, p22):
    return np.array([[p00, p01, 0.0], [p01, p11, 0.0], [0.0, 0.0, p22]])


def _entry(hb, prefix, s, src_xyzr, uts):
    cov = _cov3(hb[prefix + "p00"][s], hb[prefix + "p01"][s], hb[prefix + "p11"][s], hb[prefix + "p22"][s])
    ent = {"xyzr": src_xyzr,
           "edge_state_vector": np.array([hb[prefix + "a"][s], hb[prefix + "b"][s], hb[prefix + "c"][s]]),
           "edge_covariance": cov,
           "joint_vector": [hb[prefix + "a"][s], hb[prefix + "b"][s], hb[prefix + "tau"][s]],
           "joint_vector_covariance": cov}       # same object, as in the reference (quirk 4)
    if uts:
        ent["xy"] = (src_xyzr[0], src_xyzr[1])
        ent["zr"] = (src_xyzr[2], src_xyzr[3])
        ent["likelihood"] = hb["uts_lik"][s]
        if not np.isnan(hb["uts_lrn"][s]):
            ent["lr_layer_norm"] = hb["uts_lrn"][s]
        if hb["uts_side"][s]:
            ent["side"] = "left" if hb["uts_side"][s] == 1 else "right"
    if not np.isnan(hb[prefix + "prior"][s]):
        ent["prior"] = hb[prefix + "prior"][s]
    if not np.isnan(hb[prefix + "w"][s]):
        ent["mixture_weight"] = hb[prefix + "w"][s]
    return ent


def host_to_graphs(hb, orig_id=None, truth=None):
    """Inverse of graphs_to_host: rebuild reference-schema DiGraphs (one per in-play sub-graph) from a
    complete host batch.  Removed nodes are left out; successor order follows the out-CSR."""
    import networkx as nx
    N, S = len(hb["x"]), len(hb["sub_off"]) - 1
    oid = np.arange(N) if orig_id is None else np.asarray(orig_id)
    graphs = []
    for g in range(S):
        if hb["sub_state"][g] != 0:
            continue
        G = nx.DiGraph()
        lo, hi = int(hb["sub_off"][g]), int(hb["sub_off"][g + 1])
        for i in range(lo, hi):
            if not hb["alive"][i]:
                continue
            n = int(oid[i])
            t = -1 if truth is None else int(truth[i])
            x, y, z, r = (float(hb[k][i]) for k in "xyzr")
            G.add_node(n, GNN_Measurement=Measurement(x, y, z, r, t, n), xy=(x, y), zr=(z, r), xyzr=(x, y, z, r),
                       volume_id=int(hb["volume"][i]), in_volume_layer_id=int(hb["layer"][i]),
                       vivl_id=(int(hb["volume"][i]), int(hb["layer"][i])), module_id=np.array([n]), truth_particle=t,
                       hit_dissociation={"hit_id": np.array([n]), "particle_id": [t]}, tags=[n],
                       degree=int(hb["degree"][i]))
        for i in range(lo, hi):
            if not hb["alive"][i]:
                continue
            for o in range(int(hb["out_off"][i]), int(hb["out_off"][i + 1])):
                s = int(hb["out_slot"][o])
                v = int(hb["slot_dst"][s])
                if not hb["alive"][v]:
                    continue
                attrs = {"activated": int(hb["active"][s])}
                if not np.isnan(hb["edge_w"][s]):
                    attrs["mixture_weight"] = hb["edge_w"][s]
                G.add_edge(int(oid[i]), int(oid[v]), **attrs)
        apply_host_to_graphs(hb, [G], oid, subs=[g])
        graphs.append(G)
    return graphs


def apply_host_to_graphs(hb, graphs, orig_id=None, subs=None):
    """Write the mutable state of a host batch back into existing graphs (the drop-in 'export' step):
    edge activation / weights, both state dicts in dict order, merged state, degree."""
    N = len(hb["x"])
    oid = np.arange(N) if orig_id is None else np.asarray(orig_id)
    inplay = [g for g in range(len(hb["sub_off"]) - 1) if hb["sub_state"][g] == 0] if subs is None else subs
    for G, g in zip(graphs, inplay):
        lo, hi = int(hb["sub_off"][g]), int(hb["sub_off"][g + 1])
        dead = [int(oid[i]) for i in range(lo, hi) if not hb["alive"][i] and int(oid[i]) in G]
        G.remove_nodes_from(dead)
        for i in range(lo, hi):
            if not hb["alive"][i]:
                continue
            n = int(oid[i])
            attr = G.nodes[n]
            s0, s1 = int(hb["in_off"][i]), int(hb["in_off"][i + 1])
            tse, uts = {}, []
            for s in range(s0, s1):
                src = int(hb["in_src"][s])
                key = int(oid[src])
                xyzr = (float(hb["x"][src]), float(hb["y"][src]), float(hb["z"][src]), float(hb["r"][src]))
                if hb["tse_present"][s]:
                    tse[key] = _entry(hb, "tse_", s, xyzr, False)
                if hb["uts_present"][s]:
                    uts.append((int(hb["uts_rank"][s]), key, _entry(hb, "uts_", s, xyzr, True)))
                if G.has_edge(key, n):
                    G[key][n]["activated"] = int(hb["active"][s])
                    if not np.isnan(hb["edge_w"][s]):
                        G[key][n]["mixture_weight"] = hb["edge_w"][s]
            attr["track_state_estimates"] = tse
            if hb["has_uts"][i]:
                attr["updated_track_states"] = {k: e for _, k, e in sorted(uts, key=lambda t: t[0])}
            if hb["has_merged"][i]:
                attr["merged_state"] = np.array([hb["m_a"][i], hb["m_b"][i], hb["m_c"][i]])
                attr["merged_cov"] = _cov3(hb["m_p00"][i], hb["m_p01"][i], hb["m_p11"][i], hb["m_p22"][i])
                attr["merged_prior"] = hb["m_prior"][i]
            attr["degree"] = int(hb["degree"][i])
            attr.setdefault("angle_of_rotation", float(np.arctan2(hb["y"][i], hb["x"][i])))
            attr.setdefault("translation", (float(hb["x"][i]), float(hb["y"][i])))
    return graphs
