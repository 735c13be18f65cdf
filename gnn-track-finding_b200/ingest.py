"""Native event ingest: the reference's `nodes.csv / edges.csv` (TrackML-derived) -> event dict -> flat layout,
without pandas / networkx.

Follows utilities/helper.py:524-545 (load_nodes_edges) and :465-520 (construct_graph):
  * nodes.csv columns node_idx,layer_id,x,y,z ; volume filter layer_id in [1000*min_volume, 1000*(max_volume+1)]
    (pandas `between`, inclusive both ends) ; r = sqrt(x^2+y^2) ; volume_id = int(layer_id/1000) ;
    in_volume_layer_id = layer_id % 100
  * edges.csv: first line "<n_nodes> <n_edges>", second line the header "node2,node1,weight", then rows; an
    edge is kept when both ends survive the filter and is added in both directions, node1->node2 first
"""
import numpy as np


def radius_like_reference(x, y):
    """r = np.sqrt(row.x**2 + row.y**2) evaluated row by row (helper.py:12-13, :532): a scalar `**2` goes through libm's
    pow(), which is not always the correctly rounded x*x -- 16 of the 30,387 radii of the shipped event differ in the last
    bit from sqrt(x*x + y*y).  The radius is an INPUT of the arithmetic, so ingest reproduces the reference's bits."""
    import math
    return np.array([math.sqrt(math.pow(a, 2) + math.pow(b, 2)) for a, b in zip(x.tolist(), y.tolist())], np.float64)


def load_event_csv(event_prefix, min_volume, max_volume, truth=None):
    """event_prefix: path prefix such that prefix+'nodes.csv' / prefix+'edges.csv' exist (the reference passes
    '<dir>/event_1_filtered_graph_').  Returns a synth-style event dict (synth.event_to_host consumes it);
    node rows keep the CSV order, `node_idx` keeps the original ids."""
    nodes = np.genfromtxt(event_prefix + "nodes.csv", delimiter=",", names=True)
    lo, hi = min_volume * 1000, (max_volume + 1) * 1000
    keep = (nodes["layer_id"] >= lo) & (nodes["layer_id"] <= hi)
    nodes = nodes[keep]
    idx = nodes["node_idx"].astype(np.int64)
    layer_id = nodes["layer_id"].astype(np.int64)
    x, y, z = nodes["x"].astype(np.float64), nodes["y"].astype(np.float64), nodes["z"].astype(np.float64)
    with open(event_prefix + "edges.csv") as f:
        f.readline()                       # "<n_nodes> <n_edges>"
        header = f.readline().strip().split(",")
        e = np.loadtxt(f, delimiter=",", ndmin=2)
    c2, c1 = header.index("node2"), header.index("node1")
    n2, n1 = e[:, c2].astype(np.int64), e[:, c1].astype(np.int64)
    top = max([int(a.max()) for a in (idx, n1, n2) if a.size] + [0])
    pos = -np.ones(top + 1, np.int64)
    pos[idx] = np.arange(len(idx))
    ok = (pos[n1] >= 0) & (pos[n2] >= 0)
    ev = {
        "x": x, "y": y, "z": z, "r": radius_like_reference(x, y),
        "layer": (layer_id % 100).astype(np.int32), "volume": (layer_id // 1000).astype(np.int32),
        "layer_id_mod1000": (layer_id % 1000).astype(np.int32),
        "truth": np.full(len(idx), -1, np.int64) if truth is None else np.asarray(truth, np.int64)[keep],
        "edge_a": pos[n1[ok]].astype(np.int32), "edge_b": pos[n2[ok]].astype(np.int32),   # add_edge(node1, node2) first
        "node_idx": idx,
    }
    return ev


def write_event_csv(ev, event_prefix, layer_id=None):
    """inverse, for tests / interop: writes nodes.csv and edges.csv in the reference's format"""
    n = len(ev["x"])
    lid = (ev["volume"].astype(np.int64) * 1000 + ev["layer"]) if layer_id is None else layer_id
    with open(event_prefix + "nodes.csv", "w") as f:
        f.write("node_idx,layer_id,x,y,z\n")
        for i in range(n):
            f.write("%d,%d,%r,%r,%r\n" % (i, lid[i], float(ev["x"][i]), float(ev["y"][i]), float(ev["z"][i])))
    with open(event_prefix + "edges.csv", "w") as f:
        f.write("%d %d\n" % (n, len(ev["edge_a"])))
        f.write("node2,node1,weight\n")
        for a, b in zip(ev["edge_a"].tolist(), ev["edge_b"].tolist()):
            f.write("%d,%d,1.0\n" % (b, a))
