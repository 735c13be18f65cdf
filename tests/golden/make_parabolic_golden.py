"""tests/golden/parabolic_seed.npz: the parabolic-model seeding of the KL look-up-table training pipeline,
learn_KL_parabolic_model/src/generate_training_data/utils.py:221-299 (compute_track_state_estimates + rotate_track :197-218),
called UNMODIFIED on a seeded toy graph.  Build-container only.

Stored per directed (node, neighbour) pair, in the reference's own order (nodes in graph order, keys of the node's
`track_state_estimates` dict in dict order): node / neighbour ids and (x, y), `edge_state_vector` (3), `edge_covariance` (3 x 3);
per node: `xy_edge_gradient_mean_var`."""
import importlib.util
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")
import ref_harness as rh  # noqa: E402

rh.setup_reference()
REF = "/root/reference/learn_KL_parabolic_model/src"


def load_utils():
    sys.path.insert(0, os.path.join(REF, "GNN_Measurement"))
    spec = importlib.util.spec_from_file_location("ref_parabolic_utils", os.path.join(REF, "generate_training_data", "utils.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def toy_graph(gnn, seed=5, n_tracks=12, n_layers=8):
    """hits of n_tracks near-circular tracks on n_layers barrel layers; directed edges to the hits of the next layer within a
    window (both true and fake neighbours), plus a few in-edges: nx.all_neighbors = predecessors then successors"""
    import networkx as nx
    rng = np.random.default_rng(seed)
    G = nx.DiGraph()
    radii = np.linspace(32.0, 500.0, n_layers)
    hits = []
    for t in range(n_tracks):
        phi0 = rng.uniform(-np.pi, np.pi)
        curv = rng.normal() * 4e-4
        for l, r in enumerate(radii):
            phi = phi0 + curv * r + rng.normal() * 1e-4
            hits.append((l, t, r * np.cos(phi), r * np.sin(phi), rng.normal() * 100.0))
    for i, (l, t, x, y, z) in enumerate(hits):
        G.add_node(i, GNN_Measurement=gnn(x, y, z, np.hypot(x, y)), xy=(x, y))     # (only .x and .y are read)
    for i, (l, t, x, y, z) in enumerate(hits):
        for j, (l2, t2, x2, y2, z2) in enumerate(hits):
            if l2 == l + 1 and np.hypot(x2 - x, y2 - y) < 140.0:
                G.add_edge(i, j)
    return G


def main():
    u = load_utils()
    gnn = u.gnn if isinstance(u.gnn, type) else u.gnn.GNN_Measurement    # (`from GNN_Measurement import GNN_Measurement`: module or class,
    G = toy_graph(gnn)                                                     #  depending on which directory of the tree is on sys.path)
    with rh.quiet("/tmp"):
        u.compute_track_state_estimates([G])
    node, nbr, nxy, bxy, sv, cov, gmv, gnode = [], [], [], [], [], [], [], []
    for n in G.nodes():
        a = G.nodes[n]
        m = a["GNN_Measurement"]
        gmv.append(a["xy_edge_gradient_mean_var"])
        gnode.append(n)
        for k, e in a["track_state_estimates"].items():
            b = G.nodes[k]["GNN_Measurement"]
            node.append(n); nbr.append(k); nxy.append((m.x, m.y)); bxy.append((b.x, b.y))
            sv.append(e["edge_state_vector"]); cov.append(e["edge_covariance"])
    out = dict(node=np.array(node, np.int32), nbr=np.array(nbr, np.int32), node_xy=np.array(nxy), nbr_xy=np.array(bxy),
               state=np.array(sv), cov=np.array(cov), grad_node=np.array(gnode, np.int32), grad_mean_var=np.array(gmv, dtype=np.float64),
               edges=np.array(list(G.edges()), np.int32), n_nodes=np.int32(G.number_of_nodes()))
    np.savez_compressed(os.path.join(HERE, "parabolic_seed.npz"), **out)
    print("pairs", len(node), "nodes", len(gnode), "isolated nodes (NaN mean/var):", int(np.isnan(out["grad_mean_var"][:, 0]).sum()))


if __name__ == "__main__":
    main()
