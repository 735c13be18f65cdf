"""Known-answer fixture for KLDistance from the reference's ONE shipped golden file
  learn_KL_parabolic_model/src/output/track_sim_trackml_parabolic_model/minCurv_0.3_134/event_graph_data/1_events_training_data.csv
(7,574 rows kl_dist, emp_var, truth).  Build-container only.

The CSV was produced by extract_metadata_trackml_parabolic_model.py:14-97 from the volume-7 graph of the sibling
event_network/ CSVs, seeded with the parabolic-model variant of compute_track_state_estimates
(generate_training_data/utils.py:221-299).  This script rebuilds that graph, seeds it with the REFERENCE'S OWN
function (imported, not restated), and stores the per-node component means / full 3x3 covariances together with the
CSV's kl_dist column, so tests can check the oracle's KLDistance (element-wise trace quirk) against it as a sorted
multiset (row order in the CSV is glob/hash dependent)."""
import os
import sys
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
BASE = "/root/reference/learn_KL_parabolic_model/src"
EV = BASE + "/output/track_sim_trackml_parabolic_model/minCurv_0.3_134"
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.join(REPO, "oracle", "refshim"))
sys.path.insert(0, BASE)
sys.path.insert(0, BASE + "/generate_training_data")

import networkx as nx  # noqa: E402
import utils as ref_utils  # noqa: E402  (the reference's generate_training_data/utils.py)


class M(object):
    def __init__(self, x, y):
        self.x, self.y = x, y


def main():
    nodes = pd.read_csv(EV + "/event_network/event_1_filtered_graph_nodes.csv")
    nodes = nodes.loc[nodes["layer_id"] <= 7999]
    edges = pd.read_csv(EV + "/event_network/event_1_filtered_graph_edges.csv", skiprows=1)
    G = nx.DiGraph()
    for r in nodes.itertuples():
        G.add_node(int(r.node_idx), GNN_Measurement=M(r.x, r.y))
    inset = set(G.nodes())
    for n2, n1 in zip(edges["node2"].astype(int), edges["node1"].astype(int)):
        if n1 in inset and n2 in inset:          # utils.py:355-361
            G.add_edge(n1, n2)
            G.add_edge(n2, n1)
    ref_utils.compute_track_state_estimates([G])
    means, covs, off, emp = [], [], [0], []
    for n, attr in G.nodes(data=True):
        if ref_utils.query_node_degree_in_edges.__code__.co_argcount == 2:
            pass
        deg = G.in_degree(n)
        if deg <= 1:
            continue
        tse = attr["track_state_estimates"]
        for comp in tse.values():
            means.append(np.asarray(comp["edge_state_vector"], float))
            covs.append(np.asarray(comp["edge_covariance"], float).reshape(9))
        off.append(len(means))
        emp.append(attr["xy_edge_gradient_mean_var"][1])
    want = pd.read_csv(EV + "/event_graph_data/1_events_training_data.csv")
    npairs = sum((b - a) * (b - a - 1) // 2 for a, b in zip(off[:-1], off[1:]))
    print("nodes with >= 2 components:", len(off) - 1, "components:", len(means), "pairs:", npairs, "csv rows:", len(want))
    assert npairs == len(want)
    out = os.path.join(HERE, "kl_parabolic_known_answer.npz")
    np.savez_compressed(out, mean=np.array(means), cov=np.array(covs), off=np.array(off, np.int32), emp_var=np.array(emp),
                        kl_sorted=np.sort(want["kl_dist"].to_numpy()), emp_var_sorted=np.sort(want["emp_var"].to_numpy()))
    print("wrote", out, os.path.getsize(out) // 1024, "KB")


if __name__ == "__main__":
    main()
