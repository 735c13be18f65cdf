"""Generate tests/golden/shipped_vol79.npz: the reference's SHIPPED TrackML-derived event
(src/trackml_mod/event_network/minCurv_0.3_134, volumes 7-9: 30,387 hits / 73,230 directed edges), ingested and
processed by the UNMODIFIED reference.  Build-container only.      python tests/golden/make_shipped_golden.py

  * utilities/helper.py:524-545 load_nodes_edges on the shipped CSV files,
  * utilities/helper.py:465-520 construct_graph (the truth join needs files the repository does not ship --
    .MISSING_LARGE_BLOBS -- so a placeholder truth table is supplied: one hit per node, particle label = node id mod 3 --
    mixed labels keep the precision / recall denominators of clustering.py:349-369 non-zero; labels feed only those
    printed metrics, never the arithmetic or the graph order),
  * trackml_mod/event_conversion.py:76-96: DiGraph, weakly connected sub-graphs, seeding, activation, priors, weights,
  * the schedule of run_gnn_trackml_mod.sh:71-148 (cluster, extract, extrapolate, extract, metadata, cluster, extract).

The fixture holds the hits and doublets of the selection (csv_*: what `ingest.load_event_csv` reads), the reference's flat
topology (topo_*: node order, state-dict order = iteration order of `set(nx.all_neighbors(...))`, successor order) and the
COMPACT per-stage state (decisions + merged states)."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)
warnings.filterwarnings("ignore")

import ref_harness as rh  # noqa: E402

rh.setup_reference()
import make_golden as mg  # noqa: E402
from gtf_b200 import ingest  # noqa: E402

EVENT = os.path.join(rh.REF, "src", "trackml_mod", "event_network", "minCurv_0.3_134", "event_1_filtered_graph_")
VOLUMES = (7, 9)


def main():
    import networkx as nx
    import pandas as pd
    from utilities import helper as h
    with rh.quiet("/tmp"):
        nodes, edges = h.load_nodes_edges(EVENT, *VOLUMES)
        ids = nodes["node_idx"].astype(int).to_numpy()
        truth = pd.DataFrame({"node_idx": ids, "hit_id": ids, "particle_id": ids % 3, "module_id": 0})
        G = h.construct_graph(nx.DiGraph(), nodes, edges, truth)
        G = nx.DiGraph(G)
        graphs = [G.subgraph(c).copy() for c in nx.weakly_connected_components(G)]
    ev = ingest.load_event_csv(EVENT, *VOLUMES)
    extra = {"csv_" + k: ev[k] for k in ("x", "y", "z", "layer", "volume", "edge_a", "edge_b", "node_idx")}
    extra["csv_layer_id"] = (ev["volume"].astype(np.int64) * 1000 + ev["layer_id_mod1000"]).astype(np.int32)
    mg.make_from_graphs("shipped_vol79", extra, graphs, compact=True)


if __name__ == "__main__":
    main()
