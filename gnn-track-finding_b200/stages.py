"""Drop-in stage functions: the reference's names and argument order, executed on the GPU.

    from gtf_b200.stages import message_passing, reweight, compute_prior_probabilities, cluster, CCA ...

Each call flattens the given `nx.DiGraph` list (nxio.graphs_to_host), runs ONE C-ABI stage on the device and
writes the result back into the same graph objects, so the existing Python driver code keeps working.  For
throughput keep the data on the device instead: build one `EventBatch` and call its methods (batch.py);
these wrappers pay a host round trip per call by design (that is the reference's own calling convention).
"""
import glob
import os
import pickle
import re

import numpy as np

from . import nxio
from .batch import EventBatch


def _run(graphs, fn, geom=(0.3, 0.4, 0.6, 550.0)):
    hb = nxio.graphs_to_host(graphs)
    orig = hb.pop("orig_id")
    hb.pop("truth")
    hb.pop("in_key")
    b = EventBatch(hb, geom=geom)
    try:
        out = fn(b)
        res = b.download()
    finally:
        b.close()
    # graphs_to_host gives removed neighbours src = -1; such slots never exist on the device side
    nxio.apply_host_to_graphs(res, graphs, orig, subs=list(range(len(graphs))))
    return out


def initialize_edge_activation(GraphList):
    """helper.py:24"""
    return _run(GraphList, lambda b: b.initialize_edge_activation())


def compute_prior_probabilities(GraphList, track_state_key):
    """helper.py:30"""
    return _run(GraphList, lambda b: b.compute_prior_probabilities(track_state_key))


def compute_mixture_weights(GraphList, TRACK_STATE_KEY):
    """helper.py:76"""
    return _run(GraphList, lambda b: b.compute_mixture_weights(TRACK_STATE_KEY))


def query_node_degree_in_edges(subGraph, node_num):
    """helper.py:67 (a per-node host query in the reference; the batch form is EventBatch.query_node_degree_in_edges)"""
    return sum(1 for u, _ in subGraph.in_edges(node_num) if subGraph[u][node_num]["activated"] == 1)


def reweight(subGraphs, track_state_estimates_key):
    """helper.py:143"""
    return _run(subGraphs, lambda b: b.reweight(track_state_estimates_key))


def compute_track_state_estimates(GraphList, sigma0xy, sigma0rz, sigma0rz2, endcap_boundary):
    """helper.py:238.  The neighbour (dict) order is an input of the algorithm (set-iteration order in the
    reference); here it is the predecessor order of each node unless the graphs already carry seeded dicts."""
    _run(GraphList, lambda b: b.compute_track_state_estimates(), geom=(sigma0xy, sigma0rz, sigma0rz2, endcap_boundary))
    return GraphList


def message_passing(subGraphs, chi2CutFactor, sigma0xy, sigma0rz, sigma0rz2, endcap_boundary):
    """extrapolate_merged_states.py:406"""
    return _run(subGraphs, lambda b: b.message_passing(chi2CutFactor), geom=(sigma0xy, sigma0rz, sigma0rz2, endcap_boundary))


def _numeric_glob(pattern):
    def key(p):
        m = re.search(r"(\d+)_subgraph\.gpickle$", p)
        return int(m.group(1)) if m else -1
    return sorted(glob.glob(pattern), key=key)


def _load_dir(d):
    out = []
    for f in _numeric_glob(d + "*_subgraph.gpickle"):
        with open(f, "rb") as fh:
            out.append(pickle.load(fh))
    return out


def _save_dir(graphs, d):
    os.makedirs(d, exist_ok=True)
    for i, g in enumerate(graphs):
        with open(os.path.join(d, "%d_subgraph.gpickle" % i), "wb") as fh:
            pickle.dump(g, fh, pickle.HIGHEST_PROTOCOL)


def load_lut(path):
    """`bin kl_min kl_max` per line (learn_KL_linear_model/create_lut/plot_lut.py:6-17) -> 28 kl_max values"""
    rows = np.loadtxt(path).reshape(-1, 3)
    lut = np.full(28, rows[-1, 2])
    for b_, _, kmax in rows:
        if 0 <= int(b_) < 28:
            lut[int(b_)] = kmax
    return lut


def cluster(inputDir, outputDir, track_state_key, chi2_threshold, KL_threshold, KL_lut, iteration_num, reactivate,
            sigma0rz, sigma0rz2, endcap_boundary, sigma0xy=0.3, use_lut=False):
    """clustering.py:149 -- file-driven like the reference: reads `inputDir*_subgraph.gpickle`, writes
    `outputDir{i}_subgraph.gpickle`.  KL_lut is ignored unless use_lut=True (the reference never reads it,
    SURVEY.md §0.3); `reactivate` is not supported (broken upstream, clustering.py:141-144)."""
    if reactivate:
        raise NotImplementedError("reactivate=True is broken in the reference (clustering.py:141-144)")
    graphs = _load_dir(inputDir)
    lut = load_lut(KL_lut) if (use_lut and KL_lut) else None
    st = _run(graphs, lambda b: b.cluster(track_state_key, chi2_threshold, KL_threshold, lut),
              geom=(sigma0xy, sigma0rz, sigma0rz2, endcap_boundary))
    _save_dir(graphs, outputDir)
    return st


def CCA(subCopy):
    """extract_track_candidates.py:332 -- removes inactive edges from `subCopy` and returns the components
    as sub-graph copies (or [subCopy] when nothing was removed)."""
    inactive = [(u, v) for u, v in subCopy.edges() if subCopy[u][v]["activated"] == 0]
    if not inactive:
        return [subCopy]
    hb = nxio.graphs_to_host([subCopy])
    orig = hb.pop("orig_id")
    hb.pop("truth")
    hb.pop("in_key")
    b = EventBatch(hb)
    try:
        lab = b.CCA()
    finally:
        b.close()
    subCopy.remove_edges_from(inactive)
    comps = {}
    for i, l in enumerate(lab):
        comps.setdefault(int(l), []).append(int(orig[i]))
    return [subCopy.subgraph(nodes).copy() for _, nodes in sorted(comps.items())]


def KLDistance_pairs(means, covs, offsets, device=0):
    """Pairwise `KLDistance` (clustering.py:90-94, general 3x3 covariances) between the components of every group, on the
    GPU -- the inner loop of the reference's KL-threshold LUT training-data generator (compute_KL_distance.py:11-21).
    means: (M, 3), covs: (M, 3, 3) or (M, 9), offsets: (G + 1,) component ranges.  Returns the KL values of all pairs
    (i, j < i), group by group."""
    import ctypes
    from . import lib as L
    means = np.ascontiguousarray(means, np.float64).reshape(-1, 3)
    covs = np.ascontiguousarray(covs, np.float64).reshape(-1, 9)
    offsets = np.ascontiguousarray(offsets, np.int32)
    assert len(means) == len(covs) and len(offsets) >= 1 and offsets[-1] <= len(means)
    lib = L.lib()
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
    n = ctypes.c_int64(0)
    args = (device, means.ctypes.data_as(dp), covs.ctypes.data_as(dp), offsets.ctypes.data_as(ip), len(offsets) - 1)
    L.check(lib.gtf_kl_pairs(*args, None, 0, ctypes.byref(n)))
    out = np.empty(n.value, np.float64)
    if n.value:
        L.check(lib.gtf_kl_pairs(*args, out.ctypes.data_as(dp), n.value, ctypes.byref(n)))
    return out
