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


# ---------------------------------------------------------------------------------------------------------------------
# The reference's stand-alone helpers (clustering/clustering.py:11-124, extrapolate_merged_states.py:26) under their own
# names and argument order.  The arithmetic runs on the GPU through the C-ABI (one tiny launch per call: they exist for
# callers that use the helpers directly -- the stages above never go through them).
def _dp(a):
    import ctypes
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def calc_pairwise_distances_chi2(num_edges, edge_svs, edge_covs, node_coords, neighbour_coords, sigma0rz, sigma0rz2,
                                 endcap_boundary, device=0):
    """clustering.py:80 -- (num_edges, num_edges) matrix, lower triangle = mahalanobis_distance(i, j < i), zeros elsewhere"""
    from . import lib as L
    n = int(num_edges)
    sv = np.ascontiguousarray(np.asarray(edge_svs, np.float64).reshape(n, 3))
    cv = np.ascontiguousarray(np.asarray(edge_covs, np.float64).reshape(n, 9))
    nd = np.ascontiguousarray(np.asarray(node_coords, np.float64).reshape(4))
    nb = np.ascontiguousarray(np.asarray(neighbour_coords, np.float64).reshape(n, 4))
    out = np.zeros((n, n))
    L.check(L.lib().gtf_pairwise_chi2(device, n, _dp(sv), _dp(cv), _dp(nd), _dp(nb), sigma0rz, sigma0rz2, endcap_boundary, _dp(out)))
    return out


def mahalanobis_distance(mean1, cov1, mean2, cov2, node_coords, neighbour1_coords, neighbour2_coords, sigma0rz, sigma0rz2,
                         endcap_boundary, device=0):
    """clustering.py:11"""
    return calc_pairwise_distances_chi2(2, [mean2, mean1], [cov2, cov1], node_coords, [neighbour2_coords, neighbour1_coords],
                                        sigma0rz, sigma0rz2, endcap_boundary, device)[1, 0]


def KLDistance(mean1, cov1, mean2, cov2, device=0):
    """clustering.py:90 (the "trace" of the element-wise product, pinned by the reference's shipped CSV)"""
    return float(KLDistance_pairs([mean2, mean1], [cov2, cov1], [0, 2], device)[0])


def merge_states(mean1, cov1, mean2, cov2, device=0):
    """clustering.py:97 -- inverse-variance weighting; returns (merged_mean (3,), merged_cov (3, 3))"""
    from . import lib as L
    a = [np.ascontiguousarray(np.asarray(v, np.float64).reshape(-1)) for v in (mean1, cov1, mean2, cov2)]
    mm, mc = np.zeros(3), np.zeros((3, 3))
    L.check(L.lib().gtf_merge_states(device, _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), _dp(mm), _dp(mc)))
    return mm, mc


def seed_parabolic_pairs(node_xy, nbr_xy, sigma0=4.0, sigmaA=0.1, sigmaB=0.1, device=0):
    """learn_KL_parabolic_model/src/generate_training_data/utils.py:221-299 for n (node, neighbour) pairs on the GPU:
    (edge_state_vector (n, 3), edge_covariance (n, 3, 3))"""
    from . import lib as L
    a = np.ascontiguousarray(np.asarray(node_xy, np.float64).reshape(-1, 2))
    b = np.ascontiguousarray(np.asarray(nbr_xy, np.float64).reshape(-1, 2))
    if a.shape != b.shape:
        raise ValueError("node_xy and nbr_xy must hold the same number of pairs")
    n = a.shape[0]
    sv, cov = np.zeros((n, 3)), np.zeros((n, 3, 3))
    if n:
        L.check(L.lib().gtf_seed_parabolic_pairs(device, n, _dp(a), _dp(b), sigma0, sigmaA, sigmaB, _dp(sv), _dp(cov)))
    return sv, cov


def compute_track_state_estimates_parabolic(GraphList, device=0):
    """The seeding of the KL look-up-table training pipeline under the reference's name and attributes
    (learn_KL_parabolic_model/.../utils.py:221: `compute_track_state_estimates(GraphList)`): every node gets
    `track_state_estimates` = {neighbour: {'edge_state_vector', 'edge_covariance'}} over nx.all_neighbors (dict order = the
    REVERSED neighbour order, :254-255) and `xy_edge_gradient_mean_var`.  One kernel launch for all pairs of all graphs; the
    gradient mean / variance is two numpy reductions per node, as in the reference (:296)."""
    import networkx as nx
    pairs, nxy, bxy = [], [], []
    for gi, G in enumerate(GraphList):
        for node in G.nodes():
            m = G.nodes[node]["GNN_Measurement"]
            nbrs = list(nx.all_neighbors(G, node))
            grads = []
            for k in nbrs:
                b = G.nodes[k]["GNN_Measurement"]
                grads.append((m.y - b.y) / (m.x - b.x))
            for k in reversed(nbrs):
                b = G.nodes[k]["GNN_Measurement"]
                pairs.append((gi, node, k))
                nxy.append((m.x, m.y))
                bxy.append((b.x, b.y))
            G.nodes[node]["track_state_estimates"] = {}
            G.nodes[node]["xy_edge_gradient_mean_var"] = (np.mean(grads), np.var(grads))
    sv, cov = seed_parabolic_pairs(nxy, bxy, device=device)
    for (gi, node, k), s, c in zip(pairs, sv, cov):
        GraphList[gi].nodes[node]["track_state_estimates"][k] = {"edge_state_vector": s, "edge_covariance": c}
    return GraphList


def calc_dist_to_merged_state(num_edges, edge_svs, edge_covs, merged_mean, merged_cov, device=0):
    """clustering.py:107 -- list of KLDistance(component i, merged state)"""
    n = int(num_edges)
    if n == 0:
        return []
    means = np.empty((2 * n, 3))
    covs = np.empty((2 * n, 9))
    means[0::2], means[1::2] = np.asarray(merged_mean, np.float64).reshape(3), np.asarray(edge_svs, np.float64).reshape(n, 3)
    covs[0::2], covs[1::2] = np.asarray(merged_cov, np.float64).reshape(9), np.asarray(edge_covs, np.float64).reshape(n, 9)
    return [float(v) for v in KLDistance_pairs(means, covs, np.arange(0, 2 * n + 1, 2), device)]


def get_smallest_dist_idx(distances):
    """clustering.py:114 -- list: (min, first index of it); matrix: min over the NON-ZERO entries and every position holding
    it as [rows..., cols...] (ties give more than two indices; the caller uses the first two).  Index logic only."""
    if isinstance(distances, list):
        smallest = np.min(distances)
        return smallest, distances.index(smallest)
    distances = np.asarray(distances)
    smallest = np.min(distances[np.nonzero(distances)])
    rows, cols = np.where(distances == smallest)
    return smallest, np.concatenate((rows, cols), axis=None)


def extrapolate_validate(subGraph, node_num, node_attr, neighbour_num, neighbour_attr, chi2CutFactor, state_to_extrapolate,
                         state_cov, sigma0xy, sigma0rz, sigma0rz2, endcap_boundary, is_merged_state=False, device=0):
    """extrapolate_merged_states.py:26 for one edge node -> neighbour.  Like the reference it mutates `state_cov[1, 1]`
    (+= var_ms) and, on gate failure, sets the edge's `activated` to 0.  Returns the reference's 6-tuple."""
    import ctypes
    from . import lib as L

    def xyzr(attr):
        gm = attr["GNN_Measurement"]
        return np.array([gm.x, gm.y, gm.z, gm.r], np.float64)

    class EdgeResult(ctypes.Structure):
        _fields_ = [("pass_", ctypes.c_int32), ("pad", ctypes.c_int32), ("chi2", ctypes.c_double), ("var_ms", ctypes.c_double),
                    ("likelihood", ctypes.c_double), ("state", ctypes.c_double * 3), ("tau", ctypes.c_double),
                    ("cov", ctypes.c_double * 4)]

    nd, nb = xyzr(node_attr), xyzr(neighbour_attr)
    st = np.ascontiguousarray(np.asarray(state_to_extrapolate, np.float64).reshape(3))
    cov = np.ascontiguousarray(np.asarray(state_cov, np.float64).reshape(9))
    res = EdgeResult()
    geom = L.Geom(sigma0xy, sigma0rz, sigma0rz2, endcap_boundary)
    fn = L.lib().gtf_extrapolate_validate
    fn.restype = ctypes.c_int
    L.check(fn(ctypes.c_int(device), _dp(nd), _dp(nb), _dp(st), _dp(cov), ctypes.c_double(chi2CutFactor), ctypes.byref(geom),
               ctypes.byref(res)))
    state_cov[1, 1] = cov[4]                                  # in-place accumulation on the caller's matrix (quirk 2)
    same = subGraph.nodes[node_num]["truth_particle"] == subGraph.nodes[neighbour_num]["truth_particle"]
    if not res.pass_:
        subGraph[node_num][neighbour_num]["activated"] = 0
        return None, res.chi2, 0 if same else 1, 1, 0, 0
    updated_state = np.array(list(res.state))
    p00, p01, p11, p22 = list(res.cov)
    updated_cov = np.array([[p00, p01, 0.0], [p01, p11, 0.0], [0.0, 0.0, p22]])
    return {"xy": (nd[0], nd[1]), "zr": (nd[2], nd[3]), "xyzr": (nd[0], nd[1], nd[2], nd[3]),
            "edge_state_vector": updated_state, "edge_covariance": updated_cov,
            "joint_vector": [updated_state[0], updated_state[1], res.tau],
            "joint_vector_covariance": updated_cov,           # the same object, as in the reference (quirk 4)
            "likelihood": res.likelihood,
            "mixture_weight": subGraph.nodes[node_num]["track_state_estimates"][neighbour_num]["mixture_weight"]}, \
        res.chi2, 0, 0, 1 if same else 0, 1
