"""Run the UNMODIFIED reference (/root/reference/src) in-process, stage by stage, the way
run_gnn_trackml_mod.sh:71-148 does, and capture flat snapshots after every stage.

Test infrastructure only (build container; /root/reference does not exist on the GPU box).
Used by make_golden.py to produce tests/golden/*.npz and by timing probes.

Environment control applied around the reference (none of it changes arithmetic):
  * import shims from oracle/refshim (filterpy restatement, matplotlib stub, more_itertools.locate)
  * nx.read_gpickle / nx.write_gpickle re-created with pickle (removed in networkx 3)
  * glob.glob sorted numerically, so the list order of sub-graphs is stable between stages
  * cwd = scratch dir (the reference appends per-edge CSV lines to cwd, extrapolate...py:171-292)
  * stdout discarded (MBs of prints)
  * ZeroDivisionError from the trailing *metrics* blocks of helper.reweight (helper.py:216-223) and
    message_passing (extrapolate...py:509-516) swallowed: they fire after all mutation is done
"""
import contextlib
import glob as _glob
import io
import os
import pickle
import re
import sys
import tempfile

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
REF = "/root/reference"

_ready = False


def setup_reference():
    global _ready
    if _ready:
        return
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree not present (build container only)")
    sys.path.insert(0, os.path.join(REPO, "oracle", "refshim"))
    sys.path.insert(0, os.path.join(REF, "src"))
    import networkx as nx

    def read_gpickle(path):
        with open(path, "rb") as f:
            return pickle.load(f)

    def write_gpickle(G, path):
        with open(path, "wb") as f:
            pickle.dump(G, f, pickle.HIGHEST_PROTOCOL)

    nx.read_gpickle = read_gpickle
    nx.write_gpickle = write_gpickle

    real_glob = _glob.glob

    def sorted_glob(pat, *a, **k):
        out = real_glob(pat, *a, **k)

        def key(p):
            m = re.search(r"(\d+)_subgraph\.gpickle$", p)
            return int(m.group(1)) if m else -1
        return sorted(out, key=key)

    _glob.glob = sorted_glob

    from utilities import helper as h
    real_reweight = h.reweight

    def guarded_reweight(*a, **k):
        try:
            return real_reweight(*a, **k)
        except ZeroDivisionError:
            return None
    h.reweight = guarded_reweight

    from extrapolate import extrapolate_merged_states as em
    real_mp = em.message_passing

    def guarded_mp(*a, **k):
        try:
            return real_mp(*a, **k)
        except ZeroDivisionError:
            return None
    em.message_passing = guarded_mp
    em._real_message_passing = real_mp
    _ready = True


@contextlib.contextmanager
def quiet(cwd):
    old = os.getcwd()
    os.chdir(cwd)
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            yield
    finally:
        os.chdir(old)


def save_graphs(graphs, d):
    import networkx as nx
    os.makedirs(d, exist_ok=True)
    for i, g in enumerate(graphs):
        nx.write_gpickle(g, os.path.join(d, "%d_subgraph.gpickle" % i))


def load_graphs(d):
    import networkx as nx
    out = []
    i = 0
    while os.path.isfile(os.path.join(d, "%d_subgraph.gpickle" % i)):
        out.append(nx.read_gpickle(os.path.join(d, "%d_subgraph.gpickle" % i)))
        i += 1
    return out


PARAMS = dict(sigma0xy=0.3, sigma0rz=0.4, sigma0rz2=0.6, endcap=550.0,
              chi2_c1=1.0, kl_c1=2.0, chi2_cut=2.0, chi2_c3=1000.0, kl_c3=100.0,
              pval=0.01, numhits=4, sep3d=10.0, merge_dist=8.0)


def seed_graphs(graphs, P=PARAMS, scratch=None):
    """event_conversion.py:87-96: seed states, activation, priors, weights, degree."""
    from utilities import helper as h
    with quiet(scratch or tempfile.mkdtemp()):
        graphs = h.compute_track_state_estimates(graphs, P["sigma0xy"], P["sigma0rz"], P["sigma0rz2"], P["endcap"])
        h.initialize_edge_activation(graphs)
        h.compute_prior_probabilities(graphs, "track_state_estimates")
        h.compute_mixture_weights(graphs, "track_state_estimates")
        for s in graphs:
            for n, _ in s.nodes(data=True):
                s.nodes[n]["degree"] = h.query_node_degree_in_edges(s, n)
    return graphs


def run_cluster(in_dir, out_dir, key, chi2_thr, kl_thr, P=PARAMS, it=1):
    from clustering import clustering as cl
    os.makedirs(out_dir, exist_ok=True)
    with quiet(os.path.dirname(out_dir.rstrip("/"))):
        cl.cluster(in_dir, out_dir, key, chi2_thr, kl_thr, None, it, False,
                   P["sigma0rz"], P["sigma0rz2"], P["endcap"])
    return load_graphs(out_dir)


class NodeKLThreshold(object):
    """LUT-threshold mode (SURVEY.md 8c last row): the reference's own `cluster()` with `KL_threshold` swapped per node.
    Passed as the `KL_threshold` argument; `smallest_dist < KL_threshold` (clustering.py:261) on a numpy scalar defers to
    this object's reflected comparison (`__array_ufunc__ = None`), which reads the node being processed from the calling
    frame (`node_attr`, clustering.py:195) and looks its threshold up: bin = floor(emp_var / 0.05) clipped to 0..27,
    emp_var = node['xy_edge_gradient_mean_var'][1] (helper.py:446)."""
    __array_ufunc__ = None

    def __init__(self, lut):
        self.lut = [float(v) for v in lut]
        self.seen = 0

    def threshold_of(self, attr):
        import math
        ev = attr["xy_edge_gradient_mean_var"][1]
        b = 27 if ev != ev else int(math.floor(ev / 0.05))
        return self.lut[max(0, min(27, b))]

    def __gt__(self, other):
        self.seen += 1
        return bool(other < self.threshold_of(sys._getframe(1).f_locals["node_attr"]))


def run_cluster_lut(in_dir, out_dir, key, chi2_thr, lut, P=PARAMS, it=1):
    from clustering import clustering as cl
    os.makedirs(out_dir, exist_ok=True)
    thr = NodeKLThreshold(lut)
    with quiet(os.path.dirname(out_dir.rstrip("/"))):
        cl.cluster(in_dir, out_dir, key, chi2_thr, thr, None, it, False, P["sigma0rz"], P["sigma0rz2"], P["endcap"])
    assert thr.seen > 0
    return load_graphs(out_dir)


def run_tag_propagation(graph, scratch=None):
    """tag_propagation/tag_propagation.py executed UNMODIFIED on `graph` (saved as 0_subgraph.gpickle in a scratch cwd).
    The script is module-level code; its tail indexes unique_colours[200] (:208-209, IndexError on small graphs) after the
    labels are final, so the IndexError is caught and the result read from the script's namespace.
    Returns ({node: final tag}, number of sweeps, size of the work list)."""
    import networkx as nx
    path = os.path.join(REF, "tag_propagation", "tag_propagation.py")
    scratch = scratch or tempfile.mkdtemp(prefix="gtf_tag_")
    nx.write_gpickle(graph, os.path.join(scratch, "0_subgraph.gpickle"))
    saved = {n: getattr(nx, n) for n in ("draw_networkx_edges", "draw_networkx_nodes", "draw_networkx_labels")}
    for n in saved:
        setattr(nx, n, lambda *a, **k: None)      # they import matplotlib.collections internally
    ns = {"__name__": "__tag_propagation__", "__file__": path}
    try:
        with quiet(scratch):
            try:
                exec(compile(open(path).read(), path, "exec"), ns)
            except IndexError:
                pass
    finally:
        for n, f in saved.items():
            setattr(nx, n, f)
    g = ns["current_endcap_graph"]
    return {int(n): int(g.nodes[n]["tags"][-1]) for n in g.nodes()}, len(ns["frac_tags_flipped"]), ns["total_number_of_nodes_to_process"]


def _call_main(mod, argv, cwd):
    old = sys.argv
    sys.argv = [mod.__name__] + [str(a) for a in argv]
    try:
        with quiet(cwd):
            mod.main()
    finally:
        sys.argv = old


def run_extrapolate(in_dir, out_dir, P=PARAMS):
    from extrapolate import extrapolate_merged_states as em
    os.makedirs(out_dir, exist_ok=True)
    _call_main(em, ["-i", in_dir, "-o", out_dir, "-c", P["chi2_cut"], "-e", P["sigma0xy"],
                    "-z", P["sigma0rz"], "-m", P["sigma0rz2"], "-b", P["endcap"]],
               os.path.dirname(out_dir.rstrip("/")))
    return load_graphs(out_dir)


def run_extract(in_dir, root, it, P=PARAMS, prev_candidates=None):
    """extract_track_candidates.main(): returns (candidates, remaining, fragments) graph lists.
    `candidates` holds this iteration's candidates first, then earlier ones (extract...py:471-484)."""
    from extract import extract_track_candidates as ex
    cand, rem, frag = [os.path.join(root, d) + "/" for d in ("candidates", "remaining", "fragments")]
    for d in (cand, rem, frag):
        os.makedirs(d, exist_ok=True)
    if prev_candidates:
        save_graphs(prev_candidates, cand)
    _call_main(ex, ["-i", in_dir, "-c", cand, "-r", rem, "-f", frag, "-p", P["pval"], "-n", P["numhits"],
                    "-s", P["sep3d"], "-t", P["merge_dist"], "-a", it, "-e", P["sigma0xy"],
                    "-z", P["sigma0rz"], "-b", P["endcap"]], root)
    import pandas as pd
    pv = pd.read_csv(os.path.join(cand, "pvals.csv"))
    return load_graphs(cand), load_graphs(rem), load_graphs(frag), pv


def run_metadata(rem_dir):
    from update import remove_state_metadata as rm
    _call_main(rm, ["-r", rem_dir], os.path.dirname(rem_dir.rstrip("/")))
    return load_graphs(rem_dir)


def reference_schedule(graphs, P=PARAMS, root=None):
    """The as-is schedule: C1, X1, E+R (it 2), X2, meta, C3, X3.  Returns {stage: payload}."""
    root = root or tempfile.mkdtemp(prefix="gtf_ref_")
    out = {}
    d0 = os.path.join(root, "seed/")
    save_graphs(graphs, d0)
    # iteration 1
    it1 = os.path.join(root, "iteration_1")
    g = run_cluster(d0, os.path.join(it1, "network/"), "track_state_estimates", P["chi2_c1"], P["kl_c1"], P, 1)
    out["c1"] = g
    cand1, rem1, frag1, pv1 = run_extract(os.path.join(it1, "network/"), it1, 1, P)
    out["x1"] = (cand1, rem1, frag1, pv1)
    # iteration 2
    it2 = os.path.join(root, "iteration_2")
    g = run_extrapolate(os.path.join(it1, "remaining/"), os.path.join(it2, "network/"), P)
    out["e2"] = g
    cand2, rem2, frag2, pv2 = run_extract(os.path.join(it2, "network/"), it2, 2, P, prev_candidates=cand1)
    out["x2"] = (cand2, rem2, frag2, pv2)
    g = run_metadata(os.path.join(it2, "remaining/"))
    out["m2"] = g
    # iteration 3
    it3 = os.path.join(root, "iteration_3")
    g = run_cluster(os.path.join(it2, "remaining/"), os.path.join(it3, "network/"), "updated_track_states",
                    P["chi2_c3"], P["kl_c3"], P, 3)
    out["c3"] = g
    cand3, rem3, frag3, pv3 = run_extract(os.path.join(it3, "network/"), it3, 3, P, prev_candidates=cand2)
    out["x3"] = (cand3, rem3, frag3, pv3)
    return out
