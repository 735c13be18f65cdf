"""Generate the LUT-threshold and tag-propagation fixtures from the UNMODIFIED reference.  Build-container only.

    python tests/golden/make_lut_tag_golden.py

tests/golden/lut_barrel40.npz  (SURVEY.md 8c last row / 8f rank 4)
    The reference's own `cluster()` (clustering/clustering.py:149) with `KL_threshold` swapped per node by
    ref_harness.NodeKLThreshold: threshold = kl_max of the node's emp_var bin in the SHIPPED table
    learn_KL_linear_model/output/empvar/empvar.lut (bin = floor(emp_var / 0.05) clipped to 0..27,
    emp_var = node['xy_edge_gradient_mean_var'][1], utilities/helper.py:446).
      topo_*            static layout incl. topo_emp_var (the reference's own np.var)
      lut               the 28 kl_max values read from the shipped file
      lut_stress        a second table, 10^(2 + b/4): the shipped values (0..32) sit below almost every KL distance of these
                        events (seeds: 30..2.6e9), so they pin "nothing absorbed"; the stress table spans the KL range and
                        makes the per-node threshold decide (hundreds of flags differ from the scalar run)
      seed/*, c1lut/*, c1str/*   state before / after cluster('track_state_estimates', 1.0, table)
      m2/*, c3lut/*, c3str/*     state before / after cluster('updated_track_states', 1000, table) on the reference's own
                        iteration-2 output (C1, X1, E+R, X2, meta as in run_gnn_trackml_mod.sh:71-148)
      *_scalar_diff     number of activation flags that differ from the scalar-threshold run

tests/golden/tagprop_barrel30.npz  
    tag_propagation/tag_propagation.py executed unmodified (ref_harness.run_tag_propagation) on a seeded synthetic event
    given as ONE DiGraph with randomly permuted node ids (= initial tags): both directions for most doublets, a random third of the directed edges removed (so the
    successor-only rule :99-110 is exercised), five isolated hits (removed by :75-92).
      topo_*            flat layout of the graph (nxio.graphs_to_host)
      tags0, tags       initial tag (= node id) and final tag per node row
      sweeps, nwork     number of sweeps of the while loop (:137), size of the work list (:135)
"""
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)
warnings.filterwarnings("ignore")

import ref_harness as rh  # noqa: E402

rh.setup_reference()
import gtf_b200  # noqa: E402,F401
from gtf_b200 import nxio, synth, fields, stages  # noqa: E402
import make_golden as mg  # noqa: E402

LUT_FILE = os.path.join(rh.REF, "learn_KL_linear_model", "output", "empvar", "empvar.lut")


def flat_state(canon, graphs, prev):
    return mg.canonicalize(canon, nxio.graphs_to_host(graphs), prev, {})


def make_lut():
    P = rh.PARAMS
    lut = stages.load_lut(LUT_FILE)
    ev = synth.barrel_event(n_tracks=40, seed=2003, eta_max=1.0, target_degree=10.0)
    graphs = rh.seed_graphs(nxio.events_to_graphs(ev))
    canon = nxio.graphs_to_host(graphs)
    full = fields.complete_host_batch({k: v for k, v in canon.items() if k not in ("truth", "orig_id", "in_key")})
    seed = {f: full[f].copy() for f in mg.MUTABLE}
    root = tempfile.mkdtemp(prefix="gtf_lut_")
    d0 = os.path.join(root, "seed/")
    rh.save_graphs(graphs, d0)
    data = {"topo_" + k: canon[k] for k in mg.TOPO}
    data["lut"] = lut
    stress = 10.0 ** (2.0 + 0.25 * np.arange(28))
    data["lut_stress"] = stress

    def put(stage, st):
        for f in mg.MUTABLE:
            data["%s/%s" % (stage, f)] = st[f]

    put("seed", seed)
    # iteration 1 in LUT mode and, for the record, in scalar mode
    g_lut = rh.run_cluster_lut(d0, os.path.join(root, "c1lut/"), "track_state_estimates", P["chi2_c1"], lut, P, 1)
    c1lut = flat_state(canon, g_lut, seed)
    put("c1lut", c1lut)
    g_str = rh.run_cluster_lut(d0, os.path.join(root, "c1str/"), "track_state_estimates", P["chi2_c1"], stress, P, 1)
    c1str = flat_state(canon, g_str, seed)
    put("c1str", c1str)
    g_sc = rh.run_cluster(d0, os.path.join(root, "it1/network/"), "track_state_estimates", P["chi2_c1"], P["kl_c1"], P, 1)
    c1 = flat_state(canon, g_sc, seed)
    data["c1lut_scalar_diff"] = np.array(int((c1["active"] != c1lut["active"]).sum()))
    data["c1str_scalar_diff"] = np.array(int((c1["active"] != c1str["active"]).sum()))
    # the reference's own schedule up to the metadata update, then iteration 3 in LUT mode
    it1, it2 = os.path.join(root, "it1"), os.path.join(root, "it2")
    cand1, rem1, frag1, _ = rh.run_extract(os.path.join(it1, "network/"), it1, 1, P)
    fr = sorted(set(int(canon["sub"][list(canon["orig_id"]).index(int(next(iter(g.nodes()))))]) for g in frag1))
    x1 = mg.canonicalize(canon, nxio.graphs_to_host(rem1), c1, {"fragment": fr})
    g = rh.run_extrapolate(os.path.join(it1, "remaining/"), os.path.join(it2, "network/"), P)
    e2 = flat_state(canon, g, x1)
    cand2, rem2, frag2, _ = rh.run_extract(os.path.join(it2, "network/"), it2, 2, P, prev_candidates=cand1)
    fr = sorted(set(int(canon["sub"][list(canon["orig_id"]).index(int(next(iter(g.nodes()))))]) for g in frag2))
    x2 = mg.canonicalize(canon, nxio.graphs_to_host(rem2), e2, {"fragment": fr})
    g = rh.run_metadata(os.path.join(it2, "remaining/"))
    m2 = flat_state(canon, g, x2)
    put("m2", m2)
    g_lut = rh.run_cluster_lut(os.path.join(it2, "remaining/"), os.path.join(root, "c3lut/"), "updated_track_states",
                               P["chi2_c3"], lut, P, 3)
    c3lut = flat_state(canon, g_lut, m2)
    put("c3lut", c3lut)
    g_str = rh.run_cluster_lut(os.path.join(it2, "remaining/"), os.path.join(root, "c3str/"), "updated_track_states",
                               P["chi2_c3"], stress, P, 3)
    c3str = flat_state(canon, g_str, m2)
    put("c3str", c3str)
    g_sc = rh.run_cluster(os.path.join(it2, "remaining/"), os.path.join(root, "c3/"), "updated_track_states",
                          P["chi2_c3"], P["kl_c3"], P, 3)
    c3 = flat_state(canon, g_sc, m2)
    data["c3lut_scalar_diff"] = np.array(int((c3["active"] != c3lut["active"]).sum()))
    data["c3str_scalar_diff"] = np.array(int((c3["active"] != c3str["active"]).sum()))
    path = os.path.join(HERE, "lut_barrel40.npz")
    np.savez_compressed(path, **data)
    print("lut_barrel40: N=%d E=%d, flags differing from the scalar run: c1 %d / %d (shipped / stress table), c3 %d / %d "
          "-> %.1f KB" % (len(canon["x"]), len(canon["in_src"]), data["c1lut_scalar_diff"], data["c1str_scalar_diff"],
                          data["c3lut_scalar_diff"], data["c3str_scalar_diff"], os.path.getsize(path) / 1024))


def make_tag():
    import networkx as nx
    ev = synth.barrel_event(n_tracks=30, seed=2004, eta_max=0.5, target_degree=6.0)
    G = nx.compose_all(nxio.events_to_graphs(ev))
    rng = np.random.default_rng(2004)
    # hit ids grow with the radius along a track, so max-propagation from lower radii would never flip a tag: relabel randomly
    perm = rng.permutation(G.number_of_nodes())
    G = nx.relabel_nodes(G, {n: int(perm[n]) for n in G.nodes()}, copy=True)
    for n in G.nodes():
        G.nodes[n]["tags"] = [n]
    edges = sorted(G.edges())
    drop = [edges[k] for k in np.nonzero(rng.random(len(edges)) < 1.0 / 3.0)[0]]
    G.remove_edges_from(drop)
    n0 = G.number_of_nodes()
    for k in range(5):      # isolated hits
        i = n0 + k
        G.add_node(i, xy=(1.0 * k, 2.0), zr=(0.0, 40.0 + k), xyzr=(1.0 * k, 2.0, 0.0, 40.0 + k), volume_id=8,
                   in_volume_layer_id=2, vivl_id=(8, 2), truth_particle=-1, tags=[i])
    for u, v in G.edges():
        G[u][v]["activated"] = 1
    canon = nxio.graphs_to_host([G])
    tags, sweeps, nwork = rh.run_tag_propagation(G)
    oid = canon["orig_id"]
    final = np.array([tags.get(int(o), int(o)) for o in oid], np.int32)    # isolated nodes are removed: tag untouched
    data = {"topo_" + k: canon[k] for k in mg.TOPO}
    data.update(tags0=oid.astype(np.int32), tags=final, sweeps=np.array(sweeps), nwork=np.array(nwork))
    path = os.path.join(HERE, "tagprop_barrel30.npz")
    np.savez_compressed(path, **data)
    print("tagprop_barrel30: N=%d E=%d, %d sweeps, work list %d, %d nodes changed tag -> %.1f KB" % (
        len(oid), len(canon["in_src"]), sweeps, nwork, int((final != oid).sum()), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    which = sys.argv[1:] or ["lut", "tag"]
    if "lut" in which:
        make_lut()
    if "tag" in which:
        make_tag()
