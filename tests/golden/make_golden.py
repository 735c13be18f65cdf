"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/src) on seeded
synthetic events.  Build-container only.  Usage:  python tests/golden/make_golden.py [name ...]

Each fixture holds, in the canonical flat layout of include/gtf_fields.h (node/slot indexing of the
freshly seeded event, removed nodes keep their rows):
  ev_*          the synthetic event (synth.py dict) it was generated from
  topo_*        static topology (orders as exported by networkx: node order, dict order, successor order)
  <stage>/<f>   every mutable array after stage in  seed, c1, x1, e2, x2, m2, c3, x3
                (run_gnn_trackml_mod.sh:71-148: cluster, extract, extrapolate+reweight, extract,
                 remove_state_metadata, cluster, extract)
  <x>/accepted  u8[N] nodes extracted as track candidates in that extraction; <x>/cand_label i32[N]
  <x>/pvals     (k,2) p-values of accepted candidates, in reference order
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)
warnings.filterwarnings("ignore")

import ref_harness as rh  # noqa: E402

rh.setup_reference()
import gtf_b200  # noqa: E402
from gtf_b200 import nxio, synth, fields  # noqa: E402

TOPO = ("x", "y", "z", "r", "layer", "volume", "truth", "orig_id", "sub", "sub_off", "sub_event", "in_off",
        "in_src", "slot_dst", "out_off", "out_slot", "rev_slot", "in_key", "emp_var")
MUTABLE = [f for f, _, _ in fields.FIELDS if f not in TOPO and f not in ("label", "emp_var", "uts_chi2")]

EVENTS = {
    # name: (generator kwargs)
    "barrel40_eta1": dict(n_tracks=40, seed=2001, eta_max=1.0, target_degree=10.0),
    "barrel25_deg6": dict(n_tracks=25, seed=2002, eta_max=0.5, target_degree=6.0),
    "barrel100_cfg1": dict(n_tracks=100, seed=1000, eta_max=0.5, target_degree=10.0),
    # small and dense (mean in-degree 16, maxima near 50: mostly fake edges, long merge chains): the shape on which the randomised
    # sweep sees the largest value drift between oracle and kernels
    "barrel60_deg16": dict(n_tracks=60, seed=3002, eta_max=0.5, target_degree=16.0),
    # BASELINE configs[1] size: 1000 tracks -> 10k hits / 100k directed edges (several minutes of reference time)
    "barrel1000_cfg2": dict(n_tracks=1000, seed=2000, eta_max=0.5, target_degree=10.0),
}
COMPACT = {"barrel100_cfg1", "barrel1000_cfg2", "barrel60_deg16"}   # decisions + merged states only (keeps the fixture small)


def canonicalize(canon, snap, prev, graphs_alive_subs):
    """Scatter a snapshot (nxio.graphs_to_host of the current graph list) into canonical indexing."""
    out = {k: v.copy() for k, v in prev.items()}
    node_of = {int(o): i for i, o in enumerate(canon["orig_id"])}
    slot_of = {(int(canon["in_key"][s]), int(canon["orig_id"][canon["slot_dst"][s]])): s
               for s in range(len(canon["in_src"]))}
    nmap = np.array([node_of[int(o)] for o in snap["orig_id"]], np.int64)
    smap = np.array([slot_of[(int(snap["in_key"][s]), int(snap["orig_id"][snap["slot_dst"][s]]))]
                     for s in range(len(snap["in_src"]))], np.int64)
    N, E, S = len(canon["x"]), len(canon["in_src"]), len(canon["sub_off"]) - 1
    real = snap["alive"] > 0          # ghost rows (stale dict keys of removed nodes) are not graph nodes
    alive = np.zeros(N, np.uint8)
    alive[nmap[real]] = 1
    # nodes of sub-graphs that left the list as fragments stay alive (graph dropped, nodes not removed)
    sub_state = np.full(S, 2, np.uint8)
    subs_present = set(int(canon["sub"][i]) for i in nmap[real])
    for g in range(S):
        if g in subs_present:
            sub_state[g] = 0
    for g in graphs_alive_subs.get("fragment", []):
        sub_state[g] = 1
        alive[canon["sub_off"][g]:canon["sub_off"][g + 1]] = prev["alive"][canon["sub_off"][g]:canon["sub_off"][g + 1]]
    for g in range(S):       # fragments of earlier extractions keep their state
        if prev["sub_state"][g] == 1:
            sub_state[g] = 1
            alive[canon["sub_off"][g]:canon["sub_off"][g + 1]] = prev["alive"][canon["sub_off"][g]:canon["sub_off"][g + 1]]
    out["alive"] = alive
    out["sub_state"] = sub_state
    for f in MUTABLE:
        if f in ("alive", "sub_state", "uts_next"):
            continue
        ext = fields.FIELD_EXTENT[f]
        if ext == "N":
            out[f][nmap[real]] = snap[f][real]
        elif ext == "E":
            if f in ("tse_present", "uts_present"):
                # entries of live nodes that are no longer in the dict were popped
                live_dst = alive[canon["slot_dst"]] > 0
                inlist = np.isin(canon["sub"][canon["slot_dst"]], list(subs_present))
                out[f][live_dst & inlist] = 0
            out[f][smap] = snap[f]
    return out


def snapshot_graphs(graphs):
    return nxio.graphs_to_host(graphs)


def make(name):
    kw = EVENTS[name]
    ev = synth.barrel_event(**kw)
    make_from_graphs(name, {"ev_" + k: v for k, v in ev.items()}, nxio.events_to_graphs(ev), name in COMPACT)


def make_from_graphs(name, extra, graphs, compact):
    """seed the given (unseeded) sub-graph list with the reference, run its schedule, write tests/golden/<name>.npz"""
    t0 = time.time()
    graphs = rh.seed_graphs(graphs)
    t_seed = time.time() - t0
    canon = nxio.graphs_to_host(graphs)
    full = fields.complete_host_batch({k: v for k, v in canon.items() if k not in ("truth", "orig_id", "in_key")})
    state = {f: full[f].copy() for f in MUTABLE}
    t0 = time.time()
    out = rh.reference_schedule(graphs)
    t_sched = time.time() - t0
    data = dict(extra)
    for k in TOPO:
        data["topo_" + k] = canon[k]
    data["meta_times"] = np.array([t_seed, t_sched])

    def put(stage, st, extra=None):
        keep = MUTABLE if not compact else ["alive", "sub_state", "active", "has_merged", "m_a", "m_b", "m_c",
                                            "m_p00", "m_p01", "m_p11", "m_p22", "uts_present", "tse_present",
                                            "degree", "has_uts", "m_prior"]
        for f in keep:
            data["%s/%s" % (stage, f)] = st[f]
        for k, v in (extra or {}).items():
            data["%s/%s" % (stage, k)] = v

    put("seed", state)
    orig_sub = {int(o): int(s) for o, s in zip(canon["orig_id"], canon["sub"])}
    node_of = {int(o): i for i, o in enumerate(canon["orig_id"])}
    prev_cands = 0
    for stage in ("c1", "x1", "e2", "x2", "m2", "c3", "x3"):
        if stage.startswith("x"):
            cands, rem, frag, pv = out[stage]
            new = cands[:len(cands) - prev_cands]
            prev_cands = len(cands)
            acc = np.zeros(len(canon["x"]), np.uint8)
            lab = np.full(len(canon["x"]), -1, np.int32)
            for c in new:
                idx = sorted(node_of[int(n)] for n in c.nodes())
                acc[idx] = 1
                lab[idx] = idx[0]
            fr = sorted(set(orig_sub[int(next(iter(g.nodes())))] for g in frag))
            snap = snapshot_graphs(rem) if rem else None
            if snap is not None:
                state = canonicalize(canon, snap, state, {"fragment": fr})
                state["alive"][acc > 0] = 0      # (a sub-graph that became a fragment keeps its rows; its accepted nodes are gone)
            else:
                state = {k: v.copy() for k, v in state.items()}
                state["alive"][acc > 0] = 0
            # extraction only removes nodes: keep pre-extraction values, just apply alive/sub_state
            put(stage, state, {"accepted": acc, "cand_label": lab,
                               "pvals": pv[["pvals_xy", "pvals_zr"]].to_numpy().astype(np.float64).reshape(-1, 2)})
        else:
            snap = snapshot_graphs(out[stage])
            state = canonicalize(canon, snap, state, {})
            put(stage, state)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **data)
    print("%s: N=%d E=%d seed %.1fs schedule %.1fs -> %s (%.1f KB)" % (
        name, len(canon["x"]), len(canon["in_src"]), t_seed, t_sched, path, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(EVENTS)):
        make(nm)
