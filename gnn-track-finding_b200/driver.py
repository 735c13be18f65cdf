"""In-process drivers over an EventBatch: the reference's fixed schedule (run_gnn_trackml_mod.sh:71-148) and
the iterate-until-converged variant."""
import numpy as np

from . import synth
from .batch import EventBatch

DEFAULTS = dict(chi2_c1=1.0, kl_c1=2.0, chi2_cut=2.0, chi2_c3=1000.0, kl_c3=100.0, pval=0.01, numhits=4, sep3d=10.0,
                merge_dist=8.0)


def reference_schedule(b, P=DEFAULTS, seed=True):
    """iteration 1: cluster(seeds) + extract; iteration 2: extrapolate stage + extract + metadata update;
    iteration 3: cluster(updated states) + extract.  Returns per-extraction (n_accepted, accepted mask)."""
    if seed:
        b.seed()
    out = []
    b.cluster("track_state_estimates", P["chi2_c1"], P["kl_c1"])
    out.append(b.extract(P["pval"], P["numhits"], P["sep3d"], P["merge_dist"])[:2])
    b.extrapolate_stage(P["chi2_cut"])
    out.append(b.extract(P["pval"], P["numhits"], P["sep3d"], P["merge_dist"])[:2])
    b.remove_state_metadata()
    b.cluster("updated_track_states", P["chi2_c3"], P["kl_c3"])
    out.append(b.extract(P["pval"], P["numhits"], P["sep3d"], P["merge_dist"])[:2])
    return out


def converged_schedule(b, P=DEFAULTS, max_iter=10, seed=True):
    """seed, cluster(seeds), fused iterations until the active-edge bitmap stops changing, extract."""
    if seed:
        b.seed_cluster(P["chi2_c1"], P["kl_c1"])
    else:
        b.cluster("track_state_estimates", P["chi2_c1"], P["kl_c1"])
    stats = b.iterate(max_iter=max_iter, stop_when_converged=True, chi2_cut=P["chi2_cut"], cluster_chi2=P["chi2_c3"],
                      cluster_kl=P["kl_c3"])
    n, acc, pxy, pzr = b.extract(P["pval"], P["numhits"], P["sep3d"], P["merge_dist"])
    return stats, n, acc


def run_events(events, device=0, schedule="converged", gather=True):
    """events: list of synthetic event dicts (synth.py) owned by THIS rank (already sharded, shard.py).
    Returns the candidate table (event_id, candidate_id, node index) -- gathered on rank 0 if a process group
    is initialised."""
    from . import shard
    hbs = [synth.event_to_host(ev, eid) for eid, ev in events]
    hb = synth.concat_host_batches(hbs)
    hb.pop("truth")
    hb.pop("orig_id")
    b = EventBatch(hb, device=device)
    try:
        if schedule == "converged":
            converged_schedule(b)
        else:
            reference_schedule(b)
        rows = b.candidates()
    finally:
        b.close()
    return shard.gather_candidates(rows) if gather else rows
