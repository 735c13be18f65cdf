"""BASELINE configs[4]: scaling sweep -- edges per event 1e4..1e7, mixture components per node 1..8; fused iteration
throughput vs the HBM roofline.  Writes one JSON line per point.  Usage (on a B200):  python tools/sweep.py > sweep.jsonl"""
import json
import os
import sys

import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import gtf_b200  # noqa: E402
from gtf_b200 import synth  # noqa: E402

PEAK = 6452.8
if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))).get("hbm_gbs", PEAK))


def point(n_tracks, degree, n_events, steps=20):
    hbs = []
    for i in range(n_events):
        hb = synth.event_to_host(synth.barrel_event(n_tracks, seed=7000 + i, target_degree=float(degree)), i)
        hbs.append(hb)
    hb = synth.concat_host_batches(hbs)
    hb.pop("truth")
    hb.pop("orig_id")
    b = gtf_b200.EventBatch(hb)
    b.seed()
    b.cluster("track_state_estimates", 1.0, 2.0)
    n_active = bench.count_active(b)
    for _ in range(3):
        b.iterate_dry()
    b.set_timing(True)
    for _ in range(steps):
        b.iterate_dry()
    kt = b.timing_kernels()
    b.set_timing(False)
    kern = {k: kt[k] for k in ("k_send", "k_exec", "k_node2", "k_hv")}
    it_ms = sum(kern.values())
    # the iteration as a user runs it (CUDA graph replay): wall clock
    import time
    for _ in range(5):
        b.iterate_dry()
    b.sync()
    t0 = time.perf_counter()
    for _ in range(200):
        b.iterate_dry()
    b.sync()
    wall_ms = (time.perf_counter() - t0) / 200 * 1e3
    l0 = b.iteration_launches()
    b.iterate_dry()
    launches = b.iteration_launches() - l0
    deg = np.diff(hb["in_off"])
    out = {"tracks_per_event": n_tracks, "events": n_events, "target_degree": degree, "hits": b.N, "directed_edges": b.E,
           "mean_degree": float(deg.mean()), "active_edges": n_active, "kernels_ms": kern, "iteration_ms": it_ms,
           "iteration_ms_as_launched": wall_ms, "launches_per_iteration": launches,
           "roofline_frac_as_launched": bench.B_ALG * n_active / (wall_ms / 1e3) / 1e9 / PEAK,
           "active_edge_iterations_per_s": n_active / (it_ms / 1e3),
           "all_edges_per_s": b.E / (it_ms / 1e3),
           "roofline_frac": bench.B_ALG * n_active / (it_ms / 1e3) / 1e9 / PEAK}
    b.close()
    return out


if __name__ == "__main__":
    # edges per event sweep at degree 10 (single event: 1e4 .. 1e7 directed edges)
    for tracks in (100, 1000, 10000, 100000):
        print(json.dumps(point(tracks, 10, 1)), flush=True)
    # components-per-node sweep at ~1e6 edges in the batch
    for deg in (1, 2, 4, 8):
        ev = max(1, int(round(1e6 / (1000 * 10 * max(deg, 3.4)))))
        print(json.dumps(point(1000, deg, ev)), flush=True)
    # batch-size sweep (events per launch) at cfg2 shape
    for ev in (1, 8, 64, 256):
        print(json.dumps(point(1000, 10, ev)), flush=True)
