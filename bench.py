"""bench.py -- fused message-passing iteration throughput on synthetic TrackML-shaped events.

metric: directed edge-iterations/s (and events/s) of ONE fused iteration
        [extrapolate + chi2 gate + Kalman update, (prior, reweight, prune) x2, cluster/merge, degree, weights, priors]
        over a batch of independent cfg2-shaped events (10 layers, ~10k hits, ~100k directed edges each),
        events sharded across GPUs (weak scaling, no data-path collective).
See DESIGN.md "Measurement" for the definitions of value / e2e / roofline / cpu_baseline.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

B_ALG = 264.0   # algorithmic bytes per active directed edge-iteration


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--events", type=int, default=64, help="events per GPU")
    ap.add_argument("--tracks", type=int, default=1000, help="tracks per event (1000 = cfg2: 10k hits / 100k edges)")
    ap.add_argument("--distinct", type=int, default=16, help="distinct generated events (tiled up to --events)")
    ap.add_argument("--cpu-events", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def build_batch(n_events, n_tracks, seed0, distinct):
    from gtf_b200 import synth
    base = [synth.event_to_host(synth.barrel_event(n_tracks, seed=seed0 + i), 0) for i in range(min(distinct, n_events))]
    hbs = []
    for i in range(n_events):
        hb = dict(base[i % len(base)])
        hb["sub_event"] = np.full_like(hb["sub_event"], i)
        hbs.append(hb)
    hb = synth.concat_host_batches(hbs)
    hb.pop("truth")
    hb.pop("orig_id")
    return hb


class ClockSampler(object):
    def __init__(self, gpu):
        self.gpu, self.rows, self.stop_flag = gpu, [], False
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.t.start()

    def stop(self):
        self.stop_flag = True
        self.t.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 4 + k and r[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def cpu_oracle_rate(n_tracks, n_events, threads):
    """oracle port of one iteration on host cores; events on Python threads (ctypes releases the GIL)."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import oracle_lib as ol
    from gtf_b200 import synth
    obs = []
    for i in range(n_events):
        hb = synth.event_to_host(synth.barrel_event(n_tracks, seed=4000 + i), i)
        hb.pop("truth")
        hb.pop("orig_id")
        ob = ol.OracleBatch(hb)
        ob.seed()
        ob.cluster(0, 1.0, 2.0)
        obs.append(ob)
    import golden_util as gu
    active = sum(int((ob.hb["active"][gu.edge_exists(ob.hb)] == 1).sum()) for ob in obs)
    total = sum(ob.E for ob in obs)

    def work(ob):
        ob.extrapolate_stage(2.0)
        ob.cluster(1, 1000.0, 100.0)

    t0 = time.perf_counter()
    if threads > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(work, obs))
    else:
        for ob in obs:
            work(ob)
    dt = time.perf_counter() - t0
    return active / dt, total / dt, n_events / dt, dt, active


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    if a.impl == "reference":
        if rank != 0:
            return
        # reference arm: the oracle port of the path (the Python reference cannot travel to the GPU box)
        rates = []
        for _ in range(a.warmup + a.steps if a.steps <= 3 else 1 + min(a.steps, 3)):
            rates.append(cpu_oracle_rate(a.tracks, max(cores, a.cpu_events), cores))
        act, tot, evs, dt, n_act = rates[-1]
        print(json.dumps({
            "impl": "reference", "metric": "directed edge-iterations/s per fused message-passing iteration",
            "value": act, "unit": "edges/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "events_per_s": evs, "edges_total_per_s": tot,
            "config": {"workload": "cfg4-shaped batch of cfg2 events (%d tracks, ~%d directed edges each)" % (a.tracks, a.tracks * 100)},
            "cpu_baseline": {"value": act, "unit": "edges/s", "cores": cores, "kind": "port",
                             "sample": "%d events, one iteration each, C oracle on %d host threads" % (max(cores, a.cpu_events), cores)},
            "e2e": {"value": act, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    import torch
    import gtf_b200
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    hb = build_batch(a.events, a.tracks, 3000 + 100000 * rank, a.distinct)
    b = gtf_b200.EventBatch(hb, device=local)
    b.seed()
    b.cluster("track_state_estimates", 1.0, 2.0)
    st0 = b.iterate_dry(want_stats=True)
    import golden_util_bench as gub
    n_active = gub.count_active(b)
    stream = torch.cuda.ExternalStream(b.stream(), device=torch.device("cuda", local))
    peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else {}
    peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured" if "hbm_gbs" in peaks else "fallback"

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        b.iterate_dry()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    with torch.cuda.stream(stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.steps):
            b.iterate_dry()
        e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # per-kernel timing of the dominant kernel (tile kernel) with events around each launch pair
    # e2e: host buffers -> device -> iterate -> results back
    e2e = gub.e2e_rate(b, hb, a.steps, stream, torch) if True else None
    tot_active = torch.tensor([float(n_active), float(b.E), float(a.events)], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tot_active)
        e2 = torch.tensor([e2e["ms"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(e2, op=dist.ReduceOp.MAX)
        e2e["ms"] = float(e2.item())
    n_act_all, n_tot_all, n_ev_all = [float(v) for v in tot_active.tolist()]
    per_step_s = ms / a.steps / 1e3
    if rank == 0:
        value = n_act_all / per_step_s
        tile_ms = gub.kernel_times(b, a.steps, stream, torch)
        ach = B_ALG * n_active / (tile_ms["tile_ms"] / 1e3) / 1e9
        out = {
            "metric": "directed edge-iterations/s per fused message-passing iteration", "value": value, "unit": "edges/s",
            "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "events_per_s": n_ev_all / per_step_s, "edges_total_per_s": n_tot_all / per_step_s,
            "config": {"workload": "cfg4-shaped batch: %d cfg2 events per GPU (%d tracks, %d hits, %d directed edges, %d active after iteration 1); "
                                   "%d distinct events tiled" % (a.events, a.tracks, b.N, b.E, n_active, min(a.distinct, a.events)),
                       "l2": "inputs (%.1f MB touched per step) larger than L2" % (b.device_bytes() / 1e6),
                       "step": "gtf_iterate_dry: k_prefix + k_tile (fused E+R+R+C), idempotent"},
            "gpu_launches": 2 * a.steps,
            "clocks": clocks,
            "e2e": {"value": n_act_all / (e2e["ms"] / 1e3 / a.steps), "unit": "edges/s", "h2d_bytes_per_step": e2e["h2d"],
                    "d2h_bytes_per_step": e2e["d2h"]},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": "k_tile", "kernel_ms": tile_ms["tile_ms"],
                         "prefix_ms": tile_ms["prefix_ms"], "alg_bytes_per_launch": B_ALG * n_active},
            "stats": st0,
        }
        if not a.no_cpu:
            act, tot, evs, dt, n_a = cpu_oracle_rate(a.tracks, a.cpu_events, 1)
            out["cpu_baseline"] = {"value": act, "unit": "edges/s", "cores": 1, "kind": "port",
                                   "sample": "%d events, one iteration each, single-threaded C oracle (%.2f s)" % (a.cpu_events, dt)}
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
