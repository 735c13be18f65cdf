"""bench.py -- throughput of the fused message-passing iteration on synthetic TrackML-shaped events.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torch.distributed.run)
  python bench.py --impl reference ...                     (CPU arm: oracle port on all host threads)

metric   directed edge-iterations/s of ONE fused iteration
         [extrapolate + chi2 gate + Kalman update, (prior, side-norm, reweight, prune) x2, pairwise chi2 +
          greedy KL clustering/merge, degree, mixture weights, priors]
         over a batch of independent cfg2-shaped events; an "edge-iteration" is one ACTIVE directed edge taken
         through the iteration.  events/s is reported beside it.
step     gtf_iterate_dry: the packed pipeline k_send -> k_exec -> k_node2 -> k_hv<G> on the batch; it reads the
         committed state, rewrites the dict entries in place (same values every pass) and sends the merged states
         to shadow buffers, so every step does identical work.
value    device-resident throughput, CUDA events on the batch stream, max over ranks.
e2e      the same iteration through the C-ABI from HOST buffers: pinned H2D of the iteration's mutable inputs,
         one committed gtf_iterate, D2H of the resulting state -- copies inside the timed region.
roofline algorithmic bytes (264 B per active edge-iteration) / the summed CUDA-event durations of all
         kernels of the iteration vs the measured HBM copy bandwidth in MEASURED_PEAKS.json.
Events shard across GPUs with no data-path collective (weak scaling); NCCL only carries the timing reduction
and the final candidate-table gather (gtf_b200.shard).
"""
import argparse
import ctypes
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
METRIC = "directed edge-iterations/s per fused message-passing iteration"

# per-iteration inputs that change between iterations (the graph, hit coordinates and the seed mixture weights are
# static and stay resident, like model weights)
E2E_UP = ("active", "has_merged", "m_a", "m_b", "m_c", "m_p00", "m_p01", "m_p11", "m_p22", "m_prior",
          "uts_present", "has_uts", "uts_next")
# result of one iteration as the reference's driver consumes it: pruning decisions (activation bitmap), the merged
# state every node will send next, node degrees.  The updated-state mixture stays device-resident between iterations.
E2E_DOWN = ("active", "has_merged", "m_a", "m_b", "m_c", "m_p00", "m_p01", "m_p11", "m_p22", "m_prior", "degree")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--events", type=int, default=128, help="events per GPU")
    ap.add_argument("--tracks", type=int, default=1000, help="tracks per event (1000 = cfg2: 10k hits / 100k directed edges)")
    ap.add_argument("--distinct", type=int, default=16, help="distinct generated events per GPU (tiled up to --events)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU work budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="sub-batches of the end-to-end pipeline (copy/compute overlap)")
    return ap.parse_args()


def workload_name(a):
    return "cfg4-shaped batch: %d cfg2 events per GPU (trackml_mod synthetic barrel, %d tracks -> %d hits, ~%d directed " \
           "edges per event, mean in-degree 10)" % (a.events, a.tracks, a.tracks * 10, a.tracks * 100)


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
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu):
        self.gpu, self.rows, self.stop_flag = gpu, [], False
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def start(self):
        self.t.start()

    def stop(self):
        self.stop_flag = True
        self.t.join(timeout=6)

        def num(v):
            try:
                return float(v)
            except ValueError:
                return None
        sm = [num(r[0]) for r in self.rows if num(r[0]) is not None]
        mx = [num(r[1]) for r in self.rows if len(r) > 1 and num(r[1]) is not None]
        reasons = set()
        for r in self.rows:
            for k, nm in enumerate(self.NAMES):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_iteration_rate(n_tracks, threads, budget_s, n_events=8):
    """the oracle port of ONE iteration (message_passing, (prior, reweight) x2, cluster on updated states) on the
    host; events are independent, so `threads` Python threads each run whole events (ctypes drops the GIL).
    Returns (active edge-iterations/s, events/s, seconds timed, events processed)."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import oracle_lib as ol
    import golden_util as gu
    from gtf_b200 import synth
    from concurrent.futures import ThreadPoolExecutor
    pristine, n_active = [], []
    for i in range(n_events):
        hb = synth.event_to_host(synth.barrel_event(n_tracks, seed=4000 + i), i)
        hb.pop("truth")
        hb.pop("orig_id")
        ob = ol.OracleBatch(hb)
        ob.seed()
        ob.cluster(0, 1.0, 2.0)
        pristine.append(ob.hb)
        n_active.append(int((ob.hb["active"][gu.edge_exists(ob.hb)] == 1).sum()))

    def work(k):
        ob = ol.OracleBatch(pristine[k % n_events])     # copies the post-iteration-1 state (untimed share is small)
        t0 = time.perf_counter()
        ob.extrapolate_stage(2.0)
        ob.cluster(1, 1000.0, 100.0)
        return time.perf_counter() - t0

    done, cpu_s, wall0 = 0, 0.0, time.perf_counter()
    with ThreadPoolExecutor(max(threads, 1)) as ex:
        while True:
            ts = list(ex.map(work, range(done, done + max(threads, 1))))
            done += len(ts)
            cpu_s += sum(ts)
            if cpu_s / max(threads, 1) >= budget_s or time.perf_counter() - wall0 > 6 * budget_s:
                break
    wall = time.perf_counter() - wall0
    # throughput = work / (CPU seconds / threads): the per-event state copy is excluded from the timed share
    eff = cpu_s / max(threads, 1)
    act = sum(n_active[k % n_events] for k in range(done))
    return act / eff, done / eff, eff, done, wall


def reference_arm(a):
    cores = os.cpu_count() or 1
    rate = None
    steps = max(1, min(a.steps, 5))
    for _ in range(min(a.warmup, 1) + steps):
        rate = cpu_iteration_rate(a.tracks, cores, max(2.0, a.cpu_seconds / steps))
    act, evs, eff, done, wall = rate
    emit({
        "impl": "reference", "metric": METRIC, "value": act, "unit": "edges/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": eff * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "events_per_s": evs,
        "config": {"workload": workload_name(a),
                   "note": "the reference is Python and does not travel to the GPU box: this arm times the C oracle port "
                           "(oracle/gtf_oracle.c) of the same iteration; the Python reference itself sustains ~3e3-1e4 "
                           "edges/s/stage (BASELINE.md §2)"},
        "cpu_baseline": {"value": act, "unit": "edges/s", "cores": cores, "kind": "port",
                         "sample": "%d event-iterations (8 distinct cfg2 events) on %d threads, %.1f s" % (done, cores, wall)},
        "e2e": {"value": act, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ---------------------------------------------------------------------------------------------- GPU arm
def count_active(b):
    hb = b.download(["active", "alive", "in_src", "slot_dst"])
    ex = (hb["alive"][np.maximum(hb["in_src"], 0)] > 0) & (hb["in_src"] >= 0) & (hb["alive"][hb["slot_dst"]] > 0)
    return int(((hb["active"] == 1) & ex).sum())


class E2EChunk(object):
    """one sub-batch of the end-to-end pipeline: its own EventBatch (own CUDA stream) + pinned host buffers"""

    def __init__(self, hb, device, torch):
        import gtf_b200
        from gtf_b200 import fields as F
        self.F = F
        self.b = gtf_b200.EventBatch(hb, device=device)
        self.b.seed()
        self.b.cluster("track_state_estimates", 1.0, 2.0)
        state = self.b.download(list(E2E_UP))
        self.up = {k: torch.from_numpy(v.copy()).pin_memory() for k, v in state.items()}
        b = self.b
        self.dn = {k: torch.from_numpy(np.empty(F.extent_len(F.FIELD_EXTENT[k], b.N, b.E, b.S), F.FIELD_DTYPE[k])).pin_memory()
                   for k in E2E_DOWN}
        self.h2d = sum(t.numel() * t.element_size() for t in self.up.values())
        self.d2h = sum(t.numel() * t.element_size() for t in self.dn.values())

    def upload(self):
        from gtf_b200 import lib as L
        for k, t in self.up.items():
            L.check(self.b.lib.gtf_batch_upload(self.b.h, self.F.FIELD_ID[k], ctypes.c_void_p(t.data_ptr())))

    def compute(self):
        self.b.iterate(max_iter=1, stop_when_converged=False, want_stats=False)   # asynchronous: no counter read-back

    def download(self):
        from gtf_b200 import lib as L
        for k, t in self.dn.items():
            L.check(self.b.lib.gtf_batch_download_async(self.b.h, self.F.FIELD_ID[k], ctypes.c_void_p(t.data_ptr())))


def e2e_loop(chunks, steps, torch):
    """per step, through the C-ABI from HOST buffers: pinned H2D of every chunk's per-iteration inputs, one committed
    fused iteration per chunk, D2H of the results.  Chunks own separate streams, so chunk i+1's upload and chunk i-1's
    download overlap chunk i's kernels (events are independent)."""
    def one():
        for c in chunks:
            c.upload()
        for c in chunks:
            c.compute()
            c.download()
        for c in chunks:
            c.b.sync()

    one()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    return ms, sum(c.h2d for c in chunks), sum(c.d2h for c in chunks)


_REAL_STDOUT = None


def emit(obj):
    """the ONE JSON line, on the process's original stdout"""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, line)
    else:
        sys.stdout.write(line.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    # libraries (NCCL version banner, ...) print to fd 1: keep stdout clean for the single JSON line
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        if rank == 0:
            reference_arm(a)
        return

    import torch
    import gtf_b200
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    warm = max(a.warmup, 3)
    hb = build_batch(a.events, a.tracks, 3000 + 100000 * rank, a.distinct)
    b = gtf_b200.EventBatch(hb, device=local)
    b.seed()                                               # event_conversion.py:87-96 (untimed set-up)
    b.cluster("track_state_estimates", 1.0, 2.0)           # iteration 1 of run_gnn_trackml_mod.sh (untimed set-up)
    stats = b.iterate_dry(want_stats=True)
    n_active = count_active(b)
    stream = torch.cuda.ExternalStream(b.stream(), device=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        b.iterate_dry()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(a.steps):
            b.iterate_dry()
        e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    if len(sampler.rows) < 3:
        # the timed region is shorter than nvidia-smi's latency: keep the SAME step running (untimed) until the
        # sampler has seen the load
        t_end = time.perf_counter() + 1.5
        while time.perf_counter() < t_end and len(sampler.rows) < 4:
            for _ in range(10):
                b.iterate_dry()
            b.sync()
    clocks = sampler.stop()
    # per-kernel durations (CUDA events recorded by the library on its stream around each kernel)
    b.set_timing(True)
    for _ in range(a.steps):
        b.iterate_dry()
    kt = b.timing_kernels()
    b.set_timing(False)
    kern_ms = {k: kt[k] for k in ("k_send", "k_exec", "k_node2", "k_hv")}
    iter_ms = sum(kern_ms.values())
    e2e_ms, h2d, d2h = (None, 0, 0)
    if not a.no_e2e:
        nch = max(1, min(a.e2e_chunks, a.events))
        per = [a.events // nch + (1 if k < a.events % nch else 0) for k in range(nch)]
        chunks = [E2EChunk(build_batch(n, a.tracks, 3000 + 100000 * rank + 1000 * k, a.distinct), local, torch)
                  for k, n in enumerate(per)]
        barrier()
        e2e_ms, h2d, d2h = e2e_loop(chunks, a.steps, torch)
        for c in chunks:
            c.b.close()
    red = torch.tensor([ms, e2e_ms or 0.0], device="cuda", dtype=torch.float64)
    tot = torch.tensor([float(n_active), float(b.E), float(a.events)], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms, e2e_ms = [float(v) for v in red.tolist()]
    n_act_all, n_tot_all, n_ev_all = [float(v) for v in tot.tolist()]
    if rank == 0:
        pk_path = os.path.join(REPO, "MEASURED_PEAKS.json")
        peaks = json.load(open(pk_path)) if os.path.exists(pk_path) else {}
        peak = float(peaks.get("hbm_gbs", 6650.0))
        step_s = ms / a.steps / 1e3
        # the 264 algorithmic bytes cover the WHOLE iteration (extrapolate + update + reweight x2 + cluster), so they are
        # charged against the sum of all its kernels (CUDA events recorded by the library on its stream around each)
        ach = B_ALG * n_active / (iter_ms / 1e3) / 1e9
        traffic, kdram, tr_path = None, None, os.path.join(REPO, "profiles", "r01_pipeline_traffic.json")
        if os.path.exists(tr_path):   # dram__bytes_read+write of every pipeline kernel from the committed `ncu --set full` capture
            tj = json.load(open(tr_path))
            traffic = tj["dram_bytes_per_active_edge"] * n_active
            # measured DRAM bytes of each kernel (scaled to this launch's active edges) / its live CUDA-event time / peak
            scale = n_active / float(tj["active_edges"])
            grp = {"k_send": ("k_begin", "k_send"), "k_exec": ("k_exec",), "k_node2": ("k_node2",),
                   "k_hv": ("k_hv<8>", "k_hv<16>", "k_hv<4>", "k_hv<32>", "k_big")}
            kdram = {k: sum(tj["kernels"].get(n, {}).get("dram_bytes", 0.0) for n in names) * scale / (kern_ms[k] / 1e3) / 1e9 / peak
                     for k, names in grp.items() if kern_ms[k] > 0}
        out = {
            "metric": METRIC, "value": n_act_all / step_s, "unit": "edges/s", "n_gpus": world, "steps": a.steps,
            "warmup": warm, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "events_per_s": n_ev_all / step_s, "all_edges_per_s": n_tot_all / step_s,
            "config": {"workload": workload_name(a),
                       "per_gpu": {"hits": b.N, "directed_edges": b.E, "active_edges": n_active, "events": a.events,
                                   "distinct_events": min(a.distinct, a.events), "device_bytes": b.device_bytes()},
                       "l2": "per-step working set %.0f MB per GPU is larger than the 126 MB L2 (no flush needed)" % (
                           (B_ALG * n_active + 11.0 * b.E) / 1e6),
                       "step": "gtf_iterate_dry = k_send (message list + scattering prefix) + k_exec (extrapolate, gate, Kalman "
                               "update) + k_node2 (<= 2-component nodes: priors, reweight x2, prune) + k_hv<4|8|16|32> / k_big "
                               "(>= 3-component nodes: the same + pairwise chi2 + greedy KL merge); reads the committed state, "
                               "rewrites the dict entries in place, merged states to shadow buffers"},
            "gpu_launches": 9 * a.steps,   # k_begin, k_send, k_exec, k_node2, k_hv<4|8|16|32>, k_big (replayed from one CUDA graph)
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                         "traffic_source": "profiles/r01_pipeline_traffic.json (ncu --set full capture of all pipeline kernels at this "
                                           "workload) x active edges of this launch",
                         "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650",
                         "kernel": "whole iteration: k_send + k_exec + k_node2 + k_hv<4,8,16,32> (+ k_big)", "kernel_ms": iter_ms,
                         "kernels_ms": kern_ms, "dominant": max(kern_ms, key=kern_ms.get), "kernels_dram_frac": kdram,
                         "alg_bytes_per_launch": B_ALG * n_active},
            "iteration_stats": stats,
        }
        if e2e_ms:
            out["e2e"] = {"value": n_act_all / (e2e_ms / 1e3 / a.steps), "unit": "edges/s", "h2d_bytes_per_step": h2d,
                          "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / a.steps,
                          "chunks": min(a.e2e_chunks, a.events)}
        if not a.no_cpu and world == 1:
            act, evs, eff, done, wall = cpu_iteration_rate(a.tracks, 1, a.cpu_seconds)
            out["cpu_baseline"] = {"value": act, "unit": "edges/s", "cores": 1, "kind": "port", "events_per_s": evs,
                                   "sample": "%d event-iterations of cfg2 events (8 distinct), single-threaded C oracle, "
                                             "%.1f s of CPU work" % (done, eff)}
        emit(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
