"""torchrun check of the multi-GPU path: events sharded by edge count, one EventBatch per rank, NCCL
variable-length gather of the candidate tables to rank 0, compared there with a single-process run of all events.
Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_check.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import torch.distributed as dist
from gtf_b200 import synth, shard, driver

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_events = 12
events = [(e, synth.barrel_event(150 + 20 * (e % 4), seed=8000 + e)) for e in range(n_events)]
counts = [2 * len(ev["edge_a"]) for _, ev in events]
parts = shard.partition_events(counts, world)
mine = [events[e] for e in parts[rank]]
rows = driver.run_events(mine, device=local, schedule="converged", gather=True)
if rank == 0:
    ref = driver.run_events(events, device=local, schedule="converged", gather=False)
    # candidate ids / node indices are batch-local: compare per event the multiset of candidate sizes
    def sig(t):
        out = {}
        for ev in np.unique(t[:, 0]):
            r = t[t[:, 0] == ev]
            _, c = np.unique(r[:, 1], return_counts=True)
            out[int(ev)] = sorted(c.tolist())
        return out
    a, b = sig(rows), sig(ref)
    assert a == b, "gathered candidate tables differ from the single-process run"
    print("multi-GPU check ok: world %d, %d rows from %d events, %d candidates" % (world, len(rows), len(a), sum(len(v) for v in a.values())))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
