"""CPU: host-side multi-GPU logic (event partitioning, candidate-table gather) over gloo, world_size 2."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gtf_b200
from gtf_b200 import shard


def test_partition_balanced_and_complete():
    rng = np.random.default_rng(0)
    counts = rng.integers(50_000, 150_000, 37)
    parts = shard.partition_events(counts, 4)
    assert sorted(sum(parts, [])) == list(range(37))
    loads = [counts[p].sum() for p in parts]
    assert max(loads) - min(loads) <= counts.max()
    assert shard.partition_events([5, 1], 4)[2:] == [[], []]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    n = [7, 0][rank] if world == 2 else 3          # rank 1 contributes an empty table
    rows = np.stack([np.full(n, rank), rng.integers(0, 50, n), rng.integers(0, 1000, n)], 1).astype(np.int32)
    out = shard.gather_candidates(rows, device="cpu")
    if rank == 0:
        q.put((rows, out))
    else:
        assert out is None
        q.put((rows, None))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_candidates_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    allrows = np.concatenate([g[0] for g in got])
    want = allrows[np.lexsort((allrows[:, 2], allrows[:, 1], allrows[:, 0]))]
    out = [g[1] for g in got if g[1] is not None][0]
    assert np.array_equal(out, want)


def test_gather_is_identity_without_group():
    rows = np.array([[1, 2, 3], [0, 5, 1]], np.int32)
    out = shard.gather_candidates(rows)
    assert np.array_equal(out, rows[[1, 0]])
