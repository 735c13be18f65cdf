"""Multi-GPU: events are independent, so they shard across ranks with NO data-path
collective; the only exchange is the final gather of the candidate tables (event_id, candidate_id, node).

One process per GPU (torchrun); `torch.distributed` with NCCL on the B200 box, gloo in the CPU tests."""
import numpy as np


def partition_events(edge_counts, world_size):
    """Longest-processing-time assignment of events to ranks, balanced by directed-edge count.
    Returns a list (per rank) of event indices, each in ascending order."""
    order = np.argsort(-np.asarray(edge_counts, dtype=np.int64), kind="stable")
    load = np.zeros(world_size, np.int64)
    out = [[] for _ in range(world_size)]
    for e in order:
        r = int(np.argmin(load))
        out[r].append(int(e))
        load[r] += int(edge_counts[e])
    return [sorted(x) for x in out]


def gather_candidates(rows, device=None, dst=0):
    """Variable-length gather of (k, 3) int32 candidate tables to rank `dst` (all_gather of the row counts,
    then all_gather of the padded tables: ~12 B per hit, negligible on NVLink).  Returns the concatenated,
    lexicographically sorted table on `dst`, None elsewhere.  Without an initialised process group it is
    the identity."""
    import torch
    import torch.distributed as dist
    rows = np.ascontiguousarray(rows, np.int32).reshape(-1, 3)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rows[np.lexsort((rows[:, 2], rows[:, 1], rows[:, 0]))]
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    pad = torch.zeros((cap, 3), dtype=torch.int32, device=dev)
    if rows.shape[0]:
        pad[:rows.shape[0]] = torch.from_numpy(rows).to(dev)
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    if rank != dst:
        return None
    allrows = np.concatenate([b[:c].cpu().numpy() for b, c in zip(bufs, counts)], axis=0)
    return allrows[np.lexsort((allrows[:, 2], allrows[:, 1], allrows[:, 0]))]
