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


_PINNED = {}


def _pinned_rows(n):
    """(n, 3) int32 view of a pinned host buffer kept between calls (a pageable `.cpu()` of a 35 MB table costs milliseconds)"""
    import torch
    buf = _PINNED.get("rows")
    if buf is None or buf.shape[0] < n:
        buf = torch.empty((max(n, 1), 3), dtype=torch.int32).pin_memory()
        _PINNED["rows"] = buf
    return buf[:n]


def gather_candidates(rows, device=None, dst=0, sort=True, info=None):
    """Variable-length gather of (k, 3) int32 candidate tables to rank `dst`: all_gather of the row counts, then a GATHER
    of the tables padded to the largest count (~12 B per hit; only `dst` receives them), one copy into a pinned host buffer
    on `dst`.  `rows`: numpy array, or a torch tensor already on the device (EventBatch.candidates_device(): no host round
    trip before NCCL).  Returns the concatenated table on `dst` (lexicographically sorted unless sort=False), None
    elsewhere.  `info` (dict) receives the per-rank row counts and the bytes every rank contributed.  Without an
    initialised process group it is the identity."""
    import torch
    import torch.distributed as dist
    is_t = hasattr(rows, "is_cuda")
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = rows.detach().cpu().numpy() if is_t else np.ascontiguousarray(rows, np.int32)
        out = out.reshape(-1, 3)
        if info is not None:
            info.update(counts=[out.shape[0]], bytes=0)
        return out[np.lexsort((out[:, 2], out[:, 1], out[:, 0]))] if sort else out
    world, rank = dist.get_world_size(), dist.get_rank()
    nccl = dist.get_backend() == "nccl"
    dev = device if device is not None else ("cuda" if nccl else "cpu")
    if is_t:
        t = rows.reshape(-1, 3).to(dev)
    else:
        t = torch.from_numpy(np.ascontiguousarray(rows, np.int32).reshape(-1, 3)).to(dev)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=dev)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, n) if nccl else dist.all_gather(list(counts.unbind(0)), n[0])
    counts = [int(c) for c in counts.tolist()]
    cap = max(max(counts), 1)
    pad = torch.zeros((cap, 3), dtype=torch.int32, device=dev)
    if t.shape[0]:
        pad[:t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if info is not None:
        info.update(counts=counts, bytes=int(cap) * 12)
    if rank != dst:
        return None
    total = sum(counts)
    if nccl:
        host = _pinned_rows(total)
        o = 0
        for b, c in zip(bufs, counts):
            if c:
                host[o:o + c].copy_(b[:c], non_blocking=True)
            o += c
        torch.cuda.current_stream().synchronize()
        allrows = host.numpy()
    else:
        allrows = np.concatenate([b[:c].numpy() for b, c in zip(bufs, counts)], axis=0)
    return allrows[np.lexsort((allrows[:, 2], allrows[:, 1], allrows[:, 0]))] if sort else allrows
