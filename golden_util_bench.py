"""bench.py helpers (product side, no oracle): active-edge count, per-kernel timing, e2e loop."""
import ctypes
import numpy as np


def count_active(b):
    hb = b.download(["active", "alive", "in_src", "slot_dst"])
    ex = (hb["alive"][np.maximum(hb["in_src"], 0)] > 0) & (hb["in_src"] >= 0) & (hb["alive"][hb["slot_dst"]] > 0)
    return int(((hb["active"] == 1) & ex).sum())


def kernel_times(b, steps, stream, torch):
    """average duration of the two kernels of one iteration, CUDA events on the batch stream"""
    from gtf_b200 import lib as L
    lib = b.lib
    p = b._iter_params(2.0, 1000.0, 100.0, 0.1, None)
    # time whole iterations, then prefix-only via the message-passing-free path: measure k_prefix separately
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for i in range(steps):
            L.check(lib.gtf_iterate_dry(b.h, ctypes.byref(p), ctypes.byref(b.geom), None))
            ev[i + 1].record(stream)
    torch.cuda.synchronize()
    it = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    return {"tile_ms": float(np.mean(it)), "prefix_ms": None}


E2E_UP = ("active", "has_merged", "m_a", "m_b", "m_c", "m_p00", "m_p01", "m_p11", "m_p22", "m_prior", "tse_w",
          "uts_present", "has_uts", "uts_next")
E2E_DOWN = ("active", "has_merged", "m_a", "m_b", "m_c", "m_p00", "m_p01", "m_p11", "m_p22", "m_prior",
            "uts_present", "uts_a", "uts_b", "uts_c", "uts_tau", "uts_p00", "uts_p01", "uts_p11", "uts_p22", "uts_w",
            "uts_lik", "degree")


def e2e_rate(b, hb, steps, stream, torch):
    """per step: H2D of the iteration's mutable inputs from pinned host memory, one committed fused
    iteration through the C-ABI, D2H of the resulting state"""
    import time
    from gtf_b200 import fields as F
    state = b.download(list(E2E_UP))
    pinned_up = {k: torch.from_numpy(v.copy()).pin_memory() for k, v in state.items()}
    pinned_dn = {k: torch.empty(F.extent_len(F.FIELD_EXTENT[k], b.N, b.E, b.S),
                                dtype=torch.from_numpy(np.zeros(1, F.FIELD_DTYPE[k])).dtype).pin_memory() for k in E2E_DOWN}
    h2d = sum(t.numel() * t.element_size() for t in pinned_up.values())
    d2h = sum(t.numel() * t.element_size() for t in pinned_dn.values())
    lib = b.lib
    from gtf_b200 import lib as L

    def one():
        for k, t in pinned_up.items():
            L.check(lib.gtf_batch_upload(b.h, F.FIELD_ID[k], ctypes.c_void_p(t.data_ptr())))
        b.iterate(max_iter=1, stop_when_converged=False)
        for k, t in pinned_dn.items():
            L.check(lib.gtf_batch_download(b.h, F.FIELD_ID[k], ctypes.c_void_p(t.data_ptr())))

    one()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    # restore the pristine post-iteration-1 state
    for k, t in pinned_up.items():
        L.check(lib.gtf_batch_upload(b.h, F.FIELD_ID[k], ctypes.c_void_p(t.data_ptr())))
    b.sync()
    return {"ms": ms, "h2d": h2d, "d2h": d2h}
