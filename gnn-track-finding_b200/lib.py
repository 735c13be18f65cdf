"""ctypes binding of the CUDA C-ABI library (csrc/libgtf_b200.so, include/gtf.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is visible, loading /
compute calls raise."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO = os.environ.get("GTF_LIB") or os.path.join(CSRC, "libgtf_b200.so")   # GTF_LIB: experiment builds
SOURCES = ["gtf_b200.cu", "gtf_tile.cuh", "gtf_iter.cuh", "gtf_dev.cuh", "gtf_math.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "550"]


class Geom(ctypes.Structure):
    _fields_ = [("sigma0xy", ctypes.c_double), ("sigma0rz", ctypes.c_double),
                ("sigma0rz2", ctypes.c_double), ("endcap_boundary", ctypes.c_double)]


class Stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in ("nodes_merged", "edges_deactivated", "edges_sent", "edges_gated",
                                               "edges_reweight_off", "active_edges", "active_changed", "ref_errors",
                                               "near_threshold")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class NearRec(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("index", ctypes.c_int32), ("value", ctypes.c_double), ("threshold", ctypes.c_double)]


NEAR_KINDS = ("gate chi2 <= chi2CutFactor", "reweight < threshold", "cluster chi2 < chi2_threshold", "cluster KL < KL_threshold")


class IterParams(ctypes.Structure):
    _fields_ = [("chi2_cut", ctypes.c_double), ("cluster_chi2", ctypes.c_double), ("cluster_kl", ctypes.c_double),
                ("reweight_threshold", ctypes.c_double), ("kl_lut", ctypes.POINTER(ctypes.c_double)),
                ("record_chi2", ctypes.c_int32)]


class Events(ctypes.Structure):
    """gtf_events (include/gtf.h): host pointers to the arrays that define a batch of freshly converted events"""
    _I, _D = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double)
    _fields_ = [("n_nodes", ctypes.c_int32), ("n_slots", ctypes.c_int32), ("n_subgraphs", ctypes.c_int32),
                ("x", _D), ("y", _D), ("z", _D), ("r", _D), ("layer", _I), ("volume", _I), ("sub", _I), ("sub_off", _I),
                ("sub_event", _I), ("in_off", _I), ("in_src", _I), ("out_off", _I), ("out_slot", _I)]


EVENT_ARRAYS = (("x", "f8"), ("y", "f8"), ("z", "f8"), ("r", "f8"), ("layer", "i4"), ("volume", "i4"), ("sub", "i4"),
                ("sub_off", "i4"), ("sub_event", "i4"), ("in_off", "i4"), ("in_src", "i4"), ("out_off", "i4"), ("out_slot", "i4"))


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(_HERE, "..", "include", h) for h in ("gtf.h", "gtf_fields.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> csrc/libgtf_b200.so (in-tree, travels with gpurun)."""
    if not force and not needs_build():
        return SO
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO, os.path.join(CSRC, "gtf_b200.cu")]
    subprocess.check_call(cmd, cwd=CSRC)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        raise RuntimeError("libgtf_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                           "there is no CPU fallback")
    L = ctypes.CDLL(SO)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    pg, ps = ctypes.POINTER(Geom), ctypes.POINTER(Stats)
    dp = ctypes.POINTER(ctypes.c_double)
    sig = {
        "gtf_abi_version": (ctypes.c_int, []),
        "gtf_last_error": (ctypes.c_char_p, []),
        "gtf_device_count": (ctypes.c_int, []),
        "gtf_batch_create": (ctypes.c_int, [i32, i32, i32, ctypes.c_int, ctypes.POINTER(vp)]),
        "gtf_batch_destroy": (ctypes.c_int, [vp]),
        "gtf_field_count": (ctypes.c_int, []),
        "gtf_field_name": (ctypes.c_char_p, [ctypes.c_int]),
        "gtf_field_id": (ctypes.c_int, [ctypes.c_char_p]),
        "gtf_field_bytes": (i64, [vp, ctypes.c_int]),
        "gtf_batch_upload": (ctypes.c_int, [vp, ctypes.c_int, vp]),
        "gtf_batch_download": (ctypes.c_int, [vp, ctypes.c_int, vp]),
        "gtf_batch_download_async": (ctypes.c_int, [vp, ctypes.c_int, vp]),
        "gtf_batch_device_ptr": (ctypes.c_int, [vp, ctypes.c_int, ctypes.POINTER(vp)]),
        "gtf_batch_finalize": (ctypes.c_int, [vp]),
        "gtf_batch_load_events": (ctypes.c_int, [vp, ctypes.POINTER(Events)]),
        "gtf_candidates_device": (ctypes.c_int, [vp, ctypes.POINTER(vp), ctypes.POINTER(i64)]),
        "gtf_batch_sync": (ctypes.c_int, [vp]),
        "gtf_batch_near_threshold": (ctypes.c_int, [vp, ctypes.POINTER(NearRec), ctypes.c_int, ctypes.POINTER(i64)]),
        "gtf_batch_stream": (ctypes.c_int, [vp, ctypes.POINTER(vp)]),
        "gtf_batch_device_bytes": (i64, [vp]),
        "gtf_batch_iteration_launches": (i64, [vp]),
        "gtf_seed": (ctypes.c_int, [vp, pg]),
        "gtf_seed_all": (ctypes.c_int, [vp, pg, ps]),
        "gtf_seed_cluster": (ctypes.c_int, [vp, pg, dbl, dbl, dp, ps]),
        "gtf_initialize_edge_activation": (ctypes.c_int, [vp]),
        "gtf_compute_prior_probabilities": (ctypes.c_int, [vp, ctypes.c_int]),
        "gtf_compute_mixture_weights": (ctypes.c_int, [vp, ctypes.c_int, ps]),
        "gtf_query_node_degree": (ctypes.c_int, [vp]),
        "gtf_cluster": (ctypes.c_int, [vp, ctypes.c_int, dbl, dbl, dp, pg, ps]),
        "gtf_message_passing": (ctypes.c_int, [vp, dbl, pg, ps]),
        "gtf_reweight": (ctypes.c_int, [vp, ctypes.c_int, dbl, ps]),
        "gtf_extrapolate_stage": (ctypes.c_int, [vp, dbl, pg, ps]),
        "gtf_remove_state_metadata": (ctypes.c_int, [vp, ps]),
        "gtf_iterate": (ctypes.c_int, [vp, ctypes.POINTER(IterParams), pg, ctypes.c_int, ctypes.c_int, ps,
                                       ctypes.POINTER(ctypes.c_int)]),
        "gtf_iterate_dry": (ctypes.c_int, [vp, ctypes.POINTER(IterParams), pg, ps]),
        "gtf_batch_set_timing": (ctypes.c_int, [vp, ctypes.c_int]),
        "gtf_batch_timing": (ctypes.c_int, [vp, dp, dp, dp, ctypes.POINTER(ctypes.c_int)]),
        "gtf_batch_timing_kernels": (ctypes.c_int, [vp, dp, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
        "gtf_kl_pairs": (ctypes.c_int, [ctypes.c_int, dp, dp, ctypes.POINTER(i32), i32, dp, i64, ctypes.POINTER(i64)]),
        "gtf_pairwise_chi2": (ctypes.c_int, [ctypes.c_int, i32, dp, dp, dp, dp, dbl, dbl, dbl, dp]),
        "gtf_merge_states": (ctypes.c_int, [ctypes.c_int, dp, dp, dp, dp, dp, dp]),
        "gtf_seed_parabolic_pairs": (ctypes.c_int, [ctypes.c_int, i64, dp, dp, dbl, dbl, dbl, dp, dp]),
        "gtf_components": (ctypes.c_int, [vp]),
        "gtf_extract": (ctypes.c_int, [vp, pg, dbl, ctypes.c_int, dbl, dbl, ctypes.POINTER(i32),
                                       ctypes.POINTER(ctypes.c_uint8), dp, dp]),
        "gtf_tag_propagate": (ctypes.c_int, [vp, dbl, ctypes.POINTER(i32), ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
        "gtf_candidates": (ctypes.c_int, [vp, ctypes.POINTER(i32), i64, ctypes.POINTER(i64)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._gtf_symbols = sorted(sig)
    _lib = L
    return L


class GtfError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise GtfError("libgtf_b200 error %d: %s" % (rc, lib().gtf_last_error().decode()))
