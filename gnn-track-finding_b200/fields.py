"""Field table of the flat event-batch layout, parsed from include/gtf_fields.h (single source of
truth shared with the CUDA C-ABI and the CPU oracle)."""
import ctypes
import os
import re
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
FIELDS_H = os.path.join(_HERE, "..", "include", "gtf_fields.h")

_CT = {"double": (ctypes.c_double, np.float64), "int32_t": (ctypes.c_int32, np.int32),
       "uint8_t": (ctypes.c_uint8, np.uint8), "int8_t": (ctypes.c_int8, np.int8)}


def _parse():
    txt = open(FIELDS_H).read()
    out = []
    for m in re.finditer(r"^\s*X\((\w+),\s*(\w+),\s*(\w+)\)", txt, re.M):
        name, ct, ext = m.groups()
        if name == "name":
            continue
        out.append((name, ct, ext))
    return out


FIELDS = _parse()                      # [(name, c_type, extent)]
FIELD_ID = {f[0]: i for i, f in enumerate(FIELDS)}
FIELD_DTYPE = {f[0]: _CT[f[1]][1] for f in FIELDS}
FIELD_EXTENT = {f[0]: f[2] for f in FIELDS}


def extent_len(ext, N, E, S):
    return {"N": N, "N1": N + 1, "E": E, "S": S, "S1": S + 1}[ext]


def complete_host_batch(hb):
    """Fill in every GTF_FIELDS array a partial host batch (nxio.graphs_to_host) lacks and coerce dtypes."""
    N, E, S = len(hb["x"]), len(hb["in_src"]), len(hb["sub_off"]) - 1
    out = dict(hb)
    out.setdefault("uts_present", np.zeros(E, np.uint8))
    out.setdefault("uts_rank", np.full(E, -1, np.int32))
    hb = out
    if "has_uts" not in out:
        has = np.zeros(N, np.uint8)
        if E:
            np.maximum.at(has, hb["slot_dst"][hb["uts_present"] > 0], 1)
        out["has_uts"] = has
    if "uts_next" not in out:
        nxt = np.zeros(N, np.int32)
        if E:
            np.maximum.at(nxt, hb["slot_dst"], (hb["uts_rank"] + 1).astype(np.int32))
        out["uts_next"] = nxt
    for name, ct, ext in FIELDS:
        n = extent_len(ext, N, E, S)
        dt = FIELD_DTYPE[name]
        if name not in out:
            out[name] = np.full(n, np.nan) if dt == np.float64 else np.zeros(n, dt)
            if name == "label":
                out[name] = np.full(n, -1, np.int32)
        arr = np.ascontiguousarray(out[name], dtype=dt)
        assert arr.shape == (n,), (name, arr.shape, n)
        out[name] = arr
    return out
