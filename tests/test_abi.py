"""CPU: the C-ABI shared library loads without a GPU and exports every symbol include/gtf.h declares;
compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import golden_util as gu
import gtf_b200
from gtf_b200 import lib as L, fields as F


def declared_symbols():
    txt = open(os.path.join(gu.REPO, "include", "gtf.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gtf_[a-z_0-9]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported():
    L.build()
    dll = ctypes.CDLL(L.SO)
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(dll, s), s


def test_field_table_matches_header():
    lib = L.lib()
    assert lib.gtf_abi_version() == 3
    assert lib.gtf_field_count() == len(F.FIELDS)
    for i, (name, _, _) in enumerate(F.FIELDS):
        assert lib.gtf_field_name(i).decode() == name
        assert lib.gtf_field_id(name.encode()) == i


def test_no_cpu_fallback():
    lib = L.lib()
    if lib.gtf_device_count() > 0:
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.gtf_batch_create(10, 10, 1, 0, ctypes.byref(h))
    assert rc == -1 and b"CUDA" in lib.gtf_last_error()
    fx = gu.load("barrel25_deg6")
    with pytest.raises(L.GtfError):
        gtf_b200.EventBatch(gu.stage_batch(fx, "seed"))
