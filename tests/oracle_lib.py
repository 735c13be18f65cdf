"""ctypes binding of the CPU oracle (oracle/libgtf_oracle.so).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes
import os
import subprocess
import sys
import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, REPO)
import gtf_b200  # noqa: E402
from gtf_b200 import fields as F  # noqa: E402

_CT = {"double": ctypes.c_double, "int32_t": ctypes.c_int32, "uint8_t": ctypes.c_uint8, "int8_t": ctypes.c_int8}


class Arrays(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int32), ("E", ctypes.c_int32), ("S", ctypes.c_int32), ("pad_", ctypes.c_int32)] + \
               [(n, ctypes.POINTER(_CT[ct])) for n, ct, _ in F.FIELDS]


class Geom(ctypes.Structure):
    _fields_ = [("sigma0xy", ctypes.c_double), ("sigma0rz", ctypes.c_double),
                ("sigma0rz2", ctypes.c_double), ("endcap_boundary", ctypes.c_double)]


class Stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in ("nodes_with_state", "nodes_merged", "edges_deactivated",
                                               "edges_sent", "edges_gated", "edges_reweight_off")]


_lib = None


def build():
    so = os.path.join(REPO, "oracle", "libgtf_oracle.so")
    src = os.path.join(REPO, "oracle", "gtf_oracle.c")
    if (not os.path.exists(so)) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle")], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        dp = ctypes.POINTER(ctypes.c_double)
        L.gtfo_kl_distance.restype = ctypes.c_double
        L.gtfo_kl_distance.argtypes = [dp] * 4
        L.gtfo_mahalanobis.restype = ctypes.c_double
        L.gtfo_mahalanobis.argtypes = [dp] * 7 + [ctypes.c_double] * 3
        L.gtfo_merge_states.argtypes = [dp] * 6
        L.gtfo_chi2_sf.restype = ctypes.c_double
        L.gtfo_chi2_sf.argtypes = [ctypes.c_double, ctypes.c_double]
        L.gtfo_cluster.argtypes = [ctypes.POINTER(Arrays), ctypes.c_int, ctypes.c_double, ctypes.c_double, dp,
                                   ctypes.POINTER(Geom), ctypes.POINTER(Stats)]
        L.gtfo_message_passing.argtypes = [ctypes.POINTER(Arrays), ctypes.c_double, ctypes.POINTER(Geom),
                                           ctypes.POINTER(Stats)]
        L.gtfo_reweight.argtypes = [ctypes.POINTER(Arrays), ctypes.c_int, ctypes.c_double, ctypes.POINTER(Stats)]
        L.gtfo_remove_state_metadata.argtypes = [ctypes.POINTER(Arrays), ctypes.POINTER(Stats)]
        L.gtfo_seed.argtypes = [ctypes.POINTER(Arrays), ctypes.POINTER(Geom)]
        for fn in ("gtfo_compute_prior_probabilities", "gtfo_compute_mixture_weights"):
            getattr(L, fn).argtypes = [ctypes.POINTER(Arrays), ctypes.c_int]
        for fn in ("gtfo_initialize_edge_activation", "gtfo_query_node_degree", "gtfo_cca"):
            getattr(L, fn).argtypes = [ctypes.POINTER(Arrays)]
        L.gtfo_extract.argtypes = [ctypes.POINTER(Arrays), ctypes.POINTER(Geom), ctypes.c_double, ctypes.c_int,
                                   ctypes.c_double, ctypes.c_double, ctypes.POINTER(ctypes.c_uint8), dp, dp]
        L.gtfo_tag_propagation.argtypes = [ctypes.POINTER(Arrays), ctypes.c_double,
                                           ctypes.POINTER(ctypes.c_int32), ctypes.c_int]
        _lib = L
    return _lib


ERRORS = {1: ValueError, 2: ValueError, 4: ZeroDivisionError, 8: KeyError, 16: KeyError}


class OracleBatch(object):
    """A host batch (dict of numpy arrays, every GTF_FIELDS array present) bound to the oracle."""

    def __init__(self, hb, geom=(0.3, 0.4, 0.6, 550.0)):
        self.hb = {k: np.array(v, copy=True) for k, v in F.complete_host_batch(hb).items()}   # never alias the caller's arrays
        self.N, self.E, self.S = len(self.hb["x"]), len(self.hb["in_src"]), len(self.hb["sub_off"]) - 1
        self.A = Arrays()
        self.A.N, self.A.E, self.A.S = self.N, self.E, self.S
        for n, ct, _ in F.FIELDS:
            setattr(self.A, n, self.hb[n].ctypes.data_as(ctypes.POINTER(_CT[ct])))
        self.geom = Geom(*geom)
        self.stats = Stats()
        self.err = 0

    def _p(self):
        return ctypes.byref(self.A)

    def seed(self):
        self.err |= lib().gtfo_seed(self._p(), ctypes.byref(self.geom))
        L = lib()
        L.gtfo_initialize_edge_activation(self._p())
        L.gtfo_compute_prior_probabilities(self._p(), 0)
        self.err |= L.gtfo_compute_mixture_weights(self._p(), 0)
        L.gtfo_query_node_degree(self._p())

    def cluster(self, key, chi2_thr, kl_thr, lut=None):
        lp = None if lut is None else np.ascontiguousarray(lut, np.float64).ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        e = lib().gtfo_cluster(self._p(), key, chi2_thr, kl_thr, lp, ctypes.byref(self.geom), ctypes.byref(self.stats))
        self.err |= e
        return e

    def message_passing(self, chi2_cut):
        e = lib().gtfo_message_passing(self._p(), chi2_cut, ctypes.byref(self.geom), ctypes.byref(self.stats))
        self.err |= e
        return e

    def prior(self, key):
        return lib().gtfo_compute_prior_probabilities(self._p(), key)

    def mixture_weights(self, key):
        return lib().gtfo_compute_mixture_weights(self._p(), key)

    def reweight(self, key=1, thr=0.1):
        e = lib().gtfo_reweight(self._p(), key, thr, ctypes.byref(self.stats))
        self.err |= e
        return e

    def degree(self):
        return lib().gtfo_query_node_degree(self._p())

    def extrapolate_stage(self, chi2_cut):
        """extrapolate_merged_states.main(): message_passing, (prior, reweight) x2, degree (:552-567)."""
        self.message_passing(chi2_cut)
        for _ in range(2):
            self.prior(1)
            self.reweight(1)
        self.degree()

    def remove_state_metadata(self):
        e = lib().gtfo_remove_state_metadata(self._p(), ctypes.byref(self.stats))
        self.err |= e
        return e

    def cca(self):
        lib().gtfo_cca(self._p())
        return self.hb["label"]

    def extract(self, pval=0.01, numhits=4, sep3d=10.0, merge_dist=8.0):
        acc = np.zeros(self.N, np.uint8)
        pxy = np.zeros(self.N)
        pzr = np.zeros(self.N)
        dp = ctypes.POINTER(ctypes.c_double)
        n = lib().gtfo_extract(self._p(), ctypes.byref(self.geom), pval, numhits, sep3d, merge_dist,
                               acc.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                               pxy.ctypes.data_as(dp), pzr.ctypes.data_as(dp))
        return n, acc, pxy, pzr

    def tag_propagation(self, tags, thr=0.1, max_sweeps=1000):
        tags = np.ascontiguousarray(tags, np.int32).copy()
        n = lib().gtfo_tag_propagation(self._p(), thr, tags.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), max_sweeps)
        return n, tags
