"""EventBatch: device-resident batch of events (flat layout of include/gtf_fields.h) + stage methods.

The methods carry the reference's function names and argument meaning; each is one call
through the C-ABI.  Reference-side exceptions (ValueError / ZeroDivisionError / KeyError raised by the
Python reference at the cited lines) are re-raised from the `ref_errors` bits the kernels report."""
import ctypes
import numpy as np

from . import fields as F
from . import lib as L

KEY = {"track_state_estimates": 0, "updated_track_states": 1, 0: 0, 1: 1}
TOPOLOGY = ("x", "y", "z", "r", "layer", "volume", "sub", "alive", "sub_off", "sub_state", "sub_event",
            "in_off", "in_src", "slot_dst", "out_off", "out_slot", "rev_slot")
_REF_EXC = ((1, ValueError, "np.min of an empty array (clustering.py:116,120)"),
            (2, ValueError, "nan is not in list (clustering.py:117)"),
            (4, ZeroDivisionError, "1/len({}) (helper.py:90)"),
            (8, KeyError, "edge of the last dict key no longer exists (helper.py:131,138)"),
            (16, KeyError, "missing track_state_estimates entry (extrapolate_merged_states.py:384)"))


class _DeviceArray(object):
    """numpy-style view of device memory for `torch.as_tensor(..., device='cuda')` (__cuda_array_interface__)"""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
        self._owner = owner


class EventBatch(object):
    def __init__(self, host_batch, device=0, geom=(0.3, 0.4, 0.6, 550.0), raise_ref_errors=True):
        """Batch with an arbitrary (possibly mid-pipeline) state: every array of gtf_fields.h is uploaded.  For freshly
        converted events use `EventBatch.with_capacity(...)` + `load_events(...)`: 44 B per hit + 8 B per edge cross PCIe
        and everything else is initialised on the device."""
        hb = F.complete_host_batch(host_batch)
        self.N, self.E, self.S = len(hb["x"]), len(hb["in_src"]), len(hb["sub_off"]) - 1
        self._create(device, geom, raise_ref_errors)
        self.upload(hb)
        L.check(self.lib.gtf_batch_finalize(self.h))

    def _create(self, device, geom, raise_ref_errors):
        self.lib = L.lib()
        self.h = ctypes.c_void_p()
        self.device = device
        L.check(self.lib.gtf_batch_create(self.N, self.E, self.S, device, ctypes.byref(self.h)))
        self.geom = L.Geom(*geom)
        self.raise_ref_errors = raise_ref_errors
        self.last_stats = None

    @classmethod
    def with_capacity(cls, n_nodes, n_slots, n_subgraphs, device=0, geom=(0.3, 0.4, 0.6, 550.0), raise_ref_errors=True):
        """an empty batch that `load_events` fills (and refills: one allocation serves a stream of batches)"""
        self = cls.__new__(cls)
        self.N, self.E, self.S = int(n_nodes), int(n_slots), int(n_subgraphs)
        self._create(device, geom, raise_ref_errors)
        return self

    def load_events(self, ev):
        """event_conversion.py:40-112 on the device: `ev` holds the arrays of lib.EVENT_ARRAYS (hits + in-CSR in dict order +
        out-CSR in successor order, e.g. synth.event_to_host / ingest.load_event_csv output; pinned torch tensors or numpy
        arrays of the exact dtype are passed through without a copy).  Asynchronous on the batch stream."""
        e = L.Events()
        keep = []
        for name, dt in L.EVENT_ARRAYS:
            a = ev[name]
            if hasattr(a, "data_ptr"):                   # torch tensor (pinned host memory)
                ptr = a.data_ptr()
            else:
                a = np.ascontiguousarray(a, dtype=dt)
                ptr = a.ctypes.data
            keep.append(a)
            setattr(e, name, ctypes.cast(ctypes.c_void_p(ptr), dict(L.Events._fields_)[name]))
        e.n_nodes, e.n_slots, e.n_subgraphs = len(ev["x"]), len(ev["in_src"]), len(ev["sub_event"])
        L.check(self.lib.gtf_batch_load_events(self.h, ctypes.byref(e)))
        self.N, self.E, self.S = e.n_nodes, e.n_slots, e.n_subgraphs
        self._loaded = keep                              # the copies are asynchronous: keep the sources alive

    # ---- data movement
    def upload(self, hb, names=None):
        for name in (names or [f[0] for f in F.FIELDS]):
            if name not in hb:
                continue
            arr = np.ascontiguousarray(hb[name], dtype=F.FIELD_DTYPE[name])
            n = F.extent_len(F.FIELD_EXTENT[name], self.N, self.E, self.S)
            assert arr.shape == (n,), (name, arr.shape, n)
            if n:
                L.check(self.lib.gtf_batch_upload(self.h, F.FIELD_ID[name], arr.ctypes.data_as(ctypes.c_void_p)))
        L.check(self.lib.gtf_batch_sync(self.h))

    def download(self, names=None):
        out = {}
        for name in (names or [f[0] for f in F.FIELDS]):
            n = F.extent_len(F.FIELD_EXTENT[name], self.N, self.E, self.S)
            arr = np.empty(n, F.FIELD_DTYPE[name])
            if n:
                L.check(self.lib.gtf_batch_download(self.h, F.FIELD_ID[name], arr.ctypes.data_as(ctypes.c_void_p)))
            out[name] = arr
        return out

    def device_ptr(self, name):
        p = ctypes.c_void_p()
        L.check(self.lib.gtf_batch_device_ptr(self.h, F.FIELD_ID[name], ctypes.byref(p)))
        return p.value

    def stream(self):
        p = ctypes.c_void_p()
        L.check(self.lib.gtf_batch_stream(self.h, ctypes.byref(p)))
        return p.value or 0

    def sync(self):
        L.check(self.lib.gtf_batch_sync(self.h))

    def near_threshold(self):
        """boundary-flip candidates of the most recent stage call / iteration: (count, [(kind, index, value, threshold)])
        -- decisions taken within 1e-9 relative of their threshold; kind indexes lib.NEAR_KINDS, index is a
        slot (gate, reweight) or a node (cluster)"""
        rec = (L.NearRec * 256)()
        n = ctypes.c_int64(0)
        L.check(self.lib.gtf_batch_near_threshold(self.h, rec, 256, ctypes.byref(n)))
        return n.value, [(r.kind, r.index, r.value, r.threshold) for r in rec[:min(n.value, 256)]]

    def device_bytes(self):
        return int(self.lib.gtf_batch_device_bytes(self.h))

    def iteration_launches(self):
        return int(self.lib.gtf_batch_iteration_launches(self.h))

    def close(self):
        if self.h:
            self.lib.gtf_batch_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _done(self, st):
        self.last_stats = st.as_dict()
        if st.ref_errors & 64:
            raise L.GtfError("a kernel index left its array (GTF_STATUS_BOUNDS; debug build)")
        if self.raise_ref_errors and st.ref_errors:
            for bit, exc, msg in _REF_EXC:
                if st.ref_errors & bit:
                    raise exc(msg)
        return self.last_stats

    # ---- reference-named stages
    def compute_track_state_estimates(self):
        """helper.py:238 (sigmas / endcap boundary come from self.geom)."""
        L.check(self.lib.gtf_seed(self.h, ctypes.byref(self.geom)))

    def initialize_edge_activation(self):
        L.check(self.lib.gtf_initialize_edge_activation(self.h))

    def compute_prior_probabilities(self, track_state_key):
        L.check(self.lib.gtf_compute_prior_probabilities(self.h, KEY[track_state_key]))

    def compute_mixture_weights(self, track_state_key):
        st = L.Stats()
        L.check(self.lib.gtf_compute_mixture_weights(self.h, KEY[track_state_key], ctypes.byref(st)))
        return self._done(st)

    def query_node_degree_in_edges(self):
        L.check(self.lib.gtf_query_node_degree(self.h))

    def seed(self, want_stats=True):
        """event_conversion.py:87-96: seed, activate, priors, weights, degree (one call, three kernel launches).
        want_stats=False: no counter read-back (asynchronous)."""
        if not want_stats:
            L.check(self.lib.gtf_seed_all(self.h, ctypes.byref(self.geom), None))
            return None
        st = L.Stats()
        L.check(self.lib.gtf_seed_all(self.h, ctypes.byref(self.geom), ctypes.byref(st)))
        return self._done(st)

    def seed_cluster(self, chi2_threshold, KL_threshold, KL_lut=None):
        """seed() followed by cluster('track_state_estimates', ...) as one pass over the freshly seeded dicts"""
        st = L.Stats()
        lut = None
        if KL_lut is not None:
            lut = np.ascontiguousarray(KL_lut, np.float64)
            lut = lut.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        L.check(self.lib.gtf_seed_cluster(self.h, ctypes.byref(self.geom), chi2_threshold, KL_threshold, lut, ctypes.byref(st)))
        return self._done(st)

    def cluster(self, track_state_key, chi2_threshold, KL_threshold, KL_lut=None):
        st = L.Stats()
        lut = None
        if KL_lut is not None:
            lut = np.ascontiguousarray(KL_lut, np.float64)
            assert lut.shape == (28,)
            lut = lut.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        L.check(self.lib.gtf_cluster(self.h, KEY[track_state_key], chi2_threshold, KL_threshold, lut,
                                     ctypes.byref(self.geom), ctypes.byref(st)))
        return self._done(st)

    def message_passing(self, chi2CutFactor):
        st = L.Stats()
        L.check(self.lib.gtf_message_passing(self.h, chi2CutFactor, ctypes.byref(self.geom), ctypes.byref(st)))
        return self._done(st)

    def reweight(self, track_state_estimates_key="updated_track_states", threshold=0.1):
        st = L.Stats()
        L.check(self.lib.gtf_reweight(self.h, KEY[track_state_estimates_key], threshold, ctypes.byref(st)))
        return self._done(st)

    def extrapolate_stage(self, chi2CutFactor):
        st = L.Stats()
        L.check(self.lib.gtf_extrapolate_stage(self.h, chi2CutFactor, ctypes.byref(self.geom), ctypes.byref(st)))
        return self._done(st)

    def remove_state_metadata(self):
        st = L.Stats()
        L.check(self.lib.gtf_remove_state_metadata(self.h, ctypes.byref(st)))
        return self._done(st)

    def _iter_params(self, chi2_cut, cluster_chi2, cluster_kl, reweight_threshold, KL_lut, record_chi2=False):
        p = L.IterParams(chi2_cut, cluster_chi2, cluster_kl, reweight_threshold, None, 1 if record_chi2 else 0)
        if KL_lut is not None:
            self._lut_keep = np.ascontiguousarray(KL_lut, np.float64)
            p.kl_lut = self._lut_keep.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        return p

    def iterate(self, max_iter=10, stop_when_converged=True, chi2_cut=2.0, cluster_chi2=1000.0, cluster_kl=100.0,
                reweight_threshold=0.1, KL_lut=None, record_chi2=False, want_stats=True):
        """Fused iterations [message_passing, (prior, reweight) x2, cluster(updated states)].
        record_chi2: also keep every message's gate chi2 in `uts_chi2` (diagnostic; the reference only logs it).
        want_stats=False (with stop_when_converged=False): no counter read-back, the call returns without synchronising."""
        p = self._iter_params(chi2_cut, cluster_chi2, cluster_kl, reweight_threshold, KL_lut, record_chi2)
        n = ctypes.c_int(0)
        if not want_stats and not stop_when_converged:
            L.check(self.lib.gtf_iterate(self.h, ctypes.byref(p), ctypes.byref(self.geom), max_iter, 0, None, ctypes.byref(n)))
            return []
        stats = (L.Stats * max_iter)()
        L.check(self.lib.gtf_iterate(self.h, ctypes.byref(p), ctypes.byref(self.geom), max_iter,
                                     1 if stop_when_converged else 0, stats, ctypes.byref(n)))
        out = [stats[i].as_dict() for i in range(n.value)]
        for k, s in enumerate(out):
            if (self.raise_ref_errors and s["ref_errors"]) or (s["ref_errors"] & 64):
                self._done(stats[k])
        return out

    def iterate_dry(self, chi2_cut=2.0, cluster_chi2=1000.0, cluster_kl=100.0, reweight_threshold=0.1, KL_lut=None,
                    want_stats=False):
        p = self._iter_params(chi2_cut, cluster_chi2, cluster_kl, reweight_threshold, KL_lut)
        st = L.Stats()
        L.check(self.lib.gtf_iterate_dry(self.h, ctypes.byref(p), ctypes.byref(self.geom),
                                         ctypes.byref(st) if want_stats else None))
        return st.as_dict() if want_stats else None

    def set_timing(self, enable=True):
        L.check(self.lib.gtf_batch_set_timing(self.h, 1 if enable else 0))

    def timing(self):
        """(prefix_ms, tile_ms, heavy_ms, n): average CUDA-event durations of the kernels of the fused iteration"""
        a, b_, c, n = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_double(0), ctypes.c_int(0)
        L.check(self.lib.gtf_batch_timing(self.h, ctypes.byref(a), ctypes.byref(b_), ctypes.byref(c), ctypes.byref(n)))
        return a.value, b_.value, c.value, n.value

    def timing_kernels(self):
        """{kernel: ms} averages of the packed pipeline: k_send, k_exec, k_node2, cooperative (k_hv<*> + k_big)"""
        ms = (ctypes.c_double * 5)()
        n = ctypes.c_int(0)
        L.check(self.lib.gtf_batch_timing_kernels(self.h, ms, 5, ctypes.byref(n)))
        return {"k_send": ms[0], "k_exec": ms[1], "k_node2": ms[2], "k_hv": ms[3], "n": n.value}

    def CCA(self):
        """extract_track_candidates.py:332: component label per node (smallest node index; -1 = removed)."""
        L.check(self.lib.gtf_components(self.h))
        return self.download(["label"])["label"]

    def extract(self, pval=0.01, numhits=4, sep3d=10.0, merge_dist=8.0, want_arrays=True):
        """extract_track_candidates.py:402-467.  Returns (n accepted nodes, accepted flags, p-values xy, p-values zr);
        want_arrays=False skips the three per-node host arrays (returns None for them)."""
        n = ctypes.c_int32(0)
        dp = ctypes.POINTER(ctypes.c_double)
        acc = pxy = pzr = None
        if want_arrays:
            acc = np.zeros(self.N, np.uint8)
            pxy = np.zeros(self.N)
            pzr = np.zeros(self.N)
        L.check(self.lib.gtf_extract(self.h, ctypes.byref(self.geom), pval, numhits, sep3d, merge_dist, ctypes.byref(n),
                                     acc.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)) if want_arrays else None,
                                     pxy.ctypes.data_as(dp) if want_arrays else None,
                                     pzr.ctypes.data_as(dp) if want_arrays else None))
        L.check(self.lib.gtf_batch_sync(self.h))
        return n.value, acc, pxy, pzr

    def tag_propagation(self, tags, threshold=0.1, max_sweeps=1000):
        tags = np.ascontiguousarray(tags, np.int32).copy()
        n = ctypes.c_int(0)
        L.check(self.lib.gtf_tag_propagate(self.h, threshold, tags.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                           max_sweeps, ctypes.byref(n)))
        return n.value, tags

    def candidates_into(self, rows):
        """the candidate table into a caller-owned (cap, 3) int32 host array (pinned: one D2H copy); returns the row count"""
        n = ctypes.c_int64(0)
        ptr = rows.data_ptr() if hasattr(rows, "data_ptr") else rows.ctypes.data
        L.check(self.lib.gtf_candidates(self.h, ctypes.cast(ctypes.c_void_p(ptr), ctypes.POINTER(ctypes.c_int32)),
                                        int(rows.shape[0]), ctypes.byref(n)))
        return n.value

    def candidates_device(self):
        """the candidate table left on the device: an object exposing __cuda_array_interface__ ((n, 3) int32; valid until
        the next candidates call) -- `torch.as_tensor(t, device="cuda")` hands it to NCCL (shard.gather_candidates)"""
        n = ctypes.c_int64(0)
        p = ctypes.c_void_p()
        L.check(self.lib.gtf_candidates_device(self.h, ctypes.byref(p), ctypes.byref(n)))
        return _DeviceArray(p.value or 0, (n.value, 3), "<i4", self) if n.value else None

    def candidates(self):
        """(event_id, candidate_id, node_index) rows of all nodes accepted so far, sorted."""
        n = ctypes.c_int64(0)
        L.check(self.lib.gtf_candidates(self.h, None, 0, ctypes.byref(n)))
        rows = np.zeros((max(n.value, 1), 3), np.int32)
        L.check(self.lib.gtf_candidates(self.h, rows.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), max(n.value, 1),
                                        ctypes.byref(n)))
        return rows[:n.value]   # sorted by (event, candidate, node) on the device
