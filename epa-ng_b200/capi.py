"""ctypes binding of libepa_b200.so - the C ABI declared in include/epa_b200.h.

This is the stub a maintainer of the reference would write to call the library from Python; the
tests, bench.py and __graft_entry__ use it. It never computes anything itself and has no CPU
fallback: if the shared library is missing or no CUDA device is present, it raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EPA_B200_LIB") or os.path.join(HERE, "libepa_b200.so")   # override: developer A/B builds

EPA_OK, EPA_ERR_ARG, EPA_ERR_CUDA, EPA_ERR_NOMEM, EPA_ERR_STATE, EPA_ERR_QUERY = 0, -1, -2, -3, -4, -5
EPA_FLAG_RATE_SCALERS = 1
EPA_FLAG_BUGCOMPAT_FOCUS = 2


class EpaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libepa_b200 error {code}: {msg}")
        self.code = code


class ModelDesc(C.Structure):
    _fields_ = [("states", C.c_uint32), ("rate_cats", C.c_uint32), ("sites", C.c_uint32), ("flags", C.c_uint32),
                ("eigenvals", C.POINTER(C.c_double)), ("eigenvecs", C.POINTER(C.c_double)),
                ("inv_eigenvecs", C.POINTER(C.c_double)), ("freqs", C.POINTER(C.c_double)),
                ("rates", C.POINTER(C.c_double)), ("rate_weights", C.POINTER(C.c_double)),
                ("pinv", C.c_double)]


class EdgeDesc(C.Structure):
    _fields_ = [("distal", C.c_uint32), ("proximal", C.c_uint32), ("length", C.c_double)]


class ClvOp(C.Structure):
    _fields_ = [("parent", C.c_uint32), ("left", C.c_uint32), ("right", C.c_uint32), ("reserved", C.c_uint32),
                ("left_length", C.c_double), ("right_length", C.c_double)]


class HostClv(C.Structure):
    _fields_ = [("slot", C.c_uint32), ("reserved", C.c_uint32), ("clv", C.POINTER(C.c_double)),
                ("scaler", C.POINTER(C.c_uint32))]


class Options(C.Structure):
    _fields_ = [("prescoring", C.c_int32), ("heuristic", C.c_int32), ("prescoring_threshold", C.c_double),
                ("premasking", C.c_int32), ("sliding_blo", C.c_int32), ("filter_acc_lwr", C.c_int32),
                ("support_threshold", C.c_double), ("filter_min", C.c_uint32), ("filter_max", C.c_uint32)]


PLACEMENT_DTYPE = np.dtype([("branch_id", np.uint64), ("likelihood", np.float64), ("lwr", np.float64),
                            ("pendant_length", np.float64), ("distal_length", np.float64)])

# every symbol include/epa_b200.h declares: (name, restype, argtypes)
_dp, _u32p, _vp = C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.c_void_p
SYMBOLS = [
    ("epa_options_default", None, [C.POINTER(Options)]),
    ("epa_ctx_create", C.c_int, [C.POINTER(_vp), C.c_int, C.POINTER(ModelDesc), C.c_uint32, _u32p, C.c_uint32,
                                 C.POINTER(EdgeDesc), C.c_uint32]),
    ("epa_compute_clvs", C.c_int, [_vp, C.POINTER(ClvOp), C.c_uint32]),
    ("epa_upload_clvs", C.c_int, [_vp, C.POINTER(HostClv), C.c_uint32]),
    ("epa_build_lookup", C.c_int, [_vp]),
    ("epa_place_chunk", C.c_int, [_vp, _vp, C.c_uint32, C.POINTER(Options), _vp, _u32p]),
    ("epa_upload_queries", C.c_int, [_vp, _vp, C.c_uint32, C.c_int]),
    ("epa_hint_next_chunk", C.c_int, [_vp, _vp, C.c_uint32]),
    ("epa_set_deferred_results", C.c_int, [_vp, C.c_int]),
    ("epa_wait_results", C.c_int, [_vp]),
    ("epa_wait_older_results", C.c_int, [_vp]),
    ("epa_encode_queries_dev", C.c_int, [_vp, _vp, C.c_uint32, C.c_int]),
    ("epa_preplace", C.c_int, [_vp]),
    ("epa_hint_selection", C.c_int, [_vp, C.POINTER(Options)]),
    ("epa_select", C.c_int, [_vp, C.POINTER(Options), C.POINTER(C.c_uint64)]),
    ("epa_place_pairs", C.c_int, [_vp, C.POINTER(Options)]),
    ("epa_collect", C.c_int, [_vp, C.POINTER(Options), _vp, _u32p]),
    ("epa_collect_dev", C.c_int, [_vp, C.POINTER(Options), _vp, _vp]),
    ("epa_ctx_set_stream", C.c_int, [_vp, _vp]),
    ("epa_get_clv", C.c_int, [_vp, C.c_uint32, _dp, _u32p]),
    ("epa_get_lookup", C.c_int, [_vp, C.c_uint32, _dp]),
    ("epa_get_prescores", C.c_int, [_vp, _dp]),
    ("epa_get_pairs", C.c_int, [_vp, _u32p, _u32p, _vp, C.c_uint64]),
    ("epa_edge_loglikelihood", C.c_int, [_vp, C.c_uint32, _dp]),
    ("epa_last_timings", C.c_int, [_vp, C.POINTER(C.c_float)]),
    ("epa_last_lookup_ms", C.c_int, [_vp, C.POINTER(C.c_float)]),
    ("epa_num_pairs", C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    ("epa_synchronize", C.c_int, [_vp]),
    ("epa_measure_fp64_peak", C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    ("epa_device_pool_trim", None, []),
    ("epa_peer_alloc", C.c_int, [C.c_int, C.c_size_t, C.POINTER(_vp), C.c_char_p]),
    ("epa_peer_open", C.c_int, [C.c_int, C.c_char_p, C.POINTER(_vp)]),
    ("epa_peer_close", C.c_int, [C.c_int, _vp]),
    ("epa_peer_free", C.c_int, [C.c_int, _vp]),
    ("epa_pinned_alloc", C.c_int, [C.POINTER(_vp), C.c_size_t]),
    ("epa_pinned_free", None, [_vp]),
    ("epa_launch_count", C.c_uint64, [_vp]),
    ("epa_last_error", C.c_char_p, [_vp]),
    ("epa_ctx_destroy", None, [_vp]),
]

_LIB = None


def load():
    """Loads libepa_b200.so and binds every declared symbol; raises if anything is missing."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _LIB = lib
    return _LIB


def default_options(**kw) -> Options:
    o = Options()
    load().epa_options_default(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def measure_fp64_peak(device=0) -> float:
    """Measured DFMA throughput of the device in TFLOP/s (the thorough kernel's roofline denominator)."""
    v = C.c_double()
    rc = load().epa_measure_fp64_peak(device, C.byref(v))
    if rc != 0:
        raise EpaError(rc, "epa_measure_fp64_peak")
    return v.value


def _ptr(a, typ):
    return a.ctypes.data_as(typ)


class Context:
    """Owns one epa_ctx. Arrays are numpy; every call raises EpaError on a non-zero status."""

    def __init__(self, *, states, rate_cats, sites, eigenvals, eigenvecs, inv_eigenvecs, freqs, rates, weights,
                 tip_masks, n_clv_slots, edges, device=0, flags=0, pinv=0.0):
        self.lib = load()
        self._keep = [np.ascontiguousarray(x, dtype=np.float64) for x in
                      (eigenvals, eigenvecs, inv_eigenvecs, freqs, rates, weights)]
        md = ModelDesc(states, rate_cats, sites, flags, *[_ptr(a, _dp) for a in self._keep], pinv)
        tip_masks = np.ascontiguousarray(tip_masks, dtype=np.uint32)
        assert tip_masks.ndim == 2 and tip_masks.shape[1] == sites
        earr = (EdgeDesc * len(edges))(*[EdgeDesc(int(d), int(p), float(l)) for d, p, l in edges])
        self.states, self.rate_cats, self.sites = states, rate_cats, sites
        self.scalers_per_site = rate_cats if (flags & EPA_FLAG_RATE_SCALERS) else 1
        self.n_tips, self.n_edges = tip_masks.shape[0], len(edges)
        self.handle = _vp()
        rc = self.lib.epa_ctx_create(C.byref(self.handle), device, C.byref(md), tip_masks.shape[0],
                                     _ptr(tip_masks, _u32p), n_clv_slots, earr, len(edges))
        if rc != 0:
            raise EpaError(rc, self.lib.epa_last_error(None).decode())
        self.nq = 0

    def _check(self, rc):
        if rc != 0:
            raise EpaError(rc, self.lib.epa_last_error(self.handle).decode())

    def close(self):
        if self.handle:
            self.lib.epa_ctx_destroy(self.handle)
            self.handle = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference state ----
    def compute_clvs(self, ops):
        arr = (ClvOp * len(ops))(*[ClvOp(int(p), int(l), int(r), 0, float(ll), float(rl)) for p, l, r, ll, rl in ops])
        self._check(self.lib.epa_compute_clvs(self.handle, arr, len(ops)))

    def upload_clvs(self, clvs):
        keep, arr = [], (HostClv * len(clvs))()
        for i, (slot, clv, scaler) in enumerate(clvs):
            c = np.ascontiguousarray(clv, dtype=np.float64)
            s = None if scaler is None else np.ascontiguousarray(scaler, dtype=np.uint32)
            keep.append((c, s))
            arr[i] = HostClv(int(slot), 0, _ptr(c, _dp), _ptr(s, _u32p) if s is not None else None)
        self._check(self.lib.epa_upload_clvs(self.handle, arr, len(clvs)))

    def get_clv(self, node):
        clv = np.zeros(self.sites * self.rate_cats * self.states)
        sc = np.zeros(self.sites * getattr(self, "scalers_per_site", 1), dtype=np.uint32)
        self._check(self.lib.epa_get_clv(self.handle, node, _ptr(clv, _dp), _ptr(sc, _u32p)))
        return clv, sc

    def edge_loglikelihood(self, edge):
        v = C.c_double()
        self._check(self.lib.epa_edge_loglikelihood(self.handle, edge, C.byref(v)))
        return v.value

    def build_lookup(self):
        self._check(self.lib.epa_build_lookup(self.handle))

    def get_lookup(self, edge):
        k = 16 if self.states == 4 else 24
        out = np.zeros((self.sites, k))
        self._check(self.lib.epa_get_lookup(self.handle, edge, _ptr(out, _dp)))
        return out

    def lookup_ms(self):
        v = C.c_float()
        self._check(self.lib.epa_last_lookup_ms(self.handle, C.byref(v)))
        return v.value

    # ---- chunk pipeline ----
    @staticmethod
    def _rows(seqs, sites):
        """Accepts uint8[nq][sites] or a list of str/bytes; returns a contiguous uint8 matrix."""
        if isinstance(seqs, np.ndarray):
            a = np.ascontiguousarray(seqs, dtype=np.uint8)
        else:
            a = np.frombuffer(b"".join(s.encode() if isinstance(s, str) else s for s in seqs), dtype=np.uint8)
            a = a.reshape(len(seqs), -1) if len(seqs) else a.reshape(0, sites)
        if a.ndim != 2 or a.shape[1] != sites:
            raise ValueError(f"query rows must have {sites} columns")
        return a

    def upload_queries(self, seqs, premasking=True):
        a = self._rows(seqs, self.sites)
        self.nq = a.shape[0]
        self._check(self.lib.epa_upload_queries(self.handle, a.ctypes.data, a.shape[0], int(premasking)))

    def hint_next_chunk(self, host_ptr, nq):
        """Announces the chunk after the next upload: its H2D copy overlaps the placement."""
        self._check(self.lib.epa_hint_next_chunk(self.handle, host_ptr, nq))

    def upload_queries_ptr(self, host_ptr, nq, premasking=True):
        self.nq = nq
        self._check(self.lib.epa_upload_queries(self.handle, host_ptr, nq, int(premasking)))

    def encode_queries_dev(self, dev_ptr, nq, premasking=True):
        self.nq = nq
        self._check(self.lib.epa_encode_queries_dev(self.handle, dev_ptr, nq, int(premasking)))

    def hint_selection(self, opts):
        """Announces the options of the next select(): the preplacement may then select in its epilogue."""
        self._check(self.lib.epa_hint_selection(self.handle, C.byref(opts) if opts is not None else None))

    def preplace(self):
        self._check(self.lib.epa_preplace(self.handle))

    def get_prescores(self):
        out = np.zeros((self.nq, self.n_edges))
        self._check(self.lib.epa_get_prescores(self.handle, _ptr(out, _dp)))
        return out

    def select(self, opts):
        n = C.c_uint64()
        self._check(self.lib.epa_select(self.handle, C.byref(opts), C.byref(n)))
        return n.value

    def place_pairs(self, opts):
        self._check(self.lib.epa_place_pairs(self.handle, C.byref(opts)))

    def get_pairs(self, raw=True):
        n = C.c_uint64()
        self._check(self.lib.epa_num_pairs(self.handle, C.byref(n)))
        q = np.zeros(n.value, dtype=np.uint32)
        e = np.zeros(n.value, dtype=np.uint32)
        r = np.zeros(n.value, dtype=PLACEMENT_DTYPE) if raw else None
        self._check(self.lib.epa_get_pairs(self.handle, _ptr(q, _u32p), _ptr(e, _u32p),
                                           r.ctypes.data if raw else None, n.value))
        return q, e, r

    def collect(self, opts, out=None, counts=None):
        if out is None:
            out = np.zeros((self.nq, opts.filter_max), dtype=PLACEMENT_DTYPE)
        if counts is None:
            counts = np.zeros(self.nq, dtype=np.uint32)
        self._check(self.lib.epa_collect(self.handle, C.byref(opts), out.ctypes.data, _ptr(counts, _u32p)))
        return out, counts

    def collect_dev(self, opts, out_dev_ptr, counts_dev_ptr):
        self._check(self.lib.epa_collect_dev(self.handle, C.byref(opts), out_dev_ptr, counts_dev_ptr))

    def set_stream(self, cuda_stream):
        self._check(self.lib.epa_ctx_set_stream(self.handle, cuda_stream))

    def place_chunk(self, seqs, opts, out=None, counts=None):
        a = self._rows(seqs, self.sites)
        self.nq = a.shape[0]
        if out is None:
            out = np.zeros((self.nq, opts.filter_max), dtype=PLACEMENT_DTYPE)
        if counts is None:
            counts = np.zeros(self.nq, dtype=np.uint32)
        self._check(self.lib.epa_place_chunk(self.handle, a.ctypes.data, self.nq, C.byref(opts), out.ctypes.data,
                                             _ptr(counts, _u32p)))
        return out, counts

    def place_chunk_ptr(self, host_ptr, nq, opts, out_ptr, counts_ptr):
        self.nq = nq
        self._check(self.lib.epa_place_chunk(self.handle, host_ptr, nq, C.byref(opts), out_ptr,
                                             C.cast(counts_ptr, _u32p)))

    def timings(self):
        ms = (C.c_float * 5)()
        self._check(self.lib.epa_last_timings(self.handle, ms))
        return dict(zip(("upload_encode", "preplace", "select", "thorough", "collect"), [float(x) for x in ms]))

    def synchronize(self):
        self._check(self.lib.epa_synchronize(self.handle))

    def launch_count(self):
        return int(self.lib.epa_launch_count(self.handle))
