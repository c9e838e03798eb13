"""ctypes binding of the host layer (include/epa_b200_host.h): sessions, whole-run driver, jplace."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi

_vp, _u32p = C.c_void_p, C.POINTER(C.c_uint32)
HOST_SYMBOLS = [
    ("epa_session_open", C.c_int, [C.POINTER(_vp), C.c_char_p, C.c_uint32, C.POINTER(C.c_char_p), _vp, C.c_uint32,
                                   C.c_char_p, C.c_int]),
    ("epa_session_open_ex", C.c_int, [C.POINTER(_vp), C.c_char_p, C.c_uint32, C.POINTER(C.c_char_p), _vp, C.c_uint32,
                                   C.c_char_p, C.c_int, C.c_int, C.c_int]),
    ("epa_session_place", C.c_int, [_vp, _vp, C.c_uint64, C.POINTER(capi.Options), C.c_uint32, _vp, _vp]),
    ("epa_host_set_rate_scalers", C.c_int, [C.c_int, C.c_int]),
    ("epa_host_read_alignment", C.c_int, [C.c_char_p, _u32p, _u32p, _vp, C.c_size_t, C.c_char_p, C.c_size_t]),
    ("epa_host_fasta_to_bfast", C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]),
    ("epa_host_empirical_frequencies", C.c_int, [C.c_char_p, _u32p, C.c_uint32, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("epa_host_model_from_file", C.c_int, [C.c_char_p, C.c_char_p, C.c_size_t]),
    ("epa_session_ctx", _vp, [_vp]),
    ("epa_session_num_edges", C.c_uint32, [_vp]),
    ("epa_session_num_tips", C.c_uint32, [_vp]),
    ("epa_session_sites", C.c_uint32, [_vp]),
    ("epa_session_numbered_newick", C.c_char_p, [_vp, C.c_int]),
    ("epa_session_tree_logl", C.c_int, [_vp, C.POINTER(C.c_double)]),
    ("epa_session_set_preserve_rooting", C.c_int, [_vp, C.c_int]),
    ("epa_session_is_rooted", C.c_int, [_vp]),
    ("epa_session_close", None, [_vp]),
    ("epa_run_files", C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(capi.Options),
                                C.c_uint32, C.c_int, C.c_int, C.c_char_p]),
    ("epa_run_files_ex", C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(capi.Options),
                                   C.c_uint32, C.c_int, C.c_int, C.c_char_p, C.c_int]),
    ("epa_run_files_multi", C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(capi.Options),
                                      C.c_uint32, C.c_int, C.POINTER(C.c_int), C.c_uint32, C.c_char_p, C.c_int, C.c_int, _vp]),
    ("epa_host_release_pinned_pool", None, []),
    ("epa_host_read_alignment_mt", C.c_int, [C.c_char_p, C.c_int, C.c_int, _u32p, _u32p, _vp, C.c_size_t, C.c_char_p,
                                             C.c_size_t, _vp]),
    ("epa_host_format_fixed", C.c_int, [C.c_double, C.c_int, C.c_char_p, C.c_size_t]),
    ("epa_host_map_rooted", C.c_int, [C.c_char_p, _u32p, C.POINTER(C.c_double), C.c_uint32, C.c_char_p, C.c_size_t]),
    ("epa_write_jplace", C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_uint64, _vp, _u32p,
                                   C.c_uint32, C.c_int]),
    ("epa_host_parse_tree", C.c_int, [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t, _u32p, _u32p]),
    ("epa_host_tree_schedule", C.c_int, [C.c_char_p, _u32p, C.POINTER(capi.ClvOp), C.c_uint32, _u32p,
                                         C.POINTER(capi.EdgeDesc), C.c_uint32, _u32p, C.c_char_p, C.c_size_t]),
    ("epa_host_parse_model", C.c_int, [C.c_char_p, _u32p, _u32p] + [C.POINTER(C.c_double)] * 6),
    ("epa_host_last_error", C.c_char_p, []),
]
_bound = False


def lib():
    global _bound
    L = capi.load()
    if not _bound:
        for name, res, args in HOST_SYMBOLS:
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _bound = True
    return L


def _check(rc):
    if rc != 0:
        raise capi.EpaError(rc, lib().epa_host_last_error().decode())


class _BorrowedContext(capi.Context):
    """The session's epa_ctx seen through the staged-API wrapper (not owned)."""

    def __init__(self, handle, states, rate_cats, sites, n_tips, n_edges):
        self.lib = capi.load()
        self.handle = handle
        self.states, self.rate_cats, self.sites = states, rate_cats, sites
        self.n_tips, self.n_edges = n_tips, n_edges
        self.nq = 0

    def close(self):
        self.handle = _vp()


class Session:
    """Reference tree + MSA + model resident on one GPU (mirrors the reference's Tree object)."""

    def __init__(self, newick: str, names, ref_rows: np.ndarray, model: str, device: int = 0, states: int = 0,
                 rate_cats: int = 0, rate_scalers: str = "auto", bugcompat_focus: bool = True):
        """rate_scalers: "off", "on" or "auto" (the reference's --rate-scalers; auto = on above 2000 tips);
        bugcompat_focus reproduces the reference's scaler window offset of the thorough phase."""
        L = lib()
        if not states or not rate_cats:
            pm = parse_model(model)                 # shapes of the staged inspection helpers (get_clv, ...)
            states, rate_cats = pm["states"], pm["rate_cats"]
        ref_rows = np.ascontiguousarray(ref_rows, dtype=np.uint8)
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        self.handle = _vp()
        _check(L.epa_session_open_ex(C.byref(self.handle), newick.encode(), len(names), arr, ref_rows.ctypes.data,
                                     ref_rows.shape[1], model.encode(), device,
                                     {"off": 0, "on": 1, "auto": 2}[rate_scalers], int(bugcompat_focus)))
        self.sites = int(L.epa_session_sites(self.handle))
        self.n_edges = int(L.epa_session_num_edges(self.handle))
        self.n_tips = int(L.epa_session_num_tips(self.handle))
        self.ctx = _BorrowedContext(_vp(L.epa_session_ctx(self.handle)), states, rate_cats, self.sites, self.n_tips,
                                    self.n_edges)

    def numbered_newick(self, precision=10) -> str:
        return lib().epa_session_numbered_newick(self.handle, precision).decode()

    def tree_logl(self) -> float:
        v = C.c_double()
        _check(lib().epa_session_tree_logl(self.handle, C.byref(v)))
        return v.value

    def place(self, query_rows, opts=None, chunk_size=0, out=None, counts=None):
        """query_rows: uint8[nq][sites] numpy array or (host pointer, nq)."""
        opts = opts or capi.default_options()
        if isinstance(query_rows, tuple):
            ptr, nq = query_rows
        else:
            query_rows = np.ascontiguousarray(query_rows, dtype=np.uint8)
            assert query_rows.ndim == 2 and query_rows.shape[1] == self.sites
            ptr, nq = query_rows.ctypes.data, query_rows.shape[0]
        if out is None:
            out = np.zeros((nq, opts.filter_max), dtype=capi.PLACEMENT_DTYPE)
            counts = np.zeros(nq, dtype=np.uint32)
        optr = out if isinstance(out, int) else out.ctypes.data
        cptr = counts if isinstance(counts, int) else counts.ctypes.data
        _check(lib().epa_session_place(self.handle, ptr, nq, C.byref(opts), chunk_size, optr, cptr))
        return out, counts

    def close(self):
        if self.handle:
            lib().epa_session_close(self.handle)
            self.handle = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_files(tree_file, ref_msa, query_file, model, outdir, opts=None, chunk_size=0, precision=10, device=0,
              invocation="epa_run_files", preserve_rooting=True):
    opts = opts or capi.default_options()
    _check(lib().epa_run_files_ex(tree_file.encode(), ref_msa.encode(), query_file.encode(), model.encode(),
                                  outdir.encode(), C.byref(opts), chunk_size, precision, device, invocation.encode(),
                                  int(preserve_rooting)))


class RunStats(C.Structure):
    """epa_run_stats (include/epa_b200_host.h)"""
    _fields_ = [("n_queries", C.c_uint64), ("seconds_total", C.c_double), ("seconds_index", C.c_double),
                ("seconds_setup", C.c_double), ("seconds_place", C.c_double), ("busy_read", C.c_double),
                ("busy_write", C.c_double), ("busy_device_max", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def run_files_multi(tree_file, ref_msa, query_file, model, outdir, devices=(0,), opts=None, chunk_size=0, precision=10,
                    invocation="epa_run_files_multi", preserve_rooting=True, host_threads=0):
    """Files -> jplace as a reader / device(s) / writer pipeline over the listed GPUs; returns the run's wall-clock stats."""
    opts = opts or capi.default_options()
    devs = (C.c_int * len(devices))(*devices)
    st = RunStats()
    _check(lib().epa_run_files_multi(tree_file.encode(), ref_msa.encode(), query_file.encode(), model.encode(),
                                     outdir.encode(), C.byref(opts), chunk_size, precision, devs, len(devices),
                                     invocation.encode(), int(preserve_rooting), host_threads, C.byref(st)))
    return st.as_dict()


def _label_cap(path):
    # the labels cannot be longer than the file itself
    return os.path.getsize(path) + 1024


def read_alignment(path: str):
    """(names, uint8 rows [n][sites]) of a FASTA or bfast file, read by the host layer."""
    n, sites = C.c_uint32(), C.c_uint32()
    _check(lib().epa_host_read_alignment(path.encode(), C.byref(n), C.byref(sites), None, 0, None, 0))
    rows = np.zeros((n.value, sites.value), dtype=np.uint8)
    labels = C.create_string_buffer(_label_cap(path))
    _check(lib().epa_host_read_alignment(path.encode(), C.byref(n), C.byref(sites), rows.ctypes.data, rows.size, labels,
                                         len(labels)))
    return labels.value.decode().split("\n")[:n.value], rows


def read_alignment_mt(path: str, threads=4, want_mask=True):
    """The pipeline's reader (memory map + parallel index/decode): (names, rows, all-gap column mask)."""
    n, sites = C.c_uint32(), C.c_uint32()
    _check(lib().epa_host_read_alignment_mt(path.encode(), threads, int(want_mask), C.byref(n), C.byref(sites), None, 0,
                                            None, 0, None))
    rows = np.zeros((n.value, sites.value), dtype=np.uint8)
    mask = np.zeros(sites.value, dtype=np.uint8)
    labels = C.create_string_buffer(_label_cap(path))
    _check(lib().epa_host_read_alignment_mt(path.encode(), threads, int(want_mask), C.byref(n), C.byref(sites),
                                            rows.ctypes.data, rows.size, labels, len(labels), mask.ctypes.data))
    return labels.value.decode().split("\n")[:n.value], rows, mask


def format_fixed(value: float, precision: int) -> str:
    """printf('%.*f') as the jplace writer formats it (no printf involved)."""
    buf = C.create_string_buffer(512)
    n = lib().epa_host_format_fixed(C.c_double(value), precision, buf, len(buf))
    if n < 0:
        raise capi.EpaError(n, "format_fixed")
    return buf.raw[:n].decode()


def fasta_to_bfast(fasta_path: str, out_dir: str) -> str:
    """The reference's -c/--bfast converter; returns the path of the written file."""
    buf = C.create_string_buffer(4096)
    _check(lib().epa_host_fasta_to_bfast(fasta_path.encode(), out_dir.encode(), buf, len(buf)))
    return buf.value.decode()


def empirical_frequencies(model: str, tip_masks):
    """(freqs, eigenvals) of a +F / +FC model on the given tip state masks [n_tips][sites]."""
    m = np.ascontiguousarray(tip_masks, dtype=np.uint32)
    f, ev = np.zeros(20), np.zeros(20)
    _check(lib().epa_host_empirical_frequencies(model.encode(), m.ctypes.data_as(_u32p), m.shape[0], m.shape[1],
                                                f.ctypes.data_as(C.POINTER(C.c_double)), ev.ctypes.data_as(C.POINTER(C.c_double))))
    return f, ev


def model_from_file(path: str) -> str:
    """Model string of a RAxML 8 info file, raxml-ng .bestModel file or IQ-TREE report (the reference's -m <file>)."""
    buf = C.create_string_buffer(16384)
    _check(lib().epa_host_model_from_file(path.encode(), buf, len(buf)))
    return buf.value.decode()


def map_rooted(newick: str, edges, distal):
    """Rooted input: (edge, distal) on the unrooted working tree -> rooted tree; also returns the
    numbered newick of the working tree."""
    e = np.ascontiguousarray(edges, dtype=np.uint32).copy()
    d = np.ascontiguousarray(distal, dtype=np.float64).copy()
    buf = C.create_string_buffer(8 * len(newick) + 4096)
    _check(lib().epa_host_map_rooted(newick.encode(), e.ctypes.data_as(_u32p), d.ctypes.data_as(C.POINTER(C.c_double)),
                                     len(e), buf, len(buf)))
    return e, d, buf.value.decode()


def write_jplace(path, numbered_newick, invocation, names, recs, counts, precision=10):
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    recs = np.ascontiguousarray(recs)
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    _check(lib().epa_write_jplace(path.encode(), numbered_newick.encode(), invocation.encode(), arr, len(names),
                                  recs.ctypes.data, counts.ctypes.data_as(_u32p), recs.shape[1], precision))


# ---- device-free inspection of the host logic (CPU tests) ----
def parse_tree(newick: str, precision=10):
    """-> (numbered newick, n_tips, n_edges)"""
    buf = C.create_string_buffer(max(4096, 4 * len(newick) + 64 * newick.count(",") + 4096))
    nt, ne = C.c_uint32(), C.c_uint32()
    _check(lib().epa_host_parse_tree(newick.encode(), precision, buf, len(buf), C.byref(nt), C.byref(ne)))
    return buf.value.decode(), nt.value, ne.value


def tree_schedule(newick: str):
    """-> (n_slots, ops [(parent,left,right,llen,rlen)], edges [(distal,proximal,length)], tip labels)"""
    _, nt, ne = parse_tree(newick)
    ops = (capi.ClvOp * (3 * nt))()
    edges = (capi.EdgeDesc * ne)()
    labels = C.create_string_buffer(len(newick) + nt + 16)
    ns, no, ne2 = C.c_uint32(), C.c_uint32(), C.c_uint32()
    _check(lib().epa_host_tree_schedule(newick.encode(), C.byref(ns), ops, len(ops), C.byref(no), edges, ne, C.byref(ne2),
                                        labels, len(labels)))
    return (ns.value, [(o.parent, o.left, o.right, o.left_length, o.right_length) for o in ops[:no.value]],
            [(e.distal, e.proximal, e.length) for e in edges[:ne2.value]], labels.value.decode().split("\n")[:-1])


def parse_model(model: str):
    st, rc = C.c_uint32(), C.c_uint32()
    rates, weights, freqs, ev = np.zeros(8), np.zeros(8), np.zeros(20), np.zeros(20)
    V, Vi = np.zeros(400), np.zeros(400)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    _check(lib().epa_host_parse_model(model.encode(), C.byref(st), C.byref(rc), dp(rates), dp(weights), dp(freqs), dp(ev),
                                      dp(V), dp(Vi)))
    S, R = st.value, rc.value
    return dict(states=S, rate_cats=R, rates=rates[:R], weights=weights[:R], freqs=freqs[:S], eigenvals=ev[:S],
                eigenvecs=V[:S * S], inv_eigenvecs=Vi[:S * S])
