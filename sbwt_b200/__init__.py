"""sbwt_b200 -- B200-native (sm_100a) plain-matrix SBWT k-mer query path.

This module is only the ctypes view of the C ABI declared in include/sbwt_b200.h
(built from sbwt_b200/csrc into sbwt_b200/libsbwt_b200.so). The product is the
shared library; Python is test / bench plumbing. There is no CPU fallback: if the
library is missing, or there is no CUDA device, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# SBWT_B200_LIB: a differently-compiled build of the same library (kernel A/B measurements only)
LIB_PATH = os.environ.get("SBWT_B200_LIB") or os.path.join(HERE, "libsbwt_b200.so")

MODE_SEARCH = 0
MODE_STREAMING = 1
CASE_UPPER = 0
CASE_EXACT = 1
CASE_API = 2  # the direct API on raw bytes: streaming_search's mixed-case rule (include/sbwt_b200.h)

EXPORTS = [
    "sbwt_gpu_last_error", "sbwt_gpu_device_count", "sbwt_gpu_abi_version",
    "sbwt_gpu_index_create", "sbwt_gpu_index_load", "sbwt_gpu_index_destroy",
    "sbwt_gpu_index_k", "sbwt_gpu_index_n_nodes", "sbwt_gpu_index_n_kmers", "sbwt_gpu_index_precalc_k",
    "sbwt_gpu_index_has_streaming_support", "sbwt_gpu_index_device", "sbwt_gpu_index_C",
    "sbwt_gpu_index_device_bytes", "sbwt_gpu_index_edges_only_at_group_starts", "sbwt_gpu_index_compact_layout", "sbwt_gpu_rank",
    "sbwt_gpu_session_create", "sbwt_gpu_session_destroy", "sbwt_gpu_count_outputs",
    "sbwt_gpu_query_host", "sbwt_gpu_query_host_i32", "sbwt_gpu_query_device", "sbwt_gpu_query_device_i32", "sbwt_gpu_search_batch", "sbwt_gpu_streaming_batch",
    "sbwt_gpu_host_alloc", "sbwt_gpu_host_alloc_on", "sbwt_gpu_host_free", "sbwt_gpu_pack_device",
    "sbwt_gpu_query_device_counted", "sbwt_gpu_launch_count", "sbwt_gpu_sector_probe",
    "sbwt_gpu_session_set_timing", "sbwt_gpu_session_last_timing", "sbwt_gpu_index_get_precalc",
    "sbwt_gpu_index_set_table_length", "sbwt_gpu_index_table_length",
    "sbwt_gpu_text_capacity", "sbwt_gpu_format_device", "sbwt_gpu_query_host_text", "sbwt_gpu_widen_i32", "sbwt_gpu_expand_sparse", "sbwt_gpu_session_widen_threads", "sbwt_gpu_query_host_sharded",
    "sbwt_gpu_update_interval_batch", "sbwt_gpu_partial_search_batch", "sbwt_gpu_forward_batch", "sbwt_gpu_contains_batch",
    "sbwt_gpu_get_kmer_batch", "sbwt_gpu_ascii_export_sets", "sbwt_gpu_index_l2_set_aside", "sbwt_gpu_query_host_hits",
]

TEXT_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64)


class Stats(C.Structure):
    _fields_ = [("lookups", C.c_int64), ("hits", C.c_int64), ("rank_ops", C.c_int64),
                ("index_sectors", C.c_int64), ("kernel_launches", C.c_int64)]


_lib = None


def lib():
    """The loaded C-ABI library. Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "or `make -C sbwt_b200/csrc`. The SBWT GPU query path has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
        L.sbwt_gpu_last_error.restype = C.c_char_p
        L.sbwt_gpu_index_create.argtypes = [C.POINTER(vp), vp, i64, i64, i64, vp, vp, i64, i32, C.POINTER(vp)]
        L.sbwt_gpu_index_load.argtypes = [C.c_char_p, i32, C.POINTER(vp)]
        L.sbwt_gpu_index_destroy.argtypes = [vp]
        L.sbwt_gpu_index_destroy.restype = None
        for name in ("k", "n_nodes", "n_kmers", "precalc_k", "device_bytes", "l2_set_aside"):
            f = getattr(L, "sbwt_gpu_index_" + name)
            f.argtypes, f.restype = [vp], i64
        for name in ("has_streaming_support", "device", "edges_only_at_group_starts"):
            f = getattr(L, "sbwt_gpu_index_" + name)
            f.argtypes, f.restype = [vp], i32
        L.sbwt_gpu_index_compact_layout.argtypes = [vp, C.POINTER(C.c_double)]
        L.sbwt_gpu_index_C.argtypes = [vp, vp]
        L.sbwt_gpu_index_C.restype = None
        L.sbwt_gpu_rank.argtypes = [vp, vp, vp, i64, vp]
        L.sbwt_gpu_index_get_precalc.argtypes = [vp, vp]
        L.sbwt_gpu_index_set_table_length.argtypes = [vp, i32]
        L.sbwt_gpu_index_table_length.argtypes = [vp]
        L.sbwt_gpu_session_create.argtypes = [vp, i64, i64, C.POINTER(vp)]
        L.sbwt_gpu_session_destroy.argtypes = [vp]
        L.sbwt_gpu_session_destroy.restype = None
        L.sbwt_gpu_count_outputs.argtypes = [vp, i64, i64]
        L.sbwt_gpu_count_outputs.restype = i64
        L.sbwt_gpu_query_host.argtypes = [vp, vp, vp, i64, i32, i32, vp]
        L.sbwt_gpu_query_device.argtypes = [vp, vp, vp, i64, i64, i32, i32, vp, i64, vp]
        L.sbwt_gpu_query_host_i32.argtypes = [vp, vp, vp, i64, i32, i32, vp]
        L.sbwt_gpu_query_device_i32.argtypes = [vp, vp, vp, i64, i64, i32, i32, vp, i64, vp]
        L.sbwt_gpu_query_device_counted.argtypes = [vp, vp, vp, i64, i64, i32, i32, vp, i64, vp, C.POINTER(Stats)]
        L.sbwt_gpu_search_batch.argtypes = [vp, vp, vp, i64, vp]
        L.sbwt_gpu_streaming_batch.argtypes = [vp, vp, vp, i64, vp]
        L.sbwt_gpu_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
        L.sbwt_gpu_host_free.argtypes = [vp]
        L.sbwt_gpu_host_free.restype = None
        L.sbwt_gpu_pack_device.argtypes = [vp, i64, i32, vp, vp, vp]
        L.sbwt_gpu_session_set_timing.argtypes = [vp, i32]
        L.sbwt_gpu_session_last_timing.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.sbwt_gpu_launch_count.argtypes = [i32]
        L.sbwt_gpu_launch_count.restype = i64
        L.sbwt_gpu_sector_probe.argtypes = [i32, i64, i64, i32, i32, C.POINTER(C.c_double)]
        L.sbwt_gpu_widen_i32.argtypes = [vp, vp, i64, i32]
        L.sbwt_gpu_query_host_sharded.argtypes = [vp, i32, vp, vp, i64, i32, i32, vp]
        L.sbwt_gpu_expand_sparse.argtypes = [vp, vp, vp, i64, vp, i32, i32]
        L.sbwt_gpu_session_widen_threads.argtypes = [vp]
        L.sbwt_gpu_text_capacity.argtypes = [vp, i64, i64]
        L.sbwt_gpu_text_capacity.restype = i64
        L.sbwt_gpu_format_device.argtypes = [vp, vp, i32, vp, i64, vp, i64, vp, vp]
        L.sbwt_gpu_query_host_text.argtypes = [vp, vp, vp, i64, i32, i32, TEXT_SINK, vp, C.POINTER(i64)]
        L.sbwt_gpu_query_host_hits.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp, C.POINTER(i64)]
        L.sbwt_gpu_update_interval_batch.argtypes = [vp, vp, vp, i64, vp, vp]
        L.sbwt_gpu_partial_search_batch.argtypes = [vp, vp, vp, i64, vp, vp, vp]
        L.sbwt_gpu_forward_batch.argtypes = [vp, vp, vp, i64, vp]
        L.sbwt_gpu_contains_batch.argtypes = [vp, vp, vp, i64, vp]
        L.sbwt_gpu_get_kmer_batch.argtypes = [vp, vp, i64, vp]
        L.sbwt_gpu_ascii_export_sets.argtypes = [vp, vp, i64, C.POINTER(i64)]
        _lib = L
    return _lib


class SbwtGpuError(RuntimeError):
    pass


def _check(rc: int) -> None:
    if rc != 0:
        raise SbwtGpuError(lib().sbwt_gpu_last_error().decode())


def device_count() -> int:
    return lib().sbwt_gpu_device_count()


def launch_count(reset: bool = False) -> int:
    return lib().sbwt_gpu_launch_count(int(reset))


def pinned_empty(n: int, dtype) -> np.ndarray:
    """A numpy array in page-locked host memory (sbwt_gpu_host_alloc); freed with the array."""
    dtype = np.dtype(dtype)
    p = C.c_void_p()
    _check(lib().sbwt_gpu_host_alloc(max(1, n * dtype.itemsize), C.byref(p)))
    buf = (C.c_char * max(1, n * dtype.itemsize)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=n)

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            try:
                lib().sbwt_gpu_host_free(self.ptr)
            except Exception:
                pass

    _OWNERS[arr.ctypes.data] = _Owner(p)
    return arr


_OWNERS: dict = {}


def release_pinned(arr: np.ndarray) -> None:
    _OWNERS.pop(arr.ctypes.data, None)


class Index:
    """A plain-matrix SBWT index resident on one GPU (sbwt_gpu_index)."""

    def __init__(self, path: str | None = None, device: int = 0, *, arrays: dict | None = None):
        self._h = C.c_void_p()
        if path is not None:
            _check(lib().sbwt_gpu_index_load(os.fsencode(path), device, C.byref(self._h)))
        else:
            a = arrays
            bits = [np.ascontiguousarray(b, dtype=np.uint64) for b in a["bits"]]
            ptrs = (C.c_void_p * 4)(*[b.ctypes.data for b in bits])
            sgs = a.get("sgs")
            sgs = None if sgs is None else np.ascontiguousarray(sgs, dtype=np.uint64)
            Carr = np.ascontiguousarray(a["C"], dtype=np.int64)
            pre = a.get("precalc")
            pre = None if pre is None or a.get("precalc_k", 0) == 0 else np.ascontiguousarray(pre, dtype=np.int64)  # None: computed on device
            _check(lib().sbwt_gpu_index_create(ptrs, None if sgs is None else sgs.ctypes.data, a["n_nodes"], a["n_kmers"], a["k"],
                                               Carr.ctypes.data, None if pre is None else pre.ctypes.data, a.get("precalc_k", 0),
                                               device, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().sbwt_gpu_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    k = property(lambda s: lib().sbwt_gpu_index_k(s._h))
    n_nodes = property(lambda s: lib().sbwt_gpu_index_n_nodes(s._h))
    n_kmers = property(lambda s: lib().sbwt_gpu_index_n_kmers(s._h))
    precalc_k = property(lambda s: lib().sbwt_gpu_index_precalc_k(s._h))
    has_streaming_support = property(lambda s: bool(lib().sbwt_gpu_index_has_streaming_support(s._h)))
    device = property(lambda s: lib().sbwt_gpu_index_device(s._h))
    device_bytes = property(lambda s: lib().sbwt_gpu_index_device_bytes(s._h))
    l2_set_aside = property(lambda s: lib().sbwt_gpu_index_l2_set_aside(s._h))
    edges_only_at_group_starts = property(lambda s: bool(lib().sbwt_gpu_index_edges_only_at_group_starts(s._h)))

    @property
    def compact_layout(self) -> tuple[bool, float]:
        """(the one-hot layout is in use, fraction of its blocks answered from the classic sectors; -1 = not built)"""
        f = C.c_double(-1.0)
        used = lib().sbwt_gpu_index_compact_layout(self._h, C.byref(f))
        return bool(used), float(f.value)

    @property
    def csector_format(self) -> str | None:
        """'c64' / 'c96' when a compact one-hot layout is in use, else None."""
        return {0: None, 1: "c96", 2: "c64"}[lib().sbwt_gpu_index_compact_layout(self._h, None)]

    @property
    def C_array(self):
        out = np.zeros(4, dtype=np.int64)
        lib().sbwt_gpu_index_C(self._h, out.ctypes.data)
        return out.tolist()

    @property
    def table_length(self) -> int:
        return lib().sbwt_gpu_index_table_length(self._h)

    def set_table_length(self, tp: int) -> None:
        _check(lib().sbwt_gpu_index_set_table_length(self._h, tp))

    def precalc(self) -> np.ndarray:
        out = np.empty((4 ** self.precalc_k if self.precalc_k else 0, 2), dtype=np.int64)
        _check(lib().sbwt_gpu_index_get_precalc(self._h, out.ctypes.data))
        return out

    def rank(self, pos, chars: bytes) -> np.ndarray:
        pos = np.ascontiguousarray(pos, dtype=np.int64)
        assert len(chars) == pos.size
        out = np.empty(pos.size, dtype=np.int64)
        _check(lib().sbwt_gpu_rank(self._h, pos.ctypes.data, chars, pos.size, out.ctypes.data))
        return out


    # ---- the other read-only queries of SBWT.hh, batched (sbwt_gpu_*_batch)

    def update_interval(self, ascii_: np.ndarray, offsets: np.ndarray, l, r) -> tuple[np.ndarray, np.ndarray]:
        """SBWT::update_sbwt_interval on n (string, interval) pairs."""
        ascii_ = np.ascontiguousarray(ascii_, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        l, r = np.array(l, dtype=np.int64), np.array(r, dtype=np.int64)
        _check(lib().sbwt_gpu_update_interval_batch(self._h, ascii_.ctypes.data, offsets.ctypes.data, offsets.size - 1, l.ctypes.data, r.ctypes.data))
        return l, r

    def partial_search(self, ascii_: np.ndarray, offsets: np.ndarray) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        """SBWT::partial_search per string: (l, r, matched length)."""
        ascii_ = np.ascontiguousarray(ascii_, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = offsets.size - 1
        l, r, m = (np.empty(n, dtype=np.int64) for _ in range(3))
        _check(lib().sbwt_gpu_partial_search_batch(self._h, ascii_.ctypes.data, offsets.ctypes.data, n, l.ctypes.data, r.ctypes.data, m.ctypes.data))
        return l, r, m

    def forward(self, nodes, chars: bytes) -> np.ndarray:
        nodes = np.ascontiguousarray(nodes, dtype=np.int64)
        assert len(chars) == nodes.size
        out = np.empty(nodes.size, dtype=np.int64)
        _check(lib().sbwt_gpu_forward_batch(self._h, nodes.ctypes.data, chars, nodes.size, out.ctypes.data))
        return out

    def contains(self, pos, chars: bytes) -> np.ndarray:
        pos = np.ascontiguousarray(pos, dtype=np.int64)
        assert len(chars) == pos.size
        out = np.empty(pos.size, dtype=np.uint8)
        _check(lib().sbwt_gpu_contains_batch(self._h, pos.ctypes.data, chars, pos.size, out.ctypes.data))
        return out

    def get_kmers(self, colex_ranks) -> list[bytes]:
        rk = np.ascontiguousarray(colex_ranks, dtype=np.int64)
        out = np.empty(rk.size * self.k, dtype=np.uint8)
        _check(lib().sbwt_gpu_get_kmer_batch(self._h, rk.ctypes.data, rk.size, out.ctypes.data))
        return [bytes(out[i * self.k:(i + 1) * self.k]) for i in range(rk.size)]

    def ascii_export_sets(self) -> bytes:
        out = np.empty(4 * self.n_nodes + 1, dtype=np.uint8)
        n = C.c_int64(0)
        _check(lib().sbwt_gpu_ascii_export_sets(self._h, out.ctypes.data, out.size, C.byref(n)))
        return bytes(out[:n.value])


class Session:
    """Scratch + streams for batches against one index (sbwt_gpu_session)."""

    def __init__(self, index: Index, max_bases: int, max_reads: int):
        self.index = index
        self._h = C.c_void_p()
        _check(lib().sbwt_gpu_session_create(index._h, max_bases, max_reads, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().sbwt_gpu_session_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def count_outputs(self, offsets: np.ndarray) -> int:
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        return lib().sbwt_gpu_count_outputs(offsets.ctypes.data, offsets.size - 1, self.index.k)

    def query_host(self, ascii_: np.ndarray, offsets: np.ndarray, mode: int, case_mode: int = CASE_UPPER,
                   out: np.ndarray | None = None, n_out: int | None = None) -> np.ndarray:
        """sbwt_gpu_query_host: host buffers in, int64 results out (H2D + pack + walk + D2H). `n_out`: the number of results,
        when the caller already knows it (count_outputs is a pass over the offsets)."""
        assert ascii_.dtype == np.uint8 and ascii_.flags.c_contiguous
        assert offsets.dtype == np.int64 and offsets.flags.c_contiguous
        if n_out is None:
            n_out = self.count_outputs(offsets)
        if out is None:
            out = np.empty(n_out, dtype=np.int64)
        assert out.dtype == np.int64 and out.size >= n_out
        _check(lib().sbwt_gpu_query_host(self._h, ascii_.ctypes.data, offsets.ctypes.data, offsets.size - 1, mode, case_mode,
                                         out.ctypes.data))
        return out[:n_out]

    def widen_threads(self) -> int:
        """Host threads query_host widens int32 wire results with (0: int64 over PCIe, -1: undecided)."""
        return lib().sbwt_gpu_session_widen_threads(self._h)

    def query_host_hits(self, ascii_: np.ndarray, offsets: np.ndarray, mode: int, case_mode: int = CASE_UPPER, *, want_hits: bool = True,
                        mask: np.ndarray | None = None, hits: np.ndarray | None = None,
                        n_out: int | None = None) -> tuple[np.ndarray, np.ndarray | None, int]:
        """sbwt_gpu_query_host_hits: (membership bitmap as uint32 words, found values in order or None, number of hits)."""
        assert ascii_.dtype == np.uint8 and ascii_.flags.c_contiguous and offsets.dtype == np.int64 and offsets.flags.c_contiguous
        if n_out is None:
            n_out = self.count_outputs(offsets)
        if mask is None:
            mask = np.empty((n_out + 31) // 32, dtype=np.uint32)
        if want_hits and hits is None:
            hits = np.empty(max(1, n_out), dtype=np.int32)
        n = C.c_int64(0)
        _check(lib().sbwt_gpu_query_host_hits(self._h, ascii_.ctypes.data, offsets.ctypes.data, offsets.size - 1, mode, case_mode,
                                              mask.ctypes.data, hits.ctypes.data if want_hits else None, C.byref(n)))
        return mask, (hits[: n.value] if want_hits else None), n.value

    def query_host_i32(self, ascii_: np.ndarray, offsets: np.ndarray, mode: int, case_mode: int = CASE_UPPER,
                       out: np.ndarray | None = None, n_out: int | None = None) -> np.ndarray:
        """sbwt_gpu_query_host_i32: the same values as int32 (indexes with fewer than 2^31 columns)."""
        assert ascii_.dtype == np.uint8 and ascii_.flags.c_contiguous
        assert offsets.dtype == np.int64 and offsets.flags.c_contiguous
        if n_out is None:
            n_out = self.count_outputs(offsets)
        if out is None:
            out = np.empty(n_out, dtype=np.int32)
        assert out.dtype == np.int32 and out.size >= n_out
        _check(lib().sbwt_gpu_query_host_i32(self._h, ascii_.ctypes.data, offsets.ctypes.data, offsets.size - 1, mode, case_mode,
                                             out.ctypes.data))
        return out[:n_out]

    def query_device_i32(self, d_ascii: int, d_offsets: int, n_reads: int, n_bases: int, mode: int, d_out: int, n_out: int,
                         stream: int = 0, case_mode: int = CASE_UPPER) -> None:
        _check(lib().sbwt_gpu_query_device_i32(self._h, d_ascii, d_offsets, n_reads, n_bases, mode, case_mode, d_out, n_out, stream))

    def query_device(self, d_ascii: int, d_offsets: int, n_reads: int, n_bases: int, mode: int, d_out: int, n_out: int,
                     stream: int = 0, case_mode: int = CASE_UPPER) -> None:
        """sbwt_gpu_query_device: raw device pointers, asynchronous on `stream`."""
        _check(lib().sbwt_gpu_query_device(self._h, d_ascii, d_offsets, n_reads, n_bases, mode, case_mode, d_out, n_out, stream))

    def query_host_text(self, ascii_: np.ndarray, offsets: np.ndarray, mode: int, case_mode: int = CASE_UPPER,
                        sink=None) -> tuple[bytes | None, int]:
        """sbwt_gpu_query_host_text: the print_vector text of the batch (formatted on the device).
        Returns (text, n_lookups); with `sink(piece: bytes)` given the pieces go there and text is None."""
        assert ascii_.dtype == np.uint8 and ascii_.flags.c_contiguous
        assert offsets.dtype == np.int64 and offsets.flags.c_contiguous
        parts: list[bytes] = []

        def _cb(_user, ptr, n):
            piece = C.string_at(ptr, n)
            (sink or parts.append)(piece)
            return 0

        n = C.c_int64()
        _check(lib().sbwt_gpu_query_host_text(self._h, ascii_.ctypes.data, offsets.ctypes.data, offsets.size - 1, mode, case_mode,
                                              TEXT_SINK(_cb), None, C.byref(n)))
        return (None if sink else b"".join(parts)), n.value

    def text_capacity(self, n_values: int, n_reads: int) -> int:
        return lib().sbwt_gpu_text_capacity(self.index._h, n_values, n_reads)

    def format_device(self, d_vals: int, vals_are_i32: bool, d_offsets: int, n_reads: int, d_text: int, capacity: int,
                      d_text_bytes: int = 0, stream: int = 0) -> None:
        """sbwt_gpu_format_device: raw device pointers, asynchronous on `stream`."""
        _check(lib().sbwt_gpu_format_device(self._h, d_vals, int(vals_are_i32), d_offsets, n_reads, d_text, capacity,
                                            d_text_bytes or None, stream))

    def set_timing(self, enable: bool = True) -> None:
        _check(lib().sbwt_gpu_session_set_timing(self._h, int(enable)))

    def last_timing(self) -> tuple[float, float]:
        """(prep_ms, walk_ms) of the last device-buffer batch: packer+plan, and the walk kernel."""
        a, b = C.c_double(), C.c_double()
        _check(lib().sbwt_gpu_session_last_timing(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def query_device_counted(self, d_ascii: int, d_offsets: int, n_reads: int, n_bases: int, mode: int, d_out: int, n_out: int,
                             stream: int = 0, case_mode: int = CASE_UPPER) -> Stats:
        st = Stats()
        _check(lib().sbwt_gpu_query_device_counted(self._h, d_ascii, d_offsets, n_reads, n_bases, mode, case_mode, d_out, n_out,
                                                   stream, C.byref(st)))
        return st


def widen_i32(values: np.ndarray, threads: int = 4) -> np.ndarray:
    """The host widening step of the 32-bit result wire format alone (no device work)."""
    values = np.ascontiguousarray(values, dtype=np.int32)
    out = np.empty(values.size, dtype=np.int64)
    _check(lib().sbwt_gpu_widen_i32(values.ctypes.data, out.ctypes.data, values.size, threads))
    return out


def query_host_sharded(sessions: list, ascii_: np.ndarray, offsets: np.ndarray, mode: int, case_mode: int = CASE_UPPER,
                       out: np.ndarray | None = None) -> np.ndarray:
    """One batch over several sessions (replicas of one index, normally one per GPU): sbwt_gpu_query_host_sharded."""
    ascii_ = np.ascontiguousarray(ascii_, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n_out = lib().sbwt_gpu_count_outputs(offsets.ctypes.data, offsets.size - 1, sessions[0].index.k)
    if out is None:
        out = np.empty(n_out, dtype=np.int64)
    assert out.dtype == np.int64 and out.size >= n_out
    handles = (C.c_void_p * len(sessions))(*[s._h for s in sessions])
    _check(lib().sbwt_gpu_query_host_sharded(handles, len(sessions), ascii_.ctypes.data, offsets.ctypes.data, offsets.size - 1, mode,
                                            case_mode, out.ctypes.data))
    return out[:n_out]


def expand_sparse(masks: np.ndarray, block_base: np.ndarray, packed: np.ndarray, n: int, dtype=np.int64, threads: int = 4,
                  out: np.ndarray | None = None) -> np.ndarray:
    """Host half of the sparse result wire format (sbwt_gpu_expand_sparse; no device needed)."""
    masks = np.ascontiguousarray(masks, dtype=np.uint32)
    block_base = np.ascontiguousarray(block_base, dtype=np.uint32)
    packed = np.ascontiguousarray(packed, dtype=np.int32)
    packed = np.concatenate([packed, np.zeros(8, dtype=np.int32)])  # the entry reads 8 values at a time
    if out is None:
        out = np.empty(n, dtype=dtype)
    assert out.size >= n and out.dtype in (np.int64, np.int32)
    _check(lib().sbwt_gpu_expand_sparse(masks.ctypes.data, block_base.ctypes.data, packed.ctypes.data, n, out.ctypes.data,
                                       1 if out.dtype == np.int64 else 0, threads))
    return out[:n]


def sector_probe(device: int, buffer_bytes: int, n_loads: int, bytes_per_load: int = 32, iters: int = 3) -> float:
    """Random aligned-sector gather rate: returns loads per second."""
    ms = C.c_double()
    _check(lib().sbwt_gpu_sector_probe(device, buffer_bytes, n_loads, bytes_per_load, iters, C.byref(ms)))
    return n_loads / (ms.value * 1e-3)
