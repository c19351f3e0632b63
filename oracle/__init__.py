"""ctypes front end of the plain-C oracle (oracle/sbwt_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py. Nothing under sbwt_b200/
imports this module; the product has no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "sbwt_ref")


def build(with_ref: bool | None = None) -> None:
    """Compile liboracle.so / sbwt_oracle, and _ref/sbwt_ref when the reference tree is present."""
    subprocess.run(["make", "-s", "-C", HERE], check=True)
    if with_ref is None:
        with_ref = os.path.isdir("/root/reference/include/sbwt")
    if with_ref:
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


class _Index(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int64), ("n_kmers", C.c_int64), ("k", C.c_int64), ("precalc_k", C.c_int64),
        ("C", C.c_int64 * 4),
        ("bits_len", C.c_uint64 * 4), ("bits", C.POINTER(C.c_uint64) * 4),
        ("rs_words", C.c_uint64 * 4), ("rs", C.POINTER(C.c_uint64) * 4),
        ("sgs_len", C.c_uint64), ("sgs", C.POINTER(C.c_uint64)),
        ("n_precalc", C.c_int64), ("precalc", C.POINTER(C.c_int64)),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build(with_ref=False)
        L = C.CDLL(LIB_PATH)
        L.sbwt_oracle_load.argtypes = [C.c_char_p, C.POINTER(_Index), C.c_char_p, C.c_size_t]
        L.sbwt_oracle_load.restype = C.c_int
        L.sbwt_oracle_free.argtypes = [C.POINTER(_Index)]
        L.sbwt_oracle_from_arrays.argtypes = [C.POINTER(_Index), C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64]
        L.sbwt_oracle_rank.argtypes = [C.POINTER(_Index), C.c_int64, C.c_char]
        L.sbwt_oracle_rank.restype = C.c_int64
        L.sbwt_oracle_rank_naive.argtypes = [C.POINTER(_Index), C.c_int64, C.c_char]
        L.sbwt_oracle_rank_naive.restype = C.c_int64
        L.sbwt_oracle_search.argtypes = [C.POINTER(_Index), C.c_char_p]
        L.sbwt_oracle_search.restype = C.c_int64
        L.sbwt_oracle_query_batch.argtypes = [C.POINTER(_Index), C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        L.sbwt_oracle_query_batch.restype = C.c_int64
        L.sbwt_oracle_search_file.argtypes = [C.POINTER(_Index), C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]
        L.sbwt_oracle_search_file.restype = C.c_int64
        L.sbwt_oracle_contains.argtypes = [C.POINTER(_Index), C.c_int64, C.c_char]
        L.sbwt_oracle_forward.argtypes = [C.POINTER(_Index), C.c_int64, C.c_char]
        L.sbwt_oracle_forward.restype = C.c_int64
        L.sbwt_oracle_partial_search.argtypes = [C.POINTER(_Index), C.c_char_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.sbwt_oracle_partial_search.restype = C.c_int64
        L.sbwt_oracle_update_interval.argtypes = [C.POINTER(_Index), C.c_char_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.sbwt_oracle_update_interval.restype = None
        L.sbwt_oracle_get_kmer.argtypes = [C.POINTER(_Index), C.c_int64, C.c_char_p]
        L.sbwt_oracle_get_kmer.restype = None
        L.sbwt_oracle_export_sets.argtypes = [C.POINTER(_Index), C.c_void_p]
        L.sbwt_oracle_export_sets.restype = C.c_int64
        L.sbwt_oracle_format_line.argtypes = [C.c_void_p, C.c_int64, C.c_char_p]
        L.sbwt_oracle_format_line.restype = C.c_size_t
        _lib = L
    return _lib


class OracleIndex:
    """A loaded plain-matrix index answering queries on the CPU, reference semantics."""

    def __init__(self, path: str | None = None, *, arrays: dict | None = None):
        self._idx = _Index()
        if arrays is not None:  # {"bits": [A, C, G, T] (uint64 words), "sgs": words | None, "n_nodes", "n_kmers", "k", "precalc_k"}
            bits = [np.ascontiguousarray(b, dtype=np.uint64) for b in arrays["bits"]]
            ptrs = (C.c_void_p * 4)(*[b.ctypes.data for b in bits])
            sgs = arrays.get("sgs")
            sgs = None if sgs is None else np.ascontiguousarray(sgs, dtype=np.uint64)
            rc = lib().sbwt_oracle_from_arrays(C.byref(self._idx), ptrs, None if sgs is None else sgs.ctypes.data, arrays["n_nodes"],
                                               arrays.get("n_kmers", 0), arrays["k"], arrays.get("precalc_k", 0))
            if rc != 0:
                raise MemoryError("sbwt_oracle_from_arrays")
            self.path = None
            return
        err = C.create_string_buffer(512)
        if lib().sbwt_oracle_load(path.encode(), C.byref(self._idx), err, 512) != 0:
            raise RuntimeError(err.value.decode())
        self.path = path

    def close(self):
        if self._idx is not None:
            lib().sbwt_oracle_free(C.byref(self._idx))
            self._idx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    n_nodes = property(lambda s: s._idx.n_nodes)
    n_kmers = property(lambda s: s._idx.n_kmers)
    k = property(lambda s: s._idx.k)
    precalc_k = property(lambda s: s._idx.precalc_k)
    has_streaming_support = property(lambda s: s._idx.sgs_len > 0)
    C_array = property(lambda s: list(s._idx.C))

    def rank(self, pos: int, c: str) -> int:
        return lib().sbwt_oracle_rank(C.byref(self._idx), pos, c.encode())

    def rank_naive(self, pos: int, c: str) -> int:
        return lib().sbwt_oracle_rank_naive(C.byref(self._idx), pos, c.encode())

    def search(self, kmer: bytes) -> int:
        assert len(kmer) >= self.k
        return lib().sbwt_oracle_search(C.byref(self._idx), kmer)

    def n_outputs(self, offsets: np.ndarray) -> int:
        lens = np.diff(offsets)
        return int(np.maximum(lens - self.k + 1, 0).sum())

    def query_batch(self, ascii_: np.ndarray, offsets: np.ndarray, streaming: bool) -> np.ndarray:
        """All k-mers of all reads; reads are ascii_[offsets[i]:offsets[i+1]]. int64 results."""
        ascii_ = np.ascontiguousarray(ascii_, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        out = np.empty(self.n_outputs(offsets), dtype=np.int64)
        n = lib().sbwt_oracle_query_batch(C.byref(self._idx), ascii_.ctypes.data, offsets.ctypes.data,
                                          len(offsets) - 1, int(streaming), out.ctypes.data)
        if n == -2:
            raise RuntimeError("Error: streaming search support not built")
        assert n == out.size, (n, out.size)
        return out

    # ---- the other read-only queries (SBWT.hh:369-381, 423-437, 526-537, 701-773)

    def contains(self, pos: int, c: str) -> bool:
        return bool(lib().sbwt_oracle_contains(C.byref(self._idx), pos, c.encode()))

    def forward(self, node: int, c: str) -> int:
        r = lib().sbwt_oracle_forward(C.byref(self._idx), node, c.encode())
        if r == -2:
            raise RuntimeError("Error: Streaming support required for SBWT::forward")
        return r

    def partial_search(self, s: bytes) -> tuple[int, int, int]:
        l, r = C.c_int64(0), C.c_int64(0)
        m = lib().sbwt_oracle_partial_search(C.byref(self._idx), s, len(s), C.byref(l), C.byref(r))
        return l.value, r.value, m

    def update_interval(self, s: bytes, l: int, r: int) -> tuple[int, int]:
        lo, hi = C.c_int64(l), C.c_int64(r)
        lib().sbwt_oracle_update_interval(C.byref(self._idx), s, len(s), C.byref(lo), C.byref(hi))
        return lo.value, hi.value

    def get_kmer(self, colex_rank: int) -> bytes:
        buf = C.create_string_buffer(int(self.k) + 1)
        lib().sbwt_oracle_get_kmer(C.byref(self._idx), colex_rank, buf)
        return buf.raw[: self.k]

    def export_sets(self) -> bytes:
        buf = np.empty(4 * self.n_nodes + 1, dtype=np.uint8)
        n = lib().sbwt_oracle_export_sets(C.byref(self._idx), buf.ctypes.data)
        return bytes(buf[:n])

    def search_file(self, query_path: str, out_path: str) -> int:
        err = C.create_string_buffer(512)
        n = lib().sbwt_oracle_search_file(C.byref(self._idx), query_path.encode(), out_path.encode(), err, 512)
        if n < 0:
            raise RuntimeError(err.value.decode())
        return n


def format_lines(values: np.ndarray, counts) -> bytes:
    """print_vector text for consecutive reads holding counts[i] values each."""
    values = np.ascontiguousarray(values, dtype=np.int64)
    out = bytearray()
    pos = 0
    for c in counts:
        c = int(c)
        buf = C.create_string_buffer(21 * c + 2)
        n = lib().sbwt_oracle_format_line(values[pos:pos + c].ctypes.data if c else None, c, buf)
        out += buf.raw[:n]
        pos += c
    return bytes(out)


def ref_available() -> bool:
    return os.path.exists(REF_BIN)


def ref_run(*args: str, stdin: bytes | None = None) -> subprocess.CompletedProcess:
    """Run oracle/_ref/sbwt_ref (the reference's own classes behind a thin driver)."""
    return subprocess.run([REF_BIN, *args], input=stdin, capture_output=True, check=True)
