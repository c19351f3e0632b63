"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/sbwt_b200.h declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

import sbwt_b200
from conftest import ROOT, golden


def declared_functions():
    src = open(os.path.join(ROOT, "include", "sbwt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sbwt_gpu_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = declared_functions()
    assert len(names) >= 30
    L = ctypes.CDLL(sbwt_b200.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/sbwt_b200.h but not exported"
    assert set(sbwt_b200.EXPORTS) <= set(names)
    assert L.sbwt_gpu_abi_version() == 1


def test_count_outputs_host_logic():
    off = np.array([0, 5, 36, 36, 186, 216], dtype=np.int64)  # lens 5, 31, 0, 150, 30
    L = sbwt_b200.lib()
    assert L.sbwt_gpu_count_outputs(off.ctypes.data, 5, 31) == 0 + 1 + 0 + 120 + 0
    assert L.sbwt_gpu_count_outputs(off.ctypes.data, 5, 6) == 0 + 26 + 0 + 145 + 25
    assert L.sbwt_gpu_count_outputs(off.ctypes.data, 0, 31) == 0
    # large batches are counted on four threads (a pass over the offsets is bound by one core's memory bandwidth)
    rng = np.random.default_rng(5)
    lens = rng.integers(0, 300, size=2_500_001)
    big = np.concatenate([[7], 7 + np.cumsum(lens)]).astype(np.int64)
    for k in (1, 31, 64, 255):
        assert L.sbwt_gpu_count_outputs(big.ctypes.data, lens.size, k) == int(np.maximum(lens - k + 1, 0).sum())


def test_loader_rejects_bad_files(tmp_path):
    good = open(golden("cli_k6", "index.sbwt"), "rb").read()
    cases = {
        "version": good.replace(b"v0.1", b"v9.9"),
        "variant": good.replace(b"plain-matrix", b"plain-mAtrix"),
        "truncated": good[:-9],
        "trailing": good + b"\0",
    }
    want = {"version": "incompatible version", "variant": "plain-matrix", "truncated": "Corrupt", "trailing": "Corrupt"}
    for name, data in cases.items():
        p = tmp_path / f"{name}.sbwt"
        p.write_bytes(data)
        with pytest.raises(sbwt_b200.SbwtGpuError, match=want[name]):
            sbwt_b200.Index(str(p))
    with pytest.raises(sbwt_b200.SbwtGpuError, match="Error opening file"):
        sbwt_b200.Index(str(tmp_path / "missing.sbwt"))


@pytest.mark.skipif(sbwt_b200.device_count() > 0, reason="checks the behaviour WITHOUT a GPU")
def test_no_cpu_fallback_without_a_device():
    with pytest.raises(sbwt_b200.SbwtGpuError, match="no CUDA device available"):
        sbwt_b200.Index(golden("cli_k6", "index.sbwt"))
    with pytest.raises(sbwt_b200.SbwtGpuError, match="no CUDA device"):
        sbwt_b200.sector_probe(0, 1 << 20, 1 << 10)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under sbwt_b200/ may reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sbwt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".hh", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"(import|from)\s+oracle\b|oracle/|liboracle|sbwt_oracle", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_widen_i32_host_step():
    """The host half of the 32-bit result wire format: sign extension of every value (incl. -1, the largest
    column number below 2^31), odd sizes and unaligned destinations, any thread count."""
    rng = np.random.default_rng(7)
    for n in (0, 1, 7, 8, 9, 65536, 65537, 1_000_003):
        v = rng.integers(-1, 2**31 - 1, size=n, dtype=np.int64).astype(np.int32)
        if n > 2:
            v[0], v[-1] = -1, 2**31 - 1
        for threads in (1, 3, 16):
            assert np.array_equal(sbwt_b200.widen_i32(v, threads), v.astype(np.int64))
    # destination not 32-byte aligned
    v = rng.integers(-1, 2**31 - 1, size=100_001, dtype=np.int64).astype(np.int32)
    buf = np.zeros(v.size + 3, dtype=np.int64)
    for shift in (1, 2, 3):
        out = buf[shift:shift + v.size]
        assert sbwt_b200.lib().sbwt_gpu_widen_i32(v.ctypes.data, out.ctypes.data, v.size, 5) == 0
        assert np.array_equal(out, v.astype(np.int64))
    with pytest.raises(sbwt_b200.SbwtGpuError):
        sbwt_b200.widen_i32(v, 0)


def _sparse_pack_reference(v: np.ndarray, rng) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """What sparse_pack_kernel (aux_kernels.cuh) produces for the int32 results v, with the 4096-result blocks placed in
    `packed` in a shuffled order (on the device the order is whatever the atomicAdd gives)."""
    n = v.size
    n_groups, n_blocks = (n + 31) // 32, (n + 4095) // 4096
    hit = np.zeros(n_groups * 32, dtype=bool)
    hit[:n] = v >= 0
    masks = (hit.reshape(n_groups, 32).astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(axis=1).astype(np.uint32)
    counts = [int((v[b * 4096:(b + 1) * 4096] >= 0).sum()) for b in range(n_blocks)]
    order = rng.permutation(n_blocks)
    base = np.zeros(n_blocks, dtype=np.uint32)
    packed = np.zeros(max(1, sum(counts)), dtype=np.int32)
    pos = 0
    for b in order:
        base[b] = pos
        blk = v[b * 4096:(b + 1) * 4096]
        packed[pos:pos + counts[b]] = blk[blk >= 0]
        pos += counts[b]
    return masks, base, packed


def test_expand_sparse_host_step():
    """The host half of the sparse result wire format rebuilds exactly the device's int32 results, as int64 or int32:
    mixed groups, long runs of hits and of misses (the run primitives), ragged tails, unaligned destinations."""
    rng = np.random.default_rng(11)
    for n in (1, 31, 32, 33, 4095, 4096, 4097, 70_000, 300_007):
        for style in ("mixed", "runs", "all_hit", "all_miss"):
            v = rng.integers(0, 2**31 - 1, size=n, dtype=np.int64).astype(np.int32)
            if style == "mixed":
                v[rng.random(n) < 0.5] = -1
            elif style == "runs":  # reads of 120 results, found or absent as a whole
                for s0 in range(0, n, 120):
                    if rng.random() < 0.5:
                        v[s0:s0 + 120] = -1
            elif style == "all_miss":
                v[:] = -1
            masks, base, packed = _sparse_pack_reference(v, rng)
            for threads in (1, 5):
                assert np.array_equal(sbwt_b200.expand_sparse(masks, base, packed, n, np.int64, threads), v.astype(np.int64)), (n, style)
                assert np.array_equal(sbwt_b200.expand_sparse(masks, base, packed, n, np.int32, threads), v), (n, style)
            for shift in (1, 3):
                buf = np.full(n + 4, -7, dtype=np.int64)
                sbwt_b200.expand_sparse(masks, base, packed, n, threads=3, out=buf[shift:shift + n])
                assert np.array_equal(buf[shift:shift + n], v.astype(np.int64)) and buf[shift - 1] == -7 and buf[shift + n] == -7
                buf32 = np.full(n + 4, -7, dtype=np.int32)
                sbwt_b200.expand_sparse(masks, base, packed, n, threads=3, out=buf32[shift:shift + n])
                assert np.array_equal(buf32[shift:shift + n], v) and buf32[shift - 1] == -7 and buf32[shift + n] == -7


def test_cpp_mirror_header_compiles_and_links(tmp_path):
    """sbwt_b200/csrc/SBWT.hh (the C++ mirror of the reference's SBWT<subset_rank_t> surface) and its test driver build
    against the C ABI library here, without a GPU; the driver itself runs in the GPU suite (tests/test_cli.py)."""
    import subprocess
    exe = str(tmp_path / "test_mirror")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_mirror.cpp"),
                        "-o", exe, "-L" + os.path.join(ROOT, "sbwt_b200"), "-lsbwt_b200", "-Wl,-rpath," + os.path.join(ROOT, "sbwt_b200")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # without a device the mirror fails loudly instead of answering from somewhere else
    if sbwt_b200.device_count() == 0:
        run = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "small_k31", "index.sbwt"), str(tmp_path / "o.sbwt"), "ACGT" * 20],
                             capture_output=True, text=True)
        assert run.returncode != 0 and "no CUDA device" in (run.stdout + run.stderr)


def test_sharded_entry_rejects_bad_arguments_without_a_device():
    L = sbwt_b200.lib()
    assert L.sbwt_gpu_query_host_sharded(None, 0, None, None, 0, 0, 0, None) != 0
    assert b"n_sessions" in L.sbwt_gpu_last_error()
