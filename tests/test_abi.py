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
