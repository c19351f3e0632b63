#!/usr/bin/env python
"""Host-pipeline sweep of sbwt_gpu_query_host (int64 results, pinned buffers) at bench size: widen threads x chunk size x
D2H piece size. One line per setting (best of 3). usage: python tools/exp_e2e.py [workload] [reads] [set]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import sbwt_b200 as S  # noqa: E402
from sbwt_b200.testing import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
which = sys.argv[3] if len(sys.argv) > 3 else "main"
w = bench.WORKLOADS[name]
path, ref = bench.ensure_index(name, w)
reads = synth.sample_reads(ref, n_reads, 150, 0.5, seed=43, both_strands=w["rc"])
a, off = synth.matrix_to_batch(reads)
idx = S.Index(path)
mode = S.MODE_STREAMING if w["streaming"] else S.MODE_SEARCH
n_out = S.lib().sbwt_gpu_count_outputs(off.ctypes.data, n_reads, idx.k)
h_a, h_off = S.pinned_empty(a.size, np.uint8), S.pinned_empty(off.size, np.int64)
h_a[:], h_off[:] = a, off
h_out = S.pinned_empty(n_out, np.int64)
h_out32 = S.pinned_empty(n_out, np.int32)
ref_out = None

SETS = {
    "main": [(t, c, p) for t in (16, 12, 8, 6) for c in (96, 48, 24) for p in (4, 1, 64)],
    "quick": [(16, 96, 4), (12, 96, 4), (8, 96, 4), (12, 48, 4), (12, 96, 1), (12, 96, 64), (0, 96, 4)],
    # (threads, chunk, piece, wire)
    "wire": [(8, 48, 64, "dense"), (8, 48, 64, "sparse"), (12, 48, 64, "sparse"), (16, 48, 64, "sparse"), (6, 48, 64, "sparse"),
             (8, 96, 64, "sparse"), (8, 24, 64, "sparse"), (8, 12, 64, "sparse")],
}
for cfg in SETS[which]:
    threads, chunk_m, piece_m = cfg[:3]
    wire = cfg[3] if len(cfg) > 3 else os.environ.get("SBWT_B200_WIRE", "sparse")
    os.environ["SBWT_B200_WIRE"] = wire
    os.environ["SBWT_B200_WIDEN_THREADS"] = str(threads)
    os.environ["SBWT_B200_D2H_PIECE"] = str(piece_m << 20)
    chunk = chunk_m * 1_000_000
    ses = S.Session(idx, chunk, chunk // 148 + 16)
    ses.query_host(h_a, h_off, mode, out=h_out)
    if ref_out is None:
        ref_out = h_out.copy()
    else:
        assert np.array_equal(h_out, ref_out), "pipeline setting changed the results"
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        ses.query_host(h_a, h_off, mode, out=h_out)
        ts.append(time.perf_counter() - t0)
    t32 = []
    for _ in range(2):
        t0 = time.perf_counter()
        ses.query_host_i32(h_a, h_off, mode, out=h_out32)
        t32.append(time.perf_counter() - t0)
    print(f"{name} wire={wire:6s} widen_threads={threads:2d} chunk_Mbases={chunk_m:3d} piece_Mvalues={piece_m:3d} int64_ms={min(ts) * 1e3:8.2f} "
          f"({' '.join('%.1f' % (t * 1e3) for t in ts)}) -> {n_out / min(ts) / 1e9:6.2f} G lookups/s ; i32_ms={min(t32) * 1e3:8.2f}", flush=True)
    ses.close()
