#!/usr/bin/env python
"""Debug helper: run one workload several times (counted / plain), compare runs with each other and with the oracle."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, oracle
import sbwt_b200 as S
from sbwt_b200.testing import synth

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
tp = int(sys.argv[3]) if len(sys.argv) > 3 else -1
w = bench.WORKLOADS[name]
path, ref = bench.ensure_index(name, w)
reads = synth.sample_reads(ref, n_reads, 150, 0.5, seed=43, both_strands=w["rc"])
a, off = synth.matrix_to_batch(reads)
idx = S.Index(path)
if tp >= 0:
    idx.set_table_length(tp)
ses = S.Session(idx, a.size, n_reads)
n_out = ses.count_outputs(off)
d_a, d_off = torch.from_numpy(a).cuda(), torch.from_numpy(off).cuda()
mode = S.MODE_STREAMING if w["streaming"] else S.MODE_SEARCH
m = min(n_reads, 20000)
want = oracle.OracleIndex(path).query_batch(a[: m * 150], off[: m + 1], streaming=w["streaming"])
prev = None
outs = []
for i in range(5):
    d_out = torch.full((n_out,), -7, dtype=torch.int64, device="cuda")
    if i == 0:
        ses.query_device_counted(d_a.data_ptr(), d_off.data_ptr(), n_reads, a.size, mode, d_out.data_ptr(), n_out)
    else:
        ses.query_device(d_a.data_ptr(), d_off.data_ptr(), n_reads, a.size, mode, d_out.data_ptr(), n_out)
    torch.cuda.synchronize()
    if prev is not None:
        diff = torch.nonzero(prev != d_out).flatten()
        print(f"run {i}: differs from run {i-1} at {diff.numel()} positions; first {diff[:10].tolist()}; unwritten {(d_out == -7).sum().item()}", flush=True)
        for b in diff[:6].tolist():
            print("    ", b, "read", b // 120, "kmer", b % 120, "prev", prev[max(0, b - 2):b + 3].tolist(), "now", d_out[max(0, b - 2):b + 3].tolist())
    prev = d_out
    got = d_out[: want.size].cpu().numpy()
    bad = np.flatnonzero(got != want)
    print(f"run {i} ({'counted' if i == 0 else 'plain'}): mismatches vs oracle {bad.size} of {want.size}", flush=True)
    for b in bad[:12]:
        print(f"   out[{b}] read {b // 120} kmer {b % 120}: got {got[b]} want {want[b]}  (neighbours got {got[max(0,b-2):b+3]} want {want[max(0,b-2):b+3]})")
