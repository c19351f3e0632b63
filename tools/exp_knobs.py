#!/usr/bin/env python
"""Knob sweep of the walk kernel at bench size: every line = one setting (env knobs, table length, int32 results),
3 launches, CUDA-event walk times. Run plain for timings, and under
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:walk2_kernel
for the DRAM traffic of the same launches (launch order = print order, 3 per line after 1 counted launch).
usage: python tools/exp_knobs.py [workload] [reads] [set]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import sbwt_b200 as S  # noqa: E402
from sbwt_b200.testing import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
which = sys.argv[3] if len(sys.argv) > 3 else "l2"
w = bench.WORKLOADS[name]
path, ref = bench.ensure_index(name, w)
reads = synth.sample_reads(ref, n_reads, 150, 0.5, seed=43, both_strands=w["rc"])
a, off = synth.matrix_to_batch(reads)
idx = S.Index(path)
tp0 = idx.table_length
ses = S.Session(idx, a.size, n_reads)
n_out = ses.count_outputs(off)
d_a, d_off = torch.from_numpy(a).cuda(), torch.from_numpy(off).cuda()
d_out = torch.empty(n_out, dtype=torch.int64, device="cuda")
mode = S.MODE_STREAMING if w["streaming"] else S.MODE_SEARCH
ses.set_timing(True)
KNOBS = ["SBWT_B200_L2_EVICT_LAST", "SBWT_B200_BLOCKS_PER_SM", "SBWT_B200_L2_PERSIST", "SBWT_B200_DEBUG_NOSTORE", "SBWT_B200_PROBE", "SBWT_B200_L2_FRAC"]
SETS = {
    "l2": [dict(), dict(SBWT_B200_DEBUG_NOSTORE=1), dict(out32=1), dict(SBWT_B200_L2_EVICT_LAST=0), dict(SBWT_B200_L2_PERSIST=1),
           dict(SBWT_B200_L2_PERSIST=1, SBWT_B200_L2_EVICT_LAST=0),
           dict(tp=8), dict(tp=9), dict(tp=11), dict(tp=12), dict(SBWT_B200_BLOCKS_PER_SM=3), dict(SBWT_B200_BLOCKS_PER_SM=5),
           dict(SBWT_B200_BLOCKS_PER_SM=6), dict(SBWT_B200_PROBE=10), dict(SBWT_B200_PROBE=12), dict(SBWT_B200_PROBE=18),
           dict(SBWT_B200_DEBUG_NOSTORE=2)],
    "base": [dict()],
    "tp": [dict(), dict(tp=11), dict(tp=12), dict(tp=13)],
    "tp14": [dict(), dict(tp=12), dict(tp=14)],
    "compact": [dict(), dict(tp=9), dict(tp=11), dict(tp=12), dict(SBWT_B200_BLOCKS_PER_SM=3), dict(out32=1), dict(SBWT_B200_PROBE=14), dict(SBWT_B200_PROBE=20)],
    "frac": [dict()] + [dict(SBWT_B200_L2_EVICT_LAST=m, SBWT_B200_L2_FRAC=f) for m in (2, 3) for f in (0.9, 0.75, 0.6, 0.45)]
            + [dict(SBWT_B200_L2_EVICT_LAST=2, SBWT_B200_L2_FRAC=f, SBWT_B200_DEBUG_NOSTORE=1) for f in (1.0, 0.75, 0.5)],
    "quick": [dict(), dict(SBWT_B200_DEBUG_NOSTORE=1), dict(out32=1), dict(SBWT_B200_L2_EVICT_LAST=0), dict(tp=8), dict(tp=9)],
}
for cfg in SETS[which]:
    for kn in KNOBS:
        os.environ.pop(kn, None)
    for kk, v in cfg.items():
        if kk.startswith("SBWT_"):
            os.environ[kk] = str(v)
    idx.set_table_length(cfg.get("tp", tp0))
    ts = []
    for i in range(3):
        if cfg.get("out32"):
            ses.query_device_i32(d_a.data_ptr(), d_off.data_ptr(), n_reads, a.size, mode, d_out.data_ptr(), n_out)
        else:
            ses.query_device(d_a.data_ptr(), d_off.data_ptr(), n_reads, a.size, mode, d_out.data_ptr(), n_out)
        ts.append(ses.last_timing()[1])
    print(f"{name} reads={n_reads} {cfg} walk_ms={min(ts):8.3f} ({' '.join('%.3f' % t for t in ts)}) "
          f"lookups/s={n_out / min(ts) / 1e6:7.2f}G dev_MB={idx.device_bytes / 1e6:.1f}", flush=True)
