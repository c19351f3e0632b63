#!/usr/bin/env python
"""L2 behaviour experiment: a fixed list of (table length, evict_last, blocks/SM, persist) settings, 3 walk
launches each (prints CUDA-event walk times). Run it plain for timings and under
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:walk_kernel
for the DRAM traffic of the same launches (launch order = print order).
usage: python tools/exp_l2.py [workload] [reads]"""
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
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
w = bench.WORKLOADS[name]
path, ref = bench.ensure_index(name, w)
reads = synth.sample_reads(ref, n_reads, 150, 0.5, seed=43, both_strands=w["rc"])
a, off = synth.matrix_to_batch(reads)
idx = S.Index(path)
ses = S.Session(idx, a.size, n_reads)
n_out = ses.count_outputs(off)
d_a, d_off = torch.from_numpy(a).cuda(), torch.from_numpy(off).cuda()
d_out = torch.empty(n_out, dtype=torch.int64, device="cuda")
mode = S.MODE_STREAMING if w["streaming"] else S.MODE_SEARCH
ses.set_timing(True)
CONFIGS = [  # tp, evict_last, blocks/SM, persist
    (10, 1, 0, 0), (10, 0, 0, 0), (10, 0, 0, 1), (10, 1, 3, 0), (10, 1, 4, 0), (10, 0, 4, 0),
    (8, 1, 0, 0), (9, 1, 0, 0), (11, 1, 0, 0), (11, 0, 0, 1), (12, 1, 0, 0),
]
for tp, el, bp, pe in CONFIGS:
    idx.set_table_length(tp)
    os.environ["SBWT_B200_L2_EVICT_LAST"] = str(el)
    os.environ["SBWT_B200_BLOCKS_PER_SM"] = str(bp)
    os.environ["SBWT_B200_L2_PERSIST"] = str(pe)
    ts = []
    for i in range(3):
        ses.query_device(d_a.data_ptr(), d_off.data_ptr(), n_reads, a.size, mode, d_out.data_ptr(), n_out)
        ts.append(ses.last_timing()[1])
    print(f"{name} tp={tp:2d} evict_last={el} blocks/SM={bp} persist={pe} walk_ms={min(ts):8.3f} ({' '.join('%.3f' % t for t in ts)}) "
          f"lookups/s={n_out / min(ts) / 1e6:7.2f}G dev_MB={idx.device_bytes / 1e6:.1f}", flush=True)
