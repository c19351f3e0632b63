#!/usr/bin/env python
"""Sweep the walk kernel's runtime knobs (search-table length, L2 policy, resident blocks) on one GPU.
usage: python tools/sweep_walk.py [workload] [reads]   -- prints one line per configuration."""
import itertools
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
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
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
ref_out = None
tps = [int(x) for x in os.environ.get("SWEEP_TP", "8,10,11,12").split(",")]
els = [int(x) for x in os.environ.get("SWEEP_EL", "1,0").split(",")]
bps = [int(x) for x in os.environ.get("SWEEP_BPS", "0,4,6,8").split(",")]
for tp in tps:
    idx.set_table_length(tp)
    st = ses.query_device_counted(d_a.data_ptr(), d_off.data_ptr(), n_reads, a.size, mode, d_out.data_ptr(), n_out)
    if ref_out is None:
        ref_out = d_out.clone()
    assert torch.equal(ref_out, d_out), "results changed with the table length"
    for el, bp in itertools.product(els, bps):
        os.environ["SBWT_B200_L2_EVICT_LAST"] = str(el)
        os.environ["SBWT_B200_BLOCKS_PER_SM"] = str(bp)
        ts = []
        for i in range(5):
            ses.query_device(d_a.data_ptr(), d_off.data_ptr(), n_reads, a.size, mode, d_out.data_ptr(), n_out)
            ts.append(ses.last_timing()[1])
        ms = float(np.median(ts[2:]))
        print(f"tp={tp:2d} evict_last={el} blocks/SM={bp} walk_ms={ms:8.3f} lookups/s={n_out / ms / 1e6:7.2f}G sectors={st.index_sectors / 1e9:6.3f}G "
              f"sectors/s={st.index_sectors / ms / 1e6:7.1f}G rank_ops={st.rank_ops / 1e9:.3f}G table={idx.table_length} dev_MB={idx.device_bytes / 1e6:.1f}", flush=True)
