#!/usr/bin/env python
"""Random-sector gather ceiling vs buffer size (L2 capacity / TLB reach characterisation).
Run plain for rates; under `ncu --metrics lts__t_sector_hit_rate.pct,dram__bytes_read.sum,gpu__time_duration.sum -k regex:probe_kernel`
for the L2 hit rate of the same launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sbwt_b200 as S
n = 1 << 27
for bytes_per in (32, 64):
    for mb in (16, 32, 48, 56, 64, 72, 80, 96, 112, 128, 160, 256, 512, 1024, 2048, 4096, 8192, 16384):
        r = S.sector_probe(0, mb << 20, n, bytes_per, iters=2)
        print(f"bytes_per_load={bytes_per} buffer_MB={mb:6d} loads/s={r / 1e9:8.2f}G  GB/s={r * bytes_per / 1e9:8.1f}", flush=True)
