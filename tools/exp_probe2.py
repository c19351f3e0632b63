#!/usr/bin/env python
"""Random-sector gather rate vs buffer size; the L2 fetch granularity comes from SBWT_B200_L2_FETCH (per process).
Under ncu (-k regex:probe_kernel, dram__bytes_read.sum, lts__t_sector_hit_rate.pct) the same launches give the DRAM bytes
each random 32-byte load costs. usage: python tools/exp_probe2.py [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sbwt_b200 as S
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = 1 << 27
for mb in (32, 48, 56, 64, 80, 96, 112, 128, 192, 256, 1024, 8192):
    r = S.sector_probe(0, mb << 20, n, 32, iters=iters)
    print(f"fetch={os.environ.get('SBWT_B200_L2_FETCH', 'default')} buffer_MB={mb:6d} loads/s={r / 1e9:8.2f}G  GB/s={r * 32 / 1e9:8.1f}", flush=True)
