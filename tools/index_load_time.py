#!/usr/bin/env python
"""Index creation time on the device per search-table length: python tools/index_load_time.py [workload] [tp ...]
(file read + sector build + checks + table; the table is then rebuilt alone for each listed length)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import sbwt_b200 as S

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
tps = [int(x) for x in sys.argv[2:]] or [16, 14, 12]
path, _ = bench.ensure_index(name, bench.WORKLOADS[name])
torch.cuda.init()
S.Index(path).close()  # (context, file cache)
t0 = time.perf_counter()
idx = S.Index(path)
torch.cuda.synchronize()
print(f"{name}: index create {1e3 * (time.perf_counter() - t0):.1f} ms (default table length {idx.table_length}, {idx.device_bytes / 1e9:.2f} GB on device)")
for tp in tps:
    t0 = time.perf_counter()
    idx.set_table_length(tp)
    torch.cuda.synchronize()
    print(f"{name}: table of {tp} characters built in {1e3 * (time.perf_counter() - t0):.1f} ms")
idx.close()
