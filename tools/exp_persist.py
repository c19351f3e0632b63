#!/usr/bin/env python
"""L2 set-aside experiment: walk-kernel time on one workload for several cudaLimitPersistingL2CacheSize values (the index loads
carry an evict_last = persisting policy). usage: python tools/exp_persist.py [workload] [reads] [MB,MB,...]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import sbwt_b200 as S
from sbwt_b200.testing import synth
from cuda import cudart

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
mbs = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "0,32,40,48,64,80").split(",")]
w = bench.WORKLOADS[name]
path, ref = bench.ensure_index(name, w)
rp = os.path.join(bench.CACHE, f"reads_{name}_{n_reads}.npy")
if os.path.exists(rp):
    reads = np.load(rp)
else:
    reads = synth.sample_reads(ref, n_reads, 150, 0.5, seed=43, both_strands=w["rc"])
    np.save(rp, reads)
a, off = synth.matrix_to_batch(reads)
torch.cuda.init()
err, mx = cudart.cudaDeviceGetAttribute(cudart.cudaDeviceAttr.cudaDevAttrMaxPersistingL2CacheSize, 0)
print("max persisting L2 bytes", mx, flush=True)
idx = S.Index(path)
ses = S.Session(idx, a.size, n_reads)
n_out = ses.count_outputs(off)
d_a, d_off = torch.from_numpy(a).cuda(), torch.from_numpy(off).cuda()
d_out = torch.empty(n_out, dtype=torch.int64, device="cuda")
mode = S.MODE_STREAMING if w["streaming"] else S.MODE_SEARCH
ses.set_timing(True)
chk0 = None
for mb in mbs:
    (err,) = cudart.cudaDeviceSetLimit(cudart.cudaLimit.cudaLimitPersistingL2CacheSize, min(mb << 20, mx))
    cudart.cudaCtxResetPersistingL2Cache()
    ts = []
    for i in range(8):
        ses.query_device(d_a.data_ptr(), d_off.data_ptr(), n_reads, a.size, mode, d_out.data_ptr(), n_out)
        ts.append(ses.last_timing()[1])
    chk = int(d_out.sum().item())
    chk0 = chk if chk0 is None else chk0
    print(f"{name} layout={os.environ.get('SBWT_B200_LAYOUT')} lib={os.path.basename(os.environ.get('SBWT_B200_LIB', 'default'))} persist_MB={mb} ({err}) "
          f"walk_ms={np.median(ts[3:]):.3f} all={[round(t, 2) for t in ts]} checksum_same={chk == chk0}", flush=True)
