#!/usr/bin/env python
"""Quick walk-kernel timing: python tools/quick_walk.py [workload] [reads] [tp or -1] -> one line (CUDA-event walk time,
parity of the first reads against the oracle). Knobs come from the SBWT_B200_* environment."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, oracle
import sbwt_b200 as S
from sbwt_b200.testing import synth

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
tp = int(sys.argv[3]) if len(sys.argv) > 3 else -1
w = bench.WORKLOADS[name]
path, ref = bench.ensure_index(name, w)
rp = os.path.join(bench.CACHE, f"reads_{name}_{n_reads}.npy")  # (several variants are timed in one visit: generate once)
if os.path.exists(rp):
    reads = np.load(rp)
else:
    reads = synth.sample_reads(ref, n_reads, 150, 0.5, seed=43, both_strands=w["rc"])
    np.save(rp, reads)
a, off = synth.matrix_to_batch(reads)
idx = S.Index(path)
if tp >= 0:
    idx.set_table_length(tp)
ses = S.Session(idx, a.size, n_reads)
n_out = ses.count_outputs(off)
d_a, d_off = torch.from_numpy(a).cuda(), torch.from_numpy(off).cuda()
d_out = torch.empty(n_out, dtype=torch.int64, device="cuda")
out32 = os.environ.get("QUICK_OUT32") == "1"  # time the int32-result kernel (what the host pipeline runs on narrow indexes)
d_out32 = torch.empty(n_out, dtype=torch.int32, device="cuda") if out32 else None
mode = S.MODE_STREAMING if w["streaming"] else S.MODE_SEARCH
st = ses.query_device_counted(d_a.data_ptr(), d_off.data_ptr(), n_reads, a.size, mode, d_out.data_ptr(), n_out)
m = min(3000, n_reads)
want = oracle.OracleIndex(path).query_batch(a[: m * 150], off[: m + 1], streaming=w["streaming"])
ok = bool(np.array_equal(d_out[: want.size].cpu().numpy(), want))
chk = int(d_out.sum().item())
ses.set_timing(True)
ts, ps = [], []
for i in range(6):
    if out32:
        ses.query_device_i32(d_a.data_ptr(), d_off.data_ptr(), n_reads, a.size, mode, d_out32.data_ptr(), n_out)
    else:
        ses.query_device(d_a.data_ptr(), d_off.data_ptr(), n_reads, a.size, mode, d_out.data_ptr(), n_out)
    p_ms, w_ms = ses.last_timing()
    ts.append(w_ms); ps.append(p_ms)
ms = float(np.median(ts[2:]))
assert os.environ.get("QUICK_NOSTORE") or int((d_out32 if out32 else d_out).sum(dtype=torch.int64).item()) == chk
print(f"{name} reads={n_reads} tp={idx.table_length} parity={'OK' if ok else 'FAIL'} walk_ms={ms:.3f} prep_ms={np.median(ps[2:]):.3f} lookups/s={n_out / ms / 1e6:.2f}G "
      f"sectors/s={st.index_sectors / ms / 1e6:.1f}G sectors={st.index_sectors} rank_ops={st.rank_ops} hits={st.hits} checksum={chk} "
      f"env={ {k: v for k, v in os.environ.items() if k.startswith('SBWT_B200')} }", flush=True)
