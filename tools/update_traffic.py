#!/usr/bin/env python
"""profiles/traffic.json from the ncu captures of one GPU visit (traffic_<workload>.csv: dram__bytes_read.sum,
dram__bytes_write.sum, lts__t_sector_hit_rate.pct, lts__t_sectors_srcunit_tex_op_read.sum of one walk_kernel launch), labelled with the sha of the kernel sources.
usage: python tools/update_traffic.py gpurun_out/r02z "profiles/r02z_dram_traffic.txt" """
import csv, hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d, source = sys.argv[1], sys.argv[2]
sha = hashlib.sha256(b"".join(open(os.path.join(ROOT, "sbwt_b200", "csrc", f), "rb").read() for f in ("walk_kernel.cuh", "device_index.cuh"))).hexdigest()[:12]
path = os.path.join(ROOT, "profiles", "traffic.json")
t = json.load(open(path))
for wl in ("c2", "c3", "c4s", "c5s", "c4", "c5"):
    f = os.path.join(d, f"traffic_{wl}.csv")
    if not os.path.exists(f):
        continue
    rows = list(csv.reader(open(f, errors="replace")))
    hi = [i for i, r in enumerate(rows) if "Metric Name" in r]
    if not hi:
        continue
    h = rows[hi[0]]
    m = {r[h.index("Metric Name")]: float(r[h.index("Metric Value")].replace(",", "")) for r in rows[hi[0] + 1:] if len(r) == len(h)}
    if "dram__bytes_read.sum" not in m:
        continue
    t[wl] = {"bytes": int(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]),
             "source": f"{source} (ncu; L2 sector hit rate {m.get('lts__t_sector_hit_rate.pct', float('nan')):.1f} %)",
             "kernel_src_sha256_12": sha}
    if "lts__t_sectors_srcunit_tex_op_read.sum" in m:  # every sector the kernel requested through L1TEX
        t[wl]["tex_sector_reads"] = int(m["lts__t_sectors_srcunit_tex_op_read.sum"])
    print(wl, t[wl])
json.dump(t, open(path, "w"), indent=1)
