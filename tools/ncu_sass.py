#!/usr/bin/env python
"""Per-SASS-instruction executed counts and stall samples from an .ncu-rep (source page).
usage: python tools/ncu_sass.py file.ncu-rep [min_exec_millions]"""
import csv, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in rows[2:])
tot_samp = sum(int(r[ix["# Samples"]]) for r in rows[2:])
print(f"# total warp instructions {tot_inst/1e6:.1f} M, samples {tot_samp}")
for n, r in enumerate(rows[2:]):
    ie = int(r[ix["Instructions Executed"]])
    if ie / 1e6 < thr:
        continue
    print(f"{n:4d} {ie/1e6:8.1f}M thr {float(r[ix['Avg. Threads Executed']]):5.1f} samp {int(r[ix['# Samples']]):6d} "
          f"lsb {r[ix['stall_long_sb']]:>5s} wait {r[ix['stall_wait']]:>5s}  {r[ix['Source']].strip()}")
