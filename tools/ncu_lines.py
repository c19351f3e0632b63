#!/usr/bin/env python
"""Stall samples and executed instructions of one kernel per SOURCE LINE: joins the SASS rows of an .ncu-rep's source
page with the line table of the same kernel in the built library (nvdisasm -g), by instruction index.
usage: python tools/ncu_lines.py file.ncu-rep mangled-kernel-substring [library.so]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, kern = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sbwt_b200", "libsbwt_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, capture_output=True)
lines = []  # (file, line) per instruction of the kernel
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout.splitlines()
    inside, cur = False, ("?", 0)
    for l in sass:
        if l.startswith("//---") and ".text." in l:
            inside = kern in l
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(out.splitlines()))
ix = {h: i for i, h in enumerate(rows[1])}
body = rows[2:]
print(f"# {len(body)} SASS rows in the report, {len(lines)} instructions in the library's copy of the kernel")
agg = collections.defaultdict(lambda: [0, 0, 0])
for n, r in enumerate(body):
    key = lines[n] if n < len(lines) else ("?", 0)
    a = agg[key]
    a[0] += int(r[ix["# Samples"]]); a[1] += int(r[ix["Instructions Executed"]]); a[2] += 1
tot = sum(a[0] for a in agg.values()); toti = sum(a[1] for a in agg.values())
print(f"# total samples {tot}, warp instructions {toti/1e6:.1f} M")
for key in sorted(agg):
    a = agg[key]
    if a[0] * 200 >= tot or a[1] * 200 >= toti:
        print(f"{key[0]}:{key[1]:<5d} samples {100*a[0]/tot:5.1f}%  inst {100*a[1]/toti:5.1f}%  ({a[2]} SASS)")
