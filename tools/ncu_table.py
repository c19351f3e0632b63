#!/usr/bin/env python
"""Per-launch table of an `ncu --csv --metrics ...` log: python tools/ncu_table.py file.csv [every]"""
import csv, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
every = int(sys.argv[2]) if len(sys.argv) > 2 else 1
h = rows[0]
d = OrderedDict()
for r in rows[1:]:
    rec = dict(zip(h, r))
    d.setdefault(rec['ID'], {})[rec['Metric Name']] = rec['Metric Value']
for i, (k, v) in enumerate(d.items()):
    if i % every == every - 1:
        print(i, ' '.join(f"{a.replace('lts__t_sectors_srcunit_tex_op_', 'l2_').replace('.sum', '')}={float(b):.4g}" for a, b in v.items()))
