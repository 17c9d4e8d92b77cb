#!/usr/bin/env python3
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv)
of one forward: per kernel launches, time, share, DRAM bytes.  Usage: launch_summary.py list.csv "title" > summary.txt"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
iid, iname, imet, iunit, ival = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = collections.OrderedDict()
for r in rows[1:]:
    d = per.setdefault(r[iid], {"name": r[iname], "ms": 0.0, "bytes": 0.0})
    v = float(r[ival].replace(",", "")) * scale[r[iunit]]
    if r[imet].startswith("gpu__time"):
        d["ms"] += v
    else:
        d["bytes"] += v
agg = collections.OrderedDict()
for d in per.values():
    name = re.sub(r"\(.*", "", d["name"]).replace("void ", "").replace("sa::tc::", "").replace("sa::", "")
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += d["ms"]; a[2] += d["bytes"]
tot_ms = sum(a[1] for a in agg.values()); tot_b = sum(a[2] for a in agg.values())
print(f"{len(per)} launches = {sys.argv[2] if len(sys.argv) > 2 else 'one forward'}: ncu gpu__time_duration.sum + "
      "dram__bytes_{read,write}.sum per launch (serialised, cold cache)")
print(f"total {tot_ms:.2f} ms, DRAM traffic {tot_b / 1e9:.2f} GB")
for name, (n, ms, b) in agg.items():
    print(f"{name[-44:]:44s} x{n:2d} {ms:7.3f} ms {100 * ms / tot_ms:5.1f}%  {b / 1e9:6.2f} GB  {b / 1e9 / ms if ms else 0:5.2f} TB/s")
