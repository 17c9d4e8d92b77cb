#!/usr/bin/env python3
"""Summarise `ncu -i X.ncu-rep --page source --csv` output: per kernel, stall samples by reason, by opcode and the top SASS lines."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
kernels = []
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]; H = rows[i + 1]; j = i + 2; data = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(H): data.append(rows[j])
            j += 1
        kernels.append((name, H, data)); i = j
    else:
        i += 1
for name, H, data in kernels:
    ci = {n: k for k, n in enumerate(H)}
    S = ci['# Samples']
    tot = sum(int(r[S]) for r in data)
    print("==", name[:100], "lines", len(data), "samples", tot)
    stalls = [h for h in H if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {h: sum(int(r[ci[h]]) for r in data) for h in stalls}
    print("  by reason:", [(k, v, f"{100*v/tot:.0f}%") for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]])
    byop = collections.Counter()
    for r in data:
        op = [o for o in r[ci['Source']].split() if not o.startswith('@')]
        byop[op[0] if op else '?'] += int(r[S])
    print("  by opcode:", [(k, v, f"{100*v/tot:.1f}%") for k, v in byop.most_common(22)])
    for r in sorted(data, key=lambda r: -int(r[S]))[:ntop]:
        st = sorted(((h, int(r[ci[h]])) for h in stalls if int(r[ci[h]]) > 0), key=lambda x: -x[1])[:3]
        print("  ", r[S], r[ci['Instructions Executed']], r[ci['Source']].strip()[:80], st)
