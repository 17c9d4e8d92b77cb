#!/usr/bin/env python3
"""Per-kernel counts of the Blackwell-native SASS mnemonics in libsatools_hifigan.so (cuobjdump -sass; no GPU needed):
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk,
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops; HMMA would be the legacy mma.sync path (there is none).

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sa-toolkit_b200", "csrc", "libsatools_hifigan.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "FFMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for o in OPS:
                if op.startswith(o):
                    counts[cur][o] += 1
    print(f"SASS of {os.path.relpath(LIB, ROOT)} (sm_100a), instructions per kernel")
    print(f"{'kernel':78s} {'instr':>7s} " + " ".join(f"{o:>8s}" for o in OPS))
    tot = collections.Counter()
    for fn in order:
        c = counts[fn]
        name = re.sub(r"\(.*", "", demangle(fn)).replace("sa::tc::", "").replace("sa::", "").replace("(anonymous namespace)::", "")
        print(f"{name[:78]:78s} {c['_total']:7d} " + " ".join(f"{c[o]:8d}" for o in OPS))
        tot.update(c)
    print(f"{'TOTAL':78s} {tot['_total']:7d} " + " ".join(f"{tot[o]:8d}" for o in OPS))


if __name__ == "__main__":
    main()
