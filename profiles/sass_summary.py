#!/usr/bin/env python
"""profiles/sass_summary.py LIB.so > profiles/rNN_sass_tma.txt -- data-movement mnemonics per kernel of the built
library (cuobjdump -sass runs without a GPU).  TMA bulk copies show up as UBLKCP, mbarrier operations as SYNCS.*, the
async-proxy fence as FENCE.VIEW.ASYNC."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "dumphfdl_b200/libhfdl_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout.splitlines()
pat = re.compile(r"\b(UBLKCP[.\w]*|SYNCS[.\w]*|UTMALDG[.\w]*|LDGSTS[.\w]*|LDG\.E[.\w]*|STG\.E[.\w]*|LDS[.\w]*|STS[.\w]*|SHFL[.\w]*|MUFU[.\w]*|FFMA2?|FENCE[.\w]*|BAR[.\w]*|NANOSLEEP)\b")
cur, stats, keep = None, collections.OrderedDict(), collections.OrderedDict()
for l in txt:
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
        stats[cur], keep[cur] = collections.Counter(), []
        continue
    if cur is None:
        continue
    m = pat.search(l)
    if m:
        op = m.group(1)
        key = op.split(".")[0] if op.startswith(("LDS", "STS", "SHFL", "MUFU", "BAR", "FENCE")) else op
        stats[cur][key] += 1
        if op.startswith(("UBLKCP", "SYNCS", "FENCE")):
            keep[cur].append(l.strip())
print("# SASS of %s (cuobjdump -sass, sm_100a): data-movement mnemonics per kernel" % lib)
print("# TMA bulk copies = UBLKCP, mbarrier = SYNCS.*, async-proxy fence = FENCE.VIEW.ASYNC\n")
for k, c in stats.items():
    if not c:
        continue
    print(k)
    print("    " + ", ".join("%s x%d" % (a, b) for a, b in sorted(c.items())))
    for l in keep[k]:
        print("        " + l)
    print()
