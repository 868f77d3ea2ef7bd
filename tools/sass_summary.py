#!/usr/bin/env python
"""Static SASS mnemonic counts per kernel of the shipped library (no GPU needed):

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "REDG", "RED", "ATOMG", "LDS", "STS", "LDG", "STG", "FFMA", "DFMA", "IADD3",
        "IMAD", "SHFL", "MUFU", "VIMNMX", "HMMA", "UTCHMMA", "LDTM"]

txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "plenopticam_b200", "libpcb200.so")],
                     stdout=subprocess.PIPE, text=True).stdout
print("cuobjdump -sass plenopticam_b200/libpcb200.so (sm_100a), final round-2 build: instruction mnemonics per kernel (static counts)")
print("UTMALDG = cp.async.bulk.tensor (TMA tensor copy), UBLKCP = cp.async.bulk (1-D TMA bulk copy), SYNCS = mbarrier ops,")
print("LDGSTS = cp.async, RED/REDG = reductions to global (also on peer-mapped addresses); no HMMA / UTC*MMA / LDTM: tensor cores")
print("are not used (none of the stages is a dense contraction).\n")
cur, counts = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                counts[cur][k] += 1
                break
        else:
            base = op.split(".")[0]
            if base in KEYS:
                counts[cur][base] += 1
for name, c in counts.items():
    short = re.sub(r"^_ZN3pcb\d+", "", name)[:72]
    print("%-74s %s" % (short, {k: c[k] for k in KEYS if c[k]}))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("\nlibrary totals:", {k: tot[k] for k in KEYS if tot[k]})
