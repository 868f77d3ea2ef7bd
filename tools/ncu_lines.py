#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares from an .ncu-rep (needs -lineinfo + --import-source on).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [kernel-substring] [min_pct]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
seen = set()
cur_file = ""
for blk in txt.split('"File Path",')[1:]:
    lines = blk.splitlines()
    cur_file = lines[0].strip('"')
    fn = lines[1].split(",", 1)[1][:100]
    if want not in fn or (fn, cur_file) in seen:
        continue
    seen.add((fn, cur_file))
    rd = csv.reader(io.StringIO("\n".join(lines[2:])))
    hdr = next(rd)
    i_line, i_src, i_samp, i_inst = 0, 1, hdr.index("# Samples"), hdr.index("Instructions Executed")
    def num(x):
        try:
            return float(x)
        except ValueError:
            return 0.0
    rows = [r for r in rd if len(r) > i_inst and r[i_line]]
    tot_i = sum(num(r[i_inst]) for r in rows) or 1
    tot_s = sum(num(r[i_samp]) for r in rows) or 1
    print("=" * 110)
    print(fn, "|", cur_file.split("/")[-1], "| inst %.3g samples %d" % (tot_i, tot_s))
    for r in rows:
        pi, ps = num(r[i_inst]) / tot_i * 100, num(r[i_samp]) / tot_s * 100
        if pi >= minpct or ps >= minpct:
            print("%5s  inst %5.1f%%  samp %5.1f%%  %s" % (r[i_line], pi, ps, r[i_src][:95]))
