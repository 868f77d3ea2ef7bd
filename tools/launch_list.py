#!/usr/bin/env python
"""Per-launch table from an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]` log:
mean device time (us) and DRAM MB per kernel name.   python tools/launch_list.py gpurun_out/l.csv"""
import collections
import csv
import sys

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
acc = collections.OrderedDict()
for r in csv.DictReader(rows):
    acc.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0][-60:]})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
by = collections.OrderedDict()
for v in acc.values():
    by.setdefault(v["name"], []).append(v)
tot = sum(v.get("gpu__time_duration.sum", 0.0) for v in acc.values()) or 1.0
print("%-62s %4s %10s %7s %9s %9s" % ("kernel", "n", "mean_us", "share", "rd_MB", "wr_MB"))
for name, ls in sorted(by.items(), key=lambda kv: -sum(x.get("gpu__time_duration.sum", 0.0) for x in kv[1])):
    t = [x.get("gpu__time_duration.sum", 0.0) for x in ls]
    rd = [x.get("dram__bytes_read.sum", 0.0) for x in ls]
    wr = [x.get("dram__bytes_write.sum", 0.0) for x in ls]
    print("%-62s %4d %10.1f %6.1f%% %9.1f %9.1f" % (name, len(ls), sum(t) / len(t) / 1e3, 100 * sum(t) / tot,
                                                 sum(rd) / len(rd) / 1e6, sum(wr) / len(wr) / 1e6))
