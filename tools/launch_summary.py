#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, mean time, share.

    python tools/launch_summary.py gpurun_out/launches.csv [--steps-captured N] > profiles/..._launches_summary.txt
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    cnt, tot = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        name = r[ki]
        if not re.search(r"resample|viewpoints|refocus|select_|contrast|minmax|refine|refi_", name):
            continue
        t = float(r[vi].replace(",", ""))
        if r[ui] == "us":
            t *= 1e3
        elif r[ui] == "ms":
            t *= 1e6
        key = re.sub(r"\(.*", "", name)[:76]
        cnt[key] += 1
        tot[key] += t
    total = sum(tot.values()) or 1.0
    print("per-launch device time (ns), cold-cache and serialised (ncu --metrics gpu__time_duration.sum "
          "--clock-control none); share = of all library kernels")
    print()
    print("%-78s %5s %12s %7s" % ("kernel", "n", "mean_ns", "share"))
    for k in sorted(tot, key=lambda k: -tot[k]):
        print("%-78s %5d %12.0f %6.1f%%" % (k, cnt[k], tot[k] / cnt[k], 100 * tot[k] / total))


if __name__ == "__main__":
    main()
