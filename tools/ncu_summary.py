#!/usr/bin/env python
"""Summarise an .ncu-rep (from `ncu --set full --import-source on`) without a GPU:

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 14]

Prints, per captured launch: duration, DRAM bytes, throughput percentages, instruction count, issue rate,
then the SASS opcode mix and the warp-stall reasons aggregated over the sampled instructions.
"""
import collections
import csv
import io
import subprocess
import sys

RAW_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed_op_global_red.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct",
]


def run(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 14
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        d = dict(zip(hdr, row))
        print("=" * 100)
        print(d.get("Kernel Name", "?")[:100], " id", d.get("ID"))
        for k in RAW_KEYS:
            if k in d:
                print("  %-72s %-10s %s" % (k, units[hdr.index(k)], d[k]))
    txt = run(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"])
    blocks = txt.split('"Kernel Name",')
    seen = set()
    for blk in blocks[1:]:
        lines = blk.splitlines()
        name = lines[0][:90]
        if name in seen:
            continue
        seen.add(name)
        rows = list(csv.DictReader(lines[1:]))
        if not rows:
            continue
        tot_i = sum(float(r["Instructions Executed"] or 0) for r in rows) or 1.0
        tot_s = sum(float(r["# Samples"] or 0) for r in rows) or 1.0
        op, ops = collections.Counter(), collections.Counter()
        for r in rows:
            s = r["Source"].split()
            o = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
            op[o] += float(r["Instructions Executed"] or 0)
            ops[o] += float(r["# Samples"] or 0)
        print("-" * 100)
        print("SASS", name, " instructions %.3g  samples %d" % (tot_i, tot_s))
        for o, v in op.most_common(top):
            print("  %-12s inst %5.1f%%   stall samples %5.1f%%" % (o, v / tot_i * 100, ops[o] / tot_s * 100))
        keys = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
        st = {k: sum(float(r[k] or 0) for r in rows) for k in keys}
        print("  stalls: " + ", ".join("%s %.1f%%" % (k[6:], v / tot_s * 100)
                                       for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:7]))


if __name__ == "__main__":
    main()
