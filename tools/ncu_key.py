#!/usr/bin/env python
"""Key metrics + hottest SASS lines of one kernel from an .ncu-rep (run here, no GPU needed).

    python tools/ncu_key.py rep.ncu-rep <sites per launch> [n_lines]
"""
import collections
import csv
import subprocess
import sys


def main():
    rep, sites = sys.argv[1], float(sys.argv[2])
    nl = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size"]
    for i, h in enumerate(hdr):
        if h in want:
            print(f"{h} [{units[i]}] {vals[i]}")
        elif "warp_issue_stalled" in h and "per_issue_active" in h:
            try:
                if float(vals[i]) > 0.2:
                    print(f"{h} {vals[i]}")
            except ValueError:
                pass
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
    iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = [(int(r[iSm]), int(r[iE]), r[iS].strip()[:70], n) for n, r in enumerate(rows[rows.index(hdr) + 1:])
            if len(r) >= len(hdr) and r[iE].isdigit()]
    tot, ti = sum(d[0] for d in data), sum(d[1] for d in data)
    print(f"samples {tot}  warp instructions {ti}  thread-instructions per site {ti * 32 / sites:.2f}")
    ops = collections.Counter()
    for smp, e, sx, n in data:
        t = sx.split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).rstrip(";")
        b = op.split(".")[0]
        if b == "IMAD":
            b = ".".join(op.split(".")[:2])
        ops[b] += e
    print("  ".join(f"{k}:{v * 32 / sites:.2f}" for k, v in ops.most_common(24)))
    for d in sorted(data, reverse=True)[:nl]:
        print(f"  line {d[3]:5d} samples {d[0]:6d} ({100 * d[0] / tot:4.1f}%) exec {d[1]:9d}  {d[2]}")


if __name__ == "__main__":
    main()
