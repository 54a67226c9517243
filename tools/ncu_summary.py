#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the small CSV kept under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/rNN_name.csv
"""
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
    "launch__block_size", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    kcol = hdr.index("Kernel Name")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["launch", "kernel", "metric", "unit", "value"])
        for li, r in enumerate(rows[2:]):
            for i, h in enumerate(hdr):
                stall = "issue_stalled" in h and h.endswith("per_issue_active.ratio")
                if h in KEEP or stall:
                    if stall and float(r[i] or 0) < 0.2:
                        continue
                    w.writerow([li, r[kcol], h, units[i], r[i]])


if __name__ == "__main__":
    main()
