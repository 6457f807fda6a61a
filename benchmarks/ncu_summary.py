#!/usr/bin/env python
"""Markdown summary of an ncu report (run here, on the CPU box): the metrics the roofline
discussion in DESIGN.md cites plus the stall reasons per issued instruction.
    python benchmarks/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>_ncu.md
"""
import csv
import subprocess
import sys

WANT = ["launch__block_size", "launch__grid_size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[head.index("Kernel Name")]
        print("## %s\n" % name.split("(")[0].replace("void ", ""))
        print("| metric | unit | value |\n|---|---|---|")
        for w in WANT:
            if w in head:
                print("| %s | %s | %s |" % (w, units[head.index(w)], r[head.index(w)]))
        stalls = []
        for i, n in enumerate(head):
            if "issue_stalled" in n and n.endswith("per_issue_active.ratio"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0.2:
                    stalls.append("%s %.2f" % (n.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), v))
        print("\nstalls per issued instruction (>0.2): " + ", ".join(stalls) + "\n")


if __name__ == "__main__":
    main()
