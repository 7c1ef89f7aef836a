#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of counters the design decisions rest on."""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
STALL = "smsp__average_warps_issue_stalled_"
def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    for row in data:
        print("==", row[ki][:90])
        for i, h in enumerate(hdr):
            if h in WANT:
                print(f"   {h:75s} {row[i]:>18s} {units[i]}")
        st = [(float(row[i]), h[len(STALL):-len('_per_issue_active.ratio')]) for i, h in enumerate(hdr) if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and row[i]]
        st.sort(reverse=True)
        print("   stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in st[:7]))
if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
