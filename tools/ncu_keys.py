"""Print a compact set of metrics (time, pipes, issue, stalls, DRAM/L2) from an .ncu-rep (first kernel)."""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
vals = rows[2 + idx]
keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.per_cycle_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active", "dram__bytes_write.sum", "dram__bytes_read.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
for h, u, v in zip(hdr, units, vals):
    if h in keys or ("average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        try:
            if "stalled" in h and float(v) < 0.05: continue
        except ValueError:
            pass
        print(f"{h:88s} {v[:70]} {u}")
