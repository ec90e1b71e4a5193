"""Print the key metrics of an `ncu --page raw --csv` export (one block per profiled launch)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]
extra = [h for h in hdr if ("tensor" in h.lower() and "pct" in h) or "tmem" in h.lower() or "utc" in h.lower()]
for r in rows[2:]:
    print("-" * 100)
    for k in keys + extra[:12]:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:75s} {r[i][:70]:>20s} {units[i]}")
