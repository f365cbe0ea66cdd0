"""ncu raw CSV (ncu -i x.ncu-rep --page raw --csv) -> compact per-launch summary of the metrics DESIGN.md / bench.py quote."""
import csv
import sys

KEEP = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__cluster_dim_x", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct"]
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr = i
        break
if hdr is None:
    sys.exit("no header")
h = rows[hdr]
units = rows[hdr + 1]
ki = h.index("Kernel Name")
cols = [(h.index(k), k) for k in KEEP if k in h]
n = 0
for r in rows[hdr + 2:]:
    if len(r) != len(h):
        continue
    print("== launch %d: %s" % (n, r[ki][:150]))
    for ci, k in cols:
        print("  %-72s %s %s" % (k, r[ci], units[ci]))
    n += 1
