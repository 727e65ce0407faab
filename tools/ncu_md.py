"""Markdown table of the key `ncu --set full` metrics of every launch in a report:
   python tools/ncu_md.py <rep.ncu-rep> [name-filter]"""
import csv, subprocess, sys
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_active.avg", "launch__registers_per_thread", "launch__block_size",
    "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
data = [r for r in data if flt in r[hdr.index("Kernel Name")]]
names = [r[hdr.index("Kernel Name")].split("(")[0].replace("ccsd::", "").replace("void ", "") for r in data]
print("| metric | " + " | ".join(f"`{n}`" for n in names) + " |")
print("|---|" + "---|" * len(names))
for k in KEYS:
    if k not in hdr:
        continue
    i = hdr.index(k)
    print(f"| `{k}` ({units[i]}) | " + " | ".join(r[i] for r in data) + " |")
