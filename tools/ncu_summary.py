"""Summarise an .ncu-rep: key raw metrics + instruction mix (run here, no GPU needed)."""
import csv, subprocess, sys
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, vals = rows[0], rows[-1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_tensor", "smsp__inst_issued.avg.per_cycle_active"]
for h, v in zip(hdr, vals):
    if h in keys or h.startswith("smsp__average_warps_issue_stalled") and float(v or 0) > 0.3:
        print(f"{h} = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
for i, r in enumerate(rows):
    if r and r[0] == "Address":
        hdr, data = r, rows[i + 1:]
        break
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
c, s = Counter(), Counter()
for r in data:
    if len(r) <= ie or not (r[ie] or '0').replace('.','').isdigit(): continue
    t = r[ia].split()
    op = t[1] if t[0].startswith("@") else t[0]
    c[op] += int(r[ie] or 0); s[op] += int(r[isamp] or 0)
tot, ts = sum(c.values()), max(1, sum(s.values()))
print("total warp-instructions", tot)
for op, n in c.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
    print(f"  {op:30s} {100 * n / tot:5.1f}% inst  {100 * s[op] / ts:5.1f}% samples")
