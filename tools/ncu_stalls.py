"""Top stall locations (SASS) and headline pipe / stall ratios of one kernel in an ncu report."""
import csv, subprocess, sys, io

rep = sys.argv[1]
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
want = ("gpu__time_duration.sum", "smsp__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum")
for i, k in enumerate(h):
    if k in want or k.startswith("smsp__average_warps_issue_stalled") and float(v[i] or 0) > 0.2:
        print(f"{k:90s} {u[i]:12s} {v[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h, data = rows[1], rows[2:]
isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
f = lambda x: float(x) if x else 0.0
tot = sum(f(r[isamp]) for r in data)
print("total samples", tot)
for i in sorted(sorted(range(len(data)), key=lambda i: -f(data[i][isamp]))[:n_top]):
    r = data[i]
    print(f"{i:6d} {r[isrc][:64]:64s} {100 * f(r[isamp]) / tot:5.1f}%  exec {r[iex]}  prev: {data[i-1][isrc][:50]}")
