"""Per-launch table from an .ncu-rep captured with --set full (diagnostic): python tools/ncu_table.py report.ncu-rep"""
import csv
import subprocess
import sys

M = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "sm__cycles_elapsed.max",
     "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "sm__cycles_active.avg",
     "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
     "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv", "--metrics", ",".join(M)], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h = r[0]
idx = {n: i for i, n in enumerate(h)}
print("kernel | grid | block | regs | us | cyc elapsed | cyc active(avg SM) | warp inst | ipc active | stall long_sb | barrier | short_sb | L2 hit% | dram rd MB | dram wr MB")
for row in r[2:]:
    g = lambda n: row[idx[n]] if n in idx else "-"
    f = lambda n: float(g(n).replace(",", "")) if g(n) not in ("-", "") else float("nan")
    print(f"{g('Kernel Name')[:38]} | {g('launch__grid_size')} | {g('launch__block_size')} | {g('launch__registers_per_thread')} | "
          f"{f('gpu__time_duration.sum'):.1f} | {f('sm__cycles_elapsed.max'):.0f} | {f('sm__cycles_active.avg'):.0f} | {f('smsp__inst_executed.sum'):.0f} | "
          f"{f('sm__inst_executed.avg.per_cycle_active'):.2f} | {f('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'):.2f} | "
          f"{f('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio'):.2f} | {f('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio'):.2f} | "
          f"{f('lts__t_sector_hit_rate.pct'):.0f} | {f('dram__bytes_read.sum'):.2f} | {f('dram__bytes_write.sum'):.2f}")
