"""Key metrics per kernel from an .ncu-rep (raw page) as a markdown table.
usage: python scripts/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
M = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
     ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
     ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
     ("smsp__inst_executed.sum", "warp insts"),
     ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
     ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
     ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
     ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
     ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
     ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_throttle"),
     ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb")]
print("| kernel | " + " | ".join(n for _, n in M) + " |")
print("|---|" + "---|" * len(M))
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0]
    vals = []
    for k, _ in M:
        if k in ix:
            v = r[ix[k]]
            try:
                v = "%.4g" % float(v)
            except ValueError:
                pass
            vals.append(f"{v} {units[ix[k]]}".strip())
        else:
            vals.append("-")
    print(f"| {name} | " + " | ".join(vals) + " |")
