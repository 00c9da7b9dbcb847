"""Launch list of one bench step from `ncu --metrics ... --csv --log-file X.csv`: one markdown row per launch.
usage: python scripts/ncu_launches.py launches.csv [first_id last_id]"""
import csv, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0].isdigit()]
L = OrderedDict()
for r in rows:
    d = L.setdefault(int(r[0]), {"name": r[4].split("(")[0].replace("void ", "")[:48]})
    d[r[12]] = float(r[14].replace(",", ""))
ids = sorted(L)
if len(sys.argv) > 3:
    ids = [i for i in ids if int(sys.argv[2]) <= i <= int(sys.argv[3])]
tot = sum(L[i].get("gpu__time_duration.sum", 0) for i in ids)
print("| # | kernel | time (ms) | FP64 pipe % | issue % | warps active % | DRAM read (MB) | DRAM write (MB) |")
print("|---|---|---|---|---|---|---|---|")
for i in ids:
    d = L[i]
    t = d.get("gpu__time_duration.sum", 0)
    print("| %d | %s | %.3f (%.1f%%) | %.1f | %.1f | %.1f | %.2f | %.2f |" % (
        i, d["name"], t / 1e6, 100 * t / tot, d.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 0),
        d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0), d.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0),
        d.get("dram__bytes_read.sum", 0) / 1e6, d.get("dram__bytes_write.sum", 0) / 1e6))
print("\ntotal of the listed launches: %.3f ms" % (tot / 1e6))
