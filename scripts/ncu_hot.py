"""Summarise the ncu source page of one kernel: opcode mix and stall samples of the hot loop.
usage: python scripts/ncu_hot.py report.ncu-rep kernel_substring [frac=0.3] [--list]"""
import csv, subprocess, sys
from collections import Counter
rep, pat = sys.argv[1], sys.argv[2]
frac = float(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith('-') else 0.3
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = dict(name=r[1], rows=[]); blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    if pat not in b["name"]:
        continue
    hdr = b["rows"][0]; ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in b["rows"][1:] if len(r) > ix["Instructions Executed"]]
    ie = [int(r[ix["Instructions Executed"]]) for r in data]
    te = [int(r[ix["Thread Instructions Executed"]]) for r in data]
    smp = [int(r[ix["# Samples"]]) for r in data]
    mx = max(ie)
    print("==", b["name"][:90]); print("total warp insts", sum(ie), "thread insts", sum(te), "samples", sum(smp))
    hot = [i for i in range(len(data)) if ie[i] > frac * mx]
    c = Counter(); cs = Counter()
    for i in hot:
        t = data[i][ix["Source"]].split()
        op = t[1] if t[0].startswith("@") else t[0]
        c[op.split(".")[0]] += 1; cs[op.split(".")[0]] += smp[i]
    print("hot loop: %d static instrs (exec > %.0f%% of max=%d); share of all warp insts %.1f%%, of samples %.1f%%" % (
        len(hot), frac * 100, mx, 100 * sum(ie[i] for i in hot) / sum(ie), 100 * sum(smp[i] for i in hot) / max(1, sum(smp))))
    print("opcode mix:", c.most_common())
    print("samples by opcode:", cs.most_common(12))
    if "--list" in sys.argv:
        for i in hot:
            print("%9d %6d %5.1f  %s" % (ie[i], smp[i], te[i] / max(1, ie[i]), data[i][ix["Source"]].strip()))
