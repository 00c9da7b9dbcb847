"""Static view of a kernel's hot loop: opcode histogram of the innermost loop that holds the most FP64 work.
usage: python scripts/sass_loop.py obj.o function_substring [--list] [--nth K]
(K-th best loop by DFMA count; loops = backward branches in the SASS of cuobjdump -sass)"""
import re, subprocess, sys
from collections import Counter
obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
sel = [f for f in funcs[1:] if pat in f.split("\n", 1)[0]]
nth = int(sys.argv[sys.argv.index("--nth") + 1]) if "--nth" in sys.argv else 0
for f in sel:
    name = f.split("\n", 1)[0]
    ins = []
    for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", f):
        ins.append((int(m.group(1), 16), m.group(2).strip()))
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:\w+,\s*)?`?\(?\.?L?_?x?_?\w*\)?", t)
        if "BRA" in t:
            m2 = re.search(r"0x([0-9a-f]+)", t)
            if m2:
                tgt = int(m2.group(1), 16)
                if tgt <= a:
                    body = [x for x in ins if tgt <= x[0] <= a]
                    nd = sum(1 for x in body if re.match(r"(@!?U?P\d+\s+)?D(FMA|MUL|ADD)", x[1]))
                    loops.append((nd, -len(body), tgt, a, body))
    mx = max([l[0] for l in loops] or [0])
    loops = [l for l in loops if l[0] >= 0.3 * mx]
    loops.sort(key=lambda l: (-l[1], l[2]))          # shortest first, then by address
    print("==", name[:100])
    if not loops:
        print("no loops"); continue
    nd, nl, tgt, a, body = loops[min(nth, len(loops) - 1)]
    c = Counter()
    for _, t in body:
        tt = t.split()
        op = tt[1] if tt[0].startswith("@") else tt[0]
        c[op.split(".")[0]] += 1
    fp = c["DFMA"] + c["DMUL"] + c["DADD"]
    print("loop 0x%x..0x%x: %d instrs, FP64 %d (DFMA %d DMUL %d DADD %d), other %d" % (tgt, a, len(body), fp, c["DFMA"], c["DMUL"], c["DADD"], len(body) - fp))
    print(c.most_common())
    if "--list" in sys.argv:
        for ad, t in body:
            print("%06x  %s" % (ad, t))
