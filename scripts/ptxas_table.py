"""Summarise `nvcc -Xptxas -v` output (python -m moldy_b200.build --force -v 2> log) as a register/spill table."""
import re
import subprocess
import sys

log = open(sys.argv[1]).read()
rows = []
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                     r".*?Used (\d+) registers", log, re.S):
    rows.append((m.group(1), int(m.group(5)), int(m.group(2)), int(m.group(3)), int(m.group(4))))
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
print("| kernel | registers | stack B | spill st B | spill ld B |\n|---|---|---|---|---|")
for (mang, reg, stack, sst, sld), nm in sorted(zip(rows, names), key=lambda t: t[1]):
    nm = re.sub(r"\(.*", "", nm).replace("void ", "")
    print(f"| `{nm}` | {reg} | {stack} | {sst} | {sld} |")
