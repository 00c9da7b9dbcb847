"""Decode the scheduling control fields (stall count, write/read barrier, wait mask) of a kernel's SASS.
usage: python scripts/sass_ctrl.py sass.txt (output of cuobjdump -sass -fun NAME file) [regex]"""
import re, sys
lines = open(sys.argv[1]).read().splitlines()
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
i = 0
while i < len(lines):
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", lines[i + 1])
        if m2:
            c = int(m2.group(1), 16) >> 41
            st, wb, rb, wt = c & 0xf, (c >> 5) & 7, (c >> 8) & 7, (c >> 11) & 0x3f
            t = m.group(2).strip()
            if pat is None or pat.search(t):
                print("%s %-62s stall=%2d wbar=%s rbar=%s wait=%s" % (m.group(1), t, st, wb if wb != 7 else '-', rb if rb != 7 else '-', format(wt, '06b')))
            i += 2
            continue
    i += 1
