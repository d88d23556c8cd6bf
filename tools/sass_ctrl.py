#!/usr/bin/env python3
"""Decode the scheduling control fields (stall count, write/read scoreboard, wait mask) of every SASS
instruction of one kernel, from `cuobjdump -sass` of a built library.  Used to find out WHICH load an
instruction that ncu shows stalled on the long scoreboard is really waiting for.
usage: sass_ctrl.py lib.so <substring of the mangled kernel name> [grep-regex]"""
import re, subprocess, sys
lib, key = sys.argv[1], sys.argv[2]
pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
infun = False; rows = []; i = 0
while i < len(out):
    l = out[i]
    if l.strip().startswith("Function :"):
        infun = key in l
    elif infun:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", l)
        if m and i + 1 < len(out):
            m2 = re.match(r"\s*/\* 0x([0-9a-f]{16}) \*/", out[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                c = hi >> 41
                stall = c & 0xf; yld = (c >> 4) & 1; wr = (c >> 5) & 7; rd = (c >> 8) & 7; wait = (c >> 11) & 0x3f
                rows.append((m.group(1), m.group(2).strip(), stall, yld, wr, rd, wait))
                i += 1
    i += 1
for n, (a, t, stall, yld, wr, rd, wait) in enumerate(rows):
    s = f"{n:5d} {a} st{stall:2d} {'Y' if yld else ' '} W{wr if wr != 7 else '-'} R{rd if rd != 7 else '-'} wait[{''.join(str(b) if wait >> b & 1 else '.' for b in range(6))}] {t}"
    if pat is None or pat.search(s): print(s)
