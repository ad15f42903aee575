#!/usr/bin/env python
"""Static view of a kernel's SASS: total instructions and the backward branches (loops) with their spans.
usage: sass_loops.py <lib.so> <substring of the mangled kernel name>"""
import re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None; ins = []
for l in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", l)
    if m:
        cur = m.group(1); continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
print("instructions", len(ins))
for a, t in ins:
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a and (a - int(m.group(1), 16)) // 16 > 100:
        print(hex(int(m.group(1), 16)), "->", hex(a), "span", (a - int(m.group(1), 16)) // 16, t)
