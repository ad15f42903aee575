#!/usr/bin/env python
"""Hot SASS instructions (by executed count and stall samples) from an .ncu-rep source page."""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iI, iW, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Avg. Threads Executed")
body = []
for r in rows[2:]:
    if len(r) <= max(iW, iI, iT, iS) or r[0] == "Address": break
    try: body.append((int(r[iI] or 0), int(r[iW] or 0), r[iS], r[iT]))
    except ValueError: break
tot = sum(b[0] for b in body); tw = sum(b[1] for b in body)
print("instructions:", len(body), "executed:", tot, "stall samples:", tw)
# opcode histogram
hist = {}
for n, w, s, t in body:
    op = s.split()[0] if not s.startswith("@") else s.split()[1]
    op = op.split(".")[0]
    h = hist.setdefault(op, [0, 0]); h[0] += n; h[1] += w
print("-- by opcode (executed %, stall %)")
for op, (n, w) in sorted(hist.items(), key=lambda kv: -kv[1][0])[:25]:
    print("   %-10s %6.2f%%  %6.2f%%" % (op, 100.0 * n / tot, 100.0 * w / max(tw, 1)))
print("-- top stall instructions")
for idx, (n, w, s, t) in sorted(enumerate(body), key=lambda kv: -kv[1][1])[:topn]:
    print("   #%4d exec=%8d stall=%5.2f%% thr=%s  %s" % (idx, n, 100.0 * w / max(tw, 1), t, s[:90]))
