#!/usr/bin/env python
"""Per-source-line instruction / stall totals from an .ncu-rep captured with --import-source on."""
import csv, subprocess, sys, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
agg = collections.OrderedDict()
cur = None
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        iI = hdr.index("Instructions Executed"); iW = hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr is None or len(r) < 4: continue
    if r[0] != "":
        cur = (r[0], r[1]); agg.setdefault(cur, [0, 0, 0])
    if r[2] != "" and cur is not None:
        try:
            agg[cur][0] += int(r[iI] or 0); agg[cur][1] += int(r[iW] or 0); agg[cur][2] += 1
        except (ValueError, IndexError): pass
tot = sum(v[0] for v in agg.values()); tw = sum(v[1] for v in agg.values())
print("executed", tot, "stall samples", tw)
for (ln, src), (n, w, k) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print("%5s %6.2f%% %6.2f%% sass=%4d | %s" % (ln, 100.0 * n / tot, 100.0 * w / max(tw, 1), k, src.strip()[:120]))
