#!/usr/bin/env python
"""Split a k_perceive_rows capture (.ncu-rep with --import-source on) by how often its SASS instructions ran: once per
warp (the thread-per-ant phases A and C), once per chunk of 4 ants (the chunk loop), in between (the rock channel and
other conditional paths).  Prints instructions and warp stall samples per part and the top stall sites.
usage: ncu_phase_split.py <file.ncu-rep> [chunks per warp = 8]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
cpw = int(sys.argv[2]) if len(sys.argv) > 2 else 8
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iE, iS = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
cols = {k: hdr.index(k) for k in ("stall_long_sb", "stall_wait", "stall_short_sb", "stall_selected", "stall_not_selected",
                                  "stall_math", "stall_branch_resolving", "stall_no_inst") if k in hdr}
ins = []
for r in rows[2:]:
    try:
        ins.append((r[1].strip(), int(r[iE]), int(r[iS]), r))
    except (ValueError, IndexError):
        pass
warps = max(e for _, e, _, _ in ins[:40])          # the kernel's first instructions run once per warp
bins = collections.OrderedDict((k, collections.Counter()) for k in ("phases A/C (once per warp)", "conditional paths", "chunk loop"))
for t, e, s, r in ins:
    k = "phases A/C (once per warp)" if e <= warps * 1.03 else ("chunk loop" if e >= warps * cpw * 0.9 else "conditional paths")
    bins[k]["instructions"] += e
    bins[k]["stall samples"] += s
    for c, i in cols.items():
        try:
            bins[k][c] += int(r[i] or 0)
        except ValueError:
            pass
ti = sum(b["instructions"] for b in bins.values()); ts = sum(b["stall samples"] for b in bins.values())
print("%s: %d warps, %.1f warp instructions per ant, %d stall samples" % (rep.split("/")[-1], warps, ti / (warps * 32.0), ts))
for k, b in bins.items():
    print("  %-28s %5.1f %% of the instructions, %5.1f %% of the stall samples (%s)" % (
        k, 100.0 * b["instructions"] / ti, 100.0 * b["stall samples"] / max(ts, 1),
        ", ".join("%s %d" % (c.replace("stall_", ""), b[c]) for c in cols if b[c])))
print("  top stall sites:")
for t, e, s, r in sorted(ins, key=lambda x: -x[2])[:6]:
    print("    %5.1f %%  x%-6d %s" % (100.0 * s / max(ts, 1), e, t[:70]))
