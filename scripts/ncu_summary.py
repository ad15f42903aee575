#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / profiles/ quote."""
import csv, subprocess, sys
rep = sys.argv[1]
pats = sys.argv[2:] or None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "lts__t_sectors.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("== %s" % name[:100])
    for k in KEYS:
        if k in hdr:
            print("   %-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    stalls = [(float(r[i]) if r[i] else 0.0, h) for i, h in enumerate(hdr)
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    for v, h in sorted(stalls, reverse=True)[:8]:
        print("   stall %-62s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    if pats:
        for i, h in enumerate(hdr):
            if any(p in h for p in pats):
                print("   %-70s %s %s" % (h, r[i], units[i]))
    break
