#!/usr/bin/env python
"""Write profiles/<tag>.md from gpurun_out/<tag>_k_perceive.ncu-rep + gpurun_out/<tag>_launches.csv."""
import csv, subprocess, sys, os, json
from collections import defaultdict
tag, title = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else "k_perceive"
rep = "gpurun_out/%s_%s.ncu-rep" % (tag, kern)
lines = ["# %s" % title, "",
         "Command (under gpurun, 1x B200, after the same command exited 0 without ncu): "
         "`python bench.py --envs 128 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline` "
         "(cfg4 maps: 1024x1024, 1024 ants/env, 64 rocks, 128 envs = 131072 ants per launch).", ""]
lc = "gpurun_out/%s_launches.csv" % tag
if os.path.exists(lc):
    rows = list(csv.reader(open(lc)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    hdr = rows[hi]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value')
    d = defaultdict(list)
    for r in rows[hi + 1:]:
        if len(r) > mv:
            try: d[r[kn].split('(')[0].replace('void ', '')].append(float(r[mv].replace(',', '')))
            except ValueError: pass
    step_k = {k: v for k, v in d.items() if not k.startswith(("k_pack", "k_unpack", "k_rock_grid_build", "k_tiles_from", "k_occ_stamp", "k_absorb_sweep", "k_meta", "k_hill_mark"))}
    tot = sum(sum(v) for v in step_k.values())
    lines += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)", "",
              "| kernel | launches | avg us | share of step kernels |", "|---|---|---|---|"]
    for k, v in sorted(step_k.items(), key=lambda kv: -sum(kv[1])):
        lines.append("| %s | %d | %.1f | %.1f %% |" % (k, len(v), sum(v) / len(v) / 1000, 100 * sum(v) / tot))
    lines.append("")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]
g = lambda k: r[hdr.index(k)] if k in hdr else "n/a"
lines += ["## `%s` (`ncu --set full --clock-control none --import-source on`)" % g("Kernel Name")[:70], "",
          "| metric | value |", "|---|---|"]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio"]
for k in KEYS:
    if k in hdr:
        lines.append("| %s | %s %s |" % (k, r[hdr.index(k)], units[hdr.index(k)]))
stalls = [(float(r[i]) if r[i] else 0.0, h) for i, h in enumerate(hdr)
          if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
lines += ["", "Top warp stall reasons (cycles per issued instruction): " +
          ", ".join("%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
                    for v, h in sorted(stalls, reverse=True)[:6]), ""]
try:
    ants = 131072
    rd = float(g("dram__bytes_read.sum")); wr = float(g("dram__bytes_write.sum"))
    ur = units[hdr.index("dram__bytes_read.sum")]; uw = units[hdr.index("dram__bytes_write.sum")]
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
    tb = rd * scale[ur] + wr * scale[uw]
    dur = float(g("gpu__time_duration.sum")) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9}[units[hdr.index("gpu__time_duration.sum")]]
    lines += ["DRAM traffic per launch: %.1f MB = %.0f B per ant (%.0f warp-instructions per ant); "
              "%.2f ns per ant under ncu." % (tb / 1e6, tb / ants, float(g("smsp__inst_executed.sum")) / ants, dur / ants * 1e9), ""]
    json.dump({"kernel": kern, "tag": tag, "dram_bytes_per_ant": tb / ants, "ants": ants},
              open("profiles/%s_traffic.json" % tag, "w"))
except Exception as e:
    lines.append("(traffic summary failed: %s)" % e)
open("profiles/%s.md" % tag, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[-6:]))
