#!/usr/bin/env python
"""profiles/<tag>.md from the ncu artefacts of scripts/r2_ncu.sh: the launch list of the bench command and the full-set
captures of the two kernels of the step (k_perceive_rows, k_env) at the bench batch (512 envs x 1024 ants)."""
import csv, json, os, subprocess, sys
from collections import defaultdict
tag = sys.argv[1]
BATCH = int(sys.argv[2]) if len(sys.argv) > 2 else 524288
GROUPS = int(sys.argv[3]) if len(sys.argv) > 3 else 4     # ants_rollout splits the batch into env groups: one launch = BATCH / GROUPS ants
lines = ["# Round 2 - %s" % tag, "",
         "All under gpurun on one B200, `--clock-control none`, each command run plain (exit 0) right before its ncu run.", ""]
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
    setup = ("k_pack_f64", "k_pack_u8", "k_unpack", "k_rock_grid_build", "k_tiles_from", "k_occ_stamp", "k_absorb_sweep", "k_hill_mark", "k_plane_walls")
    step_k = {k: v for k, v in d.items() if not k.startswith(setup)}
    tot = sum(sum(v) for v in step_k.values())
    lines += ["## Launch list of `python bench.py --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline --late-start 0`",
              "(`ncu --metrics gpu__time_duration.sum`; cold-cache and serialised: compare shares, not absolute times)", "",
              "| kernel | launches | avg us | share of the step kernels |", "|---|---|---|---|"]
    for k, v in sorted(step_k.items(), key=lambda kv: -sum(kv[1])):
        lines.append("| %s | %d | %.1f | %.1f %% |" % (k, len(v), sum(v) / len(v) / 1000, 100 * sum(v) / tot))
    lines.append("")
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
for kern in ("k_perceive", "k_env", "k_perceive_whole"):
    rep = "gpurun_out/%s_%s.ncu-rep" % (tag, kern)
    if not os.path.exists(rep):
        continue
    ANTS = BATCH if kern.endswith("_whole") else BATCH // GROUPS
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    g = lambda k: r[hdr.index(k)] if k in hdr else "n/a"
    lines += ["## `%s` after 300 steps of the cfg4 shard (512 envs x 1024 ants; this launch: %d ants = %s; `ncu --set full --import-source on`)"
              % (g("Kernel Name")[:60], ANTS, "the whole batch (ANTS_ROLLOUT_GROUPS=1, like bench.py's per-kernel pass)" if ANTS == BATCH
                 else "one of %d env groups of ants_rollout" % GROUPS), "",
              "| metric | value |", "|---|---|"]
    for k in KEYS:
        if k in hdr:
            lines.append("| %s | %s %s |" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    stalls = [(float(r[i]) if r[i] else 0.0, h) for i, h in enumerate(hdr)
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    lines += ["", "Top warp stall reasons (cycles per issued instruction): " +
              ", ".join("%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
                        for v, h in sorted(stalls, reverse=True)[:6]), ""]
    rd = float(g("dram__bytes_read.sum")) * scale[units[hdr.index("dram__bytes_read.sum")]]
    wr = float(g("dram__bytes_write.sum")) * scale[units[hdr.index("dram__bytes_write.sum")]]
    lines += ["DRAM traffic of this launch: %.1f MB read + %.1f MB written = %.0f B per ant; %.0f warp-instructions per ant."
              % (rd / 1e6, wr / 1e6, (rd + wr) / ANTS, float(g("smsp__inst_executed.sum")) / ANTS), ""]
    if kern == "k_perceive_whole":
        json.dump({"kernel": "k_perceive", "tag": "%s_k_perceive_whole (%d ants per launch, after 300 steps)" % (tag, ANTS),
                   "dram_bytes_per_ant": (rd + wr) / ANTS, "ants": ANTS, "dram_bytes_read": rd, "dram_bytes_write": wr},
                  open("profiles/latest_traffic.json", "w"))
open("profiles/%s.md" % tag, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
