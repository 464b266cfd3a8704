#!/usr/bin/env python
"""Turn the ncu outputs fetched by scripts/profile_round.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/:
   <tag>_launches.csv (per-launch durations), <tag>_step_full.md (key metrics of the step kernels), dram_traffic.json."""
import csv, json, os, sys, collections
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = "gpurun_out", "profiles"
rows = [r for r in csv.reader(open(f"{G}/{tag}_launches.csv")) if r and not r[0].startswith("==")]
hdr = rows[0]
iK, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
dur = collections.defaultdict(list)
for r in rows[1:]:
    try:
        dur[r[iK].split("(")[0]].append(float(r[iV].replace(",", "")))
    except Exception:
        pass
with open(f"{P}/{tag}_launches_summary.txt", "w") as f:
    tot = sum(sum(v) for v in dur.values())
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 20 --warmup 3 --no-extra ({tag})\n")
    f.write("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
    for k, v in sorted(dur.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"{k:60s} launches {len(v):4d}  mean {sum(v)/len(v)/1e3:9.2f} us  share {100*sum(v)/tot:5.1f}%\n")
raw = list(csv.reader(open(f"{G}/{tag}_step_full_raw.csv")))
h, units = raw[0], raw[1]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
traffic = {}
with open(f"{P}/{tag}_step_full.md", "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on -k regex:k_step ({tag}; bench.py --steps 4 --warmup 5 --no-extra, B=4096)\n\n")
    for r in raw[2:]:
        name = r[h.index("Kernel Name")]
        f.write(f"## {name}\n\n| metric | value | unit |\n|---|---|---|\n")
        for k in keys:
            if k in h:
                f.write(f"| {k} | {r[h.index(k)]} | {units[h.index(k)]} |\n")
        f.write("\n")
        def val(k):
            v = float(r[h.index(k)].replace(",", "")); u = units[h.index(k)]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        short = "k_step_fwd" if "k_step_fwd" in name else "k_step_bwd"
        traffic[short] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
json.dump(traffic, open(f"{P}/dram_traffic.json", "w"), indent=1)
for fn in (f"{tag}_bench.json",):
    if os.path.exists(f"{G}/{fn}"):
        open(f"{P}/{fn}", "w").write(open(f"{G}/{fn}").read())
print(open(f"{P}/{tag}_launches_summary.txt").read()); print(json.dumps(traffic))
