#!/usr/bin/env python
"""Round-2 profile summaries: turns the ncu CSV exports fetched into gpurun_out/ (scripts/profile_r02.sh) and the in-tree libnb2.so
into the tracked files under profiles/ (r02_contact_full.md, r02_step_full.md, r02_*launches*.txt, r02_sass_summary.md, dram_traffic.json)."""
import collections, csv, json, os, re, subprocess, sys
G, P = "gpurun_out", "profiles"
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
STALLS = ["barrier", "wait", "no_instruction", "long_scoreboard", "short_scoreboard", "branch_resolving", "math_pipe_throttle", "mio_throttle", "not_selected", "dispatch_stall"]


def table(raw_csv, out_md, title, traffic_names):
    raw = list(csv.reader(open(raw_csv)))
    h, units = raw[0], raw[1]
    traffic = {}
    with open(out_md, "w") as f:
        f.write(f"# {title}\n\n")
        for r in raw[2:]:
            name = r[h.index("Kernel Name")]
            short = re.sub(r"\(.*", "", name).replace("<unnamed>::", "")
            f.write(f"## {short}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in h:
                    f.write(f"| {k} | {r[h.index(k)]} | {units[h.index(k)]} |\n")
            for s in STALLS:
                k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
                if k in h:
                    f.write(f"| stall {s} (warps per issue) | {r[h.index(k)]} | |\n")
            f.write("\n")
            def val(k):
                v = float(r[h.index(k)].replace(",", "")); u = units[h.index(k)]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            for tn in traffic_names:
                if tn in short and tn not in traffic:
                    traffic[tn] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    return traffic


def launches(csv_path, out_txt, title):
    rows = [r for r in csv.reader(open(csv_path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    iK, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
    dur = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            dur[re.sub(r"\(.*", "", r[iK]).replace("<unnamed>::", "")].append(float(r[iV].replace(",", "")))
        except Exception:
            pass
    tot = sum(sum(v) for v in dur.values())
    with open(out_txt, "w") as f:
        f.write(f"# {title}\n# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
        for k, v in sorted(dur.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k:70s} launches {len(v):4d}  mean {sum(v)/len(v)/1e3:9.2f} us  share {100*sum(v)/tot:5.1f}%\n")


traffic = {}
traffic.update(table(f"{G}/r02_contact_full_raw.csv", f"{P}/r02_contact_full.md",
                     "ncu --set full --clock-control none, Atlas + ground, B = 8192, second step (cold LCP cache): the five kernels of one fwd+bwd contact step (r02)",
                     ["k_cbuild", "k_csolve<0>", "k_csolve<1>", "k_capply", "k_cstep_bwd"]))
traffic.update(table(f"{G}/r02_step_full_raw.csv", f"{P}/r02_step_full.md",
                     "ncu --set full --clock-control none -k regex:k_step (r02; bench.py --steps 4 --warmup 3 --no-extra, B=4096, contact-free Atlas)",
                     ["k_step_fwd", "k_step_bwd"]))
launches(f"{G}/r02_launches.csv", f"{P}/r02_launches_summary.txt", "ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 20 --warmup 3 --no-extra (r02)")
launches(f"{G}/r02_contact_launches.csv", f"{P}/r02_contact_launches_summary.txt", "ncu --metrics gpu__time_duration.sum, 3 x (forward + backward) Atlas + ground steps, B = 8192 (scripts/dev/one_contact.py)")
for fn in ("r02_launches.csv", "r02_contact_launches.csv"):
    open(f"{P}/{fn}", "w").write(open(f"{G}/{fn}").read())
json.dump(traffic, open(f"{P}/dram_traffic.json", "w"), indent=1)
# SASS opcode histogram per kernel of the in-tree library
sass = subprocess.run(["cuobjdump", "-sass", "nimblephysics_b200/csrc/libnb2.so"], capture_output=True, text=True).stdout
cur, hist = None, {}
for l in sass.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("(anonymous namespace)::", "")
        hist[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m and cur:
        hist[cur][m.group(1).split(".")[0]] += 1
with open(f"{P}/r02_sass_summary.md", "w") as f:
    f.write("# cuobjdump -sass libnb2.so (sm_100a cubin): static instruction count and opcode histogram per kernel (r02)\n\n")
    f.write("UBLKCP = cp.async.bulk (TMA bulk copy), LDGSTS = cp.async, SHFL = warp shuffle, DFMA/DADD/DMUL = fp64 pipe, BAR = block barrier (lockstep phases)\n\n")
    for k, c in sorted(hist.items(), key=lambda kv: -sum(kv[1].values())):
        tot = sum(c.values())
        top = ", ".join(f"{op} {n}" for op, n in c.most_common(14))
        special = ", ".join(f"{op} {c[op]}" for op in ("UBLKCP", "LDGSTS", "SHFL", "BAR", "DFMA", "FFMA", "LDS", "STS", "LDL", "STL", "CALL") if c[op])
        f.write(f"## {k}\n{tot} instructions ({tot*16//1024} KB).  {special}\n\ntop: {top}\n\n")
print(json.dumps(traffic)); print(open(f"{P}/r02_contact_launches_summary.txt").read())
