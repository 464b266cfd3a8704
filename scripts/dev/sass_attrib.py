"""Dev: attribute the per-SASS-instruction counters of an ncu report (--page source --print-source sass --csv) to source FUNCTIONS,
using nvdisasm -gi line info of the same cubin.  usage: sass_attrib.py <ncu_sass.csv> <nvdisasm_gi.txt> <kernel substring>"""
import csv, re, sys, collections, os
sass_csv, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "nimblephysics_b200", "csrc")
# function line ranges per file
funcs = {}
for fn in os.listdir(SRC):
    if not fn.endswith((".cuh", ".cu", ".h")): continue
    lines = open(os.path.join(SRC, fn)).read().split("\n")
    starts = []
    for i, l in enumerate(lines):
        m = re.match(r"^(?:template\s*<[^>]*>\s*)?(?:NB2_HD|static|__global__|__device__|inline)[^;(]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", l)
        if m and not l.startswith(" "): starts.append((i + 1, m.group(1)))
        elif re.match(r"^k_cstep_\w+\(", l): starts.append((i + 1, l.split("(")[0]))
    funcs[fn] = starts
def func_of(path, line):
    fn = os.path.basename(path)
    best = "?"
    for s, name in funcs.get(fn, []):
        if s <= line: best = name
        else: break
    return fn.split(".")[0][4:] + ":" + best
# parse disassembly of the kernel
loc = {}
cur = []; pending = []; inside = False
for l in open(dis):
    if l.startswith(".text.") and l.rstrip().endswith(":"):
        inside = kern in l; continue
    if not inside: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        pending.append((m.group(1), int(m.group(2)))); continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/", l)
    if m:
        if pending: cur = pending; pending = []
        loc[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
ends = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
cand = [i for i in ends if kern.rstrip("E") in rows[i][1]]
sec = cand[0] if cand else ends[0]  # (the CSV holds demangled names, the disassembly mangled ones)
nxt = [i for i in ends if i > sec]
rows = rows[sec:(nxt[0] if nxt else len(rows))]
hdr = rows[1]
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
inoi = hdr.index("stall_no_inst"); iwait = hdr.index("stall_wait"); ilong = hdr.index("stall_long_sb"); ishort = hdr.index("stall_short_sb")
base = int(rows[2][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0, 0])
tot = [0, 0]
for r in rows[2:]:
    off = int(r[ia], 16) - base
    chain = loc.get(off, [])
    name = func_of(*chain[0]) if chain else "?"
    a = agg[name]
    a[0] += int(r[ie]); a[1] += int(r[isamp]); a[2] += int(r[inoi]); a[3] += int(r[iwait]); a[4] += int(r[ilong]); a[5] += int(r[ishort]); a[6] += 1
    tot[0] += int(r[ie]); tot[1] += int(r[isamp])
print("total inst %.3e  samples %d  static instr %d" % (tot[0], tot[1], len(rows) - 2))
print("%-34s %8s %7s %7s | %6s %6s %6s %6s | %s" % ("function (innermost frame)", "inst%", "samp%", "static", "noinst", "wait", "long", "short", "cycles/inst"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%-34s %8.2f %7.2f %7d | %6.2f %6.2f %6.2f %6.2f | %.1f" % (k, 100 * v[0] / tot[0], 100 * v[1] / tot[1], v[6], v[2] / max(v[1], 1), v[3] / max(v[1], 1), v[4] / max(v[1], 1), v[5] / max(v[1], 1), v[1] / max(v[0], 1) * tot[0] / tot[1]))
