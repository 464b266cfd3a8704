"""Dev: static SASS instruction count of one kernel by source function (innermost / outermost inlined frame)."""
import re, sys, collections, os
dis, kern = sys.argv[1], sys.argv[2]
sys.argv = [sys.argv[0], "x", dis, kern]
SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "nimblephysics_b200", "csrc")
funcs = {}
for fn in os.listdir(SRC):
    if not fn.endswith((".cuh", ".cu", ".h")): continue
    starts = []
    for i, l in enumerate(open(os.path.join(SRC, fn)).read().split("\n")):
        m = re.match(r"^(?:template\s*<[^>]*>\s*)?(?:NB2_HDN|NB2_HDG|NB2_HD|static|__global__|__device__|inline)[^;(]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", l)
        if m and not l.startswith(" "): starts.append((i + 1, m.group(1)))
        elif re.match(r"^k_cstep_\w+\(", l): starts.append((i + 1, l.split("(")[0]))
    funcs[fn] = starts
def func_of(path, line):
    fn = os.path.basename(path); best = "?"
    for s, name in funcs.get(fn, []):
        if s <= line: best = name
        else: break
    return fn.split(".")[0][4:] + ":" + best
inner = collections.Counter(); outer = collections.Counter(); sect = collections.Counter()
pending = []; cur = []; inside = False; section = None
for l in open(dis):
    if l.startswith(".text.") and l.rstrip().endswith(":"):
        inside = kern in l; continue
    if not inside: continue
    m = re.match(r"^(\$?[_A-Za-z][^\s:]*):\s*$", l)
    if m: section = m.group(1)[-60:]
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: pending.append((m.group(1), int(m.group(2)))); continue
    if re.match(r"\s*/\*([0-9a-f]+)\*/", l):
        if pending: cur = pending; pending = []
        if cur:
            inner[func_of(*cur[0])] += 1
            # outermost frame that is not the kernel body itself
            names = [func_of(*c) for c in cur]
            pick = names[0]
            for nme in names:
                if nme.split(":")[1] in ("pinv_psd","classify_once","pgs_solve","lcp_reduce","dantzig_solve","lcp_chain","collide_and_filter","pair_contacts","collide_box_box","collide_box_sphere","build_rows","assemble_A","impulse_response_all","net_wrenches","fk_collision_bodies","chain_up","chain_down","trsv_lower","trsv_lower_T","lcp_valid","fwd_pass1","fwd_pass2","fwd_pass3","bwd_B1","bwd_B2","bwd_B3","bwd_assemble","contact_forward","contact_backward","dz_swap","dz_factor","dz_solve1","dz_append","fwd_load","fwd_store","bwd_load","bwd_store"):
                    pick = nme; break
            outer[pick] += 1
        else: inner["?"] += 1; outer["?"] += 1
tot = sum(inner.values())
print("static instructions:", tot)
print("--- by enclosing routine"); 
for k, v in outer.most_common(30): print("  %-36s %6d" % (k, v))
print("--- by innermost frame")
for k, v in inner.most_common(25): print("  %-36s %6d" % (k, v))
