"""Dev bench of the fused contact kernels: Atlas + ground / half-cheetah + ground.  Mode A: single step from a fresh cache (cold) or
re-stepped (warm); mode B: a T-step rollout with the LCP cache flowing (the realistic workload), per-step forward time and the
solver-branch statistics of the last step."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import numpy as np
import torch
import nimblephysics_b200 as nb
from tests.util import contact_inputs, load_raw


def stats(world, B):
    c = world._lcp_cache
    st = c["status"].cpu().numpy(); m = c["m"].cpu().numpy()
    return "rows %.1f  shortcircuit %.2f dantzig %.2f dz_fail %.2f pgs %.2f fricdrop %.2f notstd %.2f" % (
        m.mean(), ((st & 1) > 0).mean(), ((st & 2) > 0).mean(), ((st & 4) > 0).mean(), ((st & 8) > 0).mean(), ((st & 16) > 0).mean(), ((st & 64) > 0).mean())


def run(name, B, reps=5, T=8):
    raw = load_raw(name)
    world = nb.World.from_raw(raw)
    s, a = contact_inputs(raw, name, B, seed=100)
    s = torch.tensor(s, device="cuda"); a = torch.tensor(a, device="cuda")
    g = torch.randn(B, 2 * raw.ndof, device="cuda")
    ev = lambda: torch.cuda.Event(enable_timing=True)
    out = []
    for grad in (False, True):
        ts = []
        for r in range(reps + 2):
            nb.reset_contact_cache(world)
            torch.cuda.synchronize()
            e0, e1 = ev(), ev()
            e0.record()
            if grad:
                x = s.clone().requires_grad_(True); u = a.clone().requires_grad_(True)
                nb.timestep(world, x, u).backward(g)
            else:
                with torch.no_grad(): nb.timestep(world, s, a)
            e1.record(); torch.cuda.synchronize()
            if r >= 2: ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        out.append("%s/cold %.3f ms = %.3e/s" % ("fwd+bwd" if grad else "fwd", ms, B / ms * 1e3))
    print(f"{name} B={B}: " + "  ".join(out), "|", stats(world, B), flush=True)
    # rollout: T steps forward (cache flowing), then backward through all of them
    for grad in (False, True):
        ts = []
        for r in range(3):
            nb.reset_contact_cache(world)
            x0 = s.clone().requires_grad_(grad)
            torch.cuda.synchronize()
            e0, e1 = ev(), ev()
            e0.record()
            x = x0
            if grad:
                for t in range(T): x = nb.timestep(world, x, a)
                (x * x).sum().backward()
            else:
                with torch.no_grad():
                    for t in range(T): x = nb.timestep(world, x, a)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        print("   rollout T=%d %s: %.3f ms/step = %.3e world-steps/s | last step: %s  sticky %s" % (
            T, "fwd+bwd" if grad else "fwd", ms / T, B * T / ms * 1e3, stats(world, B), hex(nb.check_contact_status(world))), flush=True)


def prof_dump(tag):
    import ctypes
    from nimblephysics_b200 import _cabi
    L = _cabi.lib()
    buf = (ctypes.c_ulonglong * 64)()
    if not hasattr(L, "nb2_cw_profile_read") or not L.nb2_cw_profile_read(buf, 1):
        return
    names = {0: "load+ABA", 1: "fk", 2: "collide", 3: "rows", 4: "assemble", 5: "chain", 6: "apply+out", 7: "(stage)", 8: "store",
             10: "c.guess", 11: "c.classify0", 12: "c.reduce", 13: "c.dantzig", 14: "c.reduce2", 15: "c.pgs", 16: "c.fricdrop", 17: "c.classify1",
             20: "b.load+B1B2", 21: "b.fk", 22: "b.collide", 23: "b.rows+sets", 24: "b.assemble", 25: "b.Q+pinv", 26: "b.nu+fields", 27: "b.inject",
             28: "b.dual", 29: "(stage)", 30: "b.B3+store"}
    print("  [prof %s] " % tag + "  ".join(f"{names.get(k, k)}={buf[k]/1e6:.0f}M" for k in range(64) if buf[k]), flush=True)


if __name__ == "__main__":
    for name in os.environ.get("MODELS", "atlas_ground,half_cheetah").split(","):
        for B in [int(x) for x in os.environ.get("BS", "4096").split(",")]:
            run(name, B)
            prof_dump(f"{name} {B}")
