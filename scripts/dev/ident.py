import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import numpy as np, torch
import nimblephysics_b200 as nb
from tests.util import contact_inputs, load_raw
raw = load_raw("atlas_ground"); world = nb.World.from_raw(raw)
for B in (148, 1184):
    s, a = contact_inputs(raw, "atlas_ground", B, seed=100)
    for ident in (False, True):
        ss = np.repeat(s[3:4], B, 0) if ident else s
        aa = np.repeat(a[3:4], B, 0) if ident else a
        st = torch.tensor(ss, device="cuda"); at = torch.tensor(aa, device="cuda")
        ts = []
        for r in range(6):
            nb.reset_contact_cache(world); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            with torch.no_grad(): nb.timestep(world, st, at)
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        stt = world._lcp_cache["status"].cpu().numpy()
        print(f"B={B} identical={ident}: {np.median(ts[1:]):.3f} ms  status0={stt[0]} pgs frac {((stt&8)>0).mean():.2f}")
