"""Shared-memory hazard check (compute-sanitizer --tool racecheck) of the cooperative-lane step kernels."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import nimblephysics_b200 as nb
from tests.util import load_raw, sample_inputs
for name in ("atlas", "half_cheetah"):
    raw = load_raw(name)
    w = nb.World.from_raw(raw); w._contacts_disabled = True
    for B in (13, 96):
        s, a, g = sample_inputs(raw, B, seed=B)
        st = torch.tensor(s, device="cuda", requires_grad=True); at = torch.tensor(a, device="cuda", requires_grad=True)
        nb.timestep(w, st, at).backward(torch.tensor(g, device="cuda"))
torch.cuda.synchronize()
print("racecheck run finished")
