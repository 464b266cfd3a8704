"""Dev: a few fwd(+bwd) launches of the fused contact kernels for ncu captures."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import torch
import nimblephysics_b200 as nb
from tests.util import contact_inputs, load_raw
name = os.environ.get("MODEL", "atlas_ground"); B = int(os.environ.get("B", "2048"))
raw = load_raw(name); world = nb.World.from_raw(raw)
s, a = contact_inputs(raw, name, B, seed=100)
s = torch.tensor(s, device="cuda"); a = torch.tensor(a, device="cuda"); g = torch.randn(B, 2 * raw.ndof, device="cuda")
for k in range(3):
    nb.reset_contact_cache(world)
    x = s.clone().requires_grad_(True); u = a.clone().requires_grad_(True)
    out = nb.timestep(world, x, u); out.backward(g)
torch.cuda.synchronize()
print("done", nb.check_contact_status(world))
