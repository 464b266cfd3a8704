"""Small end-to-end run for compute-sanitizer (memcheck): contact-free fwd+bwd (device + pinned-host paths, partial groups),
contact fwd+bwd, fused rollout."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import nimblephysics_b200 as nb
from tests.util import load_raw, sample_inputs, contact_inputs

raw = load_raw("atlas")
w = nb.World.from_raw(raw); w._contacts_disabled = True
for B in (7, 64, 203):
    s, a, g = sample_inputs(raw, B, seed=B)
    st = torch.tensor(s, device="cuda", requires_grad=True); at = torch.tensor(a, device="cuda", requires_grad=True)
    nb.timestep(w, st, at).backward(torch.tensor(g, device="cuda"))
    dm = nb.device_model_for(w)
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    hs, ha, hg = pin(s), pin(a), pin(g)
    o1, o2, o3 = torch.empty_like(hs).pin_memory(), torch.empty_like(hs).pin_memory(), torch.empty_like(ha).pin_memory()
    dm.forward_host(hs.numpy(), ha.numpy(), True, 0, out=o1.numpy()); dm.backward_host(hg.numpy(), 0, out_state=o2.numpy(), out_action=o3.numpy())
u = torch.tensor(np.random.default_rng(0).uniform(-5, 5, (5, 64, len(raw.action_map))).astype(np.float32), device="cuda", requires_grad=True)
x0 = torch.tensor(sample_inputs(raw, 64, seed=1)[0], device="cuda", requires_grad=True)
nb.rollout_fused(w, x0, u).sum().backward()
for name, B in (("half_cheetah", 40), ("atlas_ground", 24)):
    craw = load_raw(name); cw = nb.World.from_raw(craw)
    cs, ca = contact_inputs(craw, name, B, seed=3)
    st = torch.tensor(cs, device="cuda", requires_grad=True); at = torch.tensor(ca, device="cuda", requires_grad=True)
    out = nb.timestep(cw, st, at); out.sum().backward()
torch.cuda.synchronize()
print("sanitize run finished")
