"""compute-sanitizer run over the round-2 contact kernels (memcheck or racecheck): split forward (build / solve head / solve tail / apply),
fused forward, backward (both instantiations: with and without restitution), large-workspace retry (contact capacity 1), the C rollout
driver with checkpoints, the LCP batch entry, IKMapping and the host entry points.  Small batches: the tools slow the kernels 10-100x."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import nimblephysics_b200 as nb
from tests.util import load_raw, contact_inputs

for name, B, rest in (("half_cheetah", 21, 0.0), ("atlas_ground", 9, 0.0), ("half_cheetah", 10, 0.8)):
    raw = load_raw(name)
    raw.restitution[:] = rest
    w = nb.World.from_raw(raw)
    s, a = contact_inputs(raw, name, B, seed=3)
    if rest:
        s[:, raw.ndof + 1] -= 1.5
    for cap in (None, 1):
        dm = nb.device_model_for(w)
        if cap:
            dm.set_contact_capacity(cap)       # every world with contacts retries in the global pool
        nb.reset_contact_cache(w)
        st = torch.tensor(s, device="cuda", requires_grad=True); at = torch.tensor(a, device="cuda", requires_grad=True)
        x1 = nb.timestep(w, st, at)
        x2 = nb.timestep(w, x1, at)             # warm-started step
        x2.sum().backward()
    nb.reset_contact_cache(w)
    u = torch.tensor(a, device="cuda")[None].repeat(5, 1, 1).requires_grad_(True)
    x0 = torch.tensor(s, device="cuda", requires_grad=True)
    nb.rollout_fused(w, x0, u, checkpoint_every=2).sum().backward()
    print(name, rest, "status", hex(nb.check_contact_status(w)), flush=True)
# joint-limit rows (with and without shapes), self-collision rows inside one tree, round-shape pairs
raw = load_raw("half_cheetah"); raw.limit_enforced[:] = 1; raw.spring[:] = 0
w = nb.World.from_raw(raw)
s, a = contact_inputs(raw, "half_cheetah", 9, seed=9); a *= 0
for k in range(9):
    d = 3 + k % (raw.ndof - 3)
    s[k, d] = raw.pos_hi[d] + 0.004; s[k, raw.ndof + d] = 6.0
st = torch.tensor(s, device="cuda", requires_grad=True); at = torch.tensor(a, device="cuda", requires_grad=True)
nb.timestep(w, st, at).sum().backward()
print("limits status", hex(nb.check_contact_status(w)), flush=True)
from tests.test_contact_emul import _folding_arm_world
w = _folding_arm_world(); n = w.getNumDofs()
S = np.zeros((5, 2 * n), np.float32); S[:, 6] = 2.0944; S[:, 7] = 1.955; S[:, n:] = 0.05
st = torch.tensor(S, device="cuda", requires_grad=True); at = torch.zeros(5, w.getActionSize(), device="cuda", requires_grad=True)
nb.timestep(w, st, at).sum().backward()
print("self-collision status", hex(nb.check_contact_status(w)), "nc", w._lcp_cache["nc"].cpu().tolist(), flush=True)
# IKMapping + host entries + batched LCP
raw = load_raw("half_cheetah"); w = nb.World.from_raw(raw)
nodes = [b for sk in w.skeletons for b in sk._ordered_bodies()]
ik = nb.IKMapping(w); ik.addSpatialBodyNode(nodes[-1]); ik.addSkeletonCOM(w.skeletons[1])
st = torch.tensor(contact_inputs(raw, "half_cheetah", 7, seed=1)[0], device="cuda", requires_grad=True)
(nb.map_to_pos(w, ik, st).sum() + nb.map_to_vel(w, ik, st).sum()).backward()
dm = nb.device_model_for(w)
s, a = contact_inputs(raw, "half_cheetah", 12, seed=2)
h = dm.forward_contact_host(s, a, keep_for_backward=True, reset_cache=True)
dm.backward_contact_host(np.ones_like(s))
rng = np.random.default_rng(0)
m = 9
A = rng.normal(size=(4, m, m)); A = A @ A.transpose(0, 2, 1) + 0.1 * np.eye(m)
x, lab, stt = nb.solve_boxed_lcp_batch(torch.tensor(A, device="cuda"), torch.tensor(rng.normal(size=(4, m)), device="cuda"),
                                       torch.zeros(4, m, device="cuda", dtype=torch.float64), torch.full((4, m), float("inf"), device="cuda", dtype=torch.float64),
                                       torch.full((4, m), -1, device="cuda", dtype=torch.int32))
torch.cuda.synchronize()
print("sanitize contact run finished")
