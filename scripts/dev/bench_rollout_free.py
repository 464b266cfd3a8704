"""Dev: where the time of the contact-free fused rollout goes (C driver alone vs the autograd wrapper)."""
import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import numpy as np, torch
import nimblephysics_b200 as nb
from nimblephysics_b200.engine import FP32, device_model_for
from tests.util import load_raw
B = int(os.environ.get("B", "4096")); T = int(os.environ.get("T", "64"))
raw = load_raw("atlas"); world = nb.World.from_raw(raw); dm = device_model_for(world)
n2, na = 2 * raw.ndof, len(raw.action_map)
rng = np.random.default_rng(0)
x0 = torch.tensor(rng.uniform(-0.3, 0.3, (B, n2)).astype(np.float32), device="cuda")
u = torch.tensor(rng.uniform(-20, 20, (T, B, na)).astype(np.float32), device="cuda")
states = torch.empty((T + 1, B, n2), device="cuda"); states[0] = x0
saved = torch.empty((T, dm.saved_words, B), device="cuda")
gs = torch.zeros_like(states); ga = torch.empty_like(u)
st = torch.cuda.current_stream().cuda_stream
def ev(): return torch.cuda.Event(enable_timing=True)
for rep in range(3):
    e = [ev() for _ in range(3)]
    gs.zero_(); gs[-1] = 1.0
    e[0].record()
    dm.rollout_forward_device(B, T, states.data_ptr(), u.data_ptr(), saved.data_ptr(), st, FP32)
    e[1].record()
    dm.rollout_backward_device(B, T, states.data_ptr(), u.data_ptr(), saved.data_ptr(), gs.data_ptr(), ga.data_ptr(), st, FP32)
    e[2].record(); torch.cuda.synchronize()
    print("C driver: fwd %.3f ms  bwd %.3f ms  -> %.3e world-steps/s" % (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), B * T / e[0].elapsed_time(e[2]) * 1e3))
xr = x0.clone().requires_grad_(True); ur = u.clone().requires_grad_(True)
for rep in range(3):
    e0, e1 = ev(), ev(); e0.record()
    tr = nb.rollout_fused(world, xr, ur); (tr[-1] * tr[-1]).sum().backward()
    e1.record(); torch.cuda.synchronize()
    print("autograd wrapper: %.3f ms" % e0.elapsed_time(e1))
# single steps for comparison
nxt = torch.empty_like(x0); sv = torch.empty((dm.saved_words, B), device="cuda")
g = torch.randn(B, n2, device="cuda"); g1 = torch.empty_like(x0); g2 = torch.empty((B, na), device="cuda")
for rep in range(2):
    e0, e1 = ev(), ev(); e0.record()
    for t in range(T):
        dm.forward_device(B, x0.data_ptr(), u[t].data_ptr(), nxt.data_ptr(), sv.data_ptr(), st, FP32)
        dm.backward_device(B, x0.data_ptr(), u[t].data_ptr(), sv.data_ptr(), g.data_ptr(), g1.data_ptr(), g2.data_ptr(), st, FP32)
    e1.record(); torch.cuda.synchronize()
    print("T single steps (hot buffers): %.3f ms" % e0.elapsed_time(e1))
