import sys, time, numpy as np, torch
sys.path.insert(0, ".")  # run from the repo root
import nimblephysics_b200 as nb
from bench import make_inputs, ATLAS
raw = nb.RawModel.load(ATLAS)
n, na = raw.ndof, len(raw.action_map)
dm = nb.DeviceModel.from_raw(raw)
B = 4096
hs, ha, hg = make_inputs(raw, B, 99)
pin = lambda x: torch.from_numpy(x).pin_memory()
hs_t, ha_t, hg_t = pin(hs), pin(ha), pin(hg)
o_n, o_gs, o_ga = (torch.empty((B, 2 * n)).pin_memory(), torch.empty((B, 2 * n)).pin_memory(), torch.empty((B, na)).pin_memory())
def T(f, k=50):
    for _ in range(5): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e6
a_s, a_a, a_n, a_g, a_gs, a_ga = hs_t.numpy(), ha_t.numpy(), o_n.numpy(), hg_t.numpy(), o_gs.numpy(), o_ga.numpy()
print("fwd_host us", T(lambda: dm.forward_host(a_s, a_a, True, 0, out=a_n)))
print("fwd_host nokeep us", T(lambda: dm.forward_host(a_s, a_a, False, 0, out=a_n)))
def fb():
    dm.forward_host(a_s, a_a, True, 0, out=a_n); dm.backward_host(a_g, 0, out_state=a_gs, out_action=a_ga)
print("fwd+bwd host us", T(fb))
# raw copies
d = torch.empty((B, 2 * n), device="cuda")
print("h2d 1.08MB us", T(lambda: d.copy_(hs_t, non_blocking=True)))
print("d2h 1.08MB us", T(lambda: o_n.copy_(d, non_blocking=True)))
big = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); dbig = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
t = T(lambda: dbig.copy_(big, non_blocking=True), 10); print("h2d 64MB GB/s", (64 << 20) / t / 1e3)
t = T(lambda: big.copy_(dbig, non_blocking=True), 10); print("d2h 64MB GB/s", (64 << 20) / t / 1e3)
s = torch.cuda.Stream()
def sync_only():
    s.synchronize()
print("stream sync us", T(sync_only))
import ctypes
from nimblephysics_b200 import _cabi
print("ctypes call us", T(lambda: _cabi.lib().nb2_launch_count()))
sd, ad = hs_t.cuda(), ha_t.cuda(); nx = torch.empty_like(sd); sv = torch.empty((dm.saved_words, B), device="cuda")
st = torch.cuda.current_stream().cuda_stream
print("fwd dev us", T(lambda: dm.forward_device(B, sd.data_ptr(), ad.data_ptr(), nx.data_ptr(), sv.data_ptr(), st, 0)))
