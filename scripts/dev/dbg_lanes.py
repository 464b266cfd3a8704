import sys, numpy as np, torch
sys.path.insert(0, ".")  # run from the repo root
import nimblephysics_b200 as nb
from tests.util import load_raw, sample_inputs
name = sys.argv[1] if len(sys.argv) > 1 else "half_cheetah"
raw = load_raw(name)
dm = nb.DeviceModel.from_raw(raw, contacts=False)
print("schedules", [c.lanes for c in dm.schedules])
for c in dm.schedules:
    print(c.lanes, c.trunk_ranges, c.limb_ranges, "flags", c.flags.tolist(), "slot_self", c.slot_self.tolist(), "slot_parent", c.slot_parent.tolist(), "count", c.slot_count.tolist(), c.nslots)
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
s, a, g = sample_inputs(raw, B, seed=35)
sd, ad, gd = (torch.tensor(x, device="cuda") for x in (s, a, g))
stream = torch.cuda.current_stream().cuda_stream
res = {}
for K in [c.lanes for c in dm.schedules]:
    dm.set_lanes(K)
    nxt = torch.full_like(sd, float("nan")); saved = torch.full((dm.saved_words, B), float("nan"), device="cuda")
    gs, ga = torch.full_like(sd, float("nan")), torch.full_like(ad, float("nan"))
    dm.forward_device(B, sd.data_ptr(), ad.data_ptr(), nxt.data_ptr(), saved.data_ptr(), stream, 0)
    dm.backward_device(B, sd.data_ptr(), ad.data_ptr(), saved.data_ptr(), gd.data_ptr(), gs.data_ptr(), ga.data_ptr(), stream, 0)
    torch.cuda.synchronize()
    res[K] = dict(nxt=nxt.cpu().numpy(), gs=gs.cpu().numpy(), ga=ga.cpu().numpy(), saved=saved.cpu().numpy().T)
for K in list(res)[1:]:
    for k in res[1]:
        x, y = res[1][k], res[K][k]
        bad = ~np.isclose(x, y, rtol=1e-4, atol=1e-5, equal_nan=True)
        print(K, k, "nbad", bad.sum(), "worlds", np.unique(np.nonzero(bad)[0])[:20], "cols", np.unique(np.nonzero(bad)[1])[:40], "nan1", np.isnan(x).sum(), "nanK", np.isnan(y).sum())
