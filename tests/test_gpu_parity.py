"""Parity tests proper: the CUDA path (through the C ABI / the autograd boundary) vs the fp64 oracle on the same
seeded inputs, plus size-independent properties at BASELINE.json's full batch size.
Tolerance (north_star): fp32 results within 1e-4 relative of the fp64 reference path."""
import ctypes

import numpy as np
import pytest
import torch

import nimblephysics_b200 as nb
from tests.util import load_raw, rel_err, sample_inputs

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _world(name):
    raw = load_raw(name)
    w = nb.World.from_raw(raw)
    w._contacts_disabled = True  # contact-free step (the ground collider of half_cheetah is ignored here)
    return raw, w


@pytest.mark.parametrize("name", ["cartpole", "half_cheetah", "atlas"])
def test_autograd_boundary_matches_oracle(oracle_mod, name):
    raw, world = _world(name)
    ow = oracle_mod.OracleWorld(raw)
    B = 96
    s, a, g = sample_inputs(raw, B, seed=31)
    st = torch.tensor(s, device="cuda", requires_grad=True)
    at = torch.tensor(a, device="cuda", requires_grad=True)
    nxt = nb.timestep(world, st, at)
    nxt.backward(torch.tensor(g, device="cuda"))
    nxt, gs, ga = nxt.detach().cpu().numpy(), st.grad.cpu().numpy(), at.grad.cpu().numpy()
    for w in range(0, B, 5):
        s64, a64, g64 = s[w].astype(np.float64), a[w].astype(np.float64), g[w].astype(np.float64)
        ref = ow.step(s64, a64)
        rgs, rga = ow.backprop(s64, a64, g64)
        assert rel_err(nxt[w], ref) < TOL
        assert rel_err(gs[w], rgs) < TOL
        assert rel_err(ga[w], rga) < TOL


@pytest.mark.parametrize("precision", [0, 1])
def test_c_abi_device_and_host_entry_points(oracle_mod, precision):
    raw, world = _world("atlas")
    dm = nb.device_model_for(world)
    ow = oracle_mod.OracleWorld(raw)
    B = 40
    s, a, g = sample_inputs(raw, B, seed=32)
    # host entry points (copies inside)
    nxt_h = dm.forward_host(s, a, keep_for_backward=True, precision=precision)
    gs_h, ga_h = dm.backward_host(g, precision=precision)
    # device entry points
    sd, ad, gd = (torch.tensor(x, device="cuda") for x in (s, a, g))
    nxt = torch.empty_like(sd)
    saved = torch.empty((dm.saved_words, B), device="cuda", dtype=torch.float64 if precision else torch.float32)
    gs, ga = torch.empty_like(sd), torch.empty_like(ad)
    stream = torch.cuda.current_stream().cuda_stream
    dm.forward_device(B, sd.data_ptr(), ad.data_ptr(), nxt.data_ptr(), saved.data_ptr(), stream, precision)
    dm.backward_device(B, sd.data_ptr(), ad.data_ptr(), saved.data_ptr(), gd.data_ptr(), gs.data_ptr(), ga.data_ptr(), stream, precision)
    torch.cuda.synchronize()
    assert np.array_equal(nxt.cpu().numpy(), nxt_h) and np.array_equal(gs.cpu().numpy(), gs_h) and np.array_equal(ga.cpu().numpy(), ga_h)
    tol = TOL if precision == 0 else 2e-6  # fp64 arithmetic, fp32 I/O
    for w in range(0, B, 7):
        s64, a64, g64 = s[w].astype(np.float64), a[w].astype(np.float64), g[w].astype(np.float64)
        rgs, rga = ow.backprop(s64, a64, g64)
        assert rel_err(nxt_h[w], ow.step(s64, a64)) < tol and rel_err(gs_h[w], rgs) < tol and rel_err(ga_h[w], rga) < tol


def test_full_batch_properties_atlas_4096(oracle_mod):
    """BASELINE config 2 size.  Properties that do not need the oracle at every world:
    (a) a world's result does not depend on its position in the batch or on the batch size (bit-exact);
    (b) the backward is linear in the incoming gradient;  (c) M^-1 is symmetric: e_i . gtau(e_j) == e_j . gtau(e_i);
    (d) oracle parity on a strided sample."""
    raw, world = _world("atlas")
    ow = oracle_mod.OracleWorld(raw)
    n = raw.ndof
    B = 4096
    s, a, g = sample_inputs(raw, B, seed=33, tau_scale=30.0)
    st = torch.tensor(s, device="cuda")
    at = torch.tensor(a, device="cuda")

    def run(sx, ax, gx):
        sx = sx.clone().requires_grad_(True)
        ax = ax.clone().requires_grad_(True)
        out = nb.timestep(world, sx, ax)
        out.backward(gx)
        return out.detach(), sx.grad, ax.grad

    gt = torch.tensor(g, device="cuda")
    out, gs, ga = run(st, at, gt)
    perm = torch.randperm(B, device="cuda")[:777]
    out2, gs2, ga2 = run(st[perm], at[perm], gt[perm])
    assert torch.equal(out[perm], out2) and torch.equal(gs[perm], gs2) and torch.equal(ga[perm], ga2)  # (a)
    g2 = torch.randn_like(gt)
    _, gsa, gaa = run(st, at, gt + g2)
    _, gsb, gab = run(st, at, g2)
    assert rel_err((gs + gsb).cpu().numpy(), gsa.cpu().numpy()) < 1e-5  # (b)
    assert rel_err((ga + gab).cpu().numpy(), gaa.cpu().numpy()) < 1e-5
    ei = torch.zeros_like(gt)
    ej = torch.zeros_like(gt)
    ei[:, n + 7] = 1.0
    ej[:, n + 20] = 1.0
    _, _, gai = run(st, at, ei)
    _, _, gaj = run(st, at, ej)
    assert torch.allclose(gai[:, 20], gaj[:, 7], rtol=2e-4, atol=1e-9)  # (c)
    out, gs, ga = out.cpu().numpy(), gs.cpu().numpy(), ga.cpu().numpy()
    assert np.isfinite(out).all() and np.isfinite(gs).all() and np.isfinite(ga).all()
    for w in range(0, B, 257):  # (d)
        s64, a64, g64 = s[w].astype(np.float64), a[w].astype(np.float64), g[w].astype(np.float64)
        rgs, rga = ow.backprop(s64, a64, g64)
        assert rel_err(out[w], ow.step(s64, a64)) < TOL and rel_err(gs[w], rgs) < TOL and rel_err(ga[w], rga) < TOL


def test_edge_cases(oracle_mod):
    raw, world = _world("cartpole")
    ow = oracle_mod.OracleWorld(raw)
    # ragged batch sizes around the warp size, B=1 and the legacy 1-D call
    for B in (1, 31, 33, 100):
        s, a, g = sample_inputs(raw, B, seed=B)
        out = nb.timestep(world, torch.tensor(s, device="cuda"), torch.tensor(a, device="cuda")).cpu().numpy()
        for w in (0, B - 1):
            assert rel_err(out[w], ow.step(s[w].astype(np.float64), a[w].astype(np.float64))) < TOL
    s, a, g = sample_inputs(raw, 1, seed=9)
    st = torch.tensor(s[0].astype(np.float64), requires_grad=True)  # CPU fp64 1-D like the reference's scripts
    at = torch.tensor(a[0].astype(np.float64), requires_grad=True)
    out = nb.timestep(world, st, at)
    assert out.dtype == torch.float64 and out.shape == (4,)
    out.backward(torch.tensor(g[0].astype(np.float64)))
    rgs, rga = ow.backprop(s[0].astype(np.float64), a[0].astype(np.float64), g[0].astype(np.float64))
    assert rel_err(st.grad.numpy(), rgs) < TOL and rel_err(at.grad.numpy(), rga) < TOL
    assert np.allclose(world.getState(), out.detach().numpy())  # world left at the post-step state
    # wrong sizes raise (reference prints and ignores)
    with pytest.raises(ValueError):
        nb.timestep(world, torch.zeros(5, device="cuda"), torch.zeros(2, device="cuda"))
    # gradient clipping decision at a bound is bit-exact
    s, a, g = sample_inputs(raw, 8, seed=4)
    s[:, 0] = np.float32(raw.pos_hi[0])
    st = torch.tensor(s, device="cuda", requires_grad=True)
    nb.timestep(world, st, torch.tensor(a, device="cuda")).backward(torch.tensor(g, device="cuda"))
    for w in range(8):
        rgs, _ = ow.backprop(s[w].astype(np.float64), a[w].astype(np.float64), g[w].astype(np.float64))
        assert (st.grad[w, 0].item() == 0.0) == (rgs[0] == 0.0)


def test_rollout_through_autograd_matches_oracle_chain(oracle_mod):
    """8 chained steps, loss on the final state, backprop through the horizon (SingleShot.cpp:539-686 semantics)."""
    raw, world = _world("half_cheetah")
    ow = oracle_mod.OracleWorld(raw)
    T, B = 8, 16
    s, a, _ = sample_inputs(raw, B, seed=77, v_scale=0.3)
    acts = [np.random.default_rng(100 + t).uniform(-3, 3, a.shape).astype(np.float32) for t in range(T)]
    st = torch.tensor(s, device="cuda", requires_grad=True)
    ats = [torch.tensor(x, device="cuda", requires_grad=True) for x in acts]
    x = st
    for t in range(T):
        x = nb.timestep(world, x, ats[t])
    loss = (x * x).sum()
    loss.backward()
    w = 3
    xs = [s[w].astype(np.float64)]
    for t in range(T):
        xs.append(ow.step(xs[-1], acts[t][w].astype(np.float64)))
    gq = 2 * xs[-1]
    ga_ref = [None] * T
    for t in reversed(range(T)):
        gq, ga_ref[t] = ow.backprop(xs[t], acts[t][w].astype(np.float64), gq)
    assert rel_err(x[w].detach().cpu().numpy(), xs[-1]) < TOL
    assert rel_err(st.grad[w].cpu().numpy(), gq) < 5e-4
    assert rel_err(ats[0].grad[w].cpu().numpy(), ga_ref[0]) < 5e-4


@pytest.mark.parametrize("name", ["half_cheetah", "atlas"])
def test_cooperative_lane_schedules_agree(oracle_mod, name):
    """The library sweeps a world with 1, 2, 4 or 8 cooperating threads depending on the batch size
    (include/nb2.h nb2_model_add_schedule).  Every schedule must give the same numbers: children are accumulated in
    the same order whichever lane produced them, so the schedules differ only by fused-multiply-add contraction
    (a register handoff lets the compiler fuse the accumulation into the producing expression, a slot does not)."""
    raw, world = _world(name)
    dm = nb.device_model_for(world)
    ow = oracle_mod.OracleWorld(raw)
    B = 1000  # not a multiple of the worlds-per-warp of any schedule: exercises the partial last warp
    s, a, g = sample_inputs(raw, B, seed=35)
    sd, ad, gd = (torch.tensor(x, device="cuda") for x in (s, a, g))
    stream = torch.cuda.current_stream().cuda_stream
    res = {}
    lanes = sorted(int(c.lanes) for c in dm.schedules)
    assert lanes[0] == 1 and len(lanes) > 1
    try:
        for K in lanes:
            dm.set_lanes(K)
            assert dm.lanes_for(B) == K and dm.lanes_for(B, True) == K
            nxt = torch.full_like(sd, float("nan"))
            saved = torch.empty((dm.saved_words, B), device="cuda")
            gs, ga = torch.full_like(sd, float("nan")), torch.full_like(ad, float("nan"))
            dm.forward_device(B, sd.data_ptr(), ad.data_ptr(), nxt.data_ptr(), saved.data_ptr(), stream, 0)
            dm.backward_device(B, sd.data_ptr(), ad.data_ptr(), saved.data_ptr(), gd.data_ptr(), gs.data_ptr(), ga.data_ptr(), stream, 0)
            torch.cuda.synchronize()
            res[K] = (nxt.cpu().numpy(), gs.cpu().numpy(), ga.cpu().numpy(), saved.cpu().numpy())
    finally:
        dm.set_lanes(0)
    for K in lanes[1:]:
        for x, y in zip(res[1], res[K]):
            assert rel_err(x, y) < 2e-6, f"schedule with {K} lanes differs from the single-thread sweep"
    for w in range(0, B, 97):
        s64, a64, g64 = s[w].astype(np.float64), a[w].astype(np.float64), g[w].astype(np.float64)
        rgs, rga = ow.backprop(s64, a64, g64)
        K = lanes[-1]
        assert rel_err(res[K][0][w], ow.step(s64, a64)) < TOL and rel_err(res[K][1][w], rgs) < TOL and rel_err(res[K][2][w], rga) < TOL


def test_timestep_mass_argument_and_gradient(oracle_mod):
    """timestep(world, state, action, mass) (python/nimblephysics/timestep.py:28-35, 55-60): the masses registered with
    world.tuneMass are set before the step and lossWrtMass comes back as the fourth gradient."""
    import copy

    from nimblephysics_b200 import modelspec as ms

    raw, world = _world("atlas")
    sk = world.getSkeleton(0)
    bodies = sk._ordered_bodies()
    picks = [(bodies[2], ms.INERTIA_MASS), (bodies[8], ms.INERTIA_COM), (bodies[15], ms.INERTIA_MASS)]
    for b, kind in picks:
        world.tuneMass(b, kind)
    assert world.getMassDims() == 5
    m0 = world.getMasses()
    m1 = m0 * np.array([1.3, 1.0, 1.0, 1.0, 0.8]) + np.array([0, 0.01, -0.02, 0.005, 0])
    B = 16  # two full warp groups: the backward stages the saved stream in shared memory (its stride differs from the batch's)
    s, a, g = sample_inputs(raw, B, seed=41)
    st, at = torch.tensor(s, device="cuda"), torch.tensor(a, device="cuda")
    mt = torch.tensor(m1, dtype=torch.float64, requires_grad=True)
    nxt = nb.timestep(world, st, at, mt)
    assert np.allclose(world.getMasses(), m1)  # side effect of the reference: the masses stay set
    (nxt * torch.tensor(g, device="cuda")).sum().backward()
    gm = mt.grad.numpy()
    assert mt.grad.dtype == torch.float64 and gm.shape == (5,)
    # oracle at the new masses: forward parity and central differences of the loss
    entries = world._mass_entries()

    def raw_at(mvec):
        r = copy.deepcopy(raw)
        k = 0
        for (bi, kind) in entries:
            d = ms.WRT_MASS_DIMS[kind]
            r.mass[bi], r.com[bi], r.moment[bi] = ms._apply_mass_entry(kind, mvec[k:k + d], r.mass[bi], r.com[bi], r.moment[bi])
            k += d
        return r

    def loss(mvec):
        ow = oracle_mod.OracleWorld(raw_at(mvec))
        return sum(float(g[w].astype(np.float64) @ ow.step(s[w].astype(np.float64), a[w].astype(np.float64))) for w in range(B))

    ow1 = oracle_mod.OracleWorld(raw_at(m1))
    for w in range(B):
        assert rel_err(nxt[w].detach().cpu().numpy(), ow1.step(s[w].astype(np.float64), a[w].astype(np.float64))) < TOL
    fd = np.zeros(5)
    for j in range(5):
        e = np.zeros(5); e[j] = 1e-5
        fd[j] = (loss(m1 + e) - loss(m1 - e)) / 2e-5
    assert rel_err(gm, fd) < 1e-3, (gm, fd)  # fp32 kernels, sum over the batch


def test_fused_rollout_matches_step_by_step_rollout():
    """nb2_rollout_forward / nb2_rollout_backward (one C-ABI call per direction) vs chaining `timestep` through autograd:
    same kernels in the same order, so trajectories and gradients agree to the last bit; the loss looks at several states of
    the trajectory (the reverse sweep adds per-step loss gradients like SingleShot::backpropGradientWrt)."""
    raw, world = _world("atlas")
    B, T = 200, 12
    s, a, _ = sample_inputs(raw, B, seed=51)
    rng = np.random.default_rng(52)
    acts = rng.uniform(-10, 10, (T, B, len(raw.action_map))).astype(np.float32)
    wts = torch.tensor(rng.normal(size=(T + 1, B, 2 * raw.ndof)).astype(np.float32), device="cuda")
    wts[1:5] = 0  # some states carry no loss

    x0 = torch.tensor(s, device="cuda", requires_grad=True)
    u = torch.tensor(acts, device="cuda", requires_grad=True)
    traj = nb.rollout_fused(world, x0, u)
    assert traj.shape == (T + 1, B, 2 * raw.ndof)
    (traj * wts).sum().backward()

    x0r = torch.tensor(s, device="cuda", requires_grad=True)
    ur = [torch.tensor(acts[t], device="cuda", requires_grad=True) for t in range(T)]
    xs = [x0r]
    for t in range(T):
        xs.append(nb.timestep(world, xs[-1], ur[t]))
    (torch.stack(xs, 0) * wts).sum().backward()

    assert torch.equal(traj.detach(), torch.stack(xs, 0).detach())
    assert rel_err(x0.grad.cpu().numpy(), x0r.grad.cpu().numpy()) < 1e-6
    for t in range(T):
        assert rel_err(u.grad[t].cpu().numpy(), ur[t].grad.cpu().numpy()) < 1e-6


@pytest.mark.parametrize("name", ["half_cheetah", "atlas"])
def test_host_entry_points_with_page_locked_buffers(name):
    """nb2_step_forward_host / nb2_step_backward_host on page-locked buffers: the kernels read and write the host memory
    directly (zero-copy, include/nb2.h).  Results must equal the device entry points bit for bit; B is not a multiple of
    the group size so the partial last group and the unaligned (scalar) tail of the vector copies are exercised."""
    raw, world = _world(name)
    dm = nb.device_model_for(world)
    B = 1003
    s, a, g = sample_inputs(raw, B, seed=61)
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    hs, ha, hg = pin(s), pin(a), pin(g)
    o_n, o_gs, o_ga = torch.empty_like(hs).pin_memory(), torch.empty_like(hs).pin_memory(), torch.empty_like(ha).pin_memory()
    dm.forward_host(hs.numpy(), ha.numpy(), True, 0, out=o_n.numpy())
    dm.backward_host(hg.numpy(), 0, out_state=o_gs.numpy(), out_action=o_ga.numpy())
    sd, ad, gd = hs.cuda(), ha.cuda(), hg.cuda()
    nxt, gs, ga = torch.empty_like(sd), torch.empty_like(sd), torch.empty_like(ad)
    saved = torch.empty((dm.saved_words, B), device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    dm.forward_device(B, sd.data_ptr(), ad.data_ptr(), nxt.data_ptr(), saved.data_ptr(), stream, 0)
    dm.backward_device(B, sd.data_ptr(), ad.data_ptr(), saved.data_ptr(), gd.data_ptr(), gs.data_ptr(), ga.data_ptr(), stream, 0)
    torch.cuda.synchronize()
    assert torch.equal(o_n, nxt.cpu()) and torch.equal(o_gs, gs.cpu()) and torch.equal(o_ga, ga.cpu())
    # pageable buffers take the staged-copy path: same numbers
    n2 = dm.forward_host(s, a, True, 0)
    g2s, g2a = dm.backward_host(g, 0)
    assert np.array_equal(n2, o_n.numpy()) and np.array_equal(g2s, o_gs.numpy()) and np.array_equal(g2a, o_ga.numpy())


def test_wound_up_joint_angles_keep_parity(oracle_mod):
    """The fp32 kernels use the special-function unit's sine / cosine after an explicit 2*pi range reduction
    (csrc/nb2_math.cuh nb2_sincos): revolute joints that have wound up many turns must stay inside the 1e-4 tolerance."""
    raw, world = _world("atlas")
    ow = oracle_mod.OracleWorld(raw)
    B = 64
    s, a, g = sample_inputs(raw, B, seed=71)
    rng = np.random.default_rng(72)
    n = raw.ndof
    s[:, 6:n] = rng.uniform(-300.0, 300.0, (B, n - 6)).astype(np.float32)  # up to ~50 turns
    st = torch.tensor(s, device="cuda", requires_grad=True)
    at = torch.tensor(a, device="cuda", requires_grad=True)
    nxt = nb.timestep(world, st, at)
    nxt.backward(torch.tensor(g, device="cuda"))
    nxt, gs, ga = nxt.detach().cpu().numpy(), st.grad.cpu().numpy(), at.grad.cpu().numpy()
    worst = 0.0
    for w in range(0, B, 4):
        s64, a64, g64 = s[w].astype(np.float64), a[w].astype(np.float64), g[w].astype(np.float64)
        rgs, rga = ow.backprop(s64, a64, g64)
        worst = max(worst, rel_err(nxt[w], ow.step(s64, a64)), rel_err(gs[w], rgs), rel_err(ga[w], rga))
    assert worst < TOL, worst


def test_legacy_world_step(oracle_mod):
    """world.setAction(...); world.step() — the stateful single-world loop of the reference's examples (World.cpp:221-254):
    the stored state advances, the control forces are cleared after the step (World.cpp:297-302)."""
    raw, world = _world("cartpole")
    ow = oracle_mod.OracleWorld(raw)
    x = np.array([0.1, 0.3, -0.2, 0.4])
    world.setState(x)
    for k in range(3):
        a = np.array([0.5 * (k + 1), 0.0])
        world.setAction(a)
        world.step()
        x = ow.step(x.astype(np.float32).astype(np.float64), a)
        assert rel_err(world.getState(), x) < TOL
        assert np.all(world.getAction() == 0.0)
        x = world.getState()
