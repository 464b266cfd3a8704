"""Parity at the batch sizes BASELINE.json names, through the public autograd boundary:
  * configs[2] half-cheetah + ground, B = 4096: strided oracle sample (contact set, labels, status bit-exact; values 1e-4), fwd + bwd
  * configs[3] Atlas + ground, B = 8192 per GPU: same
  * an 8-step CONTACT rollout: gradient of a terminal loss w.r.t. x_0 and every action vs the oracle chained step by step with the
    same flowing LCP cache (BackpropSnapshot::backprop composed like SingleShot::backpropGradientWrt)
  * row a5: the pointer-style forward-dynamics entry on a gravity-free multi-link arm (unittests/comprehensive/test_SimpleFeatherstone.cpp:33-143)
"""
import numpy as np
import pytest
import torch

import nimblephysics_b200 as nb
from oracle import binding as ob
from tests.util import contact_inputs, load_raw, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,B,stride", [("half_cheetah", 4096, 97), ("atlas_ground", 8192, 331)])
def test_contact_step_at_baseline_batch_matches_oracle_sample(oracle_mod, name, B, stride):
    raw = load_raw(name)
    world = nb.World.from_raw(raw)
    ow = ob.OracleContactWorld(raw)
    s, a = contact_inputs(raw, name, B, seed=21)
    g = np.random.default_rng(5).normal(size=(B, 2 * raw.ndof)).astype(np.float32)
    st = torch.tensor(s, device="cuda", requires_grad=True)
    at = torch.tensor(a, device="cuda", requires_grad=True)
    nb.reset_contact_cache(world)
    out = nb.timestep(world, st, at)
    c = world._lcp_cache
    got = {k: c[k].cpu().numpy() for k in ("x", "m", "labels", "status", "nc")}
    out.backward(torch.tensor(g, device="cuda"))
    assert nb.check_contact_status(world) & (128 | 256 | 2048) == 0
    nxt, gs, ga = out.detach().cpu().numpy(), st.grad.cpu().numpy(), at.grad.cpu().numpy()
    assert np.isfinite(nxt).all() and np.isfinite(gs).all() and np.isfinite(ga).all()
    rows = checked = 0
    for w in range(0, B, stride):
        ro = ow.step_contact(s[w].astype(np.float64), a[w].astype(np.float64))
        mo = ro["m"]
        rows += mo
        assert got["nc"][w] == ro["nc"] and got["m"][w] == mo
        assert np.array_equal(got["labels"][w][:mo], ro["mapping"])
        assert (got["status"][w] & ~96) == (ro["status"] & ~96)   # bits 32 / 64: see tests/test_gpu_contact.py
        assert rel_err(nxt[w], ro["next_state"]) < 1e-4
        if (got["status"][w] & 64) != (ro["status"] & 64):
            continue  # marginal standardisation validity: the two sides may keep different (both valid) impulses
        rgs, rga, rc = ow.backprop_contact(s[w].astype(np.float64), a[w].astype(np.float64), g[w].astype(np.float64))
        assert rc >= 0
        assert rel_err(gs[w], rgs) < 1e-4 and rel_err(ga[w], rga) < 1e-4
        checked += 1
    assert rows > 0 and checked >= 10


def test_eight_step_contact_rollout_gradient_matches_chained_oracle(oracle_mod):
    """dL/dx_0 and dL/du_t of L = sum(x_T * c) through 8 half-cheetah steps with contacts, the LCP cache flowing on both sides."""
    raw = load_raw("half_cheetah")
    world = nb.World.from_raw(raw)
    ow = ob.OracleContactWorld(raw)
    B, T = 16, 8
    s, a = contact_inputs(raw, "half_cheetah", B, seed=13)
    rng = np.random.default_rng(3)
    acts = [(a + rng.normal(0, 0.2, a.shape)).astype(np.float32) for _ in range(T)]
    for u in acts:
        u[:, 0] = 0
    cvec = rng.normal(size=(B, 2 * raw.ndof)).astype(np.float32)
    x0 = torch.tensor(s, device="cuda", requires_grad=True)
    us = [torch.tensor(u, device="cuda", requires_grad=True) for u in acts]
    nb.reset_contact_cache(world)
    x = x0
    traj = []
    for t in range(T):
        traj.append(x.detach().cpu().numpy())
        x = nb.timestep(world, x, us[t])
    (x * torch.tensor(cvec, device="cuda")).sum().backward()
    assert nb.check_contact_status(world) & (128 | 256 | 2048) == 0
    gx0 = x0.grad.cpu().numpy()
    gus = [u.grad.cpu().numpy() for u in us]
    compared = 0
    for w in range(B):
        # forward on the oracle from the device's own fp32 trajectory (no tie flips from accumulated rounding), caches flowing
        xs, warm, ok = [], [None], True
        for t in range(T):
            st64 = traj[t][w].astype(np.float64)
            ro = ow.step_contact(st64, acts[t][w].astype(np.float64), warm[-1])
            if ro["status"] & (8 | 16 | 64):
                ok = False  # approximate forward answers (PGS / friction drop / unstandardised): the frozen map is not their derivative
            warm.append(ro["x"] if ro["m"] else None)
        if not ok:
            continue
        gbar = cvec[w].astype(np.float64)
        for t in reversed(range(T)):
            st64 = traj[t][w].astype(np.float64)
            gs, ga, rc = ow.backprop_contact(st64, acts[t][w].astype(np.float64), gbar, warm[t])
            assert rc >= 0
            assert rel_err(gus[t][w], ga) < 2e-3, (w, t)
            gbar = gs
        assert rel_err(gx0[w], gbar) < 2e-3, w
        compared += 1
    assert compared >= 3


def _multiarm(nlinks=5, length=0.2):
    """createMultiarmRobot(5, 0.2) (unittests/TestHelpers.hpp): a chain of revolute joints with alternating axes, each link a box of
    the given length hanging off the previous one."""
    w = nb.World()
    w.setGravity([0, 0, 0])
    w.setTimeStep(1e-3)
    sk = nb.Skeleton("arm")
    parent = None
    axes = [[0, 0, 1], [0, 1, 0], [1, 0, 0]]
    for k in range(nlinks):
        j, b = sk.createRevoluteJointAndBodyNodePair(parent)
        j.setAxis(axes[k % 3])
        T = nb.Isometry3()
        T.set_translation([0, 0, length if k else 0.0])
        j.setTransformFromParentBodyNode(T)
        b.setMass(1.0 + 0.1 * k)
        b.setLocalCOM([0.01 * k, 0.0, length / 2])
        b.setMomentOfInertia(0.02, 0.03, 0.01, 0.001, 0.0, 0.002)
        parent = b
    w.addSkeleton(sk)
    return w


def test_pointer_style_forward_dynamics_gravity_free_chain(oracle_mod):
    world = _multiarm()
    raw = nb.flatten_world(world)
    dm = nb.device_model_for(world)
    ow = ob.OracleWorld(raw)
    n, B = raw.ndof, 10
    rng = np.random.default_rng(0)
    q, v, tau = rng.uniform(-1, 1, (B, n)), rng.uniform(-1, 1, (B, n)), rng.uniform(-1, 1, (B, n))  # Eigen::VectorXs::Random
    acc = dm.forward_dynamics(torch.tensor(q, device="cuda"), torch.tensor(v, device="cuda"), torch.tensor(tau, device="cuda")).cpu().numpy()
    for w in range(B):
        nxt, qdd = ow.step(np.concatenate([q[w], v[w]]), tau[w], want_qdd=True)
        assert np.abs(acc[w] - qdd).max() < 1e-8 * max(1.0, np.abs(qdd).max()), (w, acc[w], qdd)
    # same entry with gravity on and a floating base: Atlas against the oracle's ABA
    rawA = load_raw("atlas")
    wa = nb.World.from_raw(rawA)
    da = nb.device_model_for(wa)
    oa = ob.OracleWorld(rawA)
    nA = rawA.ndof
    qa, va, ta = rng.uniform(-0.4, 0.4, (4, nA)), rng.uniform(-1, 1, (4, nA)), rng.uniform(-5, 5, (4, nA))
    accA = da.forward_dynamics(torch.tensor(qa, device="cuda"), torch.tensor(va, device="cuda"), torch.tensor(ta, device="cuda")).cpu().numpy()
    for w in range(4):
        _, qdd = oa.step(np.concatenate([qa[w], va[w]]), ta[w], want_qdd=True)
        assert np.abs(accA[w] - qdd).max() < 1e-7 * max(1.0, np.abs(qdd).max())


def test_model_edits_after_the_first_step_reach_the_gpu(oracle_mod):
    """BodyNode / Joint setters called AFTER a timestep (domain randomisation) must rebuild the device model (engine.device_model_for)."""
    world = _multiarm(3)
    n = world.getNumDofs()
    s = torch.zeros((4, 2 * n), device="cuda"); s[:, n:] = 1.0
    a = torch.ones((4, n), device="cuda")
    y0 = nb.timestep(world, s, a).clone()
    world.skeletons[0].getBodyNode(2).setMass(25.0)
    y1 = nb.timestep(world, s, a).clone()
    assert not torch.allclose(y0, y1)
    ow = ob.OracleWorld(nb.flatten_world(world))
    ref = ow.step(s[0].cpu().numpy().astype(np.float64), a[0].cpu().numpy().astype(np.float64))
    assert rel_err(y1[0].cpu().numpy(), ref) < 1e-5
    world.skeletons[0].getJoint(1).setDampingCoefficient(0, 3.0)
    y2 = nb.timestep(world, s, a)
    assert not torch.allclose(y1, y2)
