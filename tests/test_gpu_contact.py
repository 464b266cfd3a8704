"""Contact / boxed-LCP stage on the GPU (through the autograd boundary -> C ABI nb2_step_forward_contact) vs the oracle.
Contact set, LCP size, labels, status: bit-exact; impulses / next state: 1e-4 relative."""
import numpy as np
import pytest
import torch

import nimblephysics_b200 as nb
from oracle import binding as ob
from tests.util import contact_inputs, load_raw, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["half_cheetah", "atlas_ground"])
def test_contact_forward_matches_oracle(oracle_mod, name):
    raw = load_raw(name)
    world = nb.World.from_raw(raw)
    ow = ob.OracleContactWorld(raw)
    B, T = 64, 4
    s, a = contact_inputs(raw, name, B, seed=5)
    at = torch.tensor(a, device="cuda")
    x = torch.tensor(s, device="cuda")
    so = s.astype(np.float64).copy()
    xo = [None] * B
    nb.reset_contact_cache(world)
    total_rows = n_marginal = 0
    for t in range(T):
        with torch.no_grad():
            nxt = nb.timestep(world, x, at)
        c = nb.contact_cache(world, B, x.device)
        got = {k: c[k].cpu().numpy() for k in ("x", "m", "labels", "status", "nc", "cinfo")}
        nxt_h = nxt.cpu().numpy()
        for w in range(0, B, 3):
            ro = ow.step_contact(so[w], a[w].astype(np.float64), xo[w])
            mo = ro["m"]
            total_rows += mo
            assert got["nc"][w] == ro["nc"] and got["m"][w] == mo
            assert np.array_equal(got["labels"][w][:mo], ro["mapping"])
            # same solver branch.  Two informational bits may differ on SINGULAR problems (redundant contacts): bit 32 ("Dantzig
            # returned NaN and was reset" — on a singular factor NaN-vs-garbage is decided by 1e-17 rounding noise in A; both
            # are rejected and fall through to PGS identically) and bit 64 ("final standardisation rejected by the 1e-5
            # validity check" — pivoted Cholesky on the device vs SVD in the oracle, a residual sitting on the tolerance can
            # fall either side).  Labels, impulses and the next state are still compared strictly.
            assert (got["status"][w] & ~96) == (ro["status"] & ~96)
            marginal = (got["status"][w] & 64) != (ro["status"] & 64)
            n_marginal += int(marginal)
            if mo:
                assert np.abs(got["x"][w][:mo] - ro["x"]).max() < (2e-3 if marginal else 1e-5) * max(1.0, np.abs(ro["x"]).max())
            assert rel_err(nxt_h[w], ro["next_state"]) < 1e-4
            xo[w] = ro["x"] if mo else None
        # continue both sides from the device's fp32 rows (avoids tie flips from accumulated rounding differences)
        so = nxt_h.astype(np.float64)
        x = torch.tensor(so.astype(np.float32), device="cuda")
    assert total_rows > 0 and n_marginal <= 3


def test_full_batch_contact_properties_atlas_4096():
    """Size-independent properties at the benchmark batch: results do not depend on the position in the batch
    (bit-exact, including labels), contact counts are plausible, every status word is one of the restated branches."""
    raw = load_raw("atlas_ground")
    world = nb.World.from_raw(raw)
    B = 4096
    s, a = contact_inputs(raw, "atlas_ground", B, seed=11)
    st, at = torch.tensor(s, device="cuda"), torch.tensor(a, device="cuda")
    nb.reset_contact_cache(world)
    with torch.no_grad():
        nxt = nb.timestep(world, st, at)
    c = nb.contact_cache(world, B, st.device)
    labels, m, status, nc, xl = c["labels"].clone(), c["m"].clone(), c["status"].clone(), c["nc"].clone(), c["x"].clone()
    perm = torch.randperm(B, device="cuda")[:1000]
    world2 = nb.World.from_raw(raw)
    with torch.no_grad():
        nxt2 = nb.timestep(world2, st[perm], at[perm])
    c2 = nb.contact_cache(world2, 1000, st.device)
    assert torch.equal(nxt[perm], nxt2) and torch.equal(labels[perm], c2["labels"]) and torch.equal(m[perm], c2["m"])
    assert torch.equal(status[perm], c2["status"]) and torch.equal(xl[perm], c2["x"])
    assert int(nc.max()) <= 16 and int(nc.min()) >= 0 and float((nc > 0).float().mean()) > 0.5
    assert int((status & ~0x3FF).max()) == 0 and int((status & 0x180).max()) == 0  # no overflow / unsupported geometry
    assert torch.isfinite(nxt).all()


@pytest.mark.parametrize("name", ["half_cheetah", "atlas_ground"])
def test_contact_backward_matches_oracle(oracle_mod, name):
    """VJP of a step with active contact constraints (classification frozen at the forward solution) vs the oracle's
    dual-number Jacobian of the same fixed-classification map (oracle/nb_oracle.cpp step_jacobian_contact)."""
    raw = load_raw(name)
    world = nb.World.from_raw(raw)
    ow = ob.OracleContactWorld(raw)
    B = 48
    s, a = contact_inputs(raw, name, B, seed=8)
    g = np.random.default_rng(2).normal(size=(B, 2 * raw.ndof)).astype(np.float32)
    st = torch.tensor(s, device="cuda", requires_grad=True)
    at = torch.tensor(a, device="cuda", requires_grad=True)
    nb.reset_contact_cache(world)
    out = nb.timestep(world, st, at)
    out.backward(torch.tensor(g, device="cuda"))
    gs, ga = st.grad.cpu().numpy(), at.grad.cpu().numpy()
    checked = with_rows = 0
    for w in range(0, B, 2):
        rgs, rga, rc = ow.backprop_contact(s[w].astype(np.float64), a[w].astype(np.float64), g[w].astype(np.float64))
        assert rc >= 0
        assert rel_err(gs[w], rgs) < 1e-4 and rel_err(ga[w], rga) < 1e-4
        checked += 1
        with_rows += int(rc != 0)
    assert checked > 10 and with_rows > 3


def test_contact_rollout_backward_runs_and_is_finite():
    raw = load_raw("half_cheetah")
    world = nb.World.from_raw(raw)
    s, a = contact_inputs(raw, "half_cheetah", 32, seed=1)
    st = torch.tensor(s, device="cuda", requires_grad=True)
    acts = [torch.tensor(a, device="cuda", requires_grad=True) for _ in range(6)]
    nb.reset_contact_cache(world)
    x = st
    for t in range(6):
        x = nb.timestep(world, x, acts[t])
    (x * x).sum().backward()
    assert torch.isfinite(st.grad).all() and all(torch.isfinite(u.grad).all() for u in acts)


def test_box_stack_contacts_between_moving_bodies(oracle_mod):
    """Two free boxes stacked on a static ground (builder API): rows that act on two moving bodies, forward and backward
    through `timestep`, vs the oracle (contact set / labels bit-exact, next state and gradients within tolerance)."""
    from tests.test_contact_emul import _box_stack_inputs, _box_stack_world

    world = _box_stack_world()
    raw = nb.flatten_world(world)
    ow = ob.OracleContactWorld(raw)
    B = 24
    s, a = _box_stack_inputs(raw, B, seed=3)
    g = np.random.default_rng(2).normal(size=(B, 2 * raw.ndof)).astype(np.float32)
    st = torch.tensor(s, device="cuda", requires_grad=True)
    at = torch.tensor(a, device="cuda", requires_grad=True)
    nb.reset_contact_cache(world)
    out = nb.timestep(world, st, at)
    cache = nb.contact_cache(world, B, st.device)
    labels, mm, status = cache["labels"].cpu().numpy(), cache["m"].cpu().numpy(), cache["status"].cpu().numpy()
    out.backward(torch.tensor(g, device="cuda"))
    nxt, gs, ga = out.detach().cpu().numpy(), st.grad.cpu().numpy(), at.grad.cpu().numpy()
    both = 0
    for w in range(B):
        ro = ow.step_contact(s[w].astype(np.float64), a[w].astype(np.float64))
        m = int(mm[w])
        assert m == ro["m"] and (status[w] & ~96) == (ro["status"] & ~96)
        assert np.array_equal(labels[w][:m], ro["mapping"][:m])
        assert rel_err(nxt[w], ro["next_state"]) < 1e-4
        rgs, rga, rc = ow.backprop_contact(s[w].astype(np.float64), a[w].astype(np.float64), g[w].astype(np.float64))
        assert rc >= 0
        both += int(any(x >= 1 and y >= 1 for x, y in ro["bodies"].tolist()))
        if status[w] & 64:
            continue  # the forward could not standardise x (f_c != Q^+ b): the frozen-classification map is only approximate there
        assert rel_err(gs[w], rgs) < 1e-4 and rel_err(ga[w], rga) < 1e-4
    assert both >= B // 2


def test_mass_gradient_through_contact_matches_host_build(oracle_mod):
    """timestep(..., mass) on a world with active contacts: the GPU's lossWrtMass vs the same device code compiled for the
    host (tests/host_emul), which tests/test_contact_emul.py pins against finite differences of the oracle."""
    from nimblephysics_b200 import modelspec as ms
    from tests.host_emul.binding import EmulWorld

    raw = load_raw("half_cheetah")
    world = nb.World.from_raw(raw)
    bodies = [b for sk in world.skeletons for b in sk._ordered_bodies()]
    cm = nb.compile_model(raw)
    mobile = [i for i in range(raw.nb) if cm.body_owner[i] >= 0]
    picks = [(mobile[1], ms.INERTIA_MASS), (mobile[3], ms.INERTIA_COM), (mobile[5], ms.INERTIA_MASS)]
    for bi, kind in picks:
        world.tuneMass(bodies[bi], kind)
    B = 32
    s, a = contact_inputs(raw, "half_cheetah", B, seed=6)
    g = np.random.default_rng(4).normal(size=s.shape).astype(np.float32)
    mt = torch.tensor(world.getMasses(), dtype=torch.float64, requires_grad=True)
    st, at = torch.tensor(s, device="cuda"), torch.tensor(a, device="cuda")
    nb.reset_contact_cache(world)
    out = nb.timestep(world, st, at, mt)
    (out * torch.tensor(g, device="cuda")).sum().backward()
    gm = mt.grad.numpy()
    ew = EmulWorld(cm)
    r = ew.forward_contact(s, a)
    _, _, gi = ew.backward_contact(s, a, r["saved"], r["crec"], g, want_inertia_grad=True)
    P = ms.inertia_param_jacobian(raw, cm, picks)
    ref = P @ gi.astype(np.float64).sum(axis=1)
    assert int((r["m"] > 0).sum()) > B // 2
    assert rel_err(gm, ref) < 1e-4, (gm, ref)


def test_penetration_correction_backward_matches_oracle(oracle_mod):
    """World::setPenetrationCorrectionEnabled(true) (ContactConstraint.cpp:395-408): steps carry status bit 4096 and back-propagate —
    including the depth derivative of an uncapped correction — like the oracle's dual-number Jacobian.  (Restitution still fails loudly.)"""
    raw = load_raw("half_cheetah")
    raw.penetration_correction = True
    world = nb.World.from_raw(raw)
    ow = ob.OracleContactWorld(raw)
    B = 10
    s, a = contact_inputs(raw, "half_cheetah", B, seed=6)
    s[5:, 1] += 0.0099
    s[5, 1] += 0.00585  # first contact of world 5 shallower than 1e-4: correction under its cap, normal row clamping
    g = np.random.default_rng(2).normal(size=(B, 2 * raw.ndof)).astype(np.float32)
    st = torch.tensor(s, device="cuda", requires_grad=True)
    at = torch.tensor(a, device="cuda", requires_grad=True)
    nb.reset_contact_cache(world)
    out = nb.timestep(world, st, at)
    out.backward(torch.tensor(g, device="cuda"))
    bits = nb.check_contact_status(world)
    assert bits & 4096 and not bits & (1024 | 2048)
    gs, ga = st.grad.cpu().numpy(), at.grad.cpu().numpy()
    for w in range(B):
        rgs, rga, rc = ow.backprop_contact(s[w].astype(np.float64), a[w].astype(np.float64), g[w].astype(np.float64))
        assert rc >= 0
        assert rel_err(gs[w], rgs) < 1e-4 and rel_err(ga[w], rga) < 1e-4, w


def test_restitution_backward_matches_oracle(oracle_mod):
    """Restitution (ContactConstraint.cpp:410-436): bouncing worlds (status bit 1024) back-propagate through the second reverse sweep of
    k_cstep_bwd and match the oracle's dual-number Jacobian; non-bouncing worlds of the same batch take the usual path."""
    raw = load_raw("half_cheetah")
    raw.restitution[:] = 0.8
    world = nb.World.from_raw(raw)
    ow = ob.OracleContactWorld(raw)
    B = 16
    s, a = contact_inputs(raw, "half_cheetah", B, seed=4)
    s[::2, raw.ndof + 1] -= 1.5  # every other world falls fast enough to bounce
    g = np.random.default_rng(2).normal(size=(B, 2 * raw.ndof)).astype(np.float32)
    st = torch.tensor(s, device="cuda", requires_grad=True)
    at = torch.tensor(a, device="cuda", requires_grad=True)
    nb.reset_contact_cache(world)
    out = nb.timestep(world, st, at)
    status = world._lcp_cache["status"].cpu().numpy().copy()
    out.backward(torch.tensor(g, device="cuda"))
    bits = nb.check_contact_status(world)
    assert bits & 1024 and not bits & 2048
    assert int(((status & 1024) > 0).sum()) >= 4
    gs, ga = st.grad.cpu().numpy(), at.grad.cpu().numpy()
    for w in range(B):
        ro = ow.step_contact(s[w].astype(np.float64), a[w].astype(np.float64))
        assert rel_err(out[w].detach().cpu().numpy(), ro["next_state"]) < 1e-5
        rgs, rga, rc = ow.backprop_contact(s[w].astype(np.float64), a[w].astype(np.float64), g[w].astype(np.float64))
        assert rc >= 0
        assert rel_err(gs[w], rgs) < 1e-4 and rel_err(ga[w], rga) < 1e-4, (w, hex(status[w]))


def test_contact_host_entry_points_match_device_path():
    """nb2_step_forward_contact_host / nb2_step_backward_contact_host (host buffers, copies inside, solver cache in the model) give the bits of
    the device entry points, with the cache flowing from call to call."""
    raw = load_raw("half_cheetah")
    world = nb.World.from_raw(raw)
    dm = nb.device_model_for(world)
    B = 64
    s, a = contact_inputs(raw, "half_cheetah", B, seed=12)
    g = np.random.default_rng(3).normal(size=(B, 2 * raw.ndof)).astype(np.float32)
    # device path, two chained steps
    nb.reset_contact_cache(world)
    x0 = torch.tensor(s, device="cuda", requires_grad=True); u = torch.tensor(a, device="cuda")
    x1 = nb.timestep(world, x0, u)
    x1d = x1.detach().clone().requires_grad_(True)
    x2 = nb.timestep(world, x1d, u)
    x2.backward(torch.tensor(g, device="cuda"))
    # host path
    status = np.zeros(B, np.int32); sticky = np.zeros(B, np.int32)
    h1 = dm.forward_contact_host(s, a, keep_for_backward=False, reset_cache=True)
    h2 = dm.forward_contact_host(h1, a, keep_for_backward=True, status_out=status)
    gs, ga = dm.backward_contact_host(g, sticky_out=sticky)
    assert np.array_equal(h1, x1.detach().cpu().numpy()) and np.array_equal(h2, x2.detach().cpu().numpy())
    assert np.array_equal(gs, x1d.grad.cpu().numpy())
    assert np.array_equal(status, world._lcp_cache["status"].cpu().numpy())
    assert (sticky & 2048).sum() == 0 and sticky.any()


def test_box_box_face_face_annotation_golden_on_device():
    """unittests/unit/test_DARTCollide.cpp:554-592 (BOX_BOX_FACE_FACE_COLLISION_ANNOTATION), on the GPU: 4 contacts, (+-0.25, 0.5, 0) EDGE_EDGE,
    (+-0.25, 0.25, 0) FACE_VERTEX (the CPU twin of this test pins the oracle and the host build: tests/test_golden_collide.py)."""
    from tests.test_golden_collide import _two_body_world

    world = _two_body_world(nb.BoxShape([1.0, 1.0, 1.0]), [0.0, 0.0, -0.5], nb.BoxShape([0.5, 0.5, 0.5]))
    n = world.getNumDofs()
    s = torch.zeros(3, 2 * n, device="cuda"); s[:, 3:6] = torch.tensor([0.0, 0.5, 0.25], device="cuda")
    with torch.no_grad():
        nb.timestep(world, s, torch.zeros(3, world.getActionSize(), device="cuda"))
    c = world._lcp_cache
    assert c["nc"].cpu().tolist() == [4, 4, 4]
    ci = c["cinfo"][1, :4].cpu().numpy()
    seen = {(round(float(r[0]), 6), round(float(r[1]), 6)): int(r[9]) for r in ci}
    assert seen == {(0.25, 0.5): 3, (-0.25, 0.5): 3, (0.25, 0.25): 2, (-0.25, 0.25): 2}, seen
    assert np.allclose(ci[:, 2], 0, atol=1e-6) and np.allclose(ci[:, 6], 0, atol=1e-6)
    # the reference-style accessor (World::getLastCollisionResult, pybind World.cpp:247-251)
    res = world.getLastCollisionResult(1)
    assert res.getNumContacts() == 4 and res.isCollision()
    assert sorted(c.type for c in res.getContacts()) == [2, 2, 3, 3]
    assert all(c.bodyNodeA is not None and c.bodyNodeB is not None and abs(c.penetrationDepth) < 1e-6 for c in res.getContacts())


def test_round_shape_contacts_on_device(oracle_mod):
    """sphere-sphere, capsule-capsule, sphere-capsule pairs on the GPU against the oracle: contact set, labels, next state, gradients (the CPU twin
    with the world description: tests/test_contact_emul.py::test_round_shape_contacts_forward_and_backward)."""
    from scipy.spatial.transform import Rotation

    w = nb.World(); w.setGravity([0, 0, -9.81]); w.setTimeStep(1e-3)
    g = nb.Skeleton("fixed"); g.setMobile(False)
    j, b = g.createWeldJointAndBodyNodePair()
    sn = b.createShapeNode(nb.CapsuleShape(0.2, 2.0)); sn.createCollisionAspect()
    T = nb.Isometry3(); T.set_rotation(Rotation.from_rotvec([0, np.pi / 2, 0]).as_matrix()); sn.setRelativeTransform(T.matrix())
    sn2 = b.createShapeNode(nb.SphereShape(0.3)); sn2.createCollisionAspect()
    T2 = nb.Isometry3(); T2.set_translation([0.0, 2.0, 0.0]); sn2.setRelativeTransform(T2.matrix())
    w.addSkeleton(g)
    for k, shp in enumerate([nb.CapsuleShape(0.15, 1.0), nb.SphereShape(0.25), nb.SphereShape(0.2)]):
        s = nb.Skeleton(f"m{k}")
        j, b = s.createFreeJointAndBodyNodePair(); b.setMass(1.0 + 0.3 * k); b.setMomentOfInertia(0.05, 0.06, 0.04)
        b.createShapeNode(shp).createCollisionAspect()
        w.addSkeleton(s)
    raw = nb.flatten_world(w)
    n = raw.ndof
    rng = np.random.default_rng(7)
    B = 16
    S = np.zeros((B, 2 * n), np.float32)
    for k in range(B):
        S[k, 0:3] = [np.pi / 2 + rng.normal(0, 0.02), rng.normal(0, 0.02), rng.normal(0, 0.02)]
        S[k, 3:6] = [0.3 + rng.normal(0, 0.01), rng.normal(0, 0.01), 0.2 + 0.15 - 0.004 + rng.normal(0, 0.001)]
        S[k, 9:12] = [-0.5 + rng.normal(0, 0.01), rng.normal(0, 0.005), 0.2 + 0.25 - 0.003 + rng.normal(0, 0.001)]
        S[k, 15:18] = [rng.normal(0, 0.01), 2.0 + rng.normal(0, 0.01), 0.3 + 0.2 - 0.003 + rng.normal(0, 0.001)]
        S[k, n:] = rng.normal(0, 0.05, n)
    A = np.zeros((B, len(raw.action_map)), np.float32)
    gr = rng.normal(size=(B, 2 * n)).astype(np.float32)
    st = torch.tensor(S, device="cuda", requires_grad=True); at = torch.tensor(A, device="cuda", requires_grad=True)
    out = nb.timestep(w, st, at)
    c = w._lcp_cache
    labels, mm, ncs = c["labels"].cpu().numpy(), c["m"].cpu().numpy(), c["nc"].cpu().numpy()
    out.backward(torch.tensor(gr, device="cuda"))
    nb.check_contact_status(w)
    ow = ob.OracleContactWorld(raw)
    for k in range(B):
        ro = ow.step_contact(S[k].astype(np.float64), A[k].astype(np.float64))
        assert ncs[k] == ro["nc"] and ncs[k] >= 3 and mm[k] == ro["m"]
        assert np.array_equal(labels[k][: mm[k]], ro["mapping"][: ro["m"]])
        assert rel_err(out[k].detach().cpu().numpy(), ro["next_state"]) < 1e-5
        rgs, rga, rc = ow.backprop_contact(S[k].astype(np.float64), A[k].astype(np.float64), gr[k].astype(np.float64))
        assert rc >= 0 and rel_err(st.grad[k].cpu().numpy(), rgs) < 1e-4


def test_self_collision_on_device(oracle_mod):
    """Skeleton::enableSelfCollisionCheck(): rows between two bodies of the same tree, GPU against the oracle (world and the CPU twin:
    tests/test_contact_emul.py::test_self_collision_pairs_forward_and_backward)."""
    from tests.test_contact_emul import _folding_arm_world

    w = _folding_arm_world()
    raw = nb.flatten_world(w)
    n = raw.ndof
    rng = np.random.default_rng(3)
    B = 12
    S = np.zeros((B, 2 * n), np.float32)
    for k in range(B):
        S[k, 0:3] = rng.normal(0, 0.05, 3)
        S[k, 6] = 2.0944 + rng.normal(0, 0.003)
        S[k, 7] = 1.955 + rng.normal(0, 0.004)   # an edge of link 3 pressed 0.3 - 1.4 cm into link 1 (deeper contacts are clipped away)
        S[k, n:] = rng.normal(0, 0.1, n)
    A = rng.normal(0, 1.0, (B, len(raw.action_map))).astype(np.float32)
    gr = rng.normal(size=(B, 2 * n)).astype(np.float32)
    st = torch.tensor(S, device="cuda", requires_grad=True); at = torch.tensor(A, device="cuda", requires_grad=True)
    out = nb.timestep(w, st, at)
    c = w._lcp_cache
    labels, mm, ncs = c["labels"].cpu().numpy(), c["m"].cpu().numpy(), c["nc"].cpu().numpy()
    out.backward(torch.tensor(gr, device="cuda"))
    nb.check_contact_status(w)
    ow = ob.OracleContactWorld(raw)
    with_rows = 0
    for k in range(B):
        ro = ow.step_contact(S[k].astype(np.float64), A[k].astype(np.float64))
        assert ncs[k] == ro["nc"] and mm[k] == ro["m"] and np.array_equal(labels[k][: mm[k]], ro["mapping"][: ro["m"]])
        assert rel_err(out[k].detach().cpu().numpy(), ro["next_state"]) < 1e-5
        rgs, rga, rc = ow.backprop_contact(S[k].astype(np.float64), A[k].astype(np.float64), gr[k].astype(np.float64))
        assert rc >= 0 and rel_err(st.grad[k].cpu().numpy(), rgs) < 1e-4 and rel_err(at.grad[k].cpu().numpy(), rga) < 1e-4
        with_rows += int(ro["m"] > 0)
    assert with_rows >= 6


def test_contact_path_edge_cases():
    """Empty and ragged batches, the large-workspace retry (same bits as the shared-memory path), and the loud failure when a world
    exceeds the compiled contact limit (the reference has no such limit: dropping contacts silently would not be parity)."""
    raw = load_raw("half_cheetah")
    world = nb.World.from_raw(raw)
    n2, na = 2 * raw.ndof, len(raw.action_map)
    # B = 0
    out = nb.timestep(world, torch.zeros(0, n2, device="cuda"), torch.zeros(0, na, device="cuda"))
    assert out.shape == (0, n2)
    # ragged batches: B = 1, 33 (one warp + 1), 67 give the rows of a B = 67 run
    s, a = contact_inputs(raw, "half_cheetah", 67, seed=21)
    g = np.random.default_rng(4).normal(size=s.shape).astype(np.float32)

    def run(B, cap=None):
        w = nb.World.from_raw(raw)
        if cap:
            nb.device_model_for(w).set_contact_capacity(cap)
        st = torch.tensor(s[:B], device="cuda", requires_grad=True); at = torch.tensor(a[:B], device="cuda", requires_grad=True)
        o = nb.timestep(w, st, at)
        o.backward(torch.tensor(g[:B], device="cuda"))
        nb.check_contact_status(w)
        return o.detach().cpu().numpy(), st.grad.cpu().numpy(), at.grad.cpu().numpy()

    full = run(67)
    for B in (1, 33):
        part = run(B)
        for x, y in zip(part, full):
            assert np.array_equal(x, y[:B]), B
    pool = run(30, cap=1)  # every world with more than one contact retries in the global pool (32 slots at this batch size)
    for x, y in zip(pool, full):
        assert np.array_equal(x, y[:30])
    # ... and when the pool itself runs out (67 worlds, 32 slots) the step says so instead of dropping contacts silently
    with pytest.raises(RuntimeError, match="DROPPED|backward through the contact stage failed"):
        run(67, cap=1)
    # more contacts than NB2_MAX_CONTACTS (16): five boxes on the ground = 20 -> status bit 256, check_contact_status raises
    w = nb.World(); w.setGravity([0, -9.81, 0])
    g0 = nb.Skeleton("ground"); g0.setMobile(False)
    j, b = g0.createWeldJointAndBodyNodePair()
    b.createShapeNode(nb.BoxShape([10, 0.2, 10])).createCollisionAspect()
    T = nb.Isometry3(); T.set_translation([0, -0.1, 0]); j.setTransformFromParentBodyNode(T)
    w.addSkeleton(g0)
    for k in range(5):
        sk = nb.Skeleton(f"box{k}")
        j, b = sk.createFreeJointAndBodyNodePair(); b.setMass(1.0)
        b.createShapeNode(nb.BoxShape([0.3, 0.3, 0.3])).createCollisionAspect()
        w.addSkeleton(sk)
    n = w.getNumDofs()
    st = torch.zeros(2, 2 * n, device="cuda")
    for k in range(5):
        st[:, 6 * k + 3] = 1.0 * k; st[:, 6 * k + 4] = 0.15 - 0.002
    with torch.no_grad():
        nb.timestep(w, st, torch.zeros(2, w.getActionSize(), device="cuda"))
    with pytest.raises(RuntimeError, match="DROPPED"):
        nb.check_contact_status(w)


def test_joint_limit_rows_on_device(oracle_mod):
    """Joint limits enforced (constraint/JointLimitConstraint.cpp) on the GPU against the oracle: half-cheetah with contacts + limit rows, and the
    shape-less cartpole whose rail stops at its limit (CPU twin: tests/test_contact_emul.py::test_joint_limit_rows_forward_and_backward)."""
    raw = load_raw("half_cheetah")
    raw.limit_enforced[:] = 1
    raw.spring[:] = 0.0
    n = raw.ndof
    B = 16
    s, a = contact_inputs(raw, "half_cheetah", B, seed=9)
    rng = np.random.default_rng(2)
    for k in range(B):
        for d in rng.choice(np.arange(3, n), size=2, replace=False):
            hi = rng.random() < 0.5
            s[k, d] = (raw.pos_hi[d] + 0.01 * rng.random()) if hi else (raw.pos_lo[d] - 0.01 * rng.random())
            s[k, n + d] = (1.0 if hi else -1.0) * rng.choice([8.0, -0.5])
    a *= 0.0
    world = nb.World.from_raw(raw)
    assert all(j.isPositionLimitEnforced() for sk in world.skeletons for b in sk._ordered_bodies() for j in [b.parent_joint])
    gr = rng.normal(size=(B, 2 * n)).astype(np.float32)
    st = torch.tensor(s, device="cuda", requires_grad=True); at = torch.tensor(a, device="cuda", requires_grad=True)
    out = nb.timestep(world, st, at)
    c = world._lcp_cache
    labels, mm = c["labels"].cpu().numpy(), c["m"].cpu().numpy()
    out.backward(torch.tensor(gr, device="cuda"))
    nb.check_contact_status(world)
    ow = ob.OracleContactWorld(raw)
    lim = 0
    for k in range(B):
        ro = ow.step_contact(s[k].astype(np.float64), a[k].astype(np.float64))
        assert mm[k] == ro["m"] and np.array_equal(labels[k][: mm[k]], ro["mapping"][: ro["m"]]), k
        assert rel_err(out[k].detach().cpu().numpy(), ro["next_state"]) < 1e-5
        rgs, rga, rc = ow.backprop_contact(s[k].astype(np.float64), a[k].astype(np.float64), gr[k].astype(np.float64))
        assert rc >= 0 and rel_err(st.grad[k].cpu().numpy(), rgs) < 1e-4 and rel_err(at.grad[k].cpu().numpy(), rga) < 1e-4
        lim += int(sum(t >= 100 for t in ro["type"]))
    assert lim >= 20
    # cartpole: no shapes at all, the rail limit is the only constraint
    w2 = nb.World.from_raw(load_raw("cartpole"))
    rail = next(b.parent_joint for sk in w2.skeletons for b in sk._ordered_bodies() if b.parent_joint.getNumDofs() == 1)  # the prismatic rail
    rail.setPositionLowerLimit(0, -1.0); rail.setPositionUpperLimit(0, 1.0); rail.setPositionLimitEnforced(True)
    s2 = torch.tensor([[1.0, 0.3, 1.5, 0.2], [0.5, 0.3, 1.0, 0.2]], device="cuda")
    with torch.no_grad():
        o2 = nb.timestep(w2, s2, torch.zeros(2, w2.getActionSize(), device="cuda"))
    assert abs(float(o2[0, 2])) < 1e-7 and float(o2[1, 2]) > 0.9   # stopped at the limit / free
