"""Contact / boxed-LCP stage: the device code (csrc/nb2_cw.cuh, compiled for the host by tests/host_emul) vs the
fp64 oracle over several chained steps with the LCP cache flowing from step to step.
Bar: contact set, LCP dimension, per-row ConstraintMapping labels and the solver-branch status word are BIT-EXACT;
impulses and next state within 1e-4 relative (inputs are fp32 rows on the device side)."""
import numpy as np
import pytest

import nimblephysics_b200 as nb
from oracle import binding as ob
from tests.host_emul.binding import EmulWorld
from tests.util import contact_inputs, load_raw, rel_err


@pytest.mark.parametrize("name", ["half_cheetah", "atlas_ground"])
def test_contact_forward_chain_matches_oracle(oracle_mod, name):
    raw = load_raw(name)
    cm = nb.compile_model(raw)
    ow, ew = ob.OracleContactWorld(raw), EmulWorld(cm)
    B, T = 6, 6
    s, a = contact_inputs(raw, name, B, seed=3)
    so, se = s.astype(np.float64).copy(), s.copy()
    xo = [None] * B
    xe = me = None
    seen_status = set()
    for t in range(T):
        r = ew.forward_contact(se, a, xe, me)
        for w in range(B):
            ro = ow.step_contact(so[w], a[w].astype(np.float64), xo[w])
            mo = ro["m"]
            assert r["nc"][w] == ro["nc"] and r["m"][w] == mo
            assert np.array_equal(r["labels"][w][:mo], ro["mapping"])          # bit-exact contact set / classification
            assert (r["status"][w] & ~96) == (ro["status"] & ~96)               # same solver branch (bits 32/64: see test_gpu_contact)
            assert np.array_equal(r["cinfo"][w][: ro["nc"], 7:9].astype(int), ro["bodies"])
            assert np.array_equal(r["cinfo"][w][: ro["nc"], 9].astype(int), ro["type"])
            if mo:
                assert np.abs(r["x"][w][:mo] - ro["x"]).max() < 1e-5 * max(1.0, np.abs(ro["x"]).max())
            assert rel_err(r["next"][w], ro["next_state"]) < 1e-4
            so[w] = ro["next_state"]
            xo[w] = ro["x"] if mo else None
            seen_status.add(int(ro["status"]))
        # feed the oracle's fp64 state to both sides so rounding of the fp32 rows does not accumulate into tie flips
        se = so.astype(np.float32)
        so = se.astype(np.float64)
        xe, me = r["x"], r["m"]
    assert len(seen_status) >= 1


def test_oracle_contact_step_invariants(oracle_mod):
    """Properties the reference's own tests rely on (unittests/GradientTestUtils.hpp verifyRecoveredLCPConstraints,
    LCPUtils::isLCPSolutionValid): a solution that did not need the friction-drop fallback is a valid boxed-LCP
    solution; clamping normal rows end with ~zero relative normal velocity; positions advance with the pre-step velocity."""
    for name in ("half_cheetah", "atlas_ground"):
        raw = load_raw(name)
        ow = ob.OracleContactWorld(raw)
        s, a = contact_inputs(raw, name, 6, seed=9)
        total_contacts = 0
        for w in range(6):
            s64, a64 = s[w].astype(np.float64), a[w].astype(np.float64)
            r = ow.step_contact(s64, a64)
            n = raw.ndof
            total_contacts += r["nc"]
            if r["m"] == 0:
                continue
            if not (r["status"] & 16):
                cfm = 1e-4 if (r["status"] & 8) else 0.0
                assert ob.lcp_valid(r["A"] + cfm * np.eye(r["m"]), r["x"], r["b"], r["hi"], r["lo"], r["findex"])
            assert np.allclose(r["A"], r["A"].T, atol=1e-9)
            w_ = r["A"] @ r["x"] - r["b"]
            for j in range(r["m"]):
                if r["mapping"][j] == -2 and r["findex"][j] == -1 and not (r["status"] & 24):
                    assert abs(w_[j]) < 1e-6
            # generic joints: q+ = q + dt * v_t (World.cpp:307-322)
            q, v = s64[:n], s64[n:]
            idx = [d for d in range(n) if not (name == "atlas_ground" and d < 6)]
            assert np.allclose(r["next_state"][:n][idx], (q + raw.dt * v)[idx], atol=1e-12)
        assert total_contacts > 0


def _check_backward(ob, raw, S, A_, tol=1e-5):
    cm = nb.compile_model(raw)
    ow, ew = ob.OracleContactWorld(raw), EmulWorld(cm)
    n = raw.ndof
    B = S.shape[0]
    g = np.random.default_rng(1).normal(size=(B, 2 * n)).astype(np.float32)
    r = ew.forward_contact(S, A_)
    gs, ga = ew.backward_contact(S, A_, r["saved"], r["crec"], g)
    kinds = []
    for w in range(B):
        rgs, rga, rc = ow.backprop_contact(S[w].astype(np.float64), A_[w].astype(np.float64), g[w].astype(np.float64))
        assert rc >= 0
        lab = r["labels"][w][: r["m"][w]]
        kinds.append((int((lab == -2).sum()), int((lab >= 0).sum()), int(r["status"][w])))
        assert rel_err(gs[w], rgs) < tol and rel_err(ga[w], rga) < tol, (w, kinds[-1])
    return kinds


def test_contact_backward_adjoint_matches_oracle_jacobian(oracle_mod):
    """The device backward through the contact stage (adjoint form + dual-number contact geometry, csrc/nb2_cw.cuh)
    vs J^T g with the oracle's forward-mode Jacobian of the same frozen-classification step.  Covers: clamping rows only,
    rank-deficient Q (24 clamping rows of a standing Atlas, rank 12), fallback-cfm solutions, sliding (upper-bound) rows."""
    # seeded contact-rich batches
    for name in ("half_cheetah", "atlas_ground"):
        raw = load_raw(name)
        s, a = contact_inputs(raw, name, 6, seed=4)
        _check_backward(ob, raw, s, a)
    # standing Atlas: every row clamping, Q rank-deficient
    raw = load_raw("atlas_ground")
    n = raw.ndof
    S = np.zeros((2, 2 * n), np.float32)
    S[:, 0] = -0.5 * np.pi
    S[:, 4] = [-0.01, -0.012]
    S[1, n + 3] = 0.01
    kinds = _check_backward(ob, raw, S, np.zeros((2, n), np.float32))
    assert any(k[0] == 24 for k in kinds)
    # low friction + tangential velocity: sliding friction rows (UPPER_BOUND labels, non-symmetric Q)
    raw = load_raw("half_cheetah")
    raw.friction[:] = 0.2
    n = raw.ndof
    S = np.zeros((3, 2 * n), np.float32)
    S[:, 1], S[:, 2] = -0.1, 0.03
    S[:, n] = [0.5, 2.0, -1.5]
    kinds = _check_backward(ob, raw, S, np.zeros((3, n), np.float32))
    assert any(k[1] > 0 for k in kinds)


def test_oracle_contact_jacobian_matches_finite_differences(oracle_mod):
    """Reference-style pin of the oracle itself (GradientTestUtils.hpp verifyVelGradients): when the forward solution is an
    exact LCP solution (short-circuit or Dantzig), the frozen-classification Jacobian equals central differences of the full
    step wherever the perturbation does not flip a label."""
    cases = []
    raw = load_raw("half_cheetah")
    s, a = contact_inputs(raw, "half_cheetah", 3, seed=4)
    cases.append((raw, s, a))
    raw2 = nb.flatten_world(_box_stack_world())  # rows acting on two moving bodies
    s2, a2 = _box_stack_inputs(raw2, 12, seed=0)  # a few of these worlds get an exact (Dantzig) answer
    cases.append((raw2, s2, a2))
    checked = []
    for raw, s, a in cases:
        ow = ob.OracleContactWorld(raw)
        n = raw.ndof
        cols = 0
        for w in range(s.shape[0]):
            s64, a64 = s[w].astype(np.float64), a[w].astype(np.float64)
            r0 = ow.step_contact(s64, a64)
            if r0["status"] & 24:
                continue  # PGS / friction-drop answers are approximate: the analytic map is not their derivative
            J, rc = ow.jacobian_contact(s64, a64)
            assert rc >= 0
            eps = 1e-7
            for c in range(2 * n):
                sp, sm = s64.copy(), s64.copy()
                sp[c] += eps
                sm[c] -= eps
                rp, rm = ow.step_contact(sp, a64), ow.step_contact(sm, a64)
                if not (np.array_equal(rp["mapping"], r0["mapping"]) and np.array_equal(rm["mapping"], r0["mapping"])):
                    continue
                fd = (rp["next_state"] - rm["next_state"]) / (2 * eps)
                assert np.abs(fd - J[:, c]).max() < 2e-6 * max(1.0, np.abs(J).max())
                cols += 1
        checked.append(cols)
    assert all(c > 0 for c in checked), checked


def _box_stack_world():
    """static ground + a free box resting on it + a smaller free box resting on the first: ground-box contacts and
    contacts between two MOVING bodies of different skeletons (ConstraintSolver.cpp:723-793 puts them in one group)."""
    w = nb.World()
    w.setGravity([0, -9.81, 0])
    w.setTimeStep(1e-3)
    g = nb.Skeleton("ground")
    g.setMobile(False)
    j, b = g.createWeldJointAndBodyNodePair()
    b.createShapeNode(nb.BoxShape([4, 0.2, 4])).createCollisionAspect()
    T = nb.Isometry3()
    T.set_translation([0, -0.1, 0])
    j.setTransformFromParentBodyNode(T)
    w.addSkeleton(g)
    for k, size in enumerate(([0.6, 0.4, 0.6], [0.3, 0.3, 0.3])):
        s = nb.Skeleton(f"box{k}")
        j, b = s.createFreeJointAndBodyNodePair()
        b.setMass(2.0 - k)
        b.createShapeNode(nb.BoxShape(size)).createCollisionAspect()
        w.addSkeleton(s)
    return w


def test_contacts_between_two_moving_bodies_forward(oracle_mod):
    raw = nb.flatten_world(_box_stack_world())
    cm = nb.compile_model(raw)
    n = raw.ndof
    rng = np.random.default_rng(0)
    B = 12
    S = np.zeros((B, 2 * n), np.float32)
    for w in range(B):
        S[w, 0:3] = rng.normal(0, 0.01, 3)
        S[w, 3:6] = [rng.normal(0, 0.01), 0.2 - 0.002, rng.normal(0, 0.01)]
        S[w, 6:9] = rng.normal(0, 0.01, 3)
        S[w, 9:12] = [rng.normal(0, 0.02), 0.55 - 0.004, rng.normal(0, 0.02)]
        S[w, n:] = rng.normal(0, 0.05, n)
    A = np.zeros((B, len(raw.action_map)), np.float32)
    r = EmulWorld(cm).forward_contact(S, A)
    ow = oracle_mod.OracleContactWorld(raw)
    both = 0
    for w in range(B):
        ro = ow.step_contact(S[w].astype(np.float64), A[w].astype(np.float64))
        m = int(r["m"][w])
        assert r["nc"][w] == ro["nc"] and m == ro["m"]
        assert (r["status"][w] & ~96) == (ro["status"] & ~96)
        assert np.array_equal(r["labels"][w][:m], ro["mapping"][:m])  # contact set and labels bit-exact
        assert rel_err(r["next"][w], ro["next_state"]) < 1e-6
        both += int(any(a >= 1 and b_ >= 1 for a, b_ in ro["bodies"].tolist()))
    assert both >= 6  # box-on-box contacts were really in play


def _box_stack_inputs(raw, B, seed):
    n = raw.ndof
    rng = np.random.default_rng(seed)
    S = np.zeros((B, 2 * n), np.float32)
    for w in range(B):
        S[w, 0:3] = rng.normal(0, 0.01, 3)
        S[w, 3:6] = [rng.normal(0, 0.01), 0.2 - 0.002, rng.normal(0, 0.01)]
        S[w, 6:9] = rng.normal(0, 0.01, 3)
        S[w, 9:12] = [rng.normal(0, 0.02), 0.55 - 0.004, rng.normal(0, 0.02)]
        S[w, n:] = rng.normal(0, 0.05, n)
    return S, np.zeros((B, len(raw.action_map)), np.float32)


def test_contacts_between_two_moving_bodies_backward(oracle_mod):
    """Rows whose wrench acts on TWO moving bodies (box resting on a box): the adjoint differentiates each wrench with
    respect to the pose of either body (csrc/nb2_cw.cuh contact_backward, dual pass) — checked against
    J^T g with the oracle's dual-number Jacobian of the frozen-classification step, and the oracle against central
    differences where the forward answer is exact."""
    raw = nb.flatten_world(_box_stack_world())
    S, A_ = _box_stack_inputs(raw, 8, seed=3)
    kinds = _check_backward(ob, raw, S, A_, tol=2e-5)
    assert len(kinds) == 8
    # low friction: sliding rows between the two boxes
    raw.friction[:] = 0.15
    S, A_ = _box_stack_inputs(raw, 6, seed=5)
    n = raw.ndof
    S[:, n + 9] += 0.8  # the upper box slides along x
    kinds = _check_backward(ob, raw, S, A_, tol=2e-5)
    assert any(k[1] > 0 for k in kinds)


def test_mass_gradient_through_the_contact_stage(oracle_mod):
    """lossWrtMass with active contact constraints: the same inertia-parameter form as the contact-free step, evaluated with
    the field of w = lambda - nu and the realised acceleration (the constraint rows do not depend on the inertias).
    Checked against central differences of the oracle's full contact step on worlds whose forward answer is exact and whose
    labels do not move under the perturbation."""
    import copy

    from nimblephysics_b200 import modelspec as ms

    raw = load_raw("half_cheetah")
    cm = nb.compile_model(raw)
    ew = EmulWorld(cm)
    mobile = [i for i in range(raw.nb) if cm.body_owner[i] >= 0]
    entries = [(mobile[1], ms.INERTIA_MASS), (mobile[3], ms.INERTIA_COM), (mobile[5], ms.INERTIA_MASS), (mobile[2], ms.INERTIA_DIAGONAL)]
    P = ms.inertia_param_jacobian(raw, cm, entries)
    s, a = contact_inputs(raw, "half_cheetah", 10, seed=4)
    g = np.random.default_rng(3).normal(size=s.shape).astype(np.float32)
    g[:, :raw.ndof] = 0  # positions do not depend on the masses within one step
    r = ew.forward_contact(s, a)
    gs, ga, gi = ew.backward_contact(s, a, r["saved"], r["crec"], g, want_inertia_grad=True)
    checked = 0
    for w in range(s.shape[0]):
        if r["status"][w] & (8 | 16 | 64) or r["m"][w] == 0:
            continue  # approximate forward answers are not differentiable maps of their inputs
        s64, a64, g64 = s[w].astype(np.float64), a[w].astype(np.float64), g[w].astype(np.float64)
        lab0 = r["labels"][w][: r["m"][w]]

        def at(row, h):
            rr = copy.deepcopy(raw)
            k = 0
            for (bi, kind) in entries:
                x = ms._mass_entry_value(kind, rr.mass[bi], rr.com[bi], rr.moment[bi]).astype(np.float64)
                for j in range(len(x)):
                    if k == row:
                        x[j] += h
                    k += 1
                rr.mass[bi], rr.com[bi], rr.moment[bi] = ms._apply_mass_entry(kind, x, rr.mass[bi], rr.com[bi], rr.moment[bi])
            ro = ob.OracleContactWorld(rr).step_contact(s64, a64)
            return float(g64 @ ro["next_state"]), ro["mapping"]

        gm = P @ gi[:, w].astype(np.float64)
        ok = True
        fd = np.zeros(P.shape[0])
        for j in range(P.shape[0]):
            (lp, mp), (lm, mm_) = at(j, 1e-5), at(j, -1e-5)
            ok = ok and np.array_equal(mp, lab0) and np.array_equal(mm_, lab0)
            fd[j] = (lp - lm) / 2e-5
        if not ok:
            continue
        assert rel_err(gm, fd) < 1e-4, (w, gm, fd)
        checked += 1
    assert checked >= 2, checked


def test_restitution_forward_and_backward(oracle_mod):
    """Bounce terms (ContactConstraint.cpp:410-436): b_normal = (1 + e)(-J v*) on the rows that bounce.  Forward parity with the oracle, and
    the backward — second reverse sweep with the field of -nu_e, v* injections (csrc/nb2_cw.cuh contact_backward / bounce_pass2_*) — against
    the oracle's dual-number Jacobian of the frozen-classification step (the reference: getBounceDiagonals, BackpropSnapshot.cpp:2624-2680)."""
    raw = load_raw("half_cheetah")
    raw.restitution[:] = 0.8
    cm = nb.compile_model(raw)
    ew, ow = EmulWorld(cm), ob.OracleContactWorld(raw)
    s, a = contact_inputs(raw, "half_cheetah", 8, seed=4)
    s[:, raw.ndof + 1] -= 1.5  # falling fast: e * (relative normal velocity) > 0.1 activates the bounce term
    r = ew.forward_contact(s, a)
    bounced = 0
    for w in range(8):
        ro = ow.step_contact(s[w].astype(np.float64), a[w].astype(np.float64))
        assert (r["status"][w] & ~96) == (ro["status"] & ~96) and rel_err(r["next"][w], ro["next_state"]) < 1e-6
        bounced += int(bool(r["status"][w] & 1024))
    assert bounced >= 4
    kinds = _check_backward(ob, raw, s, a, tol=2e-5)
    assert sum(1 for k in kinds if k[2] & 1024 and k[0] > 0) >= 3  # bouncing worlds with clamping rows went through the second sweep
    # reversed lane order and the large-workspace retry give the same gradients
    g = np.random.default_rng(1).normal(size=s.shape).astype(np.float32)
    gs0, ga0 = ew.backward_contact(s, a, r["saved"], r["crec"], g)
    gs1, ga1 = ew.backward_contact(s, a, r["saved"], r["crec"], g, small_mc=1, reverse=True)
    assert np.allclose(gs0, gs1, rtol=1e-6, atol=1e-7) and np.allclose(ga0, ga1, rtol=1e-6, atol=1e-7)
    # Atlas standing on restitutive ground, dropped: rank-deficient Q, two feet
    raw = load_raw("atlas_ground")
    raw.restitution[:] = 0.5
    s, a = contact_inputs(raw, "atlas_ground", 4, seed=4)
    s[:, raw.ndof + 5] -= 1.0  # the root's body-frame z is the world's vertical in this pose
    kinds = _check_backward(ob, raw, s, a, tol=5e-5)
    assert any(k[2] & 1024 for k in kinds)


def test_penetration_correction_forward_and_backward(oracle_mod):
    """World::setPenetrationCorrectionEnabled(true): b_normal gains min((depth - allowance) * erp / dt, cap) (ContactConstraint.cpp:395-408).
    Forward parity with the oracle, status bit 4096, and the backward — including the depth derivative of the uncapped correction, which
    the dual pass of contact_backward adds — against the oracle's dual-number Jacobian of the same frozen-classification step."""
    raw = load_raw("half_cheetah")
    raw.penetration_correction = True
    cm = nb.compile_model(raw)
    ew, ow = EmulWorld(cm), ob.OracleContactWorld(raw)
    s, a = contact_inputs(raw, "half_cheetah", 10, seed=6)
    s[5:, 1] += 0.0099
    s[5, 1] += 0.00585  # the first contact of world 5 becomes shallower than 1e-4: its correction stays UNDER the cap (depth * 10 < 1e-3)
    r = ew.forward_contact(s, a)
    d5 = r["cinfo"][5, : r["nc"][5], 6]
    assert 0 < d5[0] < 1e-4 and r["labels"][5][0] == -2, (d5, r["labels"][5][:6])  # ... and its normal row is CLAMPING
    seen = 0
    for w in range(s.shape[0]):
        ro = ow.step_contact(s[w].astype(np.float64), a[w].astype(np.float64))
        assert rel_err(r["next"][w], ro["next_state"]) < 1e-6
        if r["status"][w] & 4096:
            seen += 1
        assert not (r["status"][w] & 1024)
    assert seen >= 3
    kinds = _check_backward(ob, raw, s, a, tol=2e-5)
    assert len(kinds) == s.shape[0]


def test_round_shape_contacts_forward_and_backward(oracle_mod):
    """sphere-sphere, capsule-capsule and sphere-capsule pairs (DARTCollide.cpp:1812-1882, 4183-4420): forward against the oracle (contact set,
    labels, next state) and the backward — the dual-number contact generator differentiates them like any other pair — against the oracle's
    Jacobian.  World: a static capsule lying along x and a static sphere; a free capsule crosses the static one (pipe-pipe), a free sphere
    sits on the static capsule's side (pipe-sphere) and touches the free capsule's end region, a second free sphere rests on the static sphere."""
    from scipy.spatial.transform import Rotation

    w = nb.World()
    w.setGravity([0, 0, -9.81])
    w.setTimeStep(1e-3)
    g = nb.Skeleton("fixed"); g.setMobile(False)
    j, b = g.createWeldJointAndBodyNodePair()
    sn = b.createShapeNode(nb.CapsuleShape(0.2, 2.0)); sn.createCollisionAspect()
    T = nb.Isometry3(); T.set_rotation(Rotation.from_rotvec([0, np.pi / 2, 0]).as_matrix()); sn.setRelativeTransform(T.matrix())
    sn2 = b.createShapeNode(nb.SphereShape(0.3)); sn2.createCollisionAspect()
    T2 = nb.Isometry3(); T2.set_translation([0.0, 2.0, 0.0]); sn2.setRelativeTransform(T2.matrix())
    w.addSkeleton(g)
    shapes = [nb.CapsuleShape(0.15, 1.0), nb.SphereShape(0.25), nb.SphereShape(0.2)]
    for k, shp in enumerate(shapes):
        s = nb.Skeleton(f"m{k}")
        j, b = s.createFreeJointAndBodyNodePair(); b.setMass(1.0 + 0.3 * k)
        b.setMomentOfInertia(0.05, 0.06, 0.04)
        b.createShapeNode(shp).createCollisionAspect()
        w.addSkeleton(s)
    raw = nb.flatten_world(w)
    n = raw.ndof
    rng = np.random.default_rng(7)
    B = 8
    S = np.zeros((B, 2 * n), np.float32)
    for k in range(B):
        # free capsule: axis along y (rotated about x), crossing the static capsule from above
        S[k, 0:3] = [np.pi / 2 + rng.normal(0, 0.02), rng.normal(0, 0.02), rng.normal(0, 0.02)]
        S[k, 3:6] = [0.3 + rng.normal(0, 0.01), rng.normal(0, 0.01), 0.2 + 0.15 - 0.004 + rng.normal(0, 0.001)]
        # free sphere on the static capsule's side
        S[k, 9:12] = [-0.5 + rng.normal(0, 0.01), rng.normal(0, 0.005), 0.2 + 0.25 - 0.003 + rng.normal(0, 0.001)]
        # free sphere on the static sphere
        S[k, 15:18] = [rng.normal(0, 0.01), 2.0 + rng.normal(0, 0.01), 0.3 + 0.2 - 0.003 + rng.normal(0, 0.001)]
        S[k, n:] = rng.normal(0, 0.05, n)
    A = np.zeros((B, len(raw.action_map)), np.float32)
    cm = nb.compile_model(raw)
    r = EmulWorld(cm).forward_contact(S, A)
    ow = oracle_mod.OracleContactWorld(raw)
    types = set()
    for k in range(B):
        ro = ow.step_contact(S[k].astype(np.float64), A[k].astype(np.float64))
        assert r["nc"][k] == ro["nc"] and r["nc"][k] >= 3, (k, r["nc"][k], ro["nc"])
        assert np.array_equal(r["labels"][k][: r["m"][k]], ro["mapping"][: ro["m"]])
        assert rel_err(r["next"][k], ro["next_state"]) < 1e-6
        types |= set(int(t) for t in ro["type"])
    assert {6, 15} <= types and (13 in types or 14 in types), types
    kinds = _check_backward(ob, raw, S, A, tol=2e-5)
    assert len(kinds) == B


def _folding_arm_world(adjacent=False):
    """A free-floating 3-link chain with box links, folded so that link 3 presses on link 1 (same skeleton, not adjacent), above a ground box.
    Skeleton::enableSelfCollisionCheck() (off by default in the reference) makes the link 1 - link 3 pair visible to the narrow phase.
    (Floating base: a chain pinned to the world with parallel axes could not move out of its plane and Q would be singular by construction.)"""
    w = nb.World()
    w.setGravity([0, -9.81, 0])
    w.setTimeStep(1e-3)
    g = nb.Skeleton("ground"); g.setMobile(False)
    j, b = g.createWeldJointAndBodyNodePair()
    b.createShapeNode(nb.BoxShape([4, 0.2, 4])).createCollisionAspect()
    T = nb.Isometry3(); T.set_translation([0, -1.1, 0]); j.setTransformFromParentBodyNode(T)
    w.addSkeleton(g)
    arm = nb.Skeleton("arm")
    parent = None
    L = 0.5
    for k in range(3):
        j, b = arm.createRevoluteJointAndBodyNodePair(parent) if k else arm.createFreeJointAndBodyNodePair()
        if k:
            j.setAxis([0, 0, 1])
        if k:
            T = nb.Isometry3(); T.set_translation([L, 0, 0]); j.setTransformFromParentBodyNode(T)
        b.setMass(1.0); b.setLocalCOM([L / 2, 0, 0]); b.setMomentOfInertia(0.01, 0.02, 0.02)
        sn = b.createShapeNode(nb.BoxShape([L, 0.08, 0.1])); sn.createCollisionAspect()
        Ts = nb.Isometry3(); Ts.set_translation([L / 2, 0, 0])
        if k == 2:  # link 3's box is rolled about its long axis: no two box axes are parallel, so the SAT decision is not a coin toss of rounding
            from scipy.spatial.transform import Rotation
            Ts.set_rotation(Rotation.from_rotvec([0.35, 0, 0]).as_matrix())
        sn.setRelativeTransform(Ts.matrix())
        parent = b
    arm.enableSelfCollisionCheck()
    if adjacent:
        arm.enableAdjacentBodyCheck()
    w.addSkeleton(arm)
    return w


def test_self_collision_pairs_forward_and_backward(oracle_mod):
    """Self-collision (BodyNodeCollisionFilter, CollisionFilter.cpp:105-152): rows whose two bodies sit in the SAME tree.  Pair list (default off,
    non-adjacent only unless the adjacent-body check is enabled), forward against the oracle, backward against its Jacobian."""
    from nimblephysics_b200._cabi import collision_pairs

    w_off = _folding_arm_world(); w_off.getSkeleton(1).disableSelfCollisionCheck()
    w_on, w_adj = _folding_arm_world(), _folding_arm_world(adjacent=True)
    npairs = [len(collision_pairs(nb.compile_model(nb.flatten_world(w)))[0]) for w in (w_off, w_on, w_adj)]
    assert npairs == [3, 4, 6], npairs  # ground x 3 links | + link 1 - link 3 | + the two adjacent pairs
    raw = nb.flatten_world(w_on)
    n = raw.ndof
    rng = np.random.default_rng(3)
    B = 6
    S = np.zeros((B, 2 * n), np.float32)
    for k in range(B):
        # fold: joint 2 and joint 3 turn by ~ 2 pi / 3 each so that link 3 comes back onto link 1
        S[k, 0:3] = rng.normal(0, 0.05, 3)
        S[k, 6] = 2.0944 + rng.normal(0, 0.003)
        S[k, 7] = 1.955 + rng.normal(0, 0.004)   # an edge of link 3 pressed 0.3 - 1.4 cm into link 1 (deeper contacts are clipped away)
        S[k, n:] = rng.normal(0, 0.1, n)
    A = rng.normal(0, 1.0, (B, len(raw.action_map))).astype(np.float32)
    cm = nb.compile_model(raw)
    r = EmulWorld(cm).forward_contact(S, A)
    ow = oracle_mod.OracleContactWorld(raw)
    with_self = 0
    for k in range(B):
        ro = ow.step_contact(S[k].astype(np.float64), A[k].astype(np.float64))
        assert r["nc"][k] == ro["nc"], (k, r["nc"][k], ro["nc"])
        assert np.array_equal(r["labels"][k][: r["m"][k]], ro["mapping"][: ro["m"]])
        assert rel_err(r["next"][k], ro["next_state"]) < 1e-6
        with_self += int(any(bb[0] >= 0 and bb[1] >= 0 and raw.skel_id[bb[0]] == raw.skel_id[bb[1]] for bb in ro["bodies"]))
    assert with_self >= 3, with_self
    kinds = _check_backward(ob, raw, S, A, tol=2e-5)
    assert len(kinds) == B


def test_joint_limit_rows_forward_and_backward(oracle_mod):
    """Joint::setPositionLimitEnforced (constraint/JointLimitConstraint.cpp): while a 1-dof joint sits on or beyond a position limit the LCP
    gains one row for it after the contact rows (ConstraintSolver.cpp:642-695): b = -qdot*, [0, inf) on a lower limit, (-inf, 0] on an upper
    one.  Forward against the oracle (row count, labels, next state) and backward against its Jacobian, with and without contacts."""
    raw = load_raw("half_cheetah")
    raw.limit_enforced[:] = 1
    raw.spring[:] = 0.0   # (the model's joint springs, 60-240 N m / rad on links of ~0.01 kg m^2, would throw every joint back off its limit within the step)
    n = raw.ndof
    s, a = contact_inputs(raw, "half_cheetah", 10, seed=9)
    rng = np.random.default_rng(2)
    for k in range(10):
        for d in rng.choice(np.arange(3, n), size=2, replace=False):   # two joints per world on / beyond a limit, moving into it or away
            hi = rng.random() < 0.5
            s[k, d] = (raw.pos_hi[d] + 0.01 * rng.random()) if hi else (raw.pos_lo[d] - 0.01 * rng.random())
            s[k, n + d] = (1.0 if hi else -1.0) * rng.choice([8.0, -0.5])
    a *= 0.0   # no torques: the light links would otherwise out-accelerate the velocities set above
    cm = nb.compile_model(raw)
    assert len(cm.limit_bodies) == n   # (the planar root is two prismatic joints and a revolute one with infinite limits: never active)
    r = EmulWorld(cm).forward_contact(s, a)
    ow = oracle_mod.OracleContactWorld(raw)
    limit_rows = clamping_limits = 0
    for k in range(10):
        ro = ow.step_contact(s[k].astype(np.float64), a[k].astype(np.float64))
        assert r["nc"][k] == ro["nc"] and r["m"][k] == ro["m"], (k, r["nc"][k], ro["nc"], r["m"][k], ro["m"])
        assert np.array_equal(r["labels"][k][: r["m"][k]], ro["mapping"][: ro["m"]]), (k, r["labels"][k][: r["m"][k]], ro["mapping"])
        assert rel_err(r["next"][k], ro["next_state"]) < 1e-6
        lim = [i for i, t in enumerate(ro["type"]) if t >= 100]
        limit_rows += len(lim)
        clamping_limits += int(sum(ro["mapping"][ro["m"] - len(lim) + i] == -2 for i in range(len(lim))))   # the limit rows come last
    assert limit_rows >= 15 and clamping_limits >= 4, (limit_rows, clamping_limits)
    _check_backward(ob, raw, s, a, tol=2e-5)
    # a model without a single shape: the cartpole with its rail limit enforced
    raw2 = load_raw("cartpole")
    raw2.limit_enforced[:] = 1
    raw2.pos_lo[0], raw2.pos_hi[0] = -1.0, 1.0
    s2 = np.zeros((4, 4), np.float32); a2 = np.zeros((4, len(raw2.action_map)), np.float32)
    s2[:, 0] = [1.0, 1.02, -1.0, 0.5]; s2[:, 1] = 0.3; s2[:, 2] = [1.5, 0.7, -2.0, 1.0]; s2[:, 3] = 0.2
    cm2 = nb.compile_model(raw2)
    r2 = EmulWorld(cm2).forward_contact(s2, a2)
    ow2 = oracle_mod.OracleContactWorld(raw2)
    for k in range(4):
        ro = ow2.step_contact(s2[k].astype(np.float64), a2[k].astype(np.float64))
        assert r2["m"][k] == ro["m"] and np.array_equal(r2["labels"][k][: r2["m"][k]], ro["mapping"][: ro["m"]])
        assert rel_err(r2["next"][k], ro["next_state"]) < 1e-6
    assert list(r2["m"]) == [1, 1, 1, 0]
    assert abs(r2["next"][0][2]) < 1e-9 and abs(r2["next"][2][2]) < 1e-9   # the rail stops at the limit
    _check_backward(ob, raw2, s2, a2, tol=2e-5)


def test_all_constraint_features_together(oracle_mod):
    """Restitution + penetration correction + enforced joint limits in one model, both lane orders: contact set, labels and next state against the
    oracle; gradients against its Jacobian (status bits 0x400 bounce, 0x1000 penetration correction, the solver-branch bits all occur)."""
    raw = load_raw("half_cheetah")
    raw.limit_enforced[:] = 1
    raw.spring[:] = 0
    raw.restitution[:] = 0.6
    raw.penetration_correction = True
    n = raw.ndof
    s, a = contact_inputs(raw, "half_cheetah", 12, seed=31)
    a *= 0.1
    rng = np.random.default_rng(5)
    s[::2, n + 1] -= 1.2
    for k in range(12):
        d = rng.integers(3, n); hi = rng.random() < 0.5
        s[k, d] = (raw.pos_hi[d] + 0.004) if hi else (raw.pos_lo[d] - 0.004)
        s[k, n + d] = 6.0 if hi else -6.0
    cm = nb.compile_model(raw)
    ow = oracle_mod.OracleContactWorld(raw)
    for rev in (False, True):
        r = EmulWorld(cm).forward_contact(s, a, reverse=rev)
        bits = 0
        for k in range(12):
            ro = ow.step_contact(s[k].astype(np.float64), a[k].astype(np.float64))
            assert r["m"][k] == ro["m"] and np.array_equal(r["labels"][k][: r["m"][k]], ro["mapping"][: ro["m"]]), (rev, k)
            assert rel_err(r["next"][k], ro["next_state"]) < 1e-6
            bits |= int(r["status"][k])
        assert bits & 0x400 and bits & 0x1000
    _check_backward(ob, raw, s, a, tol=3e-5)
