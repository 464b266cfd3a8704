"""Pins the fp64 oracle (oracle/nb_oracle.cpp).  The reference ships no golden vectors for this path
(SURVEY §4, §8c), so the oracle is pinned the way the reference pins itself plus independent physics:
  1. analytic (dual-number) Jacobians vs central finite differences of the step, tol 1e-7
     (reference: unittests/GradientTestUtils.hpp:637-680 verifyVelVelJacobian & co, tol 1e-8 on their side);
  2. Lagrangian mechanics recomputed from scratch in numpy (mass matrix from numeric body Jacobians, Christoffel
     Coriolis terms, potential-energy gradient) vs the recursive ABA of the oracle — joints without free roots;
  3. free rigid body vs Newton-Euler closed form;  4. Atlas: centre-of-mass acceleration == gravity.
"""
import numpy as np
import pytest

from tests.util import load_raw, rel_err, sample_inputs

import nimblephysics_b200 as nb
from nimblephysics_b200.modelspec import T_from_12
from nimblephysics_b200.world import FREE, PRISMATIC, REVOLUTE, WELD


def _rodrigues(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th**2 * K @ K


def fk(raw, q):
    """world transforms of every raw body, straight from the joint definitions (independent of the oracle code)."""
    W = []
    for i in range(raw.nb):
        Q = np.eye(4)
        o = raw.dof_off[i]
        if raw.jtype[i] == REVOLUTE:
            Q[:3, :3] = _rodrigues(raw.axis[i] * q[o])
        elif raw.jtype[i] == PRISMATIC:
            Q[:3, 3] = raw.axis[i] * q[o]
        elif raw.jtype[i] == FREE:
            Q[:3, :3] = _rodrigues(q[o:o + 3])
            Q[:3, 3] = q[o + 3:o + 6]
        T = T_from_12(raw.Tpj[i]) @ Q @ np.linalg.inv(T_from_12(raw.Tcj[i]))
        W.append(T if raw.parent[i] < 0 else W[raw.parent[i]] @ T)
    return W


def spatial_G(raw, i):
    m, c, mo = raw.mass[i], raw.com[i], raw.moment[i]
    Ic = np.array([[mo[0], mo[3], mo[4]], [mo[3], mo[1], mo[5]], [mo[4], mo[5], mo[2]]])
    C = np.array([[0, -c[2], c[1]], [c[2], 0, -c[0]], [-c[1], c[0], 0]])
    G = np.zeros((6, 6))
    G[:3, :3] = Ic + m * C @ C.T
    G[:3, 3:] = m * C
    G[3:, :3] = m * C.T
    G[3:, 3:] = m * np.eye(3)
    return G


def mass_matrix(raw, q, eps=1e-6):
    n = raw.ndof
    W0 = fk(raw, q)
    J = [np.zeros((6, n)) for _ in range(raw.nb)]
    for k in range(n):
        qp, qm = q.copy(), q.copy()
        qp[k] += eps
        qm[k] -= eps
        Wp, Wm = fk(raw, qp), fk(raw, qm)
        for i in range(raw.nb):
            dW = np.linalg.inv(W0[i]) @ (Wp[i] - Wm[i]) / (2 * eps)  # body twist hat
            J[i][:, k] = [dW[2, 1], dW[0, 2], dW[1, 0], dW[0, 3], dW[1, 3], dW[2, 3]]
    M = sum(J[i].T @ spatial_G(raw, i) @ J[i] for i in range(raw.nb) if raw.mobile[i])
    return M


def potential(raw, q):
    W = fk(raw, q)
    g = raw.gravity
    return -sum(raw.mass[i] * g @ (W[i][:3, :3] @ raw.com[i] + W[i][:3, 3]) for i in range(raw.nb) if raw.mobile[i])


def lagrangian_qdd(raw, q, v, tau, h=1e-4):
    n = raw.ndof
    M = mass_matrix(raw, q)
    dM = np.zeros((n, n, n))  # dM[k] = dM/dq_k
    gq = np.zeros(n)
    for k in range(n):
        qp, qm = q.copy(), q.copy()
        qp[k] += h
        qm[k] -= h
        dM[k] = (mass_matrix(raw, qp) - mass_matrix(raw, qm)) / (2 * h)
        gq[k] = (potential(raw, qp) - potential(raw, qm)) / (2 * h)
    C = np.zeros(n)
    for k in range(n):
        C[k] = sum(dM[i][k, j] * v[i] * v[j] for i in range(n) for j in range(n)) - 0.5 * v @ dM[k] @ v
    f = tau - C - gq - raw.damping * v - raw.spring * (q - raw.rest + v * raw.dt)
    return np.linalg.solve(M, f)


def _tree_world():
    from nimblephysics_b200.loader import _T, _rpy_to_R

    rng = np.random.default_rng(5)

    def setup(b):
        b.setMass(rng.uniform(0.5, 2))
        b.setLocalCOM(rng.uniform(-0.2, 0.2, 3))
        A = rng.normal(size=(3, 3))
        b.moment = A @ A.T * 0.05 + np.eye(3) * 0.05

    sk = nb.Skeleton()
    j, b = sk.createRevoluteJointAndBodyNodePair()
    j.setAxis([0.3, -0.5, 0.8])
    j.setTransformFromParentBodyNode(_T(_rpy_to_R([0.1, 0.2, 0.3]), [0.1, 0.2, 0.3]))
    j.setTransformFromChildBodyNode(_T(_rpy_to_R([-0.3, 0.1, 0.2]), [0.05, -0.1, 0.02]))
    j.setDampingCoefficient(0, 0.3)
    setup(b)
    j2, b2 = sk.createRevoluteJointAndBodyNodePair(b)
    j2.setAxis([0, 0, -1])
    j2.setTransformFromParentBodyNode(_T(None, [0.3, 0, 0.1]))
    j2.setSpringStiffness(0, 7.0)
    j2.setRestPosition(0, 0.1)
    setup(b2)
    j3, b3 = sk.createPrismaticJointAndBodyNodePair(b)
    j3.setAxis([0, 1, 0])
    j3.setTransformFromParentBodyNode(_T(_rpy_to_R([0.5, 0, 0]), [-0.3, 0, 0.1]))
    setup(b3)
    j4, b4 = sk.createRevoluteJointAndBodyNodePair(b3)
    j4.setAxis([1, 0, 0])
    setup(b4)
    j5, b5 = sk.createWeldJointAndBodyNodePair(b2)
    j5.setTransformFromParentBodyNode(_T(_rpy_to_R([0.5, 0.2, 0]), [-0.3, 0.2, 0.1]))
    setup(b5)
    w = nb.World()
    w.setGravity([0.3, -9.81, 0.5])
    w.addSkeleton(sk)
    return w


@pytest.mark.parametrize("name", ["cartpole", "half_cheetah", "atlas"])
def test_jacobians_match_finite_differences(oracle_mod, name):
    raw = load_raw(name)
    ow = oracle_mod.OracleWorld(raw)
    n = raw.ndof
    s, a, _ = sample_inputs(raw, 1, seed=7)
    s, a = s[0].astype(np.float64), a[0].astype(np.float64)
    J = ow.jacobian(s, a)
    eps = 1e-6
    Jfd = np.zeros_like(J)
    for c in range(2 * n):
        sp, sm = s.copy(), s.copy()
        sp[c] += eps
        sm[c] -= eps
        Jfd[:, c] = (ow.step(sp, a) - ow.step(sm, a)) / (2 * eps)
    for i, c in enumerate(raw.action_map):
        ap, am = a.copy(), a.copy()
        ap[i] += eps
        am[i] -= eps
        Jfd[:, 2 * n + c] = (ow.step(s, ap) - ow.step(s, am)) / (2 * eps)
    assert np.abs(J - Jfd).max() < 1e-7 * max(1.0, np.abs(J).max())
    # posPos = I / velPos = dt I on 1-dof joints (BackpropSnapshot.cpp:1263-1400)
    for i in range(raw.nb):
        if raw.jtype[i] in (REVOLUTE, PRISMATIC):
            o = raw.dof_off[i]
            assert J[o, o] == pytest.approx(1.0, abs=1e-12) and J[o, n + o] == pytest.approx(raw.dt, abs=1e-12)


def test_backprop_is_jacobian_transpose_and_clips(oracle_mod):
    raw = load_raw("cartpole")
    ow = oracle_mod.OracleWorld(raw)
    n = raw.ndof
    s = np.array([0.1, 0.2, -0.3, 0.4])
    a = np.array([1.0, 0.0])
    g = np.array([0.3, -0.2, 0.5, 0.7])
    J = ow.jacobian(s, a)
    gs, ga = ow.backprop(s, a, g)
    full = J.T @ g
    assert np.allclose(gs, full[: 2 * n], atol=1e-14) and np.allclose(ga, full[2 * n:][raw.action_map], atol=1e-14)
    # sitting exactly on a bound zeroes the outward-pushing component (BackpropSnapshot.cpp:425-479)
    s2 = s.copy()
    s2[0] = raw.pos_hi[0]
    gs2, _ = ow.backprop(s2, a, g)
    raw_g = (ow.jacobian(s2, a).T @ g)[0]
    assert (gs2[0] == 0.0) if raw_g < 0 else (gs2[0] == pytest.approx(raw_g))


@pytest.mark.parametrize("which", ["tree", "cartpole", "half_cheetah"])
def test_aba_matches_lagrangian_mechanics(oracle_mod, which):
    raw = nb.flatten_world(_tree_world()) if which == "tree" else load_raw(which)
    ow = oracle_mod.OracleWorld(raw)
    rng = np.random.default_rng(3)
    n = raw.ndof
    q, v = rng.uniform(-0.5, 0.5, n), rng.uniform(-1, 1, n)
    tau = rng.uniform(-3, 3, n)
    act = tau[raw.action_map]
    tau_full = np.zeros(n)
    tau_full[raw.action_map] = act
    _, qdd = ow.step(np.concatenate([q, v]), act, want_qdd=True)
    ref = lagrangian_qdd(raw, q, v, tau_full)
    assert rel_err(qdd, ref) < 2e-5, (qdd, ref)


def test_free_rigid_body_newton_euler(oracle_mod):
    sk = nb.Skeleton()
    j, b = sk.createFreeJointAndBodyNodePair()
    b.setMass(2.0)
    b.setLocalCOM([0.0, 0.0, 0.0])
    b.setMomentOfInertia(0.3, 0.5, 0.7, 0.01, -0.02, 0.03)
    w = nb.World()
    w.setGravity([0, -9.81, 0])
    w.addSkeleton(sk)
    raw = nb.flatten_world(w)
    ow = oracle_mod.OracleWorld(raw)
    q = np.array([0.2, -0.4, 0.3, 1.0, 2.0, 3.0])
    v = np.array([0.5, -0.3, 0.8, 0.1, 0.2, -0.4])
    tau = np.array([0.1, -0.2, 0.3, 1.0, -2.0, 0.5])
    _, qdd = ow.step(np.concatenate([q, v]), tau, want_qdd=True)
    I = b.moment
    R = _rodrigues(q[:3])
    wdot = np.linalg.solve(I, tau[:3] - np.cross(v[:3], I @ v[:3]))
    # body-frame linear acceleration of the origin: (f + m R^T g)/m - w x v
    vdot = tau[3:] / 2.0 + R.T @ np.array([0, -9.81, 0]) - np.cross(v[:3], v[3:])
    assert np.allclose(qdd[:3], wdot, atol=1e-10) and np.allclose(qdd[3:], vdot, atol=1e-10)


@pytest.mark.parametrize("name", ["atlas", "atlas_sdf"])
def test_atlas_centre_of_mass_accelerates_with_gravity(oracle_mod, name):
    """URDF- and SDF-loaded Atlas: whatever the internal torques, the centre of mass of a floating robot accelerates with g —
    a check of the loaders' frames (joint poses, axes, inertial offsets) together with the dynamics."""
    from scipy.spatial.transform import Rotation

    raw = load_raw(name)
    ow = oracle_mod.OracleWorld(raw)
    s, a, _ = sample_inputs(raw, 1, seed=11)
    s, a = s[0].astype(np.float64), a[0].astype(np.float64)
    a[:6] = 0.0  # no external wrench on the floating base: only internal torques + gravity
    n = raw.ndof
    q, v = s[:n], s[n:]
    _, qdd = ow.step(s, a, want_qdd=True)
    mtot = raw.mass.sum()

    def com(qq):
        W = fk(raw, qq)
        return sum(raw.mass[i] * (W[i][:3, :3] @ raw.com[i] + W[i][:3, 3]) for i in range(raw.nb)) / mtot

    def advance(qq, vv, e):  # FreeJoint.cpp:922-929 for the root, q + e v elsewhere
        out = qq + e * vv
        R = Rotation.from_rotvec(qq[:3]).as_matrix()
        out[:3] = Rotation.from_matrix(R @ Rotation.from_rotvec(vv[:3] * e).as_matrix()).as_rotvec()
        out[3:6] = qq[3:6] + R @ (vv[3:6] * e)
        return out

    def com_vel(qq, vv, e=1e-6):
        return (com(advance(qq, vv, e)) - com(advance(qq, vv, -e))) / (2 * e)

    h = 1e-5
    # d/dt v_com = d/dq[v_com] . qdot + J_com(q) qdd
    term_q = (com_vel(advance(q, v, h), v) - com_vel(advance(q, v, -h), v)) / (2 * h)
    term_a = com_vel(q, qdd)  # J_com is linear in its second argument
    acc = term_q + term_a
    assert np.allclose(acc, raw.gravity, atol=2e-3), acc
