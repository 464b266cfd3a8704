"""CPU: the IKMapping restatement of the oracle (oracle/nb_oracle.cpp ik_map; neural/IKMapping.cpp:146-237, 371-476).
Pinned like the reference pins it (unittests/comprehensive/test_IKMapping-style checks): closed-form values on a pendulum, and the
dual-number Jacobians against central finite differences of the mapped vectors."""
import numpy as np
import pytest

from tests.util import load_raw


@pytest.fixture(scope="module")
def oracle_mod():
    from oracle import binding
    binding.build()
    return binding


def test_ik_atlas_jacobians_match_finite_differences(oracle_mod):
    raw = load_raw("atlas")
    ow = oracle_mod.OracleWorld(raw)
    rng = np.random.default_rng(0)
    n = raw.ndof
    s = rng.uniform(-0.4, 0.4, 2 * n)
    names = list(raw.body_names)
    bodies = [names.index("l_hand"), names.index("r_foot"), names.index("utorso"), names.index("pelvis"), 0]
    types = [0, 1, 2, 0, 3]
    pos, vel, Jp, Jv = ow.ik(s, types, bodies)
    assert pos.shape == (6 + 3 + 3 + 6 + 3,)
    eps = 1e-6
    for j in range(n):
        d = np.zeros(2 * n); d[j] = eps
        pp, _, _, _ = ow.ik(s + d, types, bodies, want_jac=False)
        pm, _, _, _ = ow.ik(s - d, types, bodies, want_jac=False)
        assert np.allclose((pp - pm) / (2 * eps), Jp[:, j], atol=2e-6), j
        d = np.zeros(2 * n); d[n + j] = eps
        _, vp, _, _ = ow.ik(s + d, types, bodies, want_jac=False)
        _, vm, _, _ = ow.ik(s - d, types, bodies, want_jac=False)
        assert np.allclose((vp - vm) / (2 * eps), Jv[:, j], atol=2e-6), j
    # mapped velocity = geometric Jacobian * qdot (Skeleton::getWorldJacobian, IKMapping.cpp:444-462)
    assert np.allclose(Jv @ s[n:], vel, atol=1e-9)


def test_ik_pendulum_closed_form(oracle_mod):
    import nimblephysics_b200 as nb
    from nimblephysics_b200.modelspec import flatten_world

    w = nb.World()
    sk = nb.Skeleton("p")
    j, b = sk.createRevoluteJointAndBodyNodePair(None)
    j.setAxis([0, 0, 1])
    T = nb.Isometry3(); T.set_translation([0.5, 0, 0])
    j.setTransformFromChildBodyNode(nb.Isometry3())
    j2, b2 = sk.createRevoluteJointAndBodyNodePair(b)
    j2.setAxis([0, 0, 1]); j2.setTransformFromParentBodyNode(T)
    b.setMass(1.0); b2.setMass(1.0)
    w.addSkeleton(sk)
    raw = flatten_world(w)
    ow = oracle_mod.OracleWorld(raw)
    th, om = 0.3, 2.0
    pos, vel, Jp, Jv = ow.ik(np.array([th, 0.0, om, 0.0]), [0], [1])
    assert np.allclose(pos, [0, 0, th, 0.5 * np.cos(th), 0.5 * np.sin(th), 0], atol=1e-12)
    assert np.allclose(vel, [0, 0, om, -0.5 * np.sin(th) * om, 0.5 * np.cos(th) * om, 0], atol=1e-12)
