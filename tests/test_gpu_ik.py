"""GPU: IKMapping on the device (nb2_ik_forward / nb2_ik_backward through nimblephysics_b200.mapping) against the oracle's restatement of
neural/IKMapping.cpp: mapped positions / velocities of spatial, linear, angular and COM entries, and the VJP against J^T g with the oracle's
dual-number Jacobians.  fp32 in / out, fp64 inside: tolerance 2e-6 absolute on O(1) quantities."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from tests.util import load_raw  # noqa: E402


@pytest.fixture(scope="module")
def oracle_mod():
    from oracle import binding
    binding.build()
    return binding


@pytest.mark.parametrize("name", ["atlas", "half_cheetah"])
def test_ik_matches_oracle(oracle_mod, name):
    import nimblephysics_b200 as nb

    raw = load_raw(name)
    world = nb.World.from_raw(raw)
    ow = oracle_mod.OracleWorld(raw)
    names = list(raw.body_names)
    mobile = [i for i in range(raw.nb) if raw.mobile[i]]
    picks = [mobile[-1], mobile[len(mobile) // 2], mobile[0], mobile[len(mobile) // 3]]
    types = [0, 1, 2, 1, 3]
    bodies = picks + [mobile[0]]
    # node objects of the rebuilt world, in raw order
    nodes = [b for sk in world.skeletons for b in sk._ordered_bodies()]
    ik = nb.IKMapping(world)
    ik.addSpatialBodyNode(nodes[picks[0]]); ik.addLinearBodyNode(nodes[picks[1]]); ik.addAngularBodyNode(nodes[picks[2]])
    ik.addLinearBodyNode(nodes[picks[3]])
    ik.addSkeletonCOM(world.skeletons[int(raw.skel_id[mobile[0]])])
    n, B = raw.ndof, 6
    rng = np.random.default_rng(1)
    s = rng.uniform(-0.5, 0.5, (B, 2 * n)).astype(np.float32)
    st = torch.tensor(s, device="cuda", requires_grad=True)
    pos = nb.map_to_pos(world, ik, st)
    gp = torch.tensor(rng.standard_normal((B, ik.getPosDim())).astype(np.float32), device="cuda")
    pos.backward(gp)
    g_pos = st.grad.clone(); st.grad = None
    vel = nb.map_to_vel(world, ik, st)
    gv = torch.tensor(rng.standard_normal((B, ik.getVelDim())).astype(np.float32), device="cuda")
    vel.backward(gv)
    g_vel = st.grad.clone()
    for w in range(B):
        po, vo, Jp, Jv = ow.ik(s[w].astype(np.float64), types, bodies)
        assert np.allclose(pos[w].detach().cpu().numpy(), po, atol=2e-6), (name, w, np.abs(pos[w].detach().cpu().numpy() - po).max())
        assert np.allclose(vel[w].detach().cpu().numpy(), vo, atol=5e-6), (name, w)
        ref_p = Jp.T @ gp[w].cpu().numpy().astype(np.float64)
        ref_v = Jv.T @ gv[w].cpu().numpy().astype(np.float64)
        gq = g_pos[w].cpu().numpy()
        assert np.allclose(gq[:n], ref_p, atol=5e-6 * (1 + np.abs(ref_p).max())), (name, w, np.abs(gq[:n] - ref_p).max())
        assert np.all(gq[n:] == 0), "map_to_pos feeds only d/dq (mapping.py:41-47)"
        gd = g_vel[w].cpu().numpy()
        assert np.allclose(gd[n:], ref_v, atol=5e-6 * (1 + np.abs(ref_v).max())), (name, w)
        assert np.all(gd[:n] == 0), "map_to_vel feeds only d/dqdot (mapping.py:89-95)"


def test_ik_legacy_single_world_and_jacobian_getters(oracle_mod):
    import nimblephysics_b200 as nb
    from tests.util import multiarm_world
    from nimblephysics_b200.modelspec import flatten_world

    world = multiarm_world(5, 0.2)
    n = world.getNumDofs()
    world.setState(np.linspace(-0.4, 0.5, 2 * n))
    hand = world.getSkeleton(0).getBodyNode(n - 1)
    ik = nb.IKMapping(world)
    ik.addSpatialBodyNode(hand)
    ow = oracle_mod.OracleWorld(flatten_world(world))
    po, vo, Jp, Jv = ow.ik(world.getState(), [0], [n - 1])
    assert np.allclose(ik.getPositions(world), po, atol=2e-6)
    assert np.allclose(ik.getVelocities(world), vo, atol=2e-6)
    assert np.allclose(ik.getRealPosToMappedPosJac(world), Jp, atol=5e-6)
    assert np.allclose(ik.getRealVelToMappedVelJac(world), Jv, atol=5e-6)
    out = nb.map_to_pos(world, ik, torch.tensor(world.getState(), dtype=torch.float64, requires_grad=True))
    assert out.dtype == torch.float64 and out.shape == (6,)
