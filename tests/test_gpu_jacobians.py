"""GPU: Jacobian getters (BackpropSnapshot::getStateJacobian / getActionJacobian, dart/neural/BackpropSnapshot.cpp:1230-1260) built by seeding
the backward kernels with identity rows.  Checked (a) against the VJP they come from, (b) against central finite differences of the forward
kernel for the blocks the reference defines by the true derivative (vel-vel, pos-vel, force-vel; the reference itself checks its blocks
against finite differences, BackpropSnapshot.cpp:4069-4260), (c) the block structure the reference assembles (zero position rows of the
action Jacobian, pos-pos = I and vel-pos = dt I for revolute chains)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from tests.util import contact_inputs, load_raw, multiarm_world  # noqa: E402


def _fd(nb, world, s, a, wrt, eps):
    x = s if wrt == "s" else a
    cols = []
    for j in range(x.shape[1]):
        d = torch.zeros_like(x); d[:, j] = eps
        with torch.no_grad():
            if wrt == "s":
                p = nb.timestep(world, s + d, a); m = nb.timestep(world, s - d, a)
            else:
                p = nb.timestep(world, s, a + d); m = nb.timestep(world, s, a - d)
        cols.append(((p.double() - m.double()) / (2 * eps)))
    return torch.stack(cols, -1)  # [B, 2n, dim]


def test_jacobians_arm_fd_and_structure():
    import nimblephysics_b200 as nb

    world = multiarm_world(5, 0.2, gravity=(0.0, -9.81, 0.0))
    n = world.getNumDofs()
    rng = np.random.default_rng(5)
    B = 3
    s = torch.tensor(rng.uniform(-0.5, 0.5, (B, 2 * n)).astype(np.float32), device="cuda")
    a = torch.tensor(rng.uniform(-2, 2, (B, world.getActionSize())).astype(np.float32), device="cuda")
    nxt, Js, Ja = nb.step_jacobians(world, s, a)
    with torch.no_grad():
        assert torch.equal(nxt, nb.timestep(world, s, a))
    dt = world.getTimeStep()
    eye = torch.eye(n, device="cuda")
    # structure the reference assembles (BackpropSnapshot.cpp:1236-1239, 1254-1258)
    assert torch.allclose(Js[:, :n, :n], eye.expand(B, n, n), atol=1e-6)               # pos-pos
    assert torch.allclose(Js[:, :n, n:], (dt * eye).expand(B, n, n), atol=1e-7)        # vel-pos = dt I
    assert torch.count_nonzero(Ja[:, :n, :]) == 0                                       # position rows of the action Jacobian
    # true-derivative blocks against finite differences of the forward kernel (fp32 forward: loose tolerance, scale-relative)
    fd_s = _fd(nb, world, s, a, "s", 2e-2)
    fd_a = _fd(nb, world, s, a, "a", 5e-1)
    for blk, ref in ((Js[:, n:, n:], fd_s[:, n:, n:]), (Js[:, n:, :n], fd_s[:, n:, :n]), (Ja[:, n:, :], fd_a[:, n:, :])):
        scale = ref.abs().max().item() + 1e-6
        assert (blk.double() - ref).abs().max().item() < 3e-2 * scale, ((blk.double() - ref).abs().max().item(), scale)
    # and they ARE the VJP: g^T J == backward(g)
    g = torch.randn(B, 2 * n, device="cuda")
    x = s.clone().requires_grad_(True); u = a.clone().requires_grad_(True)
    nb.timestep(world, x, u).backward(g)
    assert torch.allclose(torch.einsum("bi,bij->bj", g, Js), x.grad, rtol=1e-4, atol=1e-5 * x.grad.abs().max().item())
    assert torch.allclose(torch.einsum("bi,bij->bj", g, Ja), u.grad, rtol=1e-4, atol=1e-5 * u.grad.abs().max().item())


def test_jacobians_contact_world_keep_cache_and_match_vjp():
    import nimblephysics_b200 as nb

    raw = load_raw("half_cheetah")
    world = nb.World.from_raw(raw)
    B = 4
    s_np, a_np = contact_inputs(raw, "half_cheetah", B, seed=2)
    s = torch.tensor(s_np, device="cuda"); a = torch.tensor(a_np, device="cuda")
    with torch.no_grad():
        nb.timestep(world, s, a)  # warms the world's LCP cache
    cx, cm = world._lcp_cache["x"].clone(), world._lcp_cache["m"].clone()
    nxt, Js, Ja = nb.step_jacobians(world, s, a)
    assert torch.equal(world._lcp_cache["x"], cx) and torch.equal(world._lcp_cache["m"], cm), "the getter must not disturb the solver cache"
    assert torch.isfinite(Js).all() and torch.isfinite(Ja).all()
    g = torch.randn(B, 2 * raw.ndof, device="cuda")
    x = s.clone().requires_grad_(True); u = a.clone().requires_grad_(True)
    out = nb.timestep(world, x, u)
    assert torch.equal(out.detach(), nxt)
    out.backward(g)
    assert torch.allclose(torch.einsum("bi,bij->bj", g, Js), x.grad, rtol=2e-4, atol=2e-5 * x.grad.abs().max().item())
    assert torch.allclose(torch.einsum("bi,bij->bj", g, Ja), u.grad, rtol=2e-4, atol=2e-5 * u.grad.abs().max().item())


def test_legacy_world_getters():
    import nimblephysics_b200 as nb

    world = multiarm_world(5, 0.2, gravity=(0.0, -9.81, 0.0))
    n = world.getNumDofs()
    world.setState(np.linspace(-0.3, 0.3, 2 * n))
    world.setAction(np.ones(world.getActionSize()))
    Js, Ja = world.getStateJacobian(), world.getActionJacobian()
    assert Js.shape == (2 * n, 2 * n) and Ja.shape == (2 * n, world.getActionSize())
    assert np.allclose(Js[:n, :n], np.eye(n), atol=1e-6) and np.all(Ja[:n] == 0)
    assert np.allclose(world.getState(), np.linspace(-0.3, 0.3, 2 * n)), "the getters do not advance the world"
