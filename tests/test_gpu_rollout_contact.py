"""GPU: the C rollout driver of contact worlds (nb2_rollout_forward_contact / _backward_contact, include/nb2.h) against chaining
timestep() — the per-step path the other GPU tests pin to the oracle.  States and gradients must be the SAME BITS, with the full tape
and with checkpoints (segment length dividing the horizon or not), and the world's LCP cache must end in the same state.
reference: dart/trajectory/SingleShot.cpp:539-686 (getSnapshots + backpropGradientWrt)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from tests.util import contact_inputs, load_raw  # noqa: E402


def _chain(nb, world, x0, acts, W):
    nb.reset_contact_cache(world)
    x = x0.clone().requires_grad_(True)
    us = [a.clone().requires_grad_(True) for a in acts]
    xs = [x]
    for u in us:
        xs.append(nb.timestep(world, xs[-1], u))
    loss = sum((W[t] * xs[t]).sum() for t in range(len(xs)))
    loss.backward()
    c = world._lcp_cache
    return torch.stack([v.detach() for v in xs]), x.grad.clone(), torch.stack([u.grad for u in us]), c["x"].clone(), c["m"].clone()


@pytest.mark.parametrize("name,B,T", [("half_cheetah", 96, 7), ("atlas_ground", 48, 5)])
@pytest.mark.parametrize("k", [0, 1, 3, 64])
def test_contact_rollout_matches_chained_timestep(name, B, T, k):
    import nimblephysics_b200 as nb

    raw = load_raw(name)
    world = nb.World.from_raw(raw)
    s, a = contact_inputs(raw, name, B, seed=7)
    x0 = torch.tensor(s, device="cuda")
    g = torch.Generator(device="cpu").manual_seed(3)
    acts = torch.tensor(a, device="cuda")[None].repeat(T, 1, 1) * (1.0 + 0.1 * torch.randn(T, 1, 1, generator=g).cuda())
    W = torch.randn(T + 1, B, 2 * raw.ndof, generator=g).cuda()
    xs_ref, gx_ref, ga_ref, cx_ref, cm_ref = _chain(nb, world, x0, list(acts), W)

    nb.reset_contact_cache(world)
    x = x0.clone().requires_grad_(True)
    u = acts.clone().requires_grad_(True)
    xs = nb.rollout_fused(world, x, u, checkpoint_every=k)
    (W * xs).sum().backward()
    assert torch.equal(xs.detach(), xs_ref)
    c = world._lcp_cache
    assert torch.equal(c["m"], cm_ref) and torch.equal(c["x"], cx_ref), "the solver cache must end where chaining timestep() leaves it"
    assert torch.equal(x.grad, gx_ref)
    assert torch.equal(u.grad, ga_ref)
    assert torch.isfinite(x.grad).all() and torch.isfinite(u.grad).all()
    if 0 < k < T:
        assert nb.rollout_tape_bytes(world, B, T, k) < nb.rollout_tape_bytes(world, B, T, 0)
    nb.check_contact_status(world)


def test_multishot_closed_knots_reproduce_the_single_shot():
    """MultiShot as a batch (nimblephysics_b200.multishot_rollout; MultiShot.cpp:164-213, 902-975): with the knot points placed ON the
    single-shot trajectory the shots reproduce it and every defect is zero; gradients reach the knots and the actions."""
    import nimblephysics_b200 as nb
    from tests.util import sample_inputs

    raw = load_raw("atlas")
    world = nb.World.from_raw(raw)
    B, T, L = 16, 10, 4
    s, a, _ = sample_inputs(raw, B, seed=5)
    x0 = torch.tensor(s, device="cuda")
    acts = torch.tensor(a, device="cuda")[None].repeat(T, 1, 1)
    with torch.no_grad():
        single = nb.rollout_fused(world, x0, acts)          # [T+1, B, 2n]
    S = (T + L - 1) // L
    starts = torch.stack([single[i * L] for i in range(S)], 0).clone().requires_grad_(True)
    u = acts.clone().requires_grad_(True)
    states, defects = nb.multishot_rollout(world, starts, u, L)
    assert states.shape == (T, B, 2 * raw.ndof) and defects.shape == (S - 1, B, 2 * raw.ndof)
    assert torch.allclose(states, single[1:], rtol=1e-5, atol=1e-5)
    assert defects.abs().max().item() < 1e-5
    (states[-1].pow(2).sum() + defects.pow(2).sum()).backward()
    assert torch.isfinite(starts.grad).all() and torch.isfinite(u.grad).all()
    assert starts.grad[-1].abs().max() > 0 and u.grad[-1].abs().max() > 0
