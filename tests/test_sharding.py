"""Multi-GPU path on CPU: batch sharding + the one collective (loss all-reduce) with world_size-2 gloo processes.
The timestep itself is replaced by a CPU stand-in with the same signature (the product has no CPU path); what is tested
is the host logic: disjoint complete shards, per-shard gradients equal to the unsharded ones, loss all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import nimblephysics_b200 as nb


def test_shard_range_partitions():
    for total in (0, 1, 7, 8, 4096, 65536 + 3):
        for ws in (1, 2, 3, 8):
            got = [nb.shard_range(total, r, ws) for r in range(ws)]
            assert got[0][0] == 0 and got[-1][1] == total
            assert all(got[i][1] == got[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in got]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        nb.shard_range(10, 2, 2)


def _fake_step(world, x, a):  # same signature as nb.timestep; linear toy dynamics, differentiable on CPU
    return x * 0.99 + torch.cat([a, a], dim=1) * 0.01


def _worker(rank, ws, port, B, T, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=ws)
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(B, 6, generator=g)
    acts = [torch.randn(B, 3, generator=g) for _ in range(T)]
    loss, gx, ga = nb.sharded_trajectory_loss(None, x0, acts, lambda x: (x * x).sum(), rank, ws, step_fn=_fake_step)
    out.put((rank, float(loss), gx.numpy(), [a.numpy() for a in ga]))
    dist.destroy_process_group()


def test_two_rank_gloo_rollout_matches_single_process():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    B, T, ws = 10, 4, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, B, T, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(ws)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(B, 6, generator=g).requires_grad_(True)
    acts = [torch.randn(B, 3, generator=g).requires_grad_(True) for _ in range(T)]
    x = x0
    for a in acts:
        x = _fake_step(None, x, a)
    loss = (x * x).sum()
    loss.backward()
    assert res[0][1] == pytest.approx(float(loss), rel=1e-6) and res[1][1] == pytest.approx(float(loss), rel=1e-6)
    gx = np.concatenate([r[2] for r in res])
    assert np.allclose(gx, x0.grad.numpy(), atol=1e-6)
    for t in range(T):
        ga = np.concatenate([r[3][t] for r in res])
        assert np.allclose(ga, acts[t].grad.numpy(), atol=1e-6)


def test_multishot_layout_and_defects_cpu():
    """MultiShot-as-batch bookkeeping (shots -> batch slices, getStates order, knot defects; MultiShot.cpp:164-213, 902-975) with a
    stand-in integrator: x' = x + u broadcast over the state, so every number is predictable."""
    import torch
    import nimblephysics_b200 as nb

    def fake_rollout(world, x0, acts):
        xs = [x0]
        for t in range(acts.shape[0]):
            xs.append(xs[-1] + acts[t].sum(-1, keepdim=True))
        return torch.stack(xs, 0)

    T, L, B, n2, na = 7, 3, 2, 4, 2
    S = 3
    g = torch.Generator().manual_seed(0)
    starts = torch.randn(S, B, n2, generator=g, requires_grad=True)
    acts = torch.randn(T, B, na, generator=g, requires_grad=True)
    states, defects = nb.multishot_rollout(None, starts, acts, L, rollout_fn=fake_rollout)
    assert states.shape == (T, B, n2) and defects.shape == (S - 1, B, n2)
    for t in range(T):
        s, k = divmod(t, L)
        ref = starts[s] + acts[s * L: s * L + k + 1].sum(0).sum(-1, keepdim=True)
        assert torch.allclose(states[t], ref, atol=1e-6)
    for i in range(S - 1):
        assert torch.allclose(defects[i], states[(i + 1) * L - 1] - starts[i + 1], atol=1e-6)
    (states[-1].sum() + defects.pow(2).sum()).backward()
    assert torch.isfinite(starts.grad).all() and torch.isfinite(acts.grad).all()
    # closing the knots (start of shot i+1 := end of shot i) reproduces the single-shot trajectory
    with torch.no_grad():
        x = starts[0]
        single = []
        for t in range(T):
            x = x + acts[t].sum(-1, keepdim=True)
            single.append(x)
        closed = torch.stack([starts[0].detach(), single[L - 1], single[2 * L - 1]], 0)
        st2, df2 = nb.multishot_rollout(None, closed, acts.detach(), L, rollout_fn=fake_rollout)
        assert torch.allclose(st2, torch.stack(single, 0), atol=1e-6) and df2.abs().max() < 1e-6
