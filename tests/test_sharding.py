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
