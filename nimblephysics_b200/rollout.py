"""T-step rollouts with backprop through the horizon, and batch sharding across the GPUs of one box.

reference: dart/trajectory/SingleShot.cpp:635-686 (getSnapshots: unroll T forward passes) and :539-631
(backpropGradientWrt: reverse sweep carrying dL/dq, dL/dqdot and adding per-step loss gradients);
MultiShot runs shots on std::async threads with one cloned World each (MultiShot.cpp:57-72, 1245-1275).
Here a "shot" is simply a slice of the batch: worlds are independent, so the batch shards across ranks with NO
data-path collective; the only collective is the all-reduce of the scalar loss / of gradients of parameters shared by
all worlds (torch.distributed, NCCL over NVLink), outside the timestep.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch

from .timestep import timestep


def rollout(world, state0: torch.Tensor, actions: Sequence[torch.Tensor], keep_states: bool = False):
    """Unroll len(actions) differentiable steps.  state0: [B, 2n]; actions[t]: [B, a].
    Returns the final state (and the list of intermediate states when keep_states)."""
    x = state0
    states: List[torch.Tensor] = []
    for a in actions:
        x = timestep(world, x, a)
        if keep_states:
            states.append(x)
    return (x, states) if keep_states else x


def shard_range(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of a batch of `total` worlds owned by `rank` (sizes differ by at most one)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, rem = divmod(total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    lo, hi = shard_range(t.shape[0], rank, world_size)
    return t[lo:hi]


def allreduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce when torch.distributed is initialised (no-op otherwise)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def sharded_trajectory_loss(world, state0: torch.Tensor, actions: Sequence[torch.Tensor],
                            loss_fn: Callable[[torch.Tensor], torch.Tensor], rank: int, world_size: int,
                            step_fn: Optional[Callable] = None):
    """Each rank rolls out its slice of the batch, backpropagates its part of the loss, and the scalar loss is
    all-reduced (config 5 of BASELINE.json).  Returns (global_loss, local_state0_grad, local_action_grads)."""
    step = step_fn or timestep
    lo, hi = shard_range(state0.shape[0], rank, world_size)
    x0 = state0[lo:hi].detach().clone().requires_grad_(True)
    acts = [a[lo:hi].detach().clone().requires_grad_(True) for a in actions]
    x = x0
    for a in acts:
        x = step(world, x, a)
    loss = loss_fn(x)
    loss.backward()
    total = allreduce_sum_(loss.detach().clone())
    return total, x0.grad, [a.grad for a in acts]
