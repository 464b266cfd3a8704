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


class _FusedRollout(torch.autograd.Function):
    """Whole-horizon rollout behind ONE C-ABI call per direction (include/nb2.h nb2_rollout_forward / _backward): the 2T
    kernels are queued back to back on the stream, the trajectory and the saved streams never leave the device, and
    autograd sees a single node instead of T."""

    @staticmethod
    def forward(ctx, world, state0, actions):
        from .engine import FP32, device_model_for

        dm = device_model_for(world)
        T, B = actions.shape[0], actions.shape[1]
        n2, na = 2 * dm.ndof, dm.na
        if state0.shape != (B, n2) or actions.shape[2] != na:
            raise ValueError(f"rollout(): state0 {tuple(state0.shape)} / actions {tuple(actions.shape)} do not match [B,{n2}] / [T,B,{na}]")
        dev = state0.device
        states = torch.empty((T + 1, B, n2), dtype=torch.float32, device=dev)
        states[0].copy_(state0.detach())
        acts = actions.detach().to(dtype=torch.float32).contiguous()
        need = any(ctx.needs_input_grad[1:3])
        saved = torch.empty((T, dm.saved_words, B), dtype=torch.float32, device=dev) if need else None
        with torch.cuda.device(dev):
            dm.rollout_forward_device(B, T, states.data_ptr(), acts.data_ptr(), saved.data_ptr() if need else None,
                                      torch.cuda.current_stream().cuda_stream, FP32)
        ctx.dm, ctx.T, ctx.B = dm, T, B
        ctx.dtypes = (state0.dtype, actions.dtype)
        if need:
            ctx.save_for_backward(states, acts, saved)
        return states.to(state0.dtype)

    @staticmethod
    def backward(ctx, grad_states):
        from .engine import FP32

        states, acts, saved = ctx.saved_tensors
        dm, T, B = ctx.dm, ctx.T, ctx.B
        gs = grad_states.detach().to(dtype=torch.float32).contiguous().clone()  # in: loss gradient per state; out: total dL/dx_t
        ga = torch.empty_like(acts)
        with torch.cuda.device(states.device):
            dm.rollout_backward_device(B, T, states.data_ptr(), acts.data_ptr(), saved.data_ptr(), gs.data_ptr(), ga.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream, FP32)
        return None, gs[0].to(ctx.dtypes[0]), ga.to(ctx.dtypes[1])


def rollout_fused(world, state0: torch.Tensor, actions: torch.Tensor) -> torch.Tensor:
    """states[T+1, B, 2n] of the T-step rollout x_{t+1} = timestep(x_t, actions[t]) of a contact-free world
    (SingleShot::getSnapshots, dart/trajectory/SingleShot.cpp:635-686), differentiable with respect to state0 and every
    action (SingleShot::backpropGradientWrt, :539-631); losses may look at any state of the trajectory.
    Worlds with collision pairs keep the per-step path (`rollout`), whose LCP cache flows from step to step."""
    from .engine import device_model_for

    if device_model_for(world).has_contacts:
        xs = [state0]
        for t in range(actions.shape[0]):
            xs.append(timestep(world, xs[-1], actions[t]))
        return torch.stack(xs, 0)
    return _FusedRollout.apply(world, state0, actions)


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
    if step is timestep:
        from .timestep import check_contact_status

        check_contact_status(world)  # ONE host sync per rollout: worlds that dropped contacts / could not be back-propagated raise here
    return total, x0.grad, [a.grad for a in acts]
