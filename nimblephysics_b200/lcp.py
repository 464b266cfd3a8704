"""Batched boxed-LCP solves on the GPU — the reference's pointer-style solver boundary as a tensor call.

reference: BoxedLcpSolver::solve (dart/constraint/BoxedLcpSolver.hpp:125-135), DantzigBoxedLcpSolver::solve
(DantzigBoxedLcpSolver.cpp:55-107 -> dSolveLCP, dart/external/odelcpsolver/lcp.cpp:780-1114) and the solve chain of
BoxedLcpConstraintSolver::solveLcp (BoxedLcpConstraintSolver.cpp:352-789).  One warp per problem; the same device code the contact stage
of timestep() runs (csrc/nb2_cw.cuh).  No CPU path.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _cabi


def solve_boxed_lcp_batch(A: torch.Tensor, b: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor, findex: torch.Tensor,
                          m: Optional[torch.Tensor] = None, x0: Optional[torch.Tensor] = None, chain: bool = True,
                          early_termination: bool = True, fallback_cfm: float = 1e-4):
    """A [B, mcap, mcap] (symmetric), b / lo / hi [B, mcap] float64, findex [B, mcap] int32 (-1: no friction parent), m [B] int32 problem
    sizes (default: mcap).  chain=False: Dantzig only -> (x, status) with status 1 solved / 0 early termination / -1 iteration cap;
    chain=True: the whole solve chain -> (x, labels, status) with NB2_ST_* status bits and ConstraintMapping labels."""
    if not A.is_cuda:
        raise RuntimeError("solve_boxed_lcp_batch needs CUDA tensors (there is no CPU fallback)")
    B, mcap = A.shape[0], A.shape[1]
    dev = A.device
    f64 = lambda t: t.to(device=dev, dtype=torch.float64).contiguous()
    A, b, lo, hi = f64(A), f64(b), f64(lo), f64(hi)
    fi = findex.to(device=dev, dtype=torch.int32).contiguous()
    mm = torch.full((B,), mcap, dtype=torch.int32, device=dev) if m is None else m.to(device=dev, dtype=torch.int32).contiguous()
    x0c = f64(x0) if x0 is not None else None
    x = torch.zeros((B, mcap), dtype=torch.float64, device=dev)
    labels = torch.zeros((B, mcap), dtype=torch.int32, device=dev)
    status = torch.zeros((B,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().nb2_lcp_solve_batch(B, mcap, 1 if chain else 0, int(early_termination), float(fallback_cfm), mm.data_ptr(),
                                                    A.data_ptr(), b.data_ptr(), lo.data_ptr(), hi.data_ptr(), fi.data_ptr(),
                                                    x0c.data_ptr() if x0c is not None else None, x.data_ptr(), labels.data_ptr(),
                                                    status.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return (x, labels, status) if chain else (x, status)
