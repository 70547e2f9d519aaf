"""Device-resident PyTorch front-end: the solve and its sensitivities on CUDA tensors, no host round trip.

* `solve_device(mcp, Θ, …)` — batched `solve(InteriorPoint(), mcp, θ)` (`/root/reference/src/solver.jl:35-122`) on
  tensors already in HBM, through `mcpb200_solve_batched_device` on torch's current stream.
* `MCPSolve.apply(mcp, Θ, x0, y0, opts)` — a `torch.autograd.Function`: the PyTorch counterpart of the reference's
  `ChainRulesCore.rrule(solve, …)` (`src/AutoDiff.jl:42-82`).  Forward = the solve; backward = the pullback
  `∂θ = Σ_b (∂z/∂θ)[b,:]ᵀ ∂l/∂b` (`:65-75`) computed by `mcpb200_sensitivities_device`, one adjoint per instance.
  This is the fused "solve + loss + VJP" path of the training loop (`examples/train_and_test_utils.jl:250-295`)
  with everything resident on the GPU.

Layout: tensors are `[B, n]` row-major, i.e. exactly the column-major `n×B` matrices of the C ABI.
torch is used for device memory and streams only; all numerics are libmcpb200's kernels.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import capi
from .mcp import PrimalDualMCP
from .solver import _handle


def _check(t: Optional[torch.Tensor], B: int, n: int, name: str):
    if t is None:
        return None
    if not (t.is_cuda and t.dtype == torch.float64 and t.shape == (B, n)):
        raise ValueError(f"{name} must be a CUDA float64 tensor of shape ({B}, {n}), got {tuple(t.shape)} {t.dtype} {t.device}")
    return t.contiguous()


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def solve_device(mcp: PrimalDualMCP, Θ: torch.Tensor, x0=None, y0=None, s0=None, **opts):
    """Returns dict(x, y, s, kkt_error, eps, outer_iters, status, newton_steps) of CUDA tensors (eps = ϵ after its last update).  `x0`/`y0`/`s0` are
    not modified (the kernel works on copies) unless they are passed as the outputs themselves."""
    nx, ny, nθ = mcp.unconstrained_dimension, mcp.constrained_dimension, mcp.parameter_dimension
    B = Θ.shape[0]
    Θ = _check(Θ, B, nθ, "Θ")
    x0, y0, s0 = _check(x0, B, nx, "x0"), _check(y0, B, ny, "y0"), _check(s0, B, ny, "s0")
    dev = Θ.device
    h = _handle(mcp)
    out = dict(x=torch.empty((B, nx), dtype=torch.float64, device=dev), y=torch.empty((B, ny), dtype=torch.float64, device=dev),
               s=torch.empty((B, ny), dtype=torch.float64, device=dev), kkt_error=torch.empty(B, dtype=torch.float64, device=dev),
               eps=torch.empty(B, dtype=torch.float64, device=dev), outer_iters=torch.empty(B, dtype=torch.int32, device=dev),
               status=torch.empty(B, dtype=torch.int32, device=dev), newton_steps=torch.empty(B, dtype=torch.int32, device=dev))
    o = capi.default_opts(**opts)
    with torch.cuda.device(dev):
        rc = h._lib.mcpb200_solve_batched_device(
            h.raw, B, _ptr(Θ), _ptr(x0), _ptr(y0), _ptr(s0), C.byref(o), _ptr(out["x"]), _ptr(out["y"]), _ptr(out["s"]),
            _ptr(out["kkt_error"]), _ptr(out["eps"]), _ptr(out["outer_iters"]), _ptr(out["status"]), _ptr(out["newton_steps"]),
            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    h.check(rc)
    return out


def pullback_device(mcp: PrimalDualMCP, Θ, x, y, s, ϵ, zbar: torch.Tensor) -> torch.Tensor:
    """θ̄[B, nθ] = (∂z/∂θ)ᵀ z̄ per instance, z̄ = [x̄; ȳ; s̄] as a `[B, nx+2ny]` tensor (`src/AutoDiff.jl:59-76`)."""
    nx, ny, nθ = mcp.unconstrained_dimension, mcp.constrained_dimension, mcp.parameter_dimension
    B = Θ.shape[0]
    zbar = _check(zbar, B, nx + 2 * ny, "z̄")
    θbar = torch.empty((B, nθ), dtype=torch.float64, device=Θ.device)
    h = _handle(mcp)
    with torch.cuda.device(Θ.device):
        rc = h._lib.mcpb200_sensitivities_device(
            h.raw, B, _ptr(Θ.contiguous()), _ptr(x.contiguous()), _ptr(y.contiguous()), _ptr(s.contiguous()), _ptr(ϵ.contiguous()),
            None, _ptr(zbar), _ptr(θbar), 0, None, None, None, C.c_void_p(torch.cuda.current_stream(Θ.device).cuda_stream))
    if rc == capi.ERR_NO_SENSITIVITIES:
        raise ValueError("Missing sensitivities. Set `compute_sensitivities = true` when constructing the PrimalDualMCP.")
    h.check(rc)
    return θbar


class MCPSolve(torch.autograd.Function):
    """`x, y, s = MCPSolve.apply(mcp, Θ, x0, y0, opts_dict)`; differentiable w.r.t. Θ."""

    @staticmethod
    def forward(ctx, mcp, Θ, x0=None, y0=None, opts=None):
        sol = solve_device(mcp, Θ.detach(), x0, y0, **(opts or {}))
        ctx.mcp = mcp
        ctx.save_for_backward(Θ.detach(), sol["x"], sol["y"], sol["s"], sol["eps"])
        ctx.status = sol["status"]
        ctx.mark_non_differentiable(sol["status"])
        return sol["x"], sol["y"], sol["s"], sol["status"]

    @staticmethod
    def backward(ctx, gx, gy, gs, _gstatus):
        Θ, x, y, s, ϵ = ctx.saved_tensors
        B = Θ.shape[0]
        z = lambda g, ref: torch.zeros_like(ref) if g is None else g.to(torch.float64)
        zbar = torch.cat([z(gx, x), z(gy, y), z(gs, s)], dim=1).contiguous()
        θbar = pullback_device(ctx.mcp, Θ, x, y, s, ϵ, zbar)
        return None, θbar, None, None, None
