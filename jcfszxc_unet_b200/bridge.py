"""nn.Module <-> Plan bridge: plan cache per (shape, mode) and the autograd.Function that makes a
whole-network plan differentiable, so `model(x)` is a drop-in for the reference modules
(train.py:256, evaluate.py:259-275 in the reference) while every FLOP runs in libunetk.so."""
from __future__ import annotations

import torch

from . import _lib


class _PlanFunction(torch.autograd.Function):
    """forward: plan.forward(x) -> fp32 logits; backward: plan.backward(dlogits) -> parameter gradients.
    Parameters are passed as inputs only so that autograd routes their gradients."""

    @staticmethod
    def forward(ctx, plan, x, *params):
        plan.head.labels = None
        plan.forward(x)
        ctx.plan = plan
        ctx.generation = plan.generation
        return plan.head.logits.clone()

    @staticmethod
    def backward(ctx, dlogits):
        plan = ctx.plan
        if plan.generation != ctx.generation:
            raise RuntimeError(
                "jcfszxc_unet_b200: the activations saved for this backward were overwritten by a later forward "
                "of the same shape; call backward before the next forward (static-buffer plans)")
        plan.head.labels = None
        plan.head.dlogits = dlogits.contiguous().float()
        plan.head.gscale = 1.0
        plan.backward()
        plan.head.dlogits = None
        grads = tuple(g.clone() if p.requires_grad else None for p, g in zip(plan.params, plan.grads()))
        return (None, None) + grads


def require_cuda_input(x: torch.Tensor, who: str):
    if not x.is_cuda:
        raise RuntimeError(
            f"{who}: this is the B200-native path (libunetk.so, sm_100a); it has no CPU fallback. "
            f"Got a {x.device} tensor.")
    _lib.load()  # raises if the extension is missing


def run_model(model: torch.nn.Module, builder, x: torch.Tensor) -> torch.Tensor:
    """Run `model` through its cached plan; differentiable w.r.t. the parameters when grad is enabled."""
    require_cuda_input(x, type(model).__name__)
    if x.dim() != 4:
        raise ValueError(f"expected [N,C,H,W], got {tuple(x.shape)}")
    from . import get_precision
    if get_precision() == "fp32":
        from . import engine, f32
        if builder is not engine.build_unet_plan:
            raise NotImplementedError("fp32 mode covers the vanilla UNet (BASELINE.json configs[0]); "
                                      f"{type(model).__name__} runs in bf16 only")
        return f32.run_unet_f32(model, x)
    n, _, h, w = x.shape
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in model.parameters())
    training_stats = model.training
    key = (n, h, w, x.device.index, training_stats, need_grad)
    plans = model.__dict__.setdefault("_unetk_plans", {})
    plan = plans.get(key)
    if plan is None:
        # a plan with gradient buffers only when a backward can follow; BN mode follows model.training
        plan = builder(model, n, h, w, x.device, training_stats, None, need_grad)
        plans[key] = plan
    if need_grad:
        return _PlanFunction.apply(plan, x, *plan.params)
    with torch.no_grad():
        plan.head.labels = None
        plan.forward(x)
        return plan.head.logits.clone()
