"""nn.Module <-> Plan bridge: plan cache per (shape, mode) and the autograd.Function that makes a
whole-network plan differentiable, so `model(x)` is a drop-in for the reference modules
(train.py:256, evaluate.py:259-275 in the reference) while every FLOP runs in libunetk.so."""
from __future__ import annotations

import collections
import os

import torch

from . import _lib


class PlanCache(collections.OrderedDict):
    """Per-module LRU of execution plans, keyed by (shape, device, mode).

    It lives in the module's `__dict__` (so it dies with the module: plans hold references to the module's parameters)
    but never travels: `torch.save(model)` — the reference's checkpoint format, train.py:374 — and `copy.deepcopy` see an
    EMPTY builtin dict in its place.  A plan owns CUDA streams (not picklable), every activation / gradient buffer of
    its shape (GBs) and tables keyed by id(parameter) (meaningless in another process); and a pickle that named a class
    of this package could not be loaded by the reference.  At most UNETK_PLAN_CACHE (default 4) plans are kept per
    module; the least recently used one is dropped (ragged last chunks of predict_full_image, the full-validation-set
    batch, varying batch sizes)."""

    LIMIT = max(1, int(os.environ.get("UNETK_PLAN_CACHE", "4")))

    def __reduce__(self):
        return (dict, ())

    def __reduce_ex__(self, protocol):
        return (dict, ())

    def __deepcopy__(self, memo):
        return {}

    def __copy__(self):
        return {}

    def lookup(self, key):
        plan = self.get(key)
        if plan is not None:
            self.move_to_end(key)
        return plan

    def insert(self, key, plan):
        self[key] = plan
        while len(self) > self.LIMIT:
            self.popitem(last=False)
        return plan


def plan_cache(module: torch.nn.Module) -> PlanCache:
    cache = module.__dict__.get("_unetk_plans")
    if not isinstance(cache, PlanCache):     # first use, or a plain dict left by unpickling / deepcopy
        cache = module.__dict__["_unetk_plans"] = PlanCache()
    return cache


def clear_plans(module: torch.nn.Module) -> None:
    """Drop every cached plan (and its activation / gradient buffers) of `module` and of its sub-modules."""
    for m in module.modules():
        c = m.__dict__.get("_unetk_plans")
        if c:
            c.clear()


def plan_heads(plan):
    """The output heads of a plan: one, or four with UNet++'s deep supervision (UNetPP.py:93-102)."""
    return getattr(plan, "heads", None) or [plan.head]


class _PlanFunction(torch.autograd.Function):
    """forward: plan.forward(x) -> fp32 logits (one tensor per head); backward: plan.backward(dlogits) -> parameter
    gradients.  Parameters are passed as inputs only so that autograd routes their gradients."""

    @staticmethod
    def forward(ctx, plan, x, *params):
        heads = plan_heads(plan)
        for h in heads:
            h.labels = None
        plan.forward(x)
        ctx.plan = plan
        ctx.generation = plan.generation
        outs = tuple(h.logits.clone() for h in heads)
        return outs if len(outs) > 1 else outs[0]

    @staticmethod
    def backward(ctx, *dlogits):
        plan = ctx.plan
        if plan.generation != ctx.generation:
            raise RuntimeError(
                "jcfszxc_unet_b200: the activations saved for this backward were overwritten by a later forward "
                "of the same shape; call backward before the next forward (static-buffer plans)")
        heads = plan_heads(plan)
        for h, d in zip(heads, dlogits):
            h.labels = None
            # an output the loss did not use gets a zero gradient
            h.dlogits = d.contiguous().float() if d is not None else torch.zeros_like(h.logits)
            h.gscale = 1.0
        plan.backward()
        plan.awaiting_backward = False
        for h in heads:
            h.dlogits = None
        # views of the plan's gradient buffers: autograd's AccumulateGrad copies them into .grad (it cannot steal a
        # tensor the plan still references), so no second copy is made here
        grads = tuple(g if p.requires_grad else None for p, g in zip(plan.params, plan.grads()))
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
    if need_grad and x.requires_grad:
        raise NotImplementedError(
            f"{type(model).__name__}: the input requires grad, but the fused plan does not produce d(loss)/d(image) "
            "(the stem's input gradient is not on the training path, train.py:255-301); detach the input")
    plans = plan_cache(model)
    plan = plans.lookup(key)
    if plan is None:
        # a plan with gradient buffers only when a backward can follow; BN mode follows model.training
        plan = plans.insert(key, builder(model, n, h, w, x.device, training_stats, None, need_grad))
    if need_grad:
        if getattr(plan, "awaiting_backward", False):
            # a second forward of this shape while the first one's backward is still to come (loss(model(a)) + loss(model(b)),
            # train.py:256-297 semantics): a plan's activations are static buffers, so the second forward gets a twin plan
            # instead of overwriting them.  One twin per shape: a third un-backwarded forward reuses the older plan, whose
            # stale backward then raises (generation check) rather than returning wrong gradients.
            twin = plans.lookup(key + ("twin",))
            if twin is None:
                twin = plans.insert(key + ("twin",), builder(model, n, h, w, x.device, training_stats, None, need_grad))
            if not getattr(twin, "awaiting_backward", False):
                plan = twin
        plan.awaiting_backward = True
        out = _PlanFunction.apply(plan, x, *plan.params)
        return list(out) if isinstance(out, tuple) else out
    with torch.no_grad():
        heads = plan_heads(plan)
        for h in heads:
            h.labels = None
        plan.forward(x)
        outs = [h.logits.clone() for h in heads]
        return outs if len(outs) > 1 else outs[0]
