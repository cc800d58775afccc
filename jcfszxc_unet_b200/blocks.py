"""Stand-alone execution of the shared blocks (DoubleConv / Down / Up / OutConv used outside a fused
model plan, e.g. by block-level parity tests or by user code composing them by hand).

Each call runs a small Plan of the same C-ABI ops the model plans use.  Only the layout adaptation at the
boundary (NCHW-logical torch tensor <-> NHWC bf16 plan buffer) is a torch copy; no torch kernel computes.
"""
from __future__ import annotations

import torch

from . import bridge, engine
from .engine import Act, ConvBNReLU, ConvT2x2, Head, MaxPool2x2, Plan


class _BlockPlan:
    def __init__(self, plan: Plan, inputs: list[Act], output, uses_image: bool):
        self.plan, self.inputs, self.output, self.uses_image = plan, inputs, output, uses_image


def _to_plan(dst: torch.Tensor, x: torch.Tensor):
    dst.copy_(x.permute(0, 2, 3, 1))  # boundary layout/dtype adaptation


class _BlockFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bp: _BlockPlan, n_in: int, *args):
        xs = args[:n_in]
        plan = bp.plan
        if bp.uses_image:
            plan.forward(xs[0])
        else:
            for a, x in zip(bp.inputs, xs):
                _to_plan(a.t, x)
            plan.forward(None)
        ctx.bp, ctx.n_in, ctx.generation = bp, n_in, plan.generation
        ctx.in_dtypes = [x.dtype for x in xs]
        ctx.in_needs = [x.requires_grad for x in xs]
        if isinstance(bp.output, Head):
            return bp.output.logits.clone()
        return bp.output.t.permute(0, 3, 1, 2).clone(memory_format=torch.channels_last)

    @staticmethod
    def backward(ctx, dout):
        bp, plan = ctx.bp, ctx.bp.plan
        if plan.generation != ctx.generation:
            raise RuntimeError("jcfszxc_unet_b200: saved activations were overwritten by a later forward of this block")
        if isinstance(bp.output, Head):
            bp.output.dlogits = dout.contiguous().float()
            bp.output.labels = None
        else:
            _to_plan(bp.output.g, dout)
        plan.backward()
        gin = []
        for i in range(ctx.n_in):
            if bp.uses_image or not ctx.in_needs[i]:
                gin.append(None)
            else:
                gin.append(bp.inputs[i].g.permute(0, 3, 1, 2).to(ctx.in_dtypes[i]))
        grads = tuple(g if p.requires_grad else None for p, g in zip(plan.params, plan.grads()))
        return (None, None) + tuple(gin) + grads


def _run(module, key_extra, build, xs):
    for x in xs:
        bridge.require_cuda_input(x, type(module).__name__)
    need_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in module.parameters()) or
                                             any(x.requires_grad for x in xs))
    key = (tuple(tuple(x.shape) for x in xs), xs[0].device.index, module.training, need_grad, key_extra)
    cache = bridge.plan_cache(module)
    bp = cache.lookup(key)
    if bp is None:
        bp = cache.insert(key, build(module.training, need_grad))
    if need_grad:
        return _BlockFunction.apply(bp, len(xs), *xs, *bp.plan.params)
    with torch.no_grad():
        return _BlockFunction.forward(_Ctx(), bp, len(xs), *xs)


class _Ctx:
    """Throw-away ctx for the no-grad path."""


def _emit_double_conv(P: Plan, x, dc, out: Act | None = None) -> Act:
    seq = dc.double_conv
    h, w = (P.H, P.W) if isinstance(x, engine.Image) else (x.H, x.W)
    mid = P.act(h, w, seq[0].out_channels)
    ConvBNReLU(P, x, seq[0], seq[1], mid)
    out = out if out is not None else P.act(h, w, seq[3].out_channels)
    ConvBNReLU(P, mid, seq[3], seq[4], out)
    return out


def _check_c(c, who):
    if c % 8 != 0:
        raise ValueError(f"{who}: channel count {c} must be a multiple of 8 on the tensor-core path (or <= 4 for a stem)")


def run_double_conv(mod, x):
    n, c, h, w = x.shape

    def build(training, need_grad):
        P = Plan(x.device, n, h, w, training, need_grad)
        if c <= 4:
            out = _emit_double_conv(P, P.image, mod)
            return _BlockPlan(P.finalize(), [], out, True)
        _check_c(c, "DoubleConv")
        xin = P.act(h, w, c)
        out = _emit_double_conv(P, xin, mod)
        return _BlockPlan(P.finalize(), [xin], out, False)

    return _run(mod, "dc", build, [x])


def run_down(mod, x):
    n, c, h, w = x.shape
    _check_c(c, "Down")

    def build(training, need_grad):
        P = Plan(x.device, n, h, w, training, need_grad)
        xin = P.act(h, w, c)
        pooled = P.act(h // 2, w // 2, c)
        MaxPool2x2(P, xin, pooled)
        out = _emit_double_conv(P, pooled, mod.maxpool_conv[1])
        return _BlockPlan(P.finalize(), [xin], out, False)

    return _run(mod, "down", build, [x])


def run_up(mod, x1, x2):
    n, c1, h1, w1 = x1.shape
    _, c2, h2, w2 = x2.shape
    dy_, dx_ = h2 - 2 * h1, w2 - 2 * w1
    if dy_ < 0 or dx_ < 0:
        raise ValueError(f"Up: the skip ({h2}x{w2}) is smaller than the up-sampled input ({2 * h1}x{2 * w1}); negative F.pad "
                         "(cropping) never happens inside a U-Net and is not on this path")
    cup = mod.up.out_channels
    _check_c(c1, "Up")
    _check_c(c2, "Up")

    def build(training, need_grad):
        P = Plan(x1.device, n, h2, w2, training, need_grad)
        lo = P.act(h1, w1, c1)
        cat = P.act(h2, w2, c2 + cup)
        skip = cat.slice(0, c2)            # cat([x2, up(x1)]) order of unet_parts.py:69
        if dy_ or dx_:                     # F.pad branch of unet_parts.py:64-67
            tmp = P.act(2 * h1, 2 * w1, cup)
            ConvT2x2(P, lo, mod.up, tmp)
            engine.PadInto(P, tmp, cat.slice(c2, cup), dy_ // 2, dx_ // 2)
        else:
            ConvT2x2(P, lo, mod.up, cat.slice(c2, cup))
        out = _emit_double_conv(P, cat, mod.conv)
        return _BlockPlan(P.finalize(), [lo, skip], out, False)

    return _run(mod, "up", build, [x1, x2])


def run_out_conv(mod, x):
    n, c, h, w = x.shape

    def build(training, need_grad):
        P = Plan(x.device, n, h, w, training, need_grad)
        xin = P.act(h, w, c)
        head = Head(P, xin, mod.conv)
        P.head = head
        return _BlockPlan(P.finalize(), [xin], head, False)

    return _run(mod, "outc", build, [x])


def run_emit(mod, key, xs, emit):
    """Generic stand-alone execution of a block of the variants: `emit(P, acts)` wires the block into plan P from
    the input activations (the plan's Image when the first input has <= 4 channels) and returns the output Act."""
    n = xs[0].shape[0]
    h, w = xs[0].shape[2], xs[0].shape[3]
    image = xs[0].shape[1] <= 4
    if not image:
        for x in xs:
            _check_c(x.shape[1], type(mod).__name__)

    def build(training, need_grad):
        P = Plan(xs[0].device, n, h, w, training, need_grad)
        acts = [P.image] if image else [P.act(x.shape[2], x.shape[3], x.shape[1]) for x in xs]
        out = emit(P, acts)
        return _BlockPlan(P.finalize(), [] if image else acts, out, image)

    return _run(mod, key, build, list(xs))
