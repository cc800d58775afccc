"""Shared acceptance rule of the bf16 parity tests, and the parity report.

Rule ("tighten the yardstick").  A bf16 tensor of ours must agree with the fp32 oracle within

    tol = min( max(floor, slack x the reference's OWN bf16-autocast deviation from its fp32 run), ceiling )

both as tensor-level relative error ||ours - ref|| / ||ref|| and as worst element relative to max|ref|.
`floor` is BASELINE.json's tolerance (2e-2 for logits / activations), `ceiling` is absolute (5e-2 for logits): no
assertion ever relies on more slack than the ceiling.  Where the reference's own bf16 run is further than
`ceiling / slack` from its fp32 run, the INPUT is uninformative for a whole-tensor comparison (randomly initialised
gated / recurrent variants amplify every bf16 rounding by ~1.2x per layer: AttentionUNet 0.26, R2UNet 0.65,
R2AttentionUNet 1.1 at any image size, profiles/r02_parity_report.txt) — such a case does not get a wider gate: the
caller must check the same arithmetic in a regime where the comparison means something (per-block teacher forcing,
tests/test_gpu_fullsize.py::test_blocks_teacher_forced_*), and `informative()` tells it so.

Every comparison appends one line to the parity report (gpurun_out/r02_parity_report.txt on the GPU box, copied to
profiles/ after the run), so the slack actually used is visible.
"""
from __future__ import annotations

import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.environ.get("UNETK_PARITY_REPORT", os.path.join(ROOT, "gpurun_out", "r02_parity_report.txt"))

LOGIT_FLOOR, LOGIT_CEILING = 2e-2, 5e-2        # north_star: "logits within 2e-2 relative"
GRAD_FLOOR, GRAD_CEILING = 5e-2, 1.5e-1        # gradients: bf16 noise of every layer above accumulates


def record(line: str) -> None:
    print(line)
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "a") as f:
            f.write(line + "\n")
    except OSError:
        pass


def rel(a, b) -> float:
    return (a.float() - b.float()).abs().max().item() / (b.float().abs().max().item() + 1e-12)


def l2rel(a, b) -> float:
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


def informative(ref32, ref16, slack: float, ceiling: float) -> bool:
    """True if the reference's own bf16 deviation on this input leaves room under the ceiling."""
    return slack * max(l2rel(ref16, ref32), rel(ref16, ref32)) <= ceiling


def assert_close_bf16(ours, ref32, ref16, what: str, floor: float = LOGIT_FLOOR, slack: float = 1.25,
                      ceiling: float = LOGIT_CEILING, l2_only: bool = False) -> None:
    l2, mx = l2rel(ours, ref32), rel(ours, ref32)
    l2_ref, mx_ref = l2rel(ref16, ref32), rel(ref16, ref32)
    d16 = l2rel(ours, ref16)
    tol_l2 = min(max(floor, slack * l2_ref), ceiling)
    tol_mx = min(max(floor, slack * mx_ref), ceiling)
    msg = (f"{what}: ours vs fp32 l2 {l2:.4g} max {mx:.4g} | reference bf16-autocast vs fp32 l2 {l2_ref:.4g} max {mx_ref:.4g} | "
           f"ours vs reference-bf16 l2 {d16:.4g} | tol l2 {tol_l2:.3g} max {tol_mx:.3g}")
    record(msg)
    assert l2 <= tol_l2, msg
    if not l2_only:
        assert mx <= tol_mx, msg


def check_param_grads(ours: dict, g32: dict, g16: dict, tag: str) -> tuple:
    """Whole-model parameter gradients.  A parameter whose reference-bf16 gradient is within GRAD_CEILING / 2.5 of the
    fp32 one is ASSERTED under the ceiling rule.  For the others the input is uninformative (back-propagation through
    ~20 randomly initialised layers amplifies bf16 noise: the reference's own bf16 gradients of the vanilla UNet's
    encoder are ~0.5 away from its fp32 gradients at every size tried): they only get a sanity bound (no further from
    fp32 than 1.5 x the reference's own bf16 path — catches sign / scale / missing-term bugs, claims no parity) and
    are asserted per block by the teacher-forced tests.  Returns (#asserted, #sanity-only)."""
    import torch

    ours, g32, g16 = _group_scalars(ours), _group_scalars(g32), _group_scalars(g16)
    gmax = max(float(v.abs().max()) for v in g32.values())
    n_ok = n_weak = 0
    worst = (0.0, "")
    for k in g32:
        if float(g32[k].abs().max()) < 1e-4 * gmax:
            continue   # conv bias in front of a train-mode BatchNorm: zero gradient up to rounding noise
        e, e_ref = l2rel(ours[k], g32[k]), l2rel(g16[k], g32[k])
        if 2.5 * e_ref <= GRAD_CEILING:
            tol = max(GRAD_FLOOR, 2.5 * e_ref)
            n_ok += 1
        else:
            tol = 1.5 * e_ref
            n_weak += 1
        worst = max(worst, (e / tol, f"{k}: ours {e:.4g} reference-bf16 {e_ref:.4g} tol {tol:.3g}"))
        assert e <= tol, f"{tag} d {k}: ours {e:.4g} vs reference bf16-autocast {e_ref:.4g} (tol {tol:.3g})"
    keys = [k for k in g32]
    tot = l2rel(torch.cat([ours[k].flatten().float() for k in keys]), torch.cat([g32[k].flatten() for k in keys]))
    tot_ref = l2rel(torch.cat([g16[k].flatten().float() for k in keys]), torch.cat([g32[k].flatten() for k in keys]))
    record(f"{tag} gradients: whole vector ours vs fp32 l2 {tot:.4g} | reference bf16-autocast vs fp32 l2 {tot_ref:.4g} | "
           f"{n_ok} tensors asserted under the ceiling, {n_weak} uninformative (sanity bound only) | closest to its bound: {worst[1]}")
    return n_ok, n_weak


# ---------------------------------------------------------------------------------------------------------------
# Teacher forcing: every block of a model checked on the fp32 oracle's own activations
# ---------------------------------------------------------------------------------------------------------------
BLOCK_FUNCS = ("double_conv", "down", "up", "out_conv", "conv_block", "up_conv", "rrcnn_block", "attention_block",
               "residual_conv")


class BlockTape:
    """Records (function name, prefix, input tensors, trailing arguments) of every OUTERMOST block call of an oracle
    forward (oracle/unet_oracle.py looks its block functions up in module globals at call time, so wrapping the
    module attributes is enough)."""

    def __init__(self, O):
        self.O, self.records, self._depth, self._saved = O, [], 0, {}

    def __enter__(self):
        import torch

        for name in BLOCK_FUNCS:
            fn = getattr(self.O, name)
            self._saved[name] = fn

            def wrapper(*args, _fn=fn, _name=name):
                if self._depth:
                    return _fn(*args)
                tensors = [a.detach() for a in args if torch.is_tensor(a)]
                k = next(i for i, a in enumerate(args) if isinstance(a, str))
                self._depth += 1
                try:
                    out = _fn(*args)
                finally:
                    self._depth -= 1
                self.records.append((_name, args[k], tensors, tuple(args[k + 1:])))
                return out

            setattr(self.O, name, wrapper)
        return self

    def __exit__(self, *exc):
        for name, fn in self._saved.items():
            setattr(self.O, name, fn)
        return False


def _group_scalars(grads: dict) -> dict:
    """Single-element gradients (BN of the 1-channel psi conv, output-conv bias) are sums with heavy cancellation:
    they are judged together as one vector."""
    import torch

    out, scalars = {}, []
    for k, v in grads.items():
        if v.numel() == 1:
            scalars.append(v.reshape(1).float())
        else:
            out[k] = v
    if scalars:
        out["<single-element parameters>"] = torch.cat(scalars)
    return out


def check_blocks_teacher_forced(O, model, name: str, images, backward: bool = True, tag: str = "") -> int:
    """Run the fp32 oracle forward of `name` on `images`, recording every block's inputs; then run OUR stand-alone block
    (the same sm_100a kernels and dispatch as the fused plan, at the block's real shape) and the oracle's block in fp32
    and in bf16-autocast on those SAME inputs (rounded to bf16), forward and backward, and compare under the ceiling
    rule.  Errors cannot accumulate across blocks, so the comparison is informative for every variant.
    Returns the number of blocks checked."""
    import torch

    dev = images.device
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    with torch.no_grad(), BlockTape(O) as tape:
        O.FORWARDS[name](images, {k: v.clone() for k, v in sd.items()}, True)
    assert tape.records, "no block recorded"
    model.train()
    gen = torch.Generator(device=dev).manual_seed(1234)
    for fname, prefix, tensors, extra in tape.records:
        sub = model.get_submodule(prefix.rstrip("."))
        ins = [t.float().bfloat16().float() for t in tensors]           # the values both sides see
        image_in = ins[0].shape[1] <= 4
        fn = getattr(O, fname)
        local = {k: v for k, v in sd.items() if k.startswith(prefix)}
        pnames = [k for k in O.param_names(local)]
        what = f"{tag}{name} {prefix.rstrip('.')} ({fname}, in {'+'.join(str(tuple(t.shape)) for t in ins)})"

        def oracle(bf16):
            s = {k: v.clone() for k, v in local.items()}
            for k in pnames:
                s[k].requires_grad_(backward)
            xs = [t.clone().requires_grad_(backward and not image_in) for t in ins]
            with O.autocast_ctx(dev.type, bf16):
                y = fn(*xs, s, prefix, *extra)
            return y, xs, s

        y32, x32, s32 = oracle(False)
        y16, x16, s16 = oracle(True)
        xo = [t.clone().requires_grad_(backward and not image_in) for t in ins]
        with torch.set_grad_enabled(backward):
            yo = sub(*xo)
        assert_close_bf16(yo.detach(), y32.detach(), y16.detach(), what + " output")
        if not backward:
            continue
        gy = torch.randn(y32.shape, device=dev, generator=gen).bfloat16().float()
        (yo.float() * gy).sum().backward()
        (y32.float() * gy).sum().backward()
        (y16.float() * gy).sum().backward()
        if not image_in:
            for i, (a, b, c) in enumerate(zip(xo, x32, x16)):
                assert_close_bf16(a.grad, b.grad, c.grad, what + f" d input{i}", GRAD_FLOOR, 2.0, GRAD_CEILING)
        ours = _group_scalars({k: p.grad for k, p in ((prefix + n_, p_) for n_, p_ in sub.named_parameters())})
        g32 = _group_scalars({k: s32[k].grad for k in pnames})
        g16 = _group_scalars({k: s16[k].grad.float() for k in pnames})
        gmax = max(float(v.abs().max()) for v in g32.values())
        for k in g32:
            if float(g32[k].abs().max()) < 1e-4 * gmax:
                continue   # conv bias in front of a train-mode BatchNorm: zero gradient up to rounding noise
            assert_close_bf16(ours[k], g32[k], g16[k], what + f" d {k}", GRAD_FLOOR, 2.0, GRAD_CEILING, l2_only=True)
        sub.zero_grad(set_to_none=True)
    return len(tape.records)
