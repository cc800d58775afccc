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
# gradients: sums over up to 4.2 M pixels of products of bf16 values whose noise every layer above has amplified; the
# reference's OWN bf16 weight gradients of a single RRCNN block are 0.07-0.20 away from fp32.  A gradient tensor is
# asserted within max(5e-2, 2 x the reference's own bf16 deviation), never looser than 0.25 (a missing term, a wrong
# scale or a sign error shows up as 0.3 ... 2).
GRAD_FLOOR, GRAD_CEILING, GRAD_SLACK = 5e-2, 2.5e-1, 2.0


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


def check_close(ours, ref32, ref16, what: str, floor: float = LOGIT_FLOOR, slack: float = 1.25, ceiling: float = LOGIT_CEILING,
                use_max: bool = True, info_slack: float | None = None) -> str:
    """The acceptance rule.  The oracle has two precision modes: fp32, and the same functions under
    bf16 autocast (stock cuDNN kernels on the same GPU: the like-for-like oracle).  A bf16 tensor of ours PASSES if

      (fp32)  it is within min(max(floor, slack x the reference's own bf16-vs-fp32 deviation), ceiling) of the fp32 oracle
              — possible only where the reference's own bf16 run is within ceiling / slack of its fp32 run — or
      (like)  it is within `floor` (tensor-level; worst element within `ceiling`) of the bf16-autocast oracle, which rounds
              at the same places (conv output, BatchNorm+ReLU output, pool): bf16 noise that BOTH runs amplify the same way
              through the layers cancels in this comparison.

    If neither criterion CAN apply because the reference's own two runs are further apart than the ceiling (randomly
    initialised gated / recurrent variants, back-propagated gradients through ~20 layers), the input is UNINFORMATIVE:
    nothing wider is asserted except a sanity bound that claims no parity (not further from fp32 than 1.5 x the
    reference's own bf16 run), and the same arithmetic is asserted where the comparison means something (teacher-forced
    blocks).  Returns "fp32" | "like" | "uninformative"; raises AssertionError when an informative comparison fails."""
    l2, mx = l2rel(ours, ref32), rel(ours, ref32)
    l2_ref, mx_ref = l2rel(ref16, ref32), rel(ref16, ref32)
    d_l2, d_mx = l2rel(ours, ref16), rel(ours, ref16)
    # informative: the reference's own two runs are close enough for the ceiling to leave room (info_slack x their distance)
    info_slack = slack if info_slack is None else info_slack
    info = info_slack * l2_ref <= ceiling and (not use_max or info_slack * mx_ref <= ceiling)
    tol_l2, tol_mx = min(max(floor, slack * l2_ref), ceiling), min(max(floor, slack * mx_ref), ceiling)
    fp32_ok = info and l2 <= tol_l2 and (not use_max or mx <= tol_mx)
    like_ok = d_l2 <= floor and (not use_max or d_mx <= ceiling)
    status = "fp32" if fp32_ok else ("like" if like_ok else ("FAIL" if info else "uninformative"))
    msg = (f"{what}: ours vs fp32 l2 {l2:.4g} max {mx:.4g} | reference bf16-autocast vs fp32 l2 {l2_ref:.4g} max {mx_ref:.4g} | "
           f"ours vs reference-bf16 l2 {d_l2:.4g} max {d_mx:.4g} | tol(fp32) l2 {tol_l2:.3g}"
           + (f" max {tol_mx:.3g}" if use_max else "") + f", tol(like) l2 {floor:.3g} -> {status}")
    record(msg)
    assert status != "FAIL", msg
    if status == "uninformative":
        assert l2 <= 1.5 * l2_ref + floor, "sanity bound: " + msg
    return status


def assert_close_bf16(ours, ref32, ref16, what: str, floor: float = LOGIT_FLOOR, slack: float = 1.25,
                      ceiling: float = LOGIT_CEILING, l2_only: bool = False) -> None:
    """check_close for a comparison that must be informative (whole-model outputs of the plain variants, activations)."""
    status = check_close(ours, ref32, ref16, what, floor, slack, ceiling, use_max=not l2_only)
    assert status != "uninformative", f"{what}: expected an informative comparison"


def check_param_grads(ours: dict, g32: dict, g16: dict, tag: str) -> tuple:
    """Whole-model parameter gradients, one check_close per tensor (tensor-level error only: single elements of a
    gradient are sums with cancellation).  Returns (#asserted, #uninformative)."""
    import torch

    ours, g32, g16 = _group_scalars(ours), _group_scalars(g32), _group_scalars(g16)
    gmax = max(float(v.abs().max()) for v in g32.values())
    n_ok = n_weak = 0
    for k in g32:
        if float(g32[k].abs().max()) < 1e-4 * gmax:
            continue   # conv bias in front of a train-mode BatchNorm: zero gradient up to rounding noise
        st = check_close(ours[k], g32[k], g16[k], f"{tag} d {k}", GRAD_FLOOR, GRAD_SLACK, GRAD_CEILING, use_max=False, info_slack=1.5)
        n_ok += st != "uninformative"
        n_weak += st == "uninformative"
    keys = [k for k in g32]
    tot = l2rel(torch.cat([ours[k].flatten().float() for k in keys]), torch.cat([g32[k].flatten() for k in keys]))
    tot_ref = l2rel(torch.cat([g16[k].flatten().float() for k in keys]), torch.cat([g32[k].flatten() for k in keys]))
    tot_like = l2rel(torch.cat([ours[k].flatten().float() for k in keys]), torch.cat([g16[k].flatten().float() for k in keys]))
    record(f"{tag} gradients: whole vector ours vs fp32 l2 {tot:.4g} | reference bf16-autocast vs fp32 l2 {tot_ref:.4g} | ours vs "
           f"reference-bf16 l2 {tot_like:.4g} | {n_ok} tensors asserted, {n_weak} uninformative (sanity bound only)")
    return n_ok, n_weak


# ---------------------------------------------------------------------------------------------------------------
# Teacher forcing: every block of a model checked on the fp32 oracle's own activations
# ---------------------------------------------------------------------------------------------------------------
BLOCK_FUNCS = ("double_conv", "down", "up", "out_conv", "conv_block", "up_conv", "rrcnn_block", "attention_block",
               "residual_conv", "recurrent_block")
INNER_FUNCS = ("recurrent_block",)    # also recorded one level down (inside rrcnn_block: the finer, less chaotic unit)


class BlockTape:
    """Records (function name, prefix, input tensors, trailing arguments) of every OUTERMOST block call (and of the
    INNER_FUNCS one level down) of an oracle forward (oracle/unet_oracle.py looks its block functions up in module globals at call time, so wrapping the
    module attributes is enough)."""

    def __init__(self, O):
        self.O, self.records, self._depth, self._saved = O, [], 0, {}

    def __enter__(self):
        import torch

        for name in BLOCK_FUNCS:
            fn = getattr(self.O, name)
            self._saved[name] = fn

            def wrapper(*args, _fn=fn, _name=name):
                if self._depth > (1 if _name in INNER_FUNCS else 0):
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

    from jcfszxc_unet_b200 import clear_plans

    dev = images.device
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    with torch.no_grad(), BlockTape(O) as tape:
        O.FORWARDS[name](images, {k: v.clone() for k, v in sd.items()}, True)
    assert tape.records, "no block recorded"
    model.train()
    gen = torch.Generator(device=dev).manual_seed(1234)
    counts = {"fp32": 0, "like": 0, "uninformative": 0}
    for fname, prefix, tensors, extra in tape.records:
        sub = model.get_submodule(prefix.rstrip("."))
        ins = [t.float().bfloat16().float() for t in tensors]           # the values both sides see
        image_in = ins[0].shape[1] <= 4
        fn = getattr(O, fname)
        local = {k: v for k, v in sd.items() if k.startswith(prefix)}
        pnames = [k for k in O.param_names(local)]
        what = f"{tag}{name} {prefix.rstrip('.')} ({fname}, in {'+'.join(str(tuple(t.shape)) for t in ins)})"

        def oracle(bf16):
            # the "fp32" role is played by float64 on the GPU: cuDNN's fp32 (TF32 off) convolution on channels_last input
            # returns wrong results at some benchmark shapes (profiles/r02_cudnn_fp32_channels_last_wrong.txt)
            cast = (lambda t: t.clone()) if bf16 else (lambda t: t.double() if t.is_floating_point() else t.clone())
            s = {k: cast(v) for k, v in local.items()}
            for k in pnames:
                s[k].requires_grad_(backward)
            xs = [cast(t).requires_grad_(backward and not image_in) for t in ins]
            with O.autocast_ctx(dev.type, bf16):
                y = fn(*xs, s, prefix, *extra)
            return y, xs, s

        y32, x32, s32 = oracle(False)
        y16, x16, s16 = oracle(True)
        xo = [t.clone().requires_grad_(backward and not image_in) for t in ins]
        with torch.set_grad_enabled(backward):
            yo = sub(*xo)
        counts[check_close(yo.detach(), y32.detach(), y16.detach(), what + " output")] += 1
        if not backward:
            continue
        gy = torch.randn(y32.shape, device=dev, generator=gen).bfloat16().float()
        (yo.float() * gy).sum().backward()
        (y32 * gy.double()).sum().backward()
        (y16.float() * gy).sum().backward()
        if not image_in:
            for i, (a, b, c) in enumerate(zip(xo, x32, x16)):
                counts[check_close(a.grad, b.grad, c.grad, what + f" d input{i}", GRAD_FLOOR, GRAD_SLACK, GRAD_CEILING, use_max=False, info_slack=1.5)] += 1
        ours = _group_scalars({k: p.grad for k, p in ((prefix + n_, p_) for n_, p_ in sub.named_parameters())})
        g32 = _group_scalars({k: s32[k].grad for k in pnames})
        g16 = _group_scalars({k: s16[k].grad.float() for k in pnames})
        gmax = max(float(v.abs().max()) for v in g32.values())
        for k in g32:
            if float(g32[k].abs().max()) < 1e-4 * gmax:
                continue   # conv bias in front of a train-mode BatchNorm: zero gradient up to rounding noise
            counts[check_close(ours[k], g32[k], g16[k], what + f" d {k}", GRAD_FLOOR, GRAD_SLACK, GRAD_CEILING, use_max=False, info_slack=1.5)] += 1
        sub.zero_grad(set_to_none=True)
        clear_plans(sub)
    record(f"{tag}{name}: {len(tape.records)} blocks teacher-forced; comparisons passed against the fp32 oracle: {counts['fp32']}, against "
           f"the bf16-autocast oracle: {counts['like']}, uninformative (sanity bound only): {counts['uninformative']}")
    done = counts["fp32"] + counts["like"]
    assert counts["uninformative"] <= 0.3 * (done + counts["uninformative"]), "too few informative comparisons"
    return len(tape.records)
