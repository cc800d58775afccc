"""Pin the oracle (oracle/unet_oracle.py) against the UNMODIFIED reference, and write golden vectors.

Runs only where /root/reference exists (this container, not the GPU box).  The reference is imported
under its own package name `UNetFamily` from /root/reference with the 4-line timm shim of SURVEY.md
Appendix A; this repository's own `UNetFamily` package is kept OFF sys.path in this process so the two
cannot collide (the oracle is loaded by file path).

    python oracle/pin_against_reference.py            # assert bit-equality, regenerate tests/golden/*.npz

Checks (all on CPU, fp32, bit-exact unless noted):
  1. UNet.forward (train + eval mode) and every running-stat update        == oracle.unet_forward
  2. the restated train step (train.py:255-301, bf16 autocast and fp32): loss, gradients after
     clip_grad_norm_, parameters after RMSprop.step                         == oracle.train_step
  3. DoubleConv / Down / Up / OutConv blocks at small channel counts        == oracle block functions
  4. utils/dice_score.py dice_coeff / dice_loss incl. the empty-mask branch == oracle + numpy restatement
  5. F.max_pool2d(return_indices=True) incl. ties and NaNs                  == oracle numpy restatement
  6. AttentionUNet / R2UNet / R2AttentionUNet / ResUNet / NestedUNet: train+eval forward, running stats,
     bf16-autocast forward and one restated train step (loss, clipped grads, RMSprop update)
                                                                            == oracle.FORWARDS / train_step
     and their blocks conv_block / up_conv / Recurrent_block / RRCNN_block / Attention_block / ResidualConv
Golden files hold the reference's OUTPUTS for fixed seeds; weights are regenerated from the seed by
constructing the model (default torch init), so fixtures stay small.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _import_reference():
    sys.dont_write_bytecode = True
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
    _t, _l = types.ModuleType("timm"), types.ModuleType("timm.layers")
    _l.trunc_normal_ = torch.nn.init.trunc_normal_
    _t.layers = _l
    sys.modules.setdefault("timm", _t)
    sys.modules.setdefault("timm.layers", _l)
    sys.path.insert(0, REF)
    from UNetFamily import UNet as ref_unet  # noqa
    from UNetFamily.utils import unet_parts as ref_parts  # noqa
    from utils import dice_score as ref_dice  # noqa
    return ref_unet, ref_parts, ref_dice


def _load_oracle():
    spec = importlib.util.spec_from_file_location("unet_oracle", os.path.join(HERE, "unet_oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _inputs(seed, n, h, w, c=3):
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(n, c, h, w, generator=g)
    labels = (torch.rand(n, 1, h, w, generator=g) < 0.12).float()
    return images, labels


def _ref_train_step(model, opt, images, labels, bf16):
    """Line-for-line restatement of train.py:255-301 around the UNMODIFIED reference modules
    (train_model itself cannot be imported: h5py/matplotlib missing and train.py:114-122 breaks on torch 2.11)."""
    import contextlib

    ctx = torch.autocast("cpu", dtype=torch.bfloat16) if bf16 else contextlib.nullcontext()
    criterion = torch.nn.BCEWithLogitsLoss()
    with ctx:
        masks_pred = model(images)
        masks_pred_sigmoid = torch.sigmoid(masks_pred)
        bce_loss = criterion(masks_pred, labels)
        dice = REF_DICE.dice_loss(masks_pred_sigmoid.squeeze(1), labels.squeeze(1), multiclass=False)
        alpha = 0.5
        loss = alpha * bce_loss + (1 - alpha) * dice
    opt.zero_grad(set_to_none=True)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    opt.step()
    return loss.detach(), masks_pred.detach(), bce_loss.detach(), dice.detach(), grads


def _extract(path, picker):
    """Parse a reference script that cannot be imported (h5py / matplotlib at module level) and return the AST nodes
    `picker(tree)` selects — the UNMODIFIED statements, compiled below as they stand."""
    import ast

    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    return picker(tree)


def _pin_sampler_and_tiling(O):
    """(f2) train.py:126-155 + :201-241 (patch pools, the valid-centre filter, the random draw and the slicing loop) and
    (f3) evaluate.py:28-96 (predict_full_image) are executed FROM THE REFERENCE'S SOURCE TEXT: the statements are cut
    out of the files with `ast` (train_model's body cannot be imported — it needs HDF5 files, wandb-style logging and a
    torch-1.x scheduler signature — and evaluate.py imports h5py / matplotlib at module level) and compiled unchanged
    inside thin wrappers that only supply their free variables.  oracle.patch_sample_map / patch_batch /
    predict_full_image must reproduce them bit for bit."""
    import ast

    out = []
    # ---- evaluate.py: predict_full_image is a top-level function: compile it as is ----------------------------------
    fn = _extract(os.path.join(REF, "evaluate.py"),
                  lambda t: next(n for n in t.body if isinstance(n, ast.FunctionDef) and n.name == "predict_full_image"))
    assert (fn.lineno, fn.end_lineno) == (28, 96), (fn.lineno, fn.end_lineno)
    ns = {"np": np, "torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "evaluate.py", "exec"), ns)
    ref_predict = ns["predict_full_image"]

    class Tiny(torch.nn.Module):          # any deterministic fully-convolutional model works as the probe
        def __init__(self):
            super().__init__()
            torch.manual_seed(0)
            self.c = torch.nn.Conv2d(3, 1, 3, padding=1)

        def forward(self, x):
            return self.c(x)

    model = Tiny()
    rng = np.random.RandomState(3)
    for (h, w, ps, ov, bs) in ((70, 90, 32, 0.5, 4), (64, 64, 32, 0.25, 3), (50, 47, 16, 0.5, 5), (40, 40, 64, 0.5, 2)):
        image = rng.rand(h, w, 3).astype(np.float32)
        a = ref_predict(model, torch.device("cpu"), image, patch_size=ps, overlap=ov, batch_size=bs)
        b = O.predict_full_image(lambda t: model(t), image, patch_size=ps, overlap=ov, batch_size=bs)
        assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), (h, w, ps, ov, bs)
    out.append("predict_full_image (evaluate.py:28-96, compiled from the reference's source): bit-exact on 4 tilings incl. "
               "ragged borders and a patch larger than the image")

    # ---- train.py: statements of train_model's body -------------------------------------------------------------------
    def pick(tree):
        tm = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "train_model")
        setup = [n for n in tm.body if 126 <= n.lineno <= 155]

        def find_step_loop(nodes):
            for n in nodes:
                for sub in ast.walk(n):
                    if isinstance(sub, ast.For) and isinstance(sub.target, ast.Name) and sub.target.id == "step":
                        return sub
            raise AssertionError("for step in range(steps) not found")

        loop = find_step_loop(tm.body)
        draw = [n for n in loop.body if 201 <= n.lineno <= 241]
        return setup, draw

    setup, draw = _extract(os.path.join(REF, "train.py"), pick)
    assert setup and setup[0].lineno == 127 and draw and draw[0].lineno == 201 and draw[-1].end_lineno == 241, \
        ([n.lineno for n in setup], [n.lineno for n in draw])
    src = ast.Module(body=[
        ast.FunctionDef(name="ref_setup", args=ast.arguments(posonlyargs=[], args=[ast.arg("train_dataset"), ast.arg("patch_size")],
                                                             kwonlyargs=[], kw_defaults=[], defaults=[]),
                        body=setup + [ast.parse("return images_data_pool, labels_data_pool, filtered_sample_map, half_patch").body[0]],
                        decorator_list=[], type_params=[]),
        ast.FunctionDef(name="ref_batch", args=ast.arguments(posonlyargs=[], args=[ast.arg(a) for a in (
            "images_data_pool", "labels_data_pool", "filtered_sample_map", "half_patch", "batch_size")],
            kwonlyargs=[], kw_defaults=[], defaults=[]),
                        body=draw + [ast.parse("return batch_images, batch_labels").body[0]], decorator_list=[], type_params=[]),
    ], type_ignores=[])
    ast.fix_missing_locations(src)
    ns = {"np": np, "torch": torch}
    exec(compile(src, "train.py", "exec"), ns)
    rng = np.random.RandomState(11)
    n, hw, ps, bs = 3, 48, 16, 6
    images = rng.rand(n, hw, hw, 3).astype(np.float32)              # HDF5 layout: [N, W, H, C]
    masks = (rng.rand(n, hw, hw) < 0.3).astype(np.uint8)
    labels = (rng.rand(n, hw, hw) < 0.1).astype(np.uint8)
    pool, lpool, fmap, half = ns["ref_setup"]({"images": images, "masks": masks, "labels": labels}, ps)
    omap = O.patch_sample_map(masks, ps)
    assert all(np.array_equal(a, b) for a, b in zip(fmap, omap)) and half == ps // 2
    for seed in (0, 1, 2):
        np.random.seed(seed)                                         # the reference draws from the global numpy RNG
        bi, bl = ns["ref_batch"](pool, lpool, fmap, half, bs)
        ri = torch.from_numpy(bi).to(dtype=torch.float32, memory_format=torch.channels_last)     # train.py:245-253
        rl = torch.from_numpy(bl).to(dtype=torch.float32)
        oi, ol = O.patch_batch(images, labels, omap, bs, ps, np.random.RandomState(seed))
        assert torch.equal(ri, oi) and torch.equal(rl, ol) and ri.stride() == oi.stride(), seed
    out.append("patch pools + valid-centre filter (train.py:127-155) and the random draw + slicing loop + np.stack (:201-241), compiled "
               "from the reference's source: sample map and 3 seeded batches bit-exact")
    return out


def main():
    global REF_DICE
    ref_unet, ref_parts, REF_DICE = _import_reference()
    O = _load_oracle()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    torch.use_deterministic_algorithms(False)
    os.makedirs(GOLDEN, exist_ok=True)
    report = []

    # ---- 1. forward, train and eval mode ---------------------------------------------------------
    torch.manual_seed(42)
    model = ref_unet.UNet(3, 1)
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    images, labels = _inputs(7, 2, 32, 32)
    model.train()
    with torch.no_grad():
        y_ref = model(images)
    sd = {k: v.clone() for k, v in sd0.items()}
    with torch.no_grad():
        y_or = O.unet_forward(images, sd, training=True)
    assert torch.equal(y_ref, y_or), "train-mode forward differs"
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd[k]), f"running stat {k} differs"
    model.eval()
    with torch.no_grad():
        y_ref_eval = model(images)
        y_or_eval = O.unet_forward(images, sd, training=False)
    assert torch.equal(y_ref_eval, y_or_eval), "eval-mode forward differs"
    report.append("forward train/eval + running stats: bit-exact")
    # the reference's own bf16-autocast forward on the same weights/inputs: the yardstick for "how far from
    # fp32 does stock bf16 land" that the GPU tests compare our deviation with
    torch.manual_seed(42)
    model_bf = ref_unet.UNet(3, 1).train()
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        y_ref_bf16 = model_bf(images).float()
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        y_or_bf16 = O.unet_forward(images, {k: v.clone() for k, v in sd0.items()}, training=True).float()
    assert torch.equal(y_ref_bf16, y_or_bf16), "bf16-autocast forward differs"
    np.savez_compressed(os.path.join(GOLDEN, "unet_forward_seed42.npz"), images=images.numpy(), labels=labels.numpy(),
                        logits_train_bf16_autocast=y_ref_bf16.numpy(),
                        logits_train=y_ref.numpy(), logits_eval_after_1_train_fwd=y_ref_eval.numpy(),
                        running_mean_inc1=model.state_dict()["inc.double_conv.1.running_mean"].numpy(),
                        running_var_up4_4=model.state_dict()["up4.conv.double_conv.4.running_var"].numpy(),
                        num_batches_tracked=np.int64(model.state_dict()["inc.double_conv.1.num_batches_tracked"].item()))

    # ---- 2. train step (fp32 and bf16 autocast) ---------------------------------------------------
    for bf16 in (False, True):
        torch.manual_seed(42)
        model = ref_unet.UNet(3, 1).train()
        lr = 1e-3  # larger than the reference default (1e-6) so that the parameter update is visible in fp32
        opt = torch.optim.RMSprop(model.parameters(), lr=lr, weight_decay=1e-8, momentum=0.999)
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        opt_state = {k: (torch.zeros_like(sd[k]), torch.zeros_like(sd[k])) for k in O.param_names(sd)}
        gold = {}
        for step in range(2):
            images, labels = _inputs(100 + step, 2, 32, 32)
            loss_r, logits_r, bce_r, dice_r, grads_r = _ref_train_step(model, opt, images, labels, bf16)
            loss_o, logits_o, grads_o = O.train_step(sd, opt_state, images, labels, lr, bf16)
            assert torch.equal(loss_r, loss_o), f"loss differs (bf16={bf16}, step {step}): {loss_r} vs {loss_o}"
            assert torch.equal(logits_r.float(), logits_o.float())
            for k in grads_r:
                assert torch.equal(grads_r[k], grads_o[k]), f"clipped grad {k} differs"
            for k, v in model.state_dict().items():
                assert torch.equal(v, sd[k]), f"post-step tensor {k} differs (bf16={bf16}, step {step})"
            gold[f"loss{step}"] = loss_r.float().numpy()
            gold[f"bce{step}"] = bce_r.float().numpy()
            gold[f"dice_loss{step}"] = dice_r.float().numpy()
            gold[f"logits{step}"] = logits_r.float().numpy()
            gold[f"gradnorm{step}"] = np.array([float(g.float().norm()) for g in grads_r.values()], dtype=np.float64)
            for k in ("outc.conv.weight", "outc.conv.bias", "inc.double_conv.1.weight", "inc.double_conv.1.bias",
                      "up4.conv.double_conv.4.weight", "down4.maxpool_conv.1.double_conv.4.bias", "up1.up.bias",
                      "inc.double_conv.0.weight"):
                gold[f"grad{step}:{k}"] = grads_r[k].float().numpy()
                gold[f"param{step}:{k}"] = model.state_dict()[k].detach().float().clone().numpy()
        gold["param_names"] = np.array(list(grads_r.keys()))
        np.savez_compressed(os.path.join(GOLDEN, f"unet_trainstep_seed42_{'bf16' if bf16 else 'fp32'}.npz"), **gold)
        report.append(f"train step x2 ({'bf16 autocast' if bf16 else 'fp32'}): loss, clipped grads, RMSprop update bit-exact")

    # ---- 3. blocks -------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(5)
    blocks = {}
    torch.manual_seed(3)
    dc = ref_parts.DoubleConv(8, 16).train()
    x = torch.randn(2, 8, 12, 20, generator=g)
    sd = {k: v.detach().clone() for k, v in dc.state_dict().items()}
    with torch.no_grad():
        assert torch.equal(dc(x), O.double_conv(x, sd, "", True))
    blocks.update(dc_x=x.numpy(), dc_y=dc(x).detach().numpy())
    torch.manual_seed(4)
    dn = ref_parts.Down(8, 16).train()
    sd = {k: v.detach().clone() for k, v in dn.state_dict().items()}
    with torch.no_grad():
        assert torch.equal(dn(x), O.down(x, sd, "", True))
    blocks.update(down_y=dn(x).detach().numpy())
    torch.manual_seed(5)
    upm = ref_parts.Up(16, 8).train()
    x1 = torch.randn(2, 16, 6, 10, generator=g)
    x2 = torch.randn(2, 8, 12, 20, generator=g)
    sd = {k: v.detach().clone() for k, v in upm.state_dict().items()}
    with torch.no_grad():
        assert torch.equal(upm(x1, x2), O.up(x1, x2, sd, "", True))
    blocks.update(up_x1=x1.numpy(), up_x2=x2.numpy(), up_y=upm(x1, x2).detach().numpy())
    # odd sizes exercise the F.pad branch (unet_parts.py:64-67) of the oracle
    x2o = torch.randn(2, 8, 13, 21, generator=g)
    with torch.no_grad():
        assert torch.equal(upm(x1, x2o), O.up(x1, x2o, {k: v.detach().clone() for k, v in upm.state_dict().items()}, "", True))
    torch.manual_seed(6)
    oc = ref_parts.OutConv(8, 1)
    sd = {k: v.detach().clone() for k, v in oc.state_dict().items()}
    with torch.no_grad():
        assert torch.equal(oc(x), O.out_conv(x, sd, ""))
    blocks.update(outc_y=oc(x).detach().numpy())
    np.savez_compressed(os.path.join(GOLDEN, "blocks_seeds3to6.npz"), **blocks)
    report.append("DoubleConv / Down / Up (+pad branch) / OutConv: bit-exact")

    # ---- 4. dice ---------------------------------------------------------------------------------
    p = torch.rand(3, 16, 16, generator=g)
    t = (torch.rand(3, 16, 16, generator=g) < 0.2).float()
    cases = {"rand": (p, t), "empty": (torch.zeros(3, 16, 16), torch.zeros(3, 16, 16)),
             "full": (torch.ones(3, 16, 16), torch.ones(3, 16, 16)), "out_of_range": (p * 3 - 1, t)}
    dice_gold = {}
    for name, (pp, tt) in cases.items():
        a = REF_DICE.dice_coeff(pp, tt, reduce_batch_first=True)
        b = O.dice_coeff(pp, tt, reduce_batch_first=True)
        assert torch.equal(a, b), name
        assert abs(float(a) - O.dice_coeff_numpy(pp.numpy(), tt.numpy())) < 1e-6, name
        assert torch.equal(REF_DICE.dice_loss(pp, tt), O.dice_loss(pp, tt)), name
        a4 = REF_DICE.dice_coeff(pp[:, None], tt[:, None], reduce_batch_first=False)
        assert torch.equal(a4, O.dice_coeff(pp[:, None], tt[:, None], reduce_batch_first=False)), name
        assert torch.equal(REF_DICE.multiclass_dice_coeff(pp[:, None], tt[:, None]), O.multiclass_dice_coeff(pp[:, None], tt[:, None]))
        dice_gold[f"{name}_p"] = pp.numpy()
        dice_gold[f"{name}_t"] = tt.numpy()
        dice_gold[f"{name}_coeff_batch"] = a.numpy()
        dice_gold[f"{name}_coeff_per_item"] = a4.numpy()
        dice_gold[f"{name}_loss"] = REF_DICE.dice_loss(pp, tt).numpy()
    np.savez_compressed(os.path.join(GOLDEN, "dice_cases.npz"), **dice_gold)
    report.append("dice_coeff / multiclass_dice_coeff / dice_loss incl. empty-mask branch: bit-exact (+numpy within 1e-6)")

    # ---- 5. max-pool indices ----------------------------------------------------------------------
    xm = torch.randn(2, 8, 8, 12, generator=g)
    xm = torch.relu(xm)                       # many all-zero windows -> ties
    xm[0, 0, 0, 0] = float("nan")
    xm[0, 1, 2:4, 2:4] = float("nan")         # several NaNs in one window: the last one wins
    xm[1, 2, 4, 5] = float("inf")
    vals, idx = torch.nn.functional.max_pool2d(xm, 2, return_indices=True)
    ov, oi = O.maxpool2x2_with_indices_numpy(xm.numpy())
    assert np.array_equal(idx.numpy(), oi), "maxpool indices differ"
    assert np.array_equal(np.nan_to_num(vals.numpy(), nan=-7.0), np.nan_to_num(ov, nan=-7.0))
    vcl, icl = torch.nn.functional.max_pool2d(xm.contiguous(memory_format=torch.channels_last), 2, return_indices=True)
    assert torch.equal(icl, idx)
    np.savez_compressed(os.path.join(GOLDEN, "maxpool_indices.npz"), x=xm.numpy(), values=vals.numpy(), indices=idx.numpy())
    report.append("max_pool2d indices with ties / NaN / inf: bit-exact (NCHW and channels_last)")

    # ---- 6. the variants ---------------------------------------------------------------------------
    import importlib

    variants = {"AttentionUNet": ("AttentionUNet", "AttentionUNet"), "R2UNet": ("R2UNet", "R2UNet"),
                "R2AttentionUNet": ("R2AttentionUNet", "R2AttentionUNet"), "ResUNet": ("ResUNet", "ResUNet"),
                "NestedUNet": ("UNetPP", "NestedUNet")}
    for name, (modname, clsname) in variants.items():
        cls = getattr(importlib.import_module(f"UNetFamily.{modname}"), clsname)
        torch.manual_seed(42)
        model = cls().train()
        sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
        images, labels = _inputs(21, 2, 32, 32)
        fwd = O.FORWARDS[name]
        with torch.no_grad():
            y_ref = model(images)
        sd = {k: v.clone() for k, v in sd0.items()}
        with torch.no_grad():
            y_or = fwd(images, sd, training=True)
        assert torch.equal(y_ref, y_or), f"{name}: train-mode forward differs"
        for k, v in model.state_dict().items():
            assert torch.equal(v, sd[k]), f"{name}: running stat {k} differs"
        model.eval()
        with torch.no_grad():
            y_ref_eval = model(images)
            y_or_eval = fwd(images, sd, training=False)
        assert torch.equal(y_ref_eval, y_or_eval), f"{name}: eval-mode forward differs"
        torch.manual_seed(42)
        model_bf = cls().train()
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            y_ref_bf16 = model_bf(images).float()
            y_or_bf16 = fwd(images, {k: v.clone() for k, v in sd0.items()}, training=True).float()
        assert torch.equal(y_ref_bf16, y_or_bf16), f"{name}: bf16-autocast forward differs"
        gold = dict(images=images.numpy(), labels=labels.numpy(), logits_train=y_ref.numpy(),
                    logits_eval_after_1_train_fwd=y_ref_eval.numpy(), logits_train_bf16_autocast=y_ref_bf16.numpy(),
                    state_dict_keys=np.array(list(sd0.keys())),
                    init_abs_sum=np.array([float(v.double().abs().sum()) for v in sd0.values()], dtype=np.float64))
        # one restated train step, fp32 (train.py:255-301 is model-agnostic)
        torch.manual_seed(42)
        model = cls().train()
        lr = 1e-3
        opt = torch.optim.RMSprop(model.parameters(), lr=lr, weight_decay=1e-8, momentum=0.999)
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        opt_state = {k: (torch.zeros_like(sd[k]), torch.zeros_like(sd[k])) for k in O.param_names(sd)}
        images, labels = _inputs(200, 2, 32, 32)
        loss_r, logits_r, bce_r, dice_r, grads_r = _ref_train_step(model, opt, images, labels, False)
        loss_o, logits_o, grads_o = O.train_step(sd, opt_state, images, labels, lr, False, model=name)
        assert torch.equal(loss_r, loss_o), f"{name}: loss differs {loss_r} vs {loss_o}"
        assert torch.equal(logits_r, logits_o)
        for k in grads_r:
            assert torch.equal(grads_r[k], grads_o[k]), f"{name}: clipped grad {k} differs"
        for k, v in model.state_dict().items():
            assert torch.equal(v, sd[k]), f"{name}: post-step tensor {k} differs"
        gold.update(step_images=images.numpy(), step_labels=labels.numpy(), step_loss=loss_r.numpy(),
                    step_bce=bce_r.numpy(), step_dice_loss=dice_r.numpy(), step_logits=logits_r.numpy(),
                    step_param_names=np.array(list(grads_r.keys())),
                    step_gradnorm=np.array([float(g.norm()) for g in grads_r.values()], dtype=np.float64))
        np.savez_compressed(os.path.join(GOLDEN, f"{name.lower()}_seed42.npz"), **gold)
        report.append(f"{name}: forward train/eval, running stats, bf16-autocast forward, fp32 train step: bit-exact")

    # blocks of the variants (small channel counts), forward bit-exact
    g = torch.Generator().manual_seed(9)
    vb = {}
    x16 = torch.randn(2, 16, 12, 20, generator=g)
    torch.manual_seed(11)
    m = ref_parts.conv_block(16, 24).train()
    with torch.no_grad():
        y = m(x16)
        assert torch.equal(y, O.conv_block(x16, {k: v.detach().clone() for k, v in m.state_dict().items()}, "", True))
    vb.update(x16=x16.numpy(), conv_block_y=y.numpy())
    torch.manual_seed(12)
    m = ref_parts.up_conv(16, 8).train()
    with torch.no_grad():
        y = m(x16)
        assert torch.equal(y, O.up_conv(x16, {k: v.detach().clone() for k, v in m.state_dict().items()}, "", True))
    vb.update(up_conv_y=y.numpy())
    torch.manual_seed(13)
    m = ref_parts.Recurrent_block(16, t=2).train()
    with torch.no_grad():
        y = m(x16)
        assert torch.equal(y, O.recurrent_block(x16, {k: v.detach().clone() for k, v in m.state_dict().items()}, "", True, 2))
    vb.update(recurrent_y=y.numpy())
    torch.manual_seed(14)
    m = ref_parts.RRCNN_block(16, 24, t=2).train()
    with torch.no_grad():
        y = m(x16)
        assert torch.equal(y, O.rrcnn_block(x16, {k: v.detach().clone() for k, v in m.state_dict().items()}, "", True, 2))
    vb.update(rrcnn_y=y.numpy())
    torch.manual_seed(15)
    m = ref_parts.Attention_block(16, 16, 8).train()
    gsig = torch.randn(2, 16, 12, 20, generator=g)
    with torch.no_grad():
        y = m(gsig, x16)
        assert torch.equal(y, O.attention_block(gsig, x16, {k: v.detach().clone() for k, v in m.state_dict().items()}, "", True))
    vb.update(att_g=gsig.numpy(), attention_y=y.numpy())
    for stride, seed in ((1, 16), (2, 17)):
        torch.manual_seed(seed)
        m = ref_parts.ResidualConv(16, 24, stride, 1).train()
        with torch.no_grad():
            y = m(x16)
            assert torch.equal(y, O.residual_conv(x16, {k: v.detach().clone() for k, v in m.state_dict().items()}, "", True, stride))
        vb[f"residual_conv_s{stride}_y"] = y.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "variant_blocks_seeds11to17.npz"), **vb)
    report.append("conv_block / up_conv / Recurrent_block / RRCNN_block / Attention_block / ResidualConv(s1,s2): bit-exact")

    # ---- 6b. UNet++ deep supervision: the reference class with its hard-coded attribute forced on ------------------
    from UNetFamily import UNetPP as ref_pp

    class _DS(ref_pp.NestedUNet):
        """`self.deepsupervision = False` (UNetPP.py:38) becomes True; everything else is the reference's code."""

        def __setattr__(self, k, v):
            super().__setattr__(k, True if k == "deepsupervision" else v)

    torch.manual_seed(42)
    mds = _DS()
    assert mds.deepsupervision and hasattr(mds, "final4") and not hasattr(mds, "final")
    sd0 = {k: v.detach().clone() for k, v in mds.state_dict().items()}
    images, labels = _inputs(7, 2, 32, 32)
    mds.train()
    with torch.no_grad():
        y_ref = mds(images)
        y_or = O.FORWARDS["NestedUNetDS"](images, {k: v.clone() for k, v in sd0.items()}, True)
    assert len(y_ref) == 4 and all(torch.equal(a, b) for a, b in zip(y_ref, y_or)), "deep-supervision forward differs"
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        torch.manual_seed(42)
        m2 = _DS().train()
        y_bf = [o.float() for o in m2(images)]
    mds.eval()
    with torch.no_grad():
        y_eval = mds(images)
    # the documented loss (mean over the heads of train.py:264-278) and its gradients through the reference modules
    torch.manual_seed(42)
    m3 = _DS().train()
    outs = m3(images)
    crit = torch.nn.BCEWithLogitsLoss()
    loss = sum(0.5 * crit(o, labels) + 0.5 * REF_DICE.dice_loss(torch.sigmoid(o).squeeze(1), labels.squeeze(1), multiclass=False)
               for o in outs) / 4
    loss.backward()
    s3 = {k: v.clone() for k, v in sd0.items()}
    for k in O.param_names(s3):
        s3[k].requires_grad_(True)
    _, lo, _, _ = O.forward_loss(s3, images, labels, bf16=False, training=True, model="NestedUNetDS")
    lo.backward()
    assert torch.equal(loss.detach(), lo.detach())
    for k, p_ in m3.named_parameters():
        assert torch.equal(p_.grad, s3[k].grad), k
    np.savez_compressed(os.path.join(GOLDEN, "nestedunet_ds_seed42.npz"), images=images.numpy(), labels=labels.numpy(),
                        **{f"out{k + 1}_train": y_ref[k].numpy() for k in range(4)},
                        **{f"out{k + 1}_train_bf16_autocast": y_bf[k].numpy() for k in range(4)},
                        **{f"out{k + 1}_eval": y_eval[k].numpy() for k in range(4)}, loss=np.float64(loss.item()),
                        grad_final1_weight=m3.final1.weight.grad.numpy(), grad_conv0_0_w=m3.conv0_0.conv[0].weight.grad.numpy(),
                        keys=np.array(list(sd0.keys())), final1_weight=sd0["final1.weight"].numpy())
    report.append("NestedUNet with deepsupervision forced on (UNetPP.py:65-69,93-102): 4 outputs train/eval, mean-of-heads loss and "
                  "all gradients bit-exact")

    # ---- 7. training-batch assembly and sliding-window inference: the reference's SOURCE TEXT, extracted ----------
    report += _pin_sampler_and_tiling(O)

    print("ORACLE PINNED against /root/reference:")
    for r in report:
        print("  -", r)
    print("golden vectors written to", GOLDEN)


if __name__ == "__main__":
    main()
