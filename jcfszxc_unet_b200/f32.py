"""fp32 mode of the vanilla U-Net forward (BASELINE.json configs[0]: "vanilla UNet(n_channels=3, n_classes=1) fp32
forward", north_star: "fp32 mode within 1e-4").

The reference without autocast runs UNet.py:39-55 in fp32.  This plan reproduces that accuracy on the bf16 tensor
cores: every activation is carried as three bf16 terms and every conv evaluates six exact bf16 x bf16 partial
products with fp32 accumulation (csrc/f32path.cu explains the layout); BatchNorm, ReLU, max-pool and the 1x1 head
run in fp32.  Forward only: train-mode (batch statistics, running stats updated) or eval-mode BatchNorm.

    with jcfszxc_unet_b200.precision("fp32"):
        logits = model(x)          # x fp32 [N,3,H,W] on cuda, H and W divisible by 16
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, engine, ops

BF16 = torch.bfloat16


def _s():
    return torch.cuda.current_stream().cuda_stream


class _SplitPack:
    """bf16 [T][R][6K] split pack of one fp32 master weight; rebuilt when the master changes."""

    def __init__(self, weight: torch.Tensor, transposed: bool, slices: list[int]):
        self.w, self.transposed, self.slices = weight, transposed, slices
        if transposed:   # ConvTranspose2d [Cin, Cout, 2, 2]: rows = Cout, K = Cin
            self.K, self.R, self.T = weight.shape[0], weight.shape[1], 4
            self.sr, self.sk, self.st = 4, weight.shape[1] * 4, 1
        else:            # Conv2d [Cout, Cin, 3, 3]
            self.R, self.K, self.T = weight.shape[0], weight.shape[1], 9
            self.sr, self.sk, self.st = weight.shape[1] * 9, 9, 1
        assert sum(slices) == self.K
        self.pack = torch.empty((self.T, self.R, 6 * self.K), dtype=BF16, device=weight.device)
        self._stamp = None

    def refresh(self):
        from .engine import weight_stamp

        if weight_stamp(self.w) == self._stamp:
            return
        w = self.w.detach()
        assert w.is_contiguous() and w.dtype == torch.float32
        arr = (ctypes.c_int * len(self.slices))(*self.slices)
        _lib.call("unetk_f32_pack_split3", w.data_ptr(), self.pack.data_ptr(), self.sr, self.sk, self.st, self.R, self.K,
                  self.T, arr, len(self.slices), _s())
        self._stamp = weight_stamp(self.w)


def split_activation(x: torch.Tensor) -> torch.Tensor:
    """fp32 NHWC [N,H,W,C] -> split tensor bf16 [N,H,W,6C] = [hi|hi|hi|mid|mid|lo] (kernel-level helper for tests)."""
    n, h, w, c = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    out = torch.empty((n, h, w, 6 * c), dtype=BF16, device=x.device)
    _lib.call("unetk_f32_bn_split", x.data_ptr(), c, None, None, out.data_ptr(), 6 * c, None, c, None, 0, n, h, w, c, 0, _s())
    return out


def conv_f32(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None, transposed: bool = False) -> torch.Tensor:
    """fp32-accurate conv3x3(p=1) / ConvTranspose2d(2,2) of an fp32 NHWC tensor on the bf16 tensor cores."""
    n, h, w, c = x.shape
    sp = _SplitPack(weight, transposed, [c])
    sp.refresh()
    xs = split_activation(x)
    cout = sp.R
    if transposed:
        y = torch.empty((n, 2 * h, 2 * w, cout), dtype=torch.float32, device=x.device)
        name = "unetk_f32_convT2x2"
    else:
        y = torch.empty((n, h, w, cout), dtype=torch.float32, device=x.device)
        name = "unetk_f32_conv3x3"
    _lib.call(name, xs.data_ptr(), 6 * c, sp.pack.data_ptr(), ops._f32(bias), y.data_ptr(), cout, n, h, w, 6 * c, cout, _s())
    return y


class UNetF32Plan:
    """Static buffers + op list of the fp32-mode forward of UNetFamily.UNet.UNet for one input shape."""

    def __init__(self, model, N: int, H: int, W: int, device, training: bool):
        if H % 16 or W % 16 or H < 16 or W < 16:
            raise ValueError(f"UNet fp32 plan needs H, W divisible by 16 (got {H}x{W})")
        if model.outc.conv.out_channels != 1:
            raise NotImplementedError("fp32 mode supports n_classes == 1")
        self.model, self.N, self.H, self.W, self.device, self.training = model, N, H, W, device, training
        self.steps: list = []
        self.packs: list[_SplitPack] = []
        lib = _lib.load()
        f32 = dict(dtype=torch.float32, device=device)
        dcs = [model.inc.double_conv] + [getattr(model, f"down{i}").maxpool_conv[1].double_conv for i in range(1, 5)]
        C = [dc[3].out_channels for dc in dcs]
        cmax = max(C)
        self.sums = torch.zeros(2 * cmax, dtype=torch.float64, device=device)
        need = max(lib.unetk_f32_stats_partial_doubles(N * H * W, c) for c in sorted(set(C)))
        self.partial = torch.empty(max(need, 1), dtype=torch.float64, device=device)
        self.image: torch.Tensor | None = None

        def split(h, w, c):
            return torch.empty((N, h, w, 6 * c), dtype=BF16, device=device)

        def raw(h, w, c):
            return torch.empty((N, h, w, c), **f32)

        # cat[i] = split planes of [skip_i | up_i]: 6*C_i channels each
        cats = [split(H >> i, W >> i, 2 * C[i]) for i in range(4)]
        x = None  # the image
        x_slices = None
        for i, dc in enumerate(dcs):
            h, w = H >> i, W >> i
            mid = split(h, w, dc[0].out_channels)
            self._conv_bn(x, x_slices, dc[0], dc[1], h, w, raw(h, w, dc[0].out_channels), mid, None, None)
            if i < 4:
                out, pooled = cats[i][..., : 6 * C[i]], split(h >> 1, w >> 1, C[i])
            else:
                out, pooled = split(h, w, C[i]), None
            self._conv_bn(mid, [dc[0].out_channels], dc[3], dc[4], h, w, raw(h, w, C[i]), out, pooled, None)
            x, x_slices = (pooled if pooled is not None else out), [C[i]]
        y, y_c = x, C[4]
        self.feat = None
        for j, i in enumerate((3, 2, 1, 0)):
            up = getattr(model, f"up{j + 1}")
            h, w = H >> i, W >> i
            self._convT(y, y_c, up.up, h, w, raw(h, w, C[i]), cats[i][..., 6 * C[i]:])
            dc = up.conv.double_conv
            mid = split(h, w, dc[0].out_channels)
            self._conv_bn(cats[i], [C[i], C[i]], dc[0], dc[1], h, w, raw(h, w, dc[0].out_channels), mid, None, None)
            last = (i == 0)
            yo = None if last else split(h, w, dc[3].out_channels)
            yf = raw(h, w, dc[3].out_channels) if last else None
            self._conv_bn(mid, [dc[0].out_channels], dc[3], dc[4], h, w, raw(h, w, dc[3].out_channels), yo, None, yf)
            y, y_c = yo, dc[3].out_channels
            if last:
                self.feat = yf
        self.logits = torch.empty((N, 1, H, W), **f32)
        conv = model.outc.conv
        feat = self.feat

        def head():
            w_ = conv.weight.detach().view(-1)
            b_ = conv.bias.detach() if conv.bias is not None else None
            _lib.call("unetk_f32_head", feat.data_ptr(), feat.shape[3], ops._f32(w_), ops._f32(b_), self.logits.data_ptr(),
                      N * H * W, feat.shape[3], _s())

        self.steps.append(head)

    # ---- builders -------------------------------------------------------------------------------------------------
    def _conv_bn(self, x, x_slices, conv, bn, h, w, raw, out_split, pooled, out_f32):
        N, dev = self.N, self.device
        cout = conv.out_channels
        stat = torch.zeros((4, cout), dtype=torch.float32, device=dev)
        npix = N * h * w
        if x is None:
            def conv_step():
                img = self.image
                bias = conv.bias.detach() if conv.bias is not None else None
                _lib.call("unetk_f32_stem_conv3x3", img.data_ptr(), img.stride(0), img.stride(1), img.stride(2),
                          img.stride(3), ops._f32(conv.weight.detach()), ops._f32(bias), raw.data_ptr(), cout, N, h, w,
                          conv.in_channels, cout, _s())
        else:
            sp = _SplitPack(conv.weight, False, x_slices)
            self.packs.append(sp)
            xp, xld = ops.nhwc(x)

            def conv_step():
                bias = conv.bias.detach() if conv.bias is not None else None
                _lib.call("unetk_f32_conv3x3", xp, xld, sp.pack.data_ptr(), ops._f32(bias), raw.data_ptr(), cout, N, h, w,
                          6 * conv.in_channels, cout, _s())

        def bn_step():
            sc, sh, mu, iv = stat[0], stat[1], stat[2], stat[3]
            gamma = bn.weight.detach() if bn.weight is not None else None
            beta = bn.bias.detach() if bn.bias is not None else None
            if self.training or not bn.track_running_stats:
                _lib.call("unetk_f32_stats", raw.data_ptr(), cout, npix, cout, self.partial.data_ptr(), self.sums.data_ptr(), _s())
                track = bn.track_running_stats and self.training
                ops.bn_finalize(self.sums, npix, gamma, beta, bn.eps, engine.bn_momentum(bn, self.training),
                                bn.running_mean if track else None, bn.running_var if track else None,
                                bn.num_batches_tracked if track else None, sc, sh, mu, iv)
            else:
                ops.bn_eval_fold(gamma, beta, bn.eps, bn.running_mean, bn.running_var, sc, sh, mu, iv)
            op, old = ops.nhwc(out_split) if out_split is not None else (None, 0)
            pp, pld = ops.nhwc(pooled) if pooled is not None else (None, 0)
            _lib.call("unetk_f32_bn_split", raw.data_ptr(), cout, ops._f32(sc), ops._f32(sh), op, old,
                      out_f32.data_ptr() if out_f32 is not None else None, cout, pp, pld, N, h, w, cout, 1, _s())

        self.steps += [conv_step, bn_step]

    def _convT(self, x, cin, mod, h, w, raw, out_split):
        """ConvTranspose2d(k=2,s=2) of the split tensor x [N,h/2,w/2,6*cin] -> fp32 raw [N,h,w,cout] -> split slice."""
        N = self.N
        cout = mod.out_channels
        sp = _SplitPack(mod.weight, True, [cin])
        self.packs.append(sp)
        xp, xld = ops.nhwc(x)
        op, old = ops.nhwc(out_split)

        def step():
            bias = mod.bias.detach() if mod.bias is not None else None
            _lib.call("unetk_f32_convT2x2", xp, xld, sp.pack.data_ptr(), ops._f32(bias), raw.data_ptr(), cout, N, h // 2,
                      w // 2, 6 * cin, cout, _s())
            _lib.call("unetk_f32_bn_split", raw.data_ptr(), cout, None, None, op, old, None, cout, None, 0, N, h, w, cout,
                      0, _s())

        self.steps.append(step)

    # ---- execution ------------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dtype != torch.float32:
            x = x.float()
        self.image = x
        for sp in self.packs:
            sp.refresh()
        for step in self.steps:
            step()
        return self.logits


def run_unet_f32(model, x: torch.Tensor) -> torch.Tensor:
    """model(x) in fp32 mode: cached plan per (shape, BatchNorm mode).  Forward only: the result carries no graph."""
    n, _, h, w = x.shape
    key = ("f32", n, h, w, x.device.index, model.training)
    from .bridge import plan_cache

    plans = plan_cache(model)
    plan = plans.lookup(key)
    if plan is None:
        plan = plans.insert(key, UNetF32Plan(model, n, h, w, x.device, model.training))
    with torch.no_grad():
        return plan.forward(x).clone()
