"""Tensor-level wrappers over the C ABI: torch tensors in, raw pointers + current stream out.

All activations are NHWC bf16 tensors of shape [N, H, W, C] whose channel stride is 1; the pixel stride
(`ld`) may exceed C, i.e. a tensor may be a channel slice of a wider (concat) buffer.
PyTorch here is plumbing only (device memory + streams); every FLOP runs in libunetk.so.
"""
from __future__ import annotations

import torch

from . import _lib

_ws_cache: dict[tuple[int, int], torch.Tensor] = {}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def nhwc(t: torch.Tensor) -> tuple[int, int]:
    """(data_ptr, ld) of an NHWC view; validates the layout the kernels assume."""
    if t.dim() != 4 or t.dtype != torch.bfloat16 or not t.is_cuda:
        raise ValueError(f"expected a 4-D CUDA bf16 NHWC tensor, got {tuple(t.shape)} {t.dtype} {t.device}")
    n, h, w, c = t.shape
    sn, sh, sw, sc = t.stride()
    ld = sw
    if sc != 1 or sh != w * ld or (n > 1 and sn != h * w * ld) or ld < c:
        raise ValueError(f"not an NHWC (channel-sliced) view: shape {tuple(t.shape)} stride {t.stride()}")
    return t.data_ptr(), ld


def _f32(t: torch.Tensor | None) -> int | None:
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        raise ValueError("expected a contiguous CUDA fp32 tensor")
    return t.data_ptr()


def workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    """Grow-only scratch buffer per (device, stream); owned by the host side, never by the library."""
    key = (device.index or 0, _stream())
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def pack_weight(w: torch.Tensor, want_ab: bool = True, want_ba: bool = True):
    """fp32 [A,B,kh,kw] -> (bf16 [T,A,B] | None, bf16 [T,B,A] | None)."""
    a, b = w.shape[0], w.shape[1]
    t = w.numel() // (a * b)
    ab = torch.empty((t, a, b), dtype=torch.bfloat16, device=w.device) if want_ab else None
    ba = torch.empty((t, b, a), dtype=torch.bfloat16, device=w.device) if want_ba else None
    _lib.call("unetk_pack_weight", _f32(w.detach()), ab.data_ptr() if ab is not None else None,
              ba.data_ptr() if ba is not None else None, a, b, t, _stream())
    return ab, ba


def conv_fwd(x, w_pack, bias, y, ksize: int = 3):
    """y <- conv(x): x [N,H,W,Cin], w_pack bf16 [k*k,Cout,Cin], y [N,H,W,Cout] (raw conv output)."""
    n, h, w, cin = x.shape
    cout = y.shape[3]
    assert w_pack.shape == (ksize * ksize, cout, cin) and y.shape[:3] == x.shape[:3]
    xp, xld = nhwc(x)
    yp, yld = nhwc(y)
    name = "unetk_conv3x3_fwd" if ksize == 3 else "unetk_conv1x1_fwd"
    _lib.call(name, xp, xld, w_pack.data_ptr(), _f32(bias), yp, yld, n, h, w, cin, cout, _stream())
    return y


def conv_dgrad_cols(dy, w_pack_t, col0: int, dx, accumulate: bool = False):
    """dx (+)<- the input-channel slice [col0, col0 + dx.C) of conv3x3^T(dy); w_pack_t is the whole pack [9,Cin,Cout]."""
    n, h, w, cout = dy.shape
    taps, cin_total, cout_w = w_pack_t.shape
    assert taps == 9 and cout_w == cout and dx.shape[:3] == dy.shape[:3] and col0 + dx.shape[3] <= cin_total
    dyp, dyld = nhwc(dy)
    dxp, dxld = nhwc(dx)
    _lib.call("unetk_conv3x3_dgrad_cols", dyp, dyld, w_pack_t.data_ptr(), cin_total, col0, dxp, dxld, int(accumulate),
              n, h, w, dx.shape[3], cout, _stream())
    return dx


def conv_dgrad(dy, w_pack_t, dx, ksize: int = 3, accumulate: bool = False, stride: int = 1):
    """dx (+)<- conv^T(dy): dy [N,H,W,Cout], w_pack_t bf16 [k*k,Cin,Cout], dx [N,stride*H,stride*W,Cin]."""
    n, h, w, cout = dy.shape
    cin = dx.shape[3]
    assert w_pack_t.shape == (ksize * ksize, cin, cout)
    assert dx.shape[1] == stride * h and dx.shape[2] == stride * w
    dyp, dyld = nhwc(dy)
    dxp, dxld = nhwc(dx)
    if stride == 2:
        assert ksize == 3
        name = "unetk_conv3x3s2_dgrad"
    else:
        name = "unetk_conv3x3_dgrad" if ksize == 3 else "unetk_conv1x1_dgrad"
    _lib.call(name, dyp, dyld, w_pack_t.data_ptr(), dxp, dxld, int(accumulate), n, h, w, cin, cout, _stream())
    return dx


def conv_dgrad_colsum(dy, w_pack_t, dx, partial, sums):
    """dx <- conv3x3^T(dy) and sums (fp64 [2,Cin]) <- per-channel (sum, sum sq) of the bf16 dx."""
    n, h, w, cout = dy.shape
    cin = dx.shape[3]
    assert w_pack_t.shape == (9, cin, cout) and dx.shape[:3] == dy.shape[:3]
    dyp, dyld = nhwc(dy)
    dxp, dxld = nhwc(dx)
    _lib.call("unetk_conv3x3_dgrad_colsum", dyp, dyld, w_pack_t.data_ptr(), dxp, dxld, partial.data_ptr(),
              sums.data_ptr(), n, h, w, cin, cout, _stream())
    return dx


def sums_to_f32(sums, c0: int, c: int, out, accumulate=False):
    """out[0:c] (+)= float(sums[c0:c0+c]) (sums: fp64 device vector)."""
    assert sums.dtype == torch.float64 and out.numel() == c
    _lib.call("unetk_sums_to_f32", sums.data_ptr() + 8 * c0, c, _f32(out), int(accumulate), _stream())


def conv_fwd_stats(x, w_pack, bias, y, partial, sums, ksize: int = 3, stride: int = 1):
    """y <- conv(x) and sums (fp64 [2,Cout]) <- per-channel (sum, sum sq) of y; stride 2 only for 3x3.
    partial/sums None (stride 2 only) skips the statistics."""
    n, h, w, cout = y.shape
    cin = x.shape[3]
    assert w_pack.shape == (ksize * ksize, cout, cin)
    assert x.shape[1] == stride * h and x.shape[2] == stride * w
    xp, xld = nhwc(x)
    yp, yld = nhwc(y)
    pp = partial.data_ptr() if partial is not None else None
    sp = sums.data_ptr() if sums is not None else None
    if stride == 2:
        assert ksize == 3
        name = "unetk_conv3x3s2_fwd"
    else:
        name = "unetk_conv3x3_fwd_bnstats" if ksize == 3 else "unetk_conv1x1_fwd_bnstats"
    _lib.call(name, xp, xld, w_pack.data_ptr(), _f32(bias), yp, yld, pp, sp, n, h, w, cin, cout, _stream())
    return y


def conv_wgrad(x, dy, dw, ksize: int = 3, accumulate: bool = False, stride: int = 1, ws=None):
    """dw (fp32 [Cout,Cin,k,k]) (+)<- sum_pixels dy (x) x; the reduction grid is dy's (stride 2: x is 2x larger)."""
    n, h, w, cout = dy.shape
    cin = x.shape[3]
    assert dw.shape == (cout, cin, ksize, ksize) and dw.dtype == torch.float32 and dw.is_contiguous()
    assert x.shape[1] == stride * h and x.shape[2] == stride * w
    if ws is None:
        need = _lib.load().unetk_conv_wgrad_workspace(n, h, w, cin, cout, ksize * ksize)
        ws = workspace(need, x.device)
    xp, xld = nhwc(x)
    dyp, dyld = nhwc(dy)
    if stride == 2:
        assert ksize == 3
        name = "unetk_conv3x3s2_wgrad"
    else:
        name = "unetk_conv3x3_wgrad" if ksize == 3 else "unetk_conv1x1_wgrad"
    _lib.call(name, xp, xld, dyp, dyld, dw.data_ptr(), int(accumulate), n, h, w, cin, cout, ws.data_ptr(),
              ws.numel(), _stream())
    return dw


def convT_fwd(x, w_pack, bias, y):
    """y [N,2H,2W,Cout] <- ConvTranspose2d(k=2,s=2)(x [N,H,W,Cin]); w_pack bf16 [4,Cout,Cin]."""
    n, h, w, cin = x.shape
    cout = y.shape[3]
    assert w_pack.shape == (4, cout, cin) and y.shape[1] == 2 * h and y.shape[2] == 2 * w
    xp, xld = nhwc(x)
    yp, yld = nhwc(y)
    _lib.call("unetk_convT2x2_fwd", xp, xld, w_pack.data_ptr(), _f32(bias), yp, yld, n, h, w, cin, cout, _stream())
    return y


def convT_dgrad(dy, w_pack_t, dx, accumulate: bool = False):
    """dx [N,H,W,Cin] (+)<- dy [N,2H,2W,Cout]; w_pack_t bf16 [4,Cin,Cout]."""
    n, h, w, cin = dx.shape
    cout = dy.shape[3]
    assert w_pack_t.shape == (4, cin, cout) and dy.shape[1] == 2 * h and dy.shape[2] == 2 * w
    dyp, dyld = nhwc(dy)
    dxp, dxld = nhwc(dx)
    _lib.call("unetk_convT2x2_dgrad", dyp, dyld, w_pack_t.data_ptr(), dxp, dxld, int(accumulate), n, h, w, cin, cout,
              _stream())
    return dx


def convT_wgrad(x, dy, dw, accumulate: bool = False):
    """dw (fp32 [Cin,Cout,2,2]) <- sum_pixels x (x) dy."""
    n, h, w, cin = x.shape
    cout = dy.shape[3]
    assert dw.shape == (cin, cout, 2, 2) and dw.dtype == torch.float32 and dw.is_contiguous()
    need = _lib.load().unetk_conv_wgrad_workspace(n, h, w, cin, cout, 4)
    ws = workspace(need, x.device)
    xp, xld = nhwc(x)
    dyp, dyld = nhwc(dy)
    _lib.call("unetk_convT2x2_wgrad", xp, xld, dyp, dyld, dw.data_ptr(), int(accumulate), n, h, w, cin, cout,
              ws.data_ptr(), ws.numel(), _stream())
    return dw


# ---- up_conv (nearest 2x + conv3x3) in sub-pixel form: x is the LOW-resolution tensor, y / dy are 2x larger
def pack_upconv_weight(w: torch.Tensor, want_fwd: bool = True, want_dgrad: bool = True):
    """fp32 [Cout,Cin,3,3] -> (bf16 [4 taps, 4 phases, Cout, Cin] | None, bf16 [16, Cin, Cout] | None)."""
    cout, cin = w.shape[0], w.shape[1]
    assert tuple(w.shape[2:]) == (3, 3)
    fwd = torch.empty((4, 4, cout, cin), dtype=torch.bfloat16, device=w.device) if want_fwd else None
    dg = torch.empty((16, cin, cout), dtype=torch.bfloat16, device=w.device) if want_dgrad else None
    _lib.call("unetk_pack_upconv_weight", _f32(w.detach()), fwd.data_ptr() if fwd is not None else None,
              dg.data_ptr() if dg is not None else None, cout, cin, _stream())
    return fwd, dg


def upconv_fwd(x, w_up, bias, y, partial=None, sums=None):
    """y [N,2H,2W,Cout] <- conv3x3(nearest2x(x [N,H,W,Cin])) (+bias); optional fused BatchNorm sums (fp64 [2,Cout])."""
    n, h, w, cin = x.shape
    cout = y.shape[3]
    assert w_up.numel() == 16 * cout * cin and y.shape[1] == 2 * h and y.shape[2] == 2 * w
    xp, xld = nhwc(x)
    yp, yld = nhwc(y)
    _lib.call("unetk_upconv3x3_fwd", xp, xld, w_up.data_ptr(), _f32(bias), yp, yld,
              partial.data_ptr() if partial is not None else None, sums.data_ptr() if sums is not None else None,
              n, h, w, cin, cout, _stream())
    return y


def upconv_fwd_affine(x, w_up, scale, shift, relu, y):
    n, h, w, cin = x.shape
    cout = y.shape[3]
    assert w_up.numel() == 16 * cout * cin and y.shape[1] == 2 * h and y.shape[2] == 2 * w
    xp, xld = nhwc(x)
    yp, yld = nhwc(y)
    _lib.call("unetk_upconv3x3_fwd_affine", xp, xld, w_up.data_ptr(), _f32(scale), _f32(shift), int(relu), yp, yld,
              n, h, w, cin, cout, _stream())
    return y


def upconv_dgrad(dy, w_up_t, dx, accumulate: bool = False):
    """dx [N,H,W,Cin] (+)<- dy [N,2H,2W,Cout]; w_up_t bf16 [16,Cin,Cout]."""
    n, h, w, cin = dx.shape
    cout = dy.shape[3]
    assert w_up_t.numel() == 16 * cout * cin and dy.shape[1] == 2 * h and dy.shape[2] == 2 * w
    dyp, dyld = nhwc(dy)
    dxp, dxld = nhwc(dx)
    _lib.call("unetk_upconv3x3_dgrad", dyp, dyld, w_up_t.data_ptr(), dxp, dxld, int(accumulate), n, h, w, cin, cout,
              _stream())
    return dx


def upconv_wgrad(x, dy, dw, accumulate: bool = False, ws=None):
    """dw (fp32 [Cout,Cin,3,3], the 3x3 master's gradient) (+)<- x [N,H,W,Cin], dy [N,2H,2W,Cout]."""
    n, h, w, cin = x.shape
    cout = dy.shape[3]
    assert dw.shape == (cout, cin, 3, 3) and dw.dtype == torch.float32 and dw.is_contiguous()
    assert dy.shape[1] == 2 * h and dy.shape[2] == 2 * w
    if ws is None:
        ws = workspace(_lib.load().unetk_upconv_wgrad_workspace(n, h, w, cin, cout), x.device)
    xp, xld = nhwc(x)
    dyp, dyld = nhwc(dy)
    _lib.call("unetk_upconv3x3_wgrad", xp, xld, dyp, dyld, dw.data_ptr(), int(accumulate), n, h, w, cin, cout,
              ws.data_ptr(), ws.numel(), _stream())
    return dw


# ------------------------------------------------------------------------------------------------
# stem (network-input conv), BatchNorm/ReLU/MaxPool, head + loss, optimizer
# ------------------------------------------------------------------------------------------------
def _img(x: torch.Tensor):
    """fp32 image [N,C,H,W] in any strides (NCHW or channels_last) -> (ptr, sn, sc, sh, sw)."""
    if x.dim() != 4 or x.dtype != torch.float32 or not x.is_cuda:
        raise ValueError(f"expected a CUDA fp32 [N,C,H,W] image, got {tuple(x.shape)} {x.dtype} {x.device}")
    sn, sc, sh, sw = x.stride()
    return x.data_ptr(), sn, sc, sh, sw


def stem_fwd(x, w, bias, y):
    n, cin, h, wd = x.shape
    cout = y.shape[3]
    xp, sn, sc, sh, sw = _img(x)
    yp, yld = nhwc(y)
    _lib.call("unetk_stem_conv3x3_fwd", xp, sn, sc, sh, sw, _f32(w.detach()), _f32(bias), yp, yld, n, h, wd, cin, cout,
              _stream())
    return y


def stem_wgrad(x, dy, dw, accumulate=False):
    n, cin, h, wd = x.shape
    cout = dy.shape[3]
    need = _lib.load().unetk_stem_wgrad_workspace(n, h, wd, cin)
    ws = workspace(need, x.device)
    xp, sn, sc, sh, sw = _img(x)
    dyp, dyld = nhwc(dy)
    _lib.call("unetk_stem_conv3x3_wgrad", xp, sn, sc, sh, sw, dyp, dyld, _f32(dw), int(accumulate), n, h, wd, cin, cout,
              ws.data_ptr(), ws.numel(), _stream())
    return dw


def chan_partial_floats(units: int, c: int) -> int:
    return _lib.load().unetk_chan_partial_floats(units, c)


def stem_fwd_stats(x, w, bias, y, partial, sums):
    """stem_fwd + bn_stats of its output in one launch (statistics from the conv epilogue)."""
    n, cin, h, wd = x.shape
    cout = y.shape[3]
    xp, sn, sc, sh, sw = _img(x)
    yp, yld = nhwc(y)
    _lib.call("unetk_stem_conv3x3_fwd_bnstats", xp, sn, sc, sh, sw, _f32(w.detach()), _f32(bias), yp, yld,
              partial.data_ptr(), sums.data_ptr(), n, h, wd, cin, cout, _stream())
    return y


def bn_stats(x, partial, sums):
    """sums (fp64 [2,C]) <- per-channel (sum, sum of squares) of x [N,H,W,C]."""
    n, h, w, c = x.shape
    xp, xld = nhwc(x)
    _lib.call("unetk_bn_stats", xp, xld, n * h * w, c, partial.data_ptr(), sums.data_ptr(), _stream())


def bn_finalize(sums, count, gamma, beta, eps, momentum, rm, rv, nbt, scale, shift, mean, invstd):
    c = scale.numel()
    _lib.call("unetk_bn_finalize", sums.data_ptr(), c, float(count), _f32(gamma), _f32(beta), eps, momentum,
              _f32(rm), _f32(rv), nbt.data_ptr() if nbt is not None else None, _f32(scale), _f32(shift), _f32(mean),
              _f32(invstd), _stream())


def bn_eval_fold(gamma, beta, eps, rm, rv, scale, shift, mean, invstd):
    _lib.call("unetk_bn_eval_fold", scale.numel(), _f32(gamma), _f32(beta), eps, _f32(rm), _f32(rv), _f32(scale),
              _f32(shift), _f32(mean), _f32(invstd), _stream())


def bn_eval_fold_bias(gamma, beta, eps, rm, rv, conv_bias, scale, shift):
    _lib.call("unetk_bn_eval_fold_bias", scale.numel(), _f32(gamma), _f32(beta), float(eps), _f32(rm), _f32(rv), _f32(conv_bias),
              _f32(scale), _f32(shift), _stream())


def conv_fwd_affine(x, w_pack, scale, shift, relu, y, stride: int = 1):
    """y = relu?(conv3x3(x) * scale + shift): eval-mode BatchNorm folded into the conv epilogue."""
    n, ho, wo, cout = y.shape
    cin = x.shape[3]
    assert x.shape[1] == stride * ho and x.shape[2] == stride * wo
    xp, xld = nhwc(x)
    yp, yld = nhwc(y)
    _lib.call("unetk_conv3x3_fwd_affine", xp, xld, w_pack.data_ptr(), _f32(scale), _f32(shift), int(relu), yp, yld, n, ho, wo,
              cin, cout, int(stride), _stream())


def stem_fwd_affine(x, w, scale, shift, relu, y):
    xp, sn, sc, sh, sw = _img(x)
    n, cin, h, wd = x.shape
    yp, yld = nhwc(y)
    _lib.call("unetk_stem_conv3x3_fwd_affine", xp, sn, sc, sh, sw, _f32(w), _f32(scale), _f32(shift), int(relu), yp, yld,
              n, h, wd, cin, y.shape[3], _stream())


def bn_apply(raw, scale, shift, out, pooled=None, relu=True, res=None):
    """out <- relu?(bn(raw)) [+ res]; pooled (optional) <- 2x2 max-pool of out."""
    n, h, w, c = raw.shape
    rp, rld = nhwc(raw)
    op, old = nhwc(out)
    pp, pld = nhwc(pooled) if pooled is not None else (None, 0)
    sp, sld = nhwc(res) if res is not None else (None, 0)
    _lib.call("unetk_bn_apply", rp, rld, _f32(scale), _f32(shift), sp, sld, op, old, pp, pld, n, h, w, c, int(relu),
              _stream())


def bn_apply_copies(raw, scale, shift, out, copies, pooled=None, relu=True):
    """bn_apply that also stores the activation into `copies` (1..3 more NHWC views of out's shape)."""
    n, h, w, c = raw.shape
    assert 1 <= len(copies) <= 3 and all(t.shape == out.shape for t in copies)
    rp, rld = nhwc(raw)
    op, old = nhwc(out)
    pp, pld = nhwc(pooled) if pooled is not None else (None, 0)
    args = []
    for i in range(3):
        args += list(nhwc(copies[i])) if i < len(copies) else [None, 0]
    _lib.call("unetk_bn_apply_copies", rp, rld, _f32(scale), _f32(shift), op, old, pp, pld, *args, n, h, w, c, int(relu),
              _stream())


def bn_bwd_reduce(raw, g1, gp, scale, shift, mean, invstd, partial, sums, relu=True):
    n, h, w, c = raw.shape
    rp, rld = nhwc(raw)
    g1p, g1ld = nhwc(g1) if g1 is not None else (None, 0)
    gpp, gpld = nhwc(gp) if gp is not None else (None, 0)
    _lib.call("unetk_bn_bwd_reduce", rp, rld, g1p, g1ld, gpp, gpld, _f32(scale), _f32(shift), _f32(mean), _f32(invstd),
              partial.data_ptr(), sums.data_ptr(), n, h, w, c, int(relu), _stream())


def bn_bwd_apply(raw, g1, gp, scale, shift, mean, invstd, sums, count, dgamma, dbeta, coef, draw, relu=True,
                 accumulate=False, draw_accumulate=False, dconv_bias=None, dres=None, dres_accumulate=False):
    """dres (optional, no fused pool): the gradient of the residual input of `out = relu(bn(raw)) + res` — g1 itself —
    written (or accumulated) by the same pass."""
    n, h, w, c = raw.shape
    rp, rld = nhwc(raw)
    if dres is not None:
        assert gp is None and g1 is not None and dres.shape == raw.shape
        g1p, g1ld = nhwc(g1)
        dp, dld = nhwc(draw)
        sp, sld = nhwc(dres)
        _lib.call("unetk_bn_bwd_apply_res", rp, rld, g1p, g1ld, _f32(scale), _f32(shift), _f32(mean), _f32(invstd),
                  sums.data_ptr(), float(count), _f32(dgamma), _f32(dbeta), int(accumulate), _f32(coef), _f32(dconv_bias),
                  dp, dld, int(draw_accumulate), sp, sld, int(dres_accumulate), n, h, w, c, int(relu), _stream())
        return
    g1p, g1ld = nhwc(g1) if g1 is not None else (None, 0)
    gpp, gpld = nhwc(gp) if gp is not None else (None, 0)
    dp, dld = nhwc(draw)
    _lib.call("unetk_bn_bwd_apply", rp, rld, g1p, g1ld, gpp, gpld, _f32(scale), _f32(shift), _f32(mean), _f32(invstd),
              sums.data_ptr(), float(count), _f32(dgamma), _f32(dbeta), int(accumulate), _f32(coef), _f32(dconv_bias),
              dp, dld, int(draw_accumulate), n, h, w, c, int(relu), _stream())


def bn_bwd_coef(sums, count, scale, mean, invstd, dgamma, dbeta, coef, accumulate=False, dconv_bias=None):
    c = scale.numel()
    _lib.call("unetk_bn_bwd_coef", sums.data_ptr(), c, float(count), _f32(scale), _f32(mean), _f32(invstd),
              _f32(dgamma), _f32(dbeta), int(accumulate), _f32(coef), _f32(dconv_bias), _stream())


def maxpool_fwd(x, y, idx=None):
    n, h, w, c = x.shape
    xp, xld = nhwc(x)
    yp, yld = nhwc(y)
    _lib.call("unetk_maxpool2x2_fwd", xp, xld, yp, yld, idx.data_ptr() if idx is not None else None, n, h, w, c,
              _stream())


def maxpool_bwd(x, dy, dx, accumulate=False):
    n, h, w, c = x.shape
    xp, xld = nhwc(x)
    dyp, dyld = nhwc(dy)
    dxp, dxld = nhwc(dx)
    _lib.call("unetk_maxpool2x2_bwd", xp, xld, dyp, dyld, dxp, dxld, int(accumulate), n, h, w, c, _stream())


def maxpool_fwd_codes(x, y, code):
    """MaxPool2d(2) whose arg-max leaves as byte codes: code uint8 NHWC [N, H/2, W/2, C] (window position 0..3)."""
    n, h, w, c = x.shape
    assert code.dtype == torch.uint8 and code.is_contiguous() and tuple(code.shape) == (n, h // 2, w // 2, c)
    xp, xld = nhwc(x)
    yp, yld = nhwc(y)
    _lib.call("unetk_maxpool2x2_fwd_codes", xp, xld, yp, yld, code.data_ptr(), n, h, w, c, _stream())


def _where(where, n, ho, wo, c):
    if where.dtype == torch.uint8:
        assert where.is_contiguous() and tuple(where.shape) == (n, ho, wo, c), "codes: uint8 NHWC [N, Ho, Wo, C]"
        return 0
    assert where.dtype == torch.int64 and where.is_contiguous() and tuple(where.shape) == (n, c, ho, wo), \
        "indices: int64 [N, C, Ho, Wo] as F.max_pool2d(return_indices=True) returns them"
    return 1


def max_unpool(x, where, out):
    """F.max_unpool2d(x, idx, 2, 2) on NHWC bf16 views (SegNet.py:115-138); `where`: byte codes or int64 indices."""
    n, ho, wo, c = x.shape
    assert tuple(out.shape) == (n, 2 * ho, 2 * wo, c)
    xp, xld = nhwc(x)
    op, old = nhwc(out)
    _lib.call("unetk_max_unpool2x2", xp, xld, where.data_ptr(), _where(where, n, ho, wo, c), op, old, n, ho, wo, c, _stream())


def max_unpool_bwd(dy, where, dx, accumulate=False):
    n, ho, wo, c = dx.shape
    assert tuple(dy.shape) == (n, 2 * ho, 2 * wo, c)
    gp, gld = nhwc(dy)
    dp, dld = nhwc(dx)
    _lib.call("unetk_max_unpool2x2_bwd", gp, gld, where.data_ptr(), _where(where, n, ho, wo, c), dp, dld, int(accumulate),
              n, ho, wo, c, _stream())


def shift_copy(dst, src, oy: int, ox: int):
    """dst[n,y,x,:] = src[n,y-oy,x-ox,:] inside src, else 0 (F.pad of Up, unet_parts.py:64-67; negative offsets crop)."""
    n, hd, wd, c = dst.shape
    _, hs, ws, _ = src.shape
    assert src.shape[0] == n and src.shape[3] == c
    dp, dld = nhwc(dst)
    sp, sld = nhwc(src)
    _lib.call("unetk_shift_copy", dp, dld, hd, wd, sp, sld, hs, ws, int(oy), int(ox), n, c, _stream())


def colsum(x, partial, out, accumulate=False):
    n, h, w, c = x.shape
    xp, xld = nhwc(x)
    _lib.call("unetk_colsum", xp, xld, n * h * w, c, partial.data_ptr(), _f32(out), int(accumulate), _stream())


def head_partial_floats(npix: int, c: int) -> int:
    return _lib.load().unetk_head_partial_floats(npix, c)


def head_fwd(x, w, bias, labels, logits, partial, sums, post_sigmoid=False):
    n, h, wd, c = x.shape
    xp, xld = nhwc(x)
    _lib.call("unetk_head_fwd", xp, xld, _f32(w), _f32(bias), _f32(labels), _f32(logits), int(post_sigmoid), n * h * wd, c,
              partial.data_ptr(), sums.data_ptr() if sums is not None else None, _stream())


def loss_finalize(sums, npix_total, fin):
    _lib.call("unetk_loss_finalize", sums.data_ptr(), float(npix_total), _f32(fin), _stream())


def head_bwd(x, w, labels, logits, fin, dlogits, gscale, dx, dw, db, partial, accumulate=False, post_sigmoid=False):
    n, h, wd, c = x.shape
    xp, xld = nhwc(x)
    dxp, dxld = nhwc(dx)
    _lib.call("unetk_head_bwd", xp, xld, _f32(w), _f32(labels), _f32(logits), _f32(fin), _f32(dlogits), float(gscale),
              int(post_sigmoid), dxp, dxld, _f32(dw), _f32(db), int(accumulate), n * h * wd, c, partial.data_ptr(), _stream())


def bn_head_partial_floats(npix: int, c: int) -> int:
    return _lib.load().unetk_bn_head_partial_floats(npix, c)


def bn_head_fwd(raw, scale, shift, relu, w, bias, labels, logits, partial, sums, post_sigmoid=False):
    """unetk_bn_apply + unetk_head_fwd in one pass over the conv output (the activation is never written)."""
    n, h, wd, c = raw.shape
    rp, rld = nhwc(raw)
    _lib.call("unetk_bn_head_fwd", rp, rld, _f32(scale), _f32(shift), int(relu), _f32(w), _f32(bias), _f32(labels),
              _f32(logits), int(post_sigmoid), n * h * wd, c, partial.data_ptr(),
              sums.data_ptr() if sums is not None else None, _stream())


def bn_head_bwd_reduce(raw, scale, shift, mean, relu, w, labels, logits, fin, dlogits, gscale, dz, dw, db, sums, partial,
                       accumulate=False, post_sigmoid=False):
    n, h, wd, c = raw.shape
    rp, rld = nhwc(raw)
    _lib.call("unetk_bn_head_bwd_reduce", rp, rld, _f32(scale), _f32(shift), _f32(mean), int(relu), _f32(w), _f32(labels),
              _f32(logits), _f32(fin), _f32(dlogits), float(gscale), int(post_sigmoid), _f32(dz), _f32(dw), _f32(db),
              int(accumulate), sums.data_ptr(), n * h * wd, c, partial.data_ptr(), _stream())


def bn_head_bwd_apply(raw, scale, shift, relu, w, dz, coef, draw):
    n, h, wd, c = raw.shape
    rp, rld = nhwc(raw)
    dp, dld = nhwc(draw)
    _lib.call("unetk_bn_head_bwd_apply", rp, rld, _f32(scale), _f32(shift), int(relu), _f32(w), _f32(dz), _f32(coef),
              dp, dld, n * h * wd, c, _stream())


def grad_clip_coef(g, gscale, max_norm, partial, out):
    _lib.call("unetk_grad_clip_coef", _f32(g), g.numel(), float(gscale), float(max_norm), partial.data_ptr(), _f32(out),
              _stream())


def rmsprop_step(p, g, sq, buf, lr, alpha, eps, wd, momentum, clip):
    _lib.call("unetk_rmsprop_step", _f32(p), _f32(g), _f32(sq), _f32(buf), p.numel(), float(lr), float(alpha),
              float(eps), float(wd), float(momentum), _f32(clip), _stream())


def rmsprop_step_dev(p, g, sq, buf, hyper, clip):
    """hyper: device fp32 [lr, alpha, eps, weight_decay, momentum] (read by the kernel at run time)."""
    _lib.call("unetk_rmsprop_step_dev", _f32(p), _f32(g), _f32(sq), _f32(buf), p.numel(), _f32(hyper), _f32(clip), _stream())


# ------------------------------------------------------------------------------------------------
# glue of the U-Net variants: residual adds / slice copies, 2x up-sampling, attention gate
# ------------------------------------------------------------------------------------------------
def add_n(dst, srcs, accumulate=False):
    """dst (+)<- sum(srcs) (1..4 bf16 NHWC views of the same shape; bf16 rounding after every add)."""
    n, h, w, c = dst.shape
    assert 1 <= len(srcs) <= 4 and all(t.shape == dst.shape for t in srcs)
    dp, dld = nhwc(dst)
    args = []
    for i in range(4):
        if i < len(srcs):
            args += list(nhwc(srcs[i]))
        else:
            args += [None, 0]
    _lib.call("unetk_add_n", dp, dld, int(accumulate), *args, n * h * w, c, _stream())


def upsample2x(x, y, mode="nearest"):
    """y [N,2H,2W,C] <- nn.Upsample(scale_factor=2, mode)(x); bilinear uses align_corners=True."""
    n, h, w, c = x.shape
    assert y.shape == (n, 2 * h, 2 * w, c)
    xp, xld = nhwc(x)
    yp, yld = nhwc(y)
    _lib.call(f"unetk_upsample_{mode}2x_fwd", xp, xld, yp, yld, n, h, w, c, _stream())


def upsample2x_bwd(dy, dx, mode="nearest", accumulate=False):
    n, h, w, c = dx.shape
    assert dy.shape == (n, 2 * h, 2 * w, c)
    dyp, dyld = nhwc(dy)
    dxp, dxld = nhwc(dx)
    _lib.call(f"unetk_upsample_{mode}2x_bwd", dyp, dyld, dxp, dxld, int(accumulate), n, h, w, c, _stream())


def copy_f32_strided(dst, dst_stride, src, src_stride, n, accumulate=False, dst_offset=0, src_offset=0):
    """dst.flat[dst_offset + i*dst_stride] (+)<- src.flat[src_offset + i*src_stride], i < n (fp32)."""
    assert dst.dtype == torch.float32 and src.dtype == torch.float32 and dst.is_cuda and src.is_cuda
    _lib.call("unetk_copy_f32_strided", dst.data_ptr() + 4 * dst_offset, dst_stride, src.data_ptr() + 4 * src_offset,
              src_stride, n, int(accumulate), _stream())
