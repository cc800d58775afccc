"""CPU ORACLE, second layer — the arithmetic itself, restated in numpy (float64).  TEST INFRASTRUCTURE ONLY.

oracle/unet_oracle.py restates the reference's GRAPH and calls torch.nn.functional for the arithmetic, because the
reference's arithmetic lives in an un-vendored third-party dependency: PyTorch (no pin file in the reference; this
image: torch 2.11.0+cu128).  This file restates that dependency's published semantics for every primitive on the hot
path, independently of torch, so that the checker does not rest on "the same library called twice":

    conv2d (stride 1/2, zero padding)      torch.nn.Conv2d           cross-correlation, NCHW, weight [Cout,Cin,kh,kw]
    conv_transpose2d k=2 s=2               torch.nn.ConvTranspose2d  weight [Cin,Cout,2,2], non-overlapping scatter
    batch_norm (train / eval)              torch.nn.BatchNorm2d      biased variance to normalise, unbiased for running_var
    relu, sigmoid, max_pool2d(2) + indices, nearest 2x / bilinear 2x (align_corners=True) up-sampling
    nearest 2x + conv3x3 as four 2x2-tap convs of the low-resolution input (the sub-pixel identity of the CUDA up-conv)
    bce_with_logits (mean)                 torch.nn.BCEWithLogitsLoss  max(z,0) - z*y + log1p(exp(-|z|))
    unet_forward                           UNetFamily/UNet.py:39-55 on UNetFamily/utils/unet_parts.py:17-79

tests/test_oracle_numpy.py pins it: against torch.nn.functional per primitive (float64, 1e-10) and against the golden
logits produced by the UNMODIFIED reference (tests/golden/unet_forward_seed42.npz, fp32: 1e-4 relative).
Plain numpy loops / einsum: sized for the small cases of the test-suite, not for speed.
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def conv2d(x, w, b=None, stride=1, padding=1):
    """y[n,co,i,j] = b[co] + sum_{ci,r,s} x[n,ci,stride*i+r-padding,stride*j+s-padding] * w[co,ci,r,s] (zero outside)."""
    n, cin, h, wd = x.shape
    cout, cin2, kh, kw = w.shape
    assert cin == cin2
    xp = np.zeros((n, cin, h + 2 * padding, wd + 2 * padding), dtype=np.float64)
    xp[:, :, padding:padding + h, padding:padding + wd] = x
    ho = (h + 2 * padding - kh) // stride + 1
    wo = (wd + 2 * padding - kw) // stride + 1
    y = np.zeros((n, cout, ho, wo), dtype=np.float64)
    for r in range(kh):
        for s in range(kw):
            patch = xp[:, :, r:r + stride * (ho - 1) + 1:stride, s:s + stride * (wo - 1) + 1:stride]
            y += np.einsum("nchw,oc->nohw", patch, w[:, :, r, s].astype(np.float64))
    if b is not None:
        y += np.asarray(b, dtype=np.float64).reshape(1, -1, 1, 1)
    return y


def conv_transpose2d_k2s2(x, w, b=None):
    """y[n,co,2i+a,2j+c] = b[co] + sum_ci x[n,ci,i,j] * w[ci,co,a,c]  (kernel 2, stride 2: every output has one source)."""
    n, cin, h, wd = x.shape
    cin2, cout, kh, kw = w.shape
    assert cin == cin2 and (kh, kw) == (2, 2)
    y = np.zeros((n, cout, 2 * h, 2 * wd), dtype=np.float64)
    for a in range(2):
        for c in range(2):
            y[:, :, a::2, c::2] = np.einsum("nchw,co->nohw", x.astype(np.float64), w[:, :, a, c].astype(np.float64))
    if b is not None:
        y += np.asarray(b, dtype=np.float64).reshape(1, -1, 1, 1)
    return y


def batch_norm(x, gamma, beta, running_mean, running_var, training, momentum=BN_MOMENTUM, eps=BN_EPS):
    """Returns (y, new_running_mean, new_running_var).  Training: normalise with the batch mean and BIASED variance,
    update the running statistics with the UNBIASED variance; eval: normalise with the running statistics."""
    x = x.astype(np.float64)
    if training:
        m = x.shape[0] * x.shape[2] * x.shape[3]
        mean = x.mean(axis=(0, 2, 3))
        var = x.var(axis=(0, 2, 3))                       # biased (divides by m)
        new_rm = (1 - momentum) * running_mean + momentum * mean
        new_rv = (1 - momentum) * running_var + momentum * var * m / max(m - 1, 1)
    else:
        mean, var, new_rm, new_rv = running_mean.astype(np.float64), running_var.astype(np.float64), running_mean, running_var
    y = (x - mean.reshape(1, -1, 1, 1)) / np.sqrt(var.reshape(1, -1, 1, 1) + eps)
    y = y * np.asarray(gamma, np.float64).reshape(1, -1, 1, 1) + np.asarray(beta, np.float64).reshape(1, -1, 1, 1)
    return y, new_rm, new_rv


def relu(x):
    return np.maximum(x, 0.0)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def max_pool2x2(x):
    """MaxPool2d(2): floor output size; returns (values, flat indices into H*W per (n,c) plane, first maximum wins)."""
    n, c, h, w = x.shape
    ho, wo = h // 2, w // 2
    win = np.stack([x[:, :, 0:2 * ho:2, 0:2 * wo:2], x[:, :, 0:2 * ho:2, 1:2 * wo:2],
                    x[:, :, 1:2 * ho:2, 0:2 * wo:2], x[:, :, 1:2 * ho:2, 1:2 * wo:2]], axis=-1)
    arg = win.argmax(axis=-1)                              # first maximum
    val = np.take_along_axis(win, arg[..., None], axis=-1)[..., 0]
    ii, jj = np.meshgrid(np.arange(ho), np.arange(wo), indexing="ij")
    idx = (2 * ii + arg // 2) * w + (2 * jj + arg % 2)
    return val, idx


def upsample_nearest2x(x):
    return x.repeat(2, axis=2).repeat(2, axis=3)


# ---- nearest 2x + conv3x3 in sub-pixel form: the identity the CUDA up-conv kernels compute by (csrc/pack.cu, capi.cu
# unetk_upconv3x3_*).  Restated here so that the identity itself — which 3x3 taps collapse into which 2x2 window tap of
# which output phase, and how the sixteen sub-filter gradients fold back — is pinned on the CPU against the reference's
# own formulation (unet_parts.py:103-104: nn.Upsample(scale_factor=2) then nn.Conv2d(k=3, p=1)).
def _subpixel_rows(q, u):
    """3x3 taps along one axis that read window tap u of output phase q: q=0: {0}, {1,2}; q=1: {0,1}, {2}."""
    return ((0,), (1, 2))[u] if q == 0 else ((0, 1), (2,))[u]


def subpixel_weights(w):
    """w [Cout,Cin,3,3] -> wq [qy,qx,u,v,Cout,Cin]: phase (qy,qx) of the 2x grid is a 2x2-tap conv of the LOW-resolution
    input whose window starts at (qy-1, qx-1)."""
    cout, cin = w.shape[:2]
    wq = np.zeros((2, 2, 2, 2, cout, cin), dtype=np.float64)
    for qy in range(2):
        for qx in range(2):
            for u in range(2):
                for v in range(2):
                    for kh in _subpixel_rows(qy, u):
                        for kw in _subpixel_rows(qx, v):
                            wq[qy, qx, u, v] += w[:, :, kh, kw]
    return wq


def upconv_subpixel(x, w, b=None):
    """conv2d(upsample_nearest2x(x), w, b) computed on the low-resolution x: y[n,:,2i+qy,2j+qx] = b + sum_{u,v}
    wq[qy,qx,u,v] @ x[n,:,i+qy-1+u,j+qx-1+v] (zero outside)."""
    n, cin, h, wd = x.shape
    wq = subpixel_weights(w)
    cout = w.shape[0]
    xp = np.zeros((n, cin, h + 2, wd + 2), dtype=np.float64)
    xp[:, :, 1:-1, 1:-1] = x
    y = np.zeros((n, cout, 2 * h, 2 * wd), dtype=np.float64)
    for qy in range(2):
        for qx in range(2):
            for u in range(2):
                for v in range(2):
                    patch = xp[:, :, qy + u:qy + u + h, qx + v:qx + v + wd]
                    y[:, :, qy::2, qx::2] += np.einsum("nchw,oc->nohw", patch, wq[qy, qx, u, v])
    if b is not None:
        y += np.asarray(b, dtype=np.float64).reshape(1, -1, 1, 1)
    return y


def fold_subpixel_wgrad(g):
    """g [qy,qx,u,v,Cout,Cin] (gradients of the sixteen sub-filters) -> dw [Cout,Cin,3,3]: tap (kh,kw) was summed into
    window tap (U(qy,kh), U(qx,kw)) of every phase, U(0,k) = k>0, U(1,k) = k>1."""
    dw = np.zeros(g.shape[4:] + (3, 3), dtype=np.float64)
    for kh in range(3):
        for kw in range(3):
            for qy in range(2):
                for qx in range(2):
                    u = int(kh > 0) if qy == 0 else int(kh > 1)
                    v = int(kw > 0) if qx == 0 else int(kw > 1)
                    dw[:, :, kh, kw] += g[qy, qx, u, v]
    return dw


def upsample_bilinear2x_align_corners(x):
    """nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True): source coordinate = dst * (in-1)/(out-1)."""
    n, c, h, w = x.shape

    def taps(size):
        out = 2 * size
        src = np.arange(out) * ((size - 1) / (out - 1)) if out > 1 else np.zeros(out)
        lo = np.minimum(np.floor(src).astype(int), size - 1)
        hi = np.minimum(lo + 1, size - 1)
        return lo, hi, src - lo

    y0, y1, fy = taps(h)
    x0, x1, fx = taps(w)
    x = x.astype(np.float64)
    top = x[:, :, y0][:, :, :, x0] * (1 - fx) + x[:, :, y0][:, :, :, x1] * fx
    bot = x[:, :, y1][:, :, :, x0] * (1 - fx) + x[:, :, y1][:, :, :, x1] * fx
    return top * (1 - fy).reshape(1, 1, -1, 1) + bot * fy.reshape(1, 1, -1, 1)


def bce_with_logits(z, y):
    z, y = z.astype(np.float64), y.astype(np.float64)
    return float(np.mean(np.maximum(z, 0) - z * y + np.log1p(np.exp(-np.abs(z)))))


# ---- the vanilla U-Net graph on the primitives above (UNetFamily/UNet.py:39-55) --------------------------------
def _bn(x, sd, prefix, training):
    y, rm, rv = batch_norm(x, sd[prefix + "weight"], sd[prefix + "bias"], sd[prefix + "running_mean"],
                           sd[prefix + "running_var"], training)
    if training:
        sd[prefix + "running_mean"], sd[prefix + "running_var"] = rm, rv
        sd[prefix + "num_batches_tracked"] = sd[prefix + "num_batches_tracked"] + 1
    return y


def double_conv(x, sd, prefix, training):
    """unet_parts.py:17-34"""
    x = relu(_bn(conv2d(x, sd[prefix + "double_conv.0.weight"]), sd, prefix + "double_conv.1.", training))
    return relu(_bn(conv2d(x, sd[prefix + "double_conv.3.weight"]), sd, prefix + "double_conv.4.", training))


def unet_forward(x, sd, training=True):
    """sd: dict of numpy arrays with the reference's state_dict keys (modified in place in training mode)."""
    x1 = double_conv(x, sd, "inc.", training)
    skips = [x1]
    y = x1
    for i in range(1, 5):
        y = double_conv(max_pool2x2(y)[0], sd, f"down{i}.maxpool_conv.1.", training)      # unet_parts.py:37-47
        skips.append(y)
    for j, i in enumerate((3, 2, 1, 0)):
        p = f"up{j + 1}."
        u = conv_transpose2d_k2s2(y, sd[p + "up.weight"], sd[p + "up.bias"])               # unet_parts.py:56-58,62
        s = skips[i]
        dy, dx = s.shape[2] - u.shape[2], s.shape[3] - u.shape[3]                          # :64-68 (zero at even sizes)
        u = np.pad(u, ((0, 0), (0, 0), (dy // 2, dy - dy // 2), (dx // 2, dx - dx // 2)))
        y = double_conv(np.concatenate([s, u], axis=1), sd, p + "conv.", training)         # :69-70
    return conv2d(y, sd["outc.conv.weight"], sd["outc.conv.bias"], padding=0)              # :73-79
