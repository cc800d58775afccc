"""CPU ORACLE, second layer — the arithmetic itself, restated in numpy (float64).  TEST INFRASTRUCTURE ONLY.

oracle/unet_oracle.py restates the reference's GRAPH and calls torch.nn.functional for the arithmetic, because the
reference's arithmetic lives in an un-vendored third-party dependency: PyTorch (no pin file in the reference; this
image: torch 2.11.0+cu128).  This file restates that dependency's published semantics for every primitive on the hot
path, independently of torch, so that the checker does not rest on "the same library called twice":

    conv2d (stride 1/2, zero padding)      torch.nn.Conv2d           cross-correlation, NCHW, weight [Cout,Cin,kh,kw]
    conv_transpose2d k=2 s=2               torch.nn.ConvTranspose2d  weight [Cin,Cout,2,2], non-overlapping scatter
    batch_norm (train / eval)              torch.nn.BatchNorm2d      biased variance to normalise, unbiased for running_var
    relu, sigmoid, max_pool2d(2) + indices, nearest 2x / bilinear 2x (align_corners=True) up-sampling
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
