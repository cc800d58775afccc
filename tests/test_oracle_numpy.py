"""The numpy restatement of the arithmetic (oracle/numpy_ops.py) pinned against torch.nn.functional per primitive and
against the golden logits of the UNMODIFIED reference — so that the torch-based oracle is not its own only witness."""
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import numpy_ops as N

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _r(*shape, seed=0):
    return np.random.RandomState(seed).randn(*shape)


def _close(a, b, tol=1e-10):
    assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max()), np.abs(a - b).max()


def test_conv2d_and_transpose_vs_torch_float64():
    x, w, b = _r(2, 5, 9, 7, seed=1), _r(6, 5, 3, 3, seed=2), _r(6, seed=3)
    _close(N.conv2d(x, w, b), F.conv2d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), padding=1).numpy())
    _close(N.conv2d(x, w, None, stride=2), F.conv2d(torch.from_numpy(x), torch.from_numpy(w), None, stride=2, padding=1).numpy())
    w1 = _r(4, 5, 1, 1, seed=4)
    _close(N.conv2d(x, w1, None, padding=0), F.conv2d(torch.from_numpy(x), torch.from_numpy(w1)).numpy())
    wt, bt = _r(5, 3, 2, 2, seed=5), _r(3, seed=6)
    _close(N.conv_transpose2d_k2s2(x, wt, bt),
           F.conv_transpose2d(torch.from_numpy(x), torch.from_numpy(wt), torch.from_numpy(bt), stride=2).numpy())


def test_batch_norm_train_and_eval_vs_torch_float64():
    x = _r(3, 4, 6, 5, seed=7) * 2 + 1
    g, b = _r(4, seed=8), _r(4, seed=9)
    rm, rv = _r(4, seed=10) * 0.1, np.abs(_r(4, seed=11)) + 0.5
    trm, trv = torch.from_numpy(rm.copy()), torch.from_numpy(rv.copy())
    ref = F.batch_norm(torch.from_numpy(x), trm, trv, torch.from_numpy(g), torch.from_numpy(b), True, 0.1, 1e-5).numpy()
    y, nrm, nrv = N.batch_norm(x, g, b, rm, rv, True)
    _close(y, ref)
    _close(nrm, trm.numpy())
    _close(nrv, trv.numpy())
    ref_e = F.batch_norm(torch.from_numpy(x), torch.from_numpy(rm), torch.from_numpy(rv), torch.from_numpy(g),
                         torch.from_numpy(b), False, 0.1, 1e-5).numpy()
    _close(N.batch_norm(x, g, b, rm, rv, False)[0], ref_e)


def test_pool_upsample_loss_vs_torch():
    x = _r(2, 3, 8, 6, seed=12)
    x[0, 0, 0, 0] = x[0, 0, 0, 1] = 5.0                        # a tie: the first maximum wins
    v, i = F.max_pool2d(torch.from_numpy(x), 2, return_indices=True)
    nv, ni = N.max_pool2x2(x)
    assert np.array_equal(nv, v.numpy()) and np.array_equal(ni, i.numpy())
    xo = _r(1, 2, 7, 5, seed=13)                               # odd size: the last row / column is dropped
    assert np.array_equal(N.max_pool2x2(xo)[0], F.max_pool2d(torch.from_numpy(xo), 2).numpy())
    _close(N.upsample_nearest2x(x), F.interpolate(torch.from_numpy(x), scale_factor=2, mode="nearest").numpy())
    _close(N.upsample_bilinear2x_align_corners(x),
           F.interpolate(torch.from_numpy(x), scale_factor=2, mode="bilinear", align_corners=True).numpy())
    z, y = _r(4, 1, 5, 5, seed=14) * 4, (np.random.RandomState(15).rand(4, 1, 5, 5) < 0.3).astype(np.float64)
    assert abs(N.bce_with_logits(z, y) - float(F.binary_cross_entropy_with_logits(torch.from_numpy(z), torch.from_numpy(y)))) <= 1e-12
    _close(N.sigmoid(z), torch.sigmoid(torch.from_numpy(z)).numpy())


def test_numpy_unet_forward_matches_reference_golden():
    """Golden logits were produced by the UNMODIFIED reference (fp32, CPU, seed 42); weights are regenerated from the
    seed by our drop-in module, whose default initialisation reproduces the reference's draw for draw."""
    from UNetFamily.UNet import UNet

    g = np.load(os.path.join(GOLDEN, "unet_forward_seed42.npz"))
    torch.manual_seed(42)
    m = UNet(3, 1)
    sd = {k: v.detach().numpy().astype(np.float64) for k, v in m.state_dict().items()}
    x = g["images"].astype(np.float64)
    y = N.unet_forward(x, sd, training=True)
    ref = g["logits_train"]
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= 1e-4 * np.abs(ref).max(), np.abs(y - ref).max() / np.abs(ref).max()
    assert np.allclose(sd["inc.double_conv.1.running_mean"], g["running_mean_inc1"], rtol=1e-4, atol=1e-6)
    assert np.allclose(sd["up4.conv.double_conv.4.running_var"], g["running_var_up4_4"], rtol=1e-4, atol=1e-6)
    ye = N.unet_forward(x, sd, training=False)
    ref_e = g["logits_eval_after_1_train_fwd"]
    assert np.abs(ye - ref_e).max() <= 1e-4 * np.abs(ref_e).max()


def test_subpixel_upconv_identity_vs_torch_float64():
    """nearest 2x + conv3x3(pad 1) == four 2x2-tap convs of the low-resolution input with pre-summed taps (the form the
    CUDA up-conv kernels compute in), forward and weight-gradient fold, against the reference's formulation
    (unet_parts.py:103-104) in torch float64; odd sizes, so that every border case of the window is hit."""
    x, w, b = _r(2, 5, 5, 7, seed=21), _r(4, 5, 3, 3, seed=22), _r(4, seed=23)
    xt = torch.from_numpy(x)
    wt = torch.from_numpy(w).requires_grad_(True)
    ref = F.conv2d(F.interpolate(xt, scale_factor=2, mode="nearest"), wt, torch.from_numpy(b), padding=1)
    _close(N.upconv_subpixel(x, w, b), ref.detach().numpy())
    _close(N.upconv_subpixel(x, w), N.conv2d(N.upsample_nearest2x(x), w))
    # gradients of the sixteen sub-filters (autograd through the restated forward), folded back to the 3x3 filter
    dy = torch.from_numpy(_r(*ref.shape, seed=24))
    ref.backward(dy)
    wq = torch.from_numpy(N.subpixel_weights(w)).requires_grad_(True)
    xp = F.pad(xt, (1, 1, 1, 1))
    y = torch.zeros_like(ref)
    for qy in range(2):
        for qx in range(2):
            for u in range(2):
                for v in range(2):
                    patch = xp[:, :, qy + u:qy + u + 5, qx + v:qx + v + 7]
                    y[:, :, qy::2, qx::2] += torch.einsum("nchw,oc->nohw", patch, wq[qy, qx, u, v])
    y.backward(dy)
    _close(N.fold_subpixel_wgrad(wq.grad.numpy()), wt.grad.numpy())
