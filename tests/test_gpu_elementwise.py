"""Parity of the HBM-bound kernels (BatchNorm stats/apply/backward, MaxPool2d(2) with indices, stem conv,
head + loss, optimizer) through the C ABI against torch.nn.functional / the oracle on the same inputs."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module", autouse=True)
def _exact_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _ops():
    from jcfszxc_unet_b200 import ops

    return ops


DEV = "cuda:0"


def _scratch(ops, units, c):
    partial = torch.empty(max(ops.chan_partial_floats(units, c), 4096), device=DEV)
    sums = torch.zeros(2 * c, dtype=torch.float64, device=DEV)
    return partial, sums


@pytest.mark.parametrize("n,h,w,c,pad", [(2, 16, 16, 64, 0), (1, 32, 24, 128, 64), (3, 8, 8, 1024, 0), (1, 10, 14, 72, 8)])
def test_bn_train_forward(n, h, w, c, pad):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(c + h)
    buf = torch.zeros(n, h, w, c + pad, device=DEV, dtype=torch.bfloat16)
    raw = buf[..., pad:]
    raw.copy_(torch.randn(n, h, w, c, device=DEV, generator=g) * 2 + 0.5)
    gamma = torch.rand(c, device=DEV, generator=g) + 0.5
    beta = torch.randn(c, device=DEV, generator=g)
    rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    nbt = torch.zeros((), dtype=torch.int64, device=DEV)
    partial, sums = _scratch(ops, n * h * w, c)
    stat = torch.zeros(4, c, device=DEV)
    ops.bn_stats(raw, partial, sums)
    ops.bn_finalize(sums, n * h * w, gamma, beta, 1e-5, 0.1, rm, rv, nbt, stat[0], stat[1], stat[2], stat[3])
    out = torch.empty(n, h, w, c, device=DEV, dtype=torch.bfloat16)
    ops.bn_apply(raw, stat[0], stat[1], out, None, True)
    # checker: nn.BatchNorm2d arithmetic in fp32 on the same bf16 input
    x = raw.float().permute(0, 3, 1, 2)
    rm_ref, rv_ref = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    ref = F.relu(F.batch_norm(x, rm_ref, rv_ref, gamma, beta, True, 0.1, 1e-5)).permute(0, 2, 3, 1)
    assert (out.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item() + 1e-3
    assert torch.allclose(rm, rm_ref, rtol=1e-5, atol=1e-6) and torch.allclose(rv, rv_ref, rtol=1e-4, atol=1e-6)
    assert int(nbt) == 1
    assert torch.allclose(stat[2], x.mean(dim=(0, 2, 3)), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("n,h,w,c", [(2, 16, 16, 64), (1, 8, 12, 256)])
def test_bn_apply_fused_pool_equals_separate_and_matches_torch(n, h, w, c):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(1)
    raw = (torch.randn(n, h, w, c, device=DEV, generator=g)).bfloat16()
    scale = torch.rand(c, device=DEV, generator=g) + 0.5
    shift = torch.randn(c, device=DEV, generator=g) * 0.5
    out = torch.empty_like(raw)
    pooled = torch.empty(n, h // 2, w // 2, c, device=DEV, dtype=torch.bfloat16)
    ops.bn_apply(raw, scale, shift, out, pooled, True)
    out2 = torch.empty_like(raw)
    ops.bn_apply(raw, scale, shift, out2, None, True)
    assert torch.equal(out, out2)
    ref_pool = F.max_pool2d(out.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert torch.equal(pooled.float(), ref_pool)
    pooled2 = torch.empty_like(pooled)
    ops.maxpool_fwd(out, pooled2)
    assert torch.equal(pooled, pooled2)


@pytest.mark.parametrize("n,h,w,c,pool,ncopy", [(2, 16, 16, 32, False, 3), (1, 8, 12, 64, True, 2), (1, 6, 10, 24, True, 1),
                                                 (3, 5, 7, 128, False, 1)])
def test_bn_apply_with_extra_destinations(n, h, w, c, pool, ncopy):
    """unetk_bn_apply_copies = unetk_bn_apply + one unetk_add_n copy per extra destination (the members of NestedUNet's
    torch.cat's, UNetPP.py:80-97): same bits in every destination, nothing written outside the channel slices."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(n * c + h)
    raw = (torch.randn(n, h, w, c, device=DEV, generator=g) * 2).bfloat16()
    sc = torch.rand(c, device=DEV, generator=g) + 0.5
    sh = torch.randn(c, device=DEV, generator=g) * 0.3
    ref = torch.empty(n, h, w, c, device=DEV, dtype=torch.bfloat16)
    ref_pool = torch.empty(n, h // 2, w // 2, c, device=DEV, dtype=torch.bfloat16) if pool else None
    ops.bn_apply(raw, sc, sh, ref, ref_pool, True)
    out = torch.empty_like(ref)
    out_pool = torch.empty_like(ref_pool) if pool else None
    bufs = [torch.full((n, h, w, c + 8 * (e + 1) + 16), 7.0, device=DEV, dtype=torch.bfloat16) for e in range(ncopy)]
    views = [b[..., 8 * (e + 1): 8 * (e + 1) + c] for e, b in enumerate(bufs)]
    ops.bn_apply_copies(raw, sc, sh, out, views, out_pool, True)
    assert torch.equal(out, ref)
    if pool:
        assert torch.equal(out_pool, ref_pool)
    for e, (b, v) in enumerate(zip(bufs, views)):
        assert torch.equal(v, ref)
        assert bool((b[..., : 8 * (e + 1)] == 7.0).all()) and bool((b[..., 8 * (e + 1) + c:] == 7.0).all())


def test_maxpool_indices_bit_exact_golden_and_live():
    """Indices must equal F.max_pool2d(..., return_indices=True) bit for bit: ties (first max wins),
    NaN (always taken, last wins), +inf.  Golden = reference torch output on CPU; live = torch on this GPU."""
    ops = _ops()
    gold = np.load(os.path.join(GOLDEN, "maxpool_indices.npz"))
    x = torch.from_numpy(gold["x"]).to(DEV)                      # [2, 8, 8, 12] NCHW, all values bf16-exact? no:
    xb = x.bfloat16()                                            # kernels are bf16; round first, compare on the rounded tensor
    xn = xb.permute(0, 2, 3, 1).contiguous()
    n, h, w, c = xn.shape
    y = torch.empty(n, h // 2, w // 2, c, device=DEV, dtype=torch.bfloat16)
    idx = torch.full((n, c, h // 2, w // 2), -1, dtype=torch.int64, device=DEV)
    ops.maxpool_fwd(xn, y, idx)
    ref_v, ref_i = F.max_pool2d(xb.float(), 2, return_indices=True)
    assert torch.equal(idx, ref_i)
    assert torch.equal(torch.nan_to_num(y.float().permute(0, 3, 1, 2), nan=-7.0), torch.nan_to_num(ref_v, nan=-7.0))
    # ties and NaNs sit at bf16-exact positions, so the golden indices (from fp32) must agree wherever rounding
    # did not reorder a window; the NaN / all-zero windows are the ones the rule is about:
    gi = torch.from_numpy(gold["indices"]).to(DEV)
    special = torch.isnan(ref_v) | (ref_v == 0)
    assert torch.equal(idx[special], gi[special])
    # larger random case with heavy ties (post-ReLU zeros)
    g = torch.Generator(device=DEV).manual_seed(3)
    xr = torch.relu(torch.randn(4, 64, 32, 64, device=DEV, generator=g)).bfloat16()   # NCHW
    xrn = xr.permute(0, 2, 3, 1).contiguous()
    y = torch.empty(4, 16, 32, 64, device=DEV, dtype=torch.bfloat16)
    idx = torch.empty(4, 64, 16, 32, dtype=torch.int64, device=DEV)
    ops.maxpool_fwd(xrn, y, idx)
    rv, ri = F.max_pool2d(xr.float(), 2, return_indices=True)
    assert torch.equal(idx, ri) and torch.equal(y.float().permute(0, 3, 1, 2), rv)
    # backward = scatter through the same argmax
    dy = torch.randn(4, 16, 32, 64, device=DEV, generator=g).bfloat16()
    dx = torch.empty_like(xrn)
    ops.maxpool_bwd(xrn, dy, dx)
    xr32 = xr.float().requires_grad_(True)
    F.max_pool2d(xr32, 2).backward(dy.float().permute(0, 3, 1, 2))
    assert torch.equal(dx.float().permute(0, 3, 1, 2), xr32.grad)


@pytest.mark.parametrize("n,h,w,c,pool,skip", [(2, 16, 16, 64, False, True), (2, 16, 16, 64, True, True),
                                               (1, 8, 8, 512, True, False), (1, 12, 20, 128, False, True)])
def test_bn_relu_backward(n, h, w, c, pool, skip):
    """dL/d(raw), dgamma, dbeta of out = relu(bn(raw)) [+ maxpool consumer] against autograd on the same graph."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(17)
    raw = torch.randn(n, h, w, c, device=DEV, generator=g).bfloat16()
    gamma = (torch.rand(c, device=DEV, generator=g) + 0.5)
    beta = torch.randn(c, device=DEV, generator=g) * 0.3
    partial, sums = _scratch(ops, n * h * w, c)
    stat = torch.zeros(4, c, device=DEV)
    ops.bn_stats(raw, partial, sums)
    ops.bn_finalize(sums, n * h * w, gamma, beta, 1e-5, 0.1, None, None, None, stat[0], stat[1], stat[2], stat[3])
    out = torch.empty_like(raw)
    ops.bn_apply(raw, stat[0], stat[1], out, None, True)
    g1 = torch.randn(n, h, w, c, device=DEV, generator=g).bfloat16() if skip else None
    gp = torch.randn(n, h // 2, w // 2, c, device=DEV, generator=g).bfloat16() if pool else None
    dgamma, dbeta = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
    coef = torch.zeros(2 * c, device=DEV)
    draw = torch.empty_like(raw)
    ops.bn_bwd_reduce(raw, g1, gp, stat[0], stat[1], stat[2], stat[3], partial, sums, True)
    ops.bn_bwd_apply(raw, g1, gp, stat[0], stat[1], stat[2], stat[3], sums, n * h * w, dgamma, dbeta, coef, draw, True)
    # autograd reference; the bf16 rounding of the BN output is emulated with a straight-through cast so that the
    # ReLU mask and the pool argmax are decided on the same values the kernel saw
    x = raw.float().permute(0, 3, 1, 2).requires_grad_(True)
    gm, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    z = F.batch_norm(x, None, None, gm, bt, True, 0.1, 1e-5)
    z = z + (z.detach().bfloat16().float() - z.detach())
    a = F.relu(z)
    if skip and pool:
        # two consumers: autograd adds their bf16 gradients into one bf16 tensor (one rounding of the sum)
        a.register_hook(lambda gr: gr.bfloat16().float())
    loss = 0
    if skip:
        loss = loss + (a * g1.float().permute(0, 3, 1, 2)).sum()
    if pool:
        loss = loss + (F.max_pool2d(a, 2) * gp.float().permute(0, 3, 1, 2)).sum()
    loss.backward()
    ref = x.grad.permute(0, 2, 3, 1)
    scale = ref.abs().max().item()
    assert (draw.float() - ref).abs().max().item() <= 1.5e-2 * scale
    assert torch.allclose(dgamma, gm.grad, rtol=2e-3, atol=2e-3 * gm.grad.abs().max().item())
    assert torch.allclose(dbeta, bt.grad, rtol=2e-3, atol=2e-3 * bt.grad.abs().max().item())


@pytest.mark.parametrize("n,h,w,c,pad", [(2, 16, 24, 64, 0), (1, 9, 7, 40, 24), (1, 33, 65, 8, 0)])
def test_bn_backward_with_residual_gradient(n, h, w, c, pad):
    """out = relu(bn(raw)) + res (Recurrent_block's x + x1, unet_parts.py:125-132): d(res) = the incoming gradient, delivered by
    the BatchNorm backward's apply pass (write and accumulate, channel-sliced destination) — bit-identical with the separate
    add pass it replaces, d(raw) / dgamma / dbeta bit-identical with the plain pass."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(23 + c)
    raw = torch.randn(n, h, w, c, device=DEV, generator=g).bfloat16()
    gamma = torch.rand(c, device=DEV, generator=g) + 0.5
    beta = torch.randn(c, device=DEV, generator=g) * 0.3
    partial, sums = _scratch(ops, n * h * w, c)
    stat = torch.zeros(4, c, device=DEV)
    ops.bn_stats(raw, partial, sums)
    ops.bn_finalize(sums, n * h * w, gamma, beta, 1e-5, 0.1, None, None, None, stat[0], stat[1], stat[2], stat[3])
    g1 = torch.randn(n, h, w, c, device=DEV, generator=g).bfloat16()
    coef = torch.zeros(2 * c, device=DEV)

    def run(dres, acc):
        dgamma, dbeta = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
        draw = torch.empty_like(raw)
        ops.bn_bwd_reduce(raw, g1, None, stat[0], stat[1], stat[2], stat[3], partial, sums, True)
        ops.bn_bwd_apply(raw, g1, None, stat[0], stat[1], stat[2], stat[3], sums, n * h * w, dgamma, dbeta, coef, draw, True,
                         dres=dres, dres_accumulate=acc)
        return draw, dgamma, dbeta

    plain = run(None, False)
    buf = torch.full((n, h, w, c + pad), 3.0, device=DEV, dtype=torch.bfloat16)
    view = buf[..., pad:]
    fused = run(view, False)
    for a, b in zip(plain, fused):
        assert torch.equal(a, b)
    assert torch.equal(view, g1) and (pad == 0 or bool((buf[..., :pad] == 3.0).all()))
    base = torch.randn(n, h, w, c, device=DEV, generator=g).bfloat16()
    view.copy_(base)
    ref = base.clone()
    ops.add_n(ref, [g1], accumulate=True)
    fused = run(view, True)
    for a, b in zip(plain, fused):
        assert torch.equal(a, b)
    assert torch.equal(view, ref)


@pytest.mark.parametrize("layout", ["nchw", "channels_last"])
@pytest.mark.parametrize("cout,bias", [(64, False), (32, True)])
def test_stem_conv(layout, cout, bias):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(2)
    x = torch.rand(2, 3, 36, 40, device=DEV, generator=g)
    if layout == "channels_last":
        x = x.contiguous(memory_format=torch.channels_last)
    w = torch.randn(cout, 3, 3, 3, device=DEV, generator=g) * 0.2
    b = torch.randn(cout, device=DEV, generator=g) if bias else None
    y = torch.empty(2, 36, 40, cout, device=DEV, dtype=torch.bfloat16)
    ops.stem_fwd(x, w, b, y)
    ref = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), b, padding=1).permute(0, 2, 3, 1)
    assert (y.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
    dy = torch.randn(2, 36, 40, cout, device=DEV, generator=g).bfloat16()
    dw = torch.empty(cout, 3, 3, 3, device=DEV)
    ops.stem_wgrad(x, dy, dw)
    refw = torch.nn.grad.conv2d_weight(x.bfloat16().float(), (cout, 3, 3, 3), dy.float().permute(0, 3, 1, 2), padding=1)
    assert (dw - refw).abs().max().item() <= 1e-3 * refw.abs().max().item()


@pytest.mark.parametrize("n,h,w,cout,bias", [(2, 36, 40, 64, False), (1, 20, 200, 128, True), (3, 16, 128, 64, True),
                                             (2, 24, 72, 32, True)])
def test_stem_conv_with_fused_bn_statistics(n, h, w, cout, bias):
    """unetk_stem_conv3x3_fwd_bnstats = unetk_stem_conv3x3_fwd + unetk_bn_stats (statistics of the bf16 output taken
    in the conv epilogue; Cout = 32 takes the two-pass route inside the library): same output bits, same sums."""
    ops = _ops()
    from jcfszxc_unet_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(cout + w)
    x = torch.rand(n, 3, h, w, device=DEV, generator=g)
    wt = torch.randn(cout, 3, 3, 3, device=DEV, generator=g) * 0.2
    b = torch.randn(cout, device=DEV, generator=g) if bias else None
    npart = max(_lib.load().unetk_stem_stats_partial_floats(n, h, w, cout), ops.chan_partial_floats(n * h * w, cout), 4096)
    partial = torch.empty(npart, device=DEV)
    y_a = torch.empty(n, h, w, cout, device=DEV, dtype=torch.bfloat16)
    y_b = torch.empty_like(y_a)
    s_a = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    s_b = torch.zeros_like(s_a)
    ops.stem_fwd(x, wt, b, y_a)
    ops.bn_stats(y_a, partial, s_a)
    ops.stem_fwd_stats(x, wt, b, y_b, partial, s_b)
    assert torch.equal(y_a, y_b)
    ref = torch.stack([y_a.double().sum(dim=(0, 1, 2)), (y_a.double() ** 2).sum(dim=(0, 1, 2))]).view(-1)
    assert torch.allclose(s_b, ref, rtol=2e-6, atol=1e-4)
    assert torch.allclose(s_b, s_a, rtol=2e-6, atol=1e-4)


@pytest.mark.parametrize("c", [64, 32])
def test_head_loss_forward_backward_vs_oracle(c):
    from oracle import unet_oracle as O

    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(4)
    n, h, w = 2, 24, 40
    x = torch.randn(n, h, w, c, device=DEV, generator=g).bfloat16()
    wt = torch.randn(1, c, 1, 1, device=DEV, generator=g) * 0.3
    b = torch.randn(1, device=DEV, generator=g)
    labels = (torch.rand(n, 1, h, w, device=DEV, generator=g) < 0.12).float()
    npix = n * h * w
    partial = torch.empty(max(ops.head_partial_floats(npix, c), 4096), device=DEV)
    sums = torch.zeros(4, dtype=torch.float64, device=DEV)
    fin = torch.zeros(8, device=DEV)
    logits = torch.empty(n, 1, h, w, device=DEV)
    ops.head_fwd(x, wt.view(-1), b, labels, logits, partial, sums)
    ops.loss_finalize(sums, npix, fin)
    # oracle on the same values (train.py:264-278 + dice_score.py)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr, br = wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    z = F.conv2d(xr, wr, br)
    loss, bce, dice_l = O.segmentation_loss(z, labels)
    assert torch.allclose(logits, z.detach(), rtol=1e-4, atol=1e-4)
    assert abs(float(fin[0]) - float(loss)) <= 1e-5 and abs(float(fin[1]) - float(bce)) <= 1e-5
    assert abs(float(fin[2]) - (1 - float(dice_l))) <= 1e-5                       # Dice within 1e-3 (north_star): here 1e-5
    loss.backward()
    dx = torch.empty_like(x)
    dw, db = torch.zeros(c, device=DEV), torch.zeros(1, device=DEV)
    ops.head_bwd(x, wt.view(-1), labels, logits, fin, None, 1.0, dx, dw, db, partial)
    ref = xr.grad.permute(0, 2, 3, 1)
    assert (dx.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
    assert torch.allclose(dw, wr.grad.view(-1), rtol=1e-3, atol=1e-3 * wr.grad.abs().max().item())
    assert torch.allclose(db, br.grad, rtol=1e-3, atol=1e-6)
    # explicit dlogits path (loss computed outside, e.g. by the reference's train.py)
    dlog = torch.randn(n, 1, h, w, device=DEV, generator=g)
    ops.head_bwd(x, wt.view(-1), None, None, None, dlog, 1.0, dx, dw, db, partial)
    ref = (dlog.permute(0, 2, 3, 1) * wt.view(1, 1, 1, c))
    assert (dx.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
    # empty mask: dice == 1 branch
    zero_labels = torch.zeros_like(labels)
    xs = (x * 0 - 0).bfloat16()
    ops.head_fwd(xs, wt.view(-1), torch.full((1,), -40.0, device=DEV), zero_labels, logits, partial, sums)
    ops.loss_finalize(sums, npix, fin)
    ref_loss, _, ref_dl = O.segmentation_loss(torch.full((n, 1, h, w), -40.0, device=DEV), zero_labels)
    assert abs(float(fin[2]) - (1 - float(ref_dl))) <= 1e-4 and abs(float(fin[0]) - float(ref_loss)) <= 1e-4
    # truly empty: 32 pixels * 1e-7 < eps -> sets_sum := inter branch of dice_score.py:35 (dice == 1, zero dice gradient)
    xt = torch.zeros(1, 4, 8, c, device=DEV, dtype=torch.bfloat16)
    lt = torch.zeros(1, 1, 4, 8, device=DEV)
    zt = torch.empty(1, 1, 4, 8, device=DEV)
    ops.head_fwd(xt, wt.view(-1), torch.full((1,), -40.0, device=DEV), lt, zt, partial, sums)
    ops.loss_finalize(sums, 32, fin)
    ref_loss, _, ref_dl = O.segmentation_loss(torch.full((1, 1, 4, 8), -40.0, device=DEV), lt)
    assert float(ref_dl) == 0.0 and abs(float(fin[2]) - 1.0) <= 1e-6 and abs(float(fin[0]) - float(ref_loss)) <= 1e-6
    assert float(fin[4]) == 0.0 and float(fin[5]) == 0.0


@pytest.mark.parametrize("n,h,w,c,relu,post_sigmoid,pad", [(2, 24, 40, 64, True, False, 0), (1, 20, 12, 32, True, True, 32),
                                                          (3, 8, 8, 64, False, False, 0), (1, 3, 5, 32, True, False, 8),
                                                          (1, 64, 96, 32, True, False, 0)])
def test_bn_head_fused_equals_separate_passes_and_oracle(n, h, w, c, relu, post_sigmoid, pad):
    """DoubleConv's last BatchNorm+ReLU folded into OutConv + loss (unet_parts.py:24-31,73-79; train.py:264-278):
    forward logits bit-identical to unetk_bn_apply -> unetk_head_fwd, backward equal to unetk_head_bwd ->
    unetk_bn_bwd_reduce/apply up to the summation order, and both against autograd on the same graph."""
    from oracle import unet_oracle as O

    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(c + h)
    buf = torch.zeros(n, h, w, c + pad, device=DEV, dtype=torch.bfloat16)
    raw = buf[..., pad:]
    raw.copy_(torch.randn(n, h, w, c, device=DEV, generator=g) * 1.5 + 0.2)
    gamma = torch.rand(c, device=DEV, generator=g) + 0.5
    beta = torch.randn(c, device=DEV, generator=g) * 0.3
    wt = torch.randn(c, device=DEV, generator=g) * 0.3
    b = torch.randn(1, device=DEV, generator=g)
    labels = (torch.rand(n, 1, h, w, device=DEV, generator=g) < 0.12).float()
    npix = n * h * w
    partial = torch.empty(max(ops.bn_head_partial_floats(npix, c), ops.chan_partial_floats(npix, c), 4096), device=DEV)
    sums = torch.zeros(2 * c, dtype=torch.float64, device=DEV)
    stat = torch.zeros(4, c, device=DEV)
    ops.bn_stats(raw, partial, sums)
    ops.bn_finalize(sums, npix, gamma, beta, 1e-5, 0.1, None, None, None, stat[0], stat[1], stat[2], stat[3])
    # --- separate passes
    out = torch.empty(n, h, w, c, device=DEV, dtype=torch.bfloat16)
    ops.bn_apply(raw, stat[0], stat[1], out, None, relu)
    ls_a, ls_b = torch.zeros(4, dtype=torch.float64, device=DEV), torch.zeros(4, dtype=torch.float64, device=DEV)
    fin = torch.zeros(8, device=DEV)
    logits_a, logits_b = torch.empty(n, 1, h, w, device=DEV), torch.empty(n, 1, h, w, device=DEV)
    ops.head_fwd(out, wt, b, labels, logits_a, partial, ls_a, post_sigmoid)
    ops.bn_head_fwd(raw, stat[0], stat[1], relu, wt, b, labels, logits_b, partial, ls_b, post_sigmoid)
    assert torch.equal(logits_a, logits_b)
    assert torch.allclose(ls_a, ls_b, rtol=1e-5, atol=1e-7)   # fp32 block partials in a different order
    ops.loss_finalize(ls_b, npix, fin)
    dx = torch.empty_like(out)
    dw_a, db_a = torch.zeros(c, device=DEV), torch.zeros(1, device=DEV)
    ops.head_bwd(out, wt, labels, logits_a, fin, None, 1.0, dx, dw_a, db_a, partial, post_sigmoid=post_sigmoid)
    dgamma_a, dbeta_a = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
    coef = torch.zeros(2 * c, device=DEV)
    draw_a = torch.empty(n, h, w, c, device=DEV, dtype=torch.bfloat16)
    ops.bn_bwd_reduce(raw, dx, None, stat[0], stat[1], stat[2], stat[3], partial, sums, relu)
    ops.bn_bwd_apply(raw, dx, None, stat[0], stat[1], stat[2], stat[3], sums, npix, dgamma_a, dbeta_a, coef, draw_a, relu)
    # --- fused
    dz = torch.empty(npix, device=DEV)
    dw_b, db_b = torch.full((c,), 2.0, device=DEV), torch.full((1,), 3.0, device=DEV)
    dgamma_b, dbeta_b = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
    dbuf = torch.zeros(n, h, w, c + pad, device=DEV, dtype=torch.bfloat16)
    draw_b = dbuf[..., pad:]
    ops.bn_head_bwd_reduce(raw, stat[0], stat[1], stat[2], relu, wt, labels, logits_b, fin, None, 1.0, dz, dw_b, db_b, sums,
                           partial, accumulate=True, post_sigmoid=post_sigmoid)
    ops.bn_bwd_coef(sums, npix, stat[0], stat[2], stat[3], dgamma_b, dbeta_b, coef)
    ops.bn_head_bwd_apply(raw, stat[0], stat[1], relu, wt, dz, coef, draw_b)
    assert float(dbuf[..., :pad].abs().sum()) == 0.0
    tol = lambda t: 1e-4 * t.abs().max().item() + 1e-12
    assert torch.allclose(dw_b - 2.0, dw_a, rtol=1e-4, atol=max(tol(dw_a), 1e-6))          # accumulate=True adds
    assert torch.allclose(db_b - 3.0, db_a, rtol=1e-4, atol=1e-6)
    assert torch.allclose(dgamma_b, dgamma_a, rtol=1e-4, atol=tol(dgamma_a))
    assert torch.allclose(dbeta_b, dbeta_a, rtol=1e-4, atol=tol(dbeta_a))
    # d(raw): same formula, sums differ in the last bits -> at most one bf16 ulp apart
    d = (draw_a.float() - draw_b.float()).abs()
    assert (d <= 2 ** -7 * draw_a.float().abs() + 1e-12).all()
    # --- autograd on the same graph (bf16 rounding of the activation emulated straight-through)
    x = raw.float().permute(0, 3, 1, 2).requires_grad_(True)
    gm, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    wr, br = wt.view(1, c, 1, 1).clone().requires_grad_(True), b.clone().requires_grad_(True)
    z = F.batch_norm(x, None, None, gm, bt, True, 0.1, 1e-5)
    z = z + (z.detach().bfloat16().float() - z.detach())
    a = F.relu(z) if relu else z
    y = F.conv2d(a, wr, br)
    if post_sigmoid:
        y = torch.sigmoid(y)
    loss, _, _ = O.segmentation_loss(y, labels)
    loss.backward()
    # torch's batch_norm and our fma differ in the last fp32 bit, so a few of the activations round to the other bf16
    # neighbour (one flip moves a logit by <= 2^-8 |a| |w|): tight on average, bounded per pixel
    dl = (logits_b - y.detach()).abs()
    assert dl.mean().item() <= 1e-4 and dl.max().item() <= 3e-2
    assert abs(float(fin[0]) - float(loss)) <= 1e-4
    ref = x.grad.permute(0, 2, 3, 1)
    assert (draw_b.float() - ref).abs().max().item() <= 1.5e-2 * ref.abs().max().item()
    assert torch.allclose(dw_b - 2.0, wr.grad.view(-1), rtol=2e-3, atol=2e-3 * wr.grad.abs().max().item())
    assert torch.allclose(db_b - 3.0, br.grad, rtol=2e-3, atol=1e-6)
    # (autograd keeps d(a) = dz*w in fp32 where both kernel paths round it to bf16, as the bf16 activation gradient is)
    assert torch.allclose(dgamma_b, gm.grad, rtol=1e-2, atol=1e-2 * gm.grad.abs().max().item())
    assert torch.allclose(dbeta_b, bt.grad, rtol=1e-2, atol=1e-2 * bt.grad.abs().max().item())
    # gradient handed in by autograd (loss computed by the caller)
    dlog = torch.randn(n, 1, h, w, device=DEV, generator=g)
    ops.bn_head_bwd_reduce(raw, stat[0], stat[1], stat[2], relu, wt, None, logits_b if post_sigmoid else None, None, dlog,
                           0.5, dz, dw_b, db_b, sums, partial, post_sigmoid=post_sigmoid)
    exp = 0.5 * dlog.view(-1) * ((logits_b * (1 - logits_b)).view(-1) if post_sigmoid else 1.0)
    assert torch.allclose(dz, exp, rtol=1e-6, atol=1e-9)
    assert torch.allclose(db_b, exp.sum().view(1), rtol=1e-4, atol=1e-5)


def test_clip_and_rmsprop_vs_oracle():
    from oracle import unet_oracle as O

    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(6)
    n = 1_000_003                       # not a multiple of 4: the vector kernel plus the scalar tail
    p = torch.randn(n, device=DEV, generator=g)
    grad = torch.randn(n, device=DEV, generator=g) * 0.01
    sq, buf = torch.zeros_like(p), torch.zeros_like(p)
    from jcfszxc_unet_b200 import _lib

    partial = torch.empty(_lib.load().unetk_sqnorm_partial_floats(p.numel()), device=DEV)
    clip = torch.zeros(2, device=DEV)
    pr, sqr, bufr = p.clone(), sq.clone(), buf.clone()
    for it in range(3):
        ops.grad_clip_coef(grad, 1.0, 1.0, partial, clip)
        ops.rmsprop_step(p, grad, sq, buf, 1e-3, 0.99, 1e-8, 1e-8, 0.999, clip)
        (gc,), total = O.clip_grad_norm([grad], 1.0)
        O.rmsprop_step(pr, gc, sqr, bufr, 1e-3)
        assert abs(float(clip[0]) - float(total)) <= 1e-4 * float(total)
    assert torch.allclose(p, pr, rtol=1e-5, atol=1e-6)
    assert torch.allclose(sq, sqr, rtol=1e-4, atol=1e-12) and torch.allclose(buf, bufr, rtol=1e-4, atol=1e-6)


def test_colsum():
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(8)
    x = torch.randn(2, 20, 12, 128, device=DEV, generator=g).bfloat16()
    partial = torch.empty(max(ops.chan_partial_floats(2 * 20 * 12, 128), 4096), device=DEV)
    out = torch.zeros(128, device=DEV)
    ops.colsum(x, partial, out)
    ref = x.float().sum(dim=(0, 1, 2))
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-3)
    ops.colsum(x, partial, out, accumulate=True)
    assert torch.allclose(out, 2 * ref, rtol=1e-4, atol=2e-3)


# ------------------------------------------------------------------------------------------------ pool / unpool pair
@pytest.mark.parametrize("n,h,w,c,pad", [(2, 8, 12, 16, (0, 0)), (1, 64, 96, 64, (8, 24)), (3, 2, 2, 8, (0, 0)), (1, 130, 70, 40, (0, 8))])
def test_max_unpool_bit_exact(n, h, w, c, pad):
    """SegNet-style pool / unpool (SegNet.py:89-138): byte codes and int64 indices both reproduce F.max_unpool2d bit
    for bit (ties, NaN and inf included), into channel slices of wider buffers; backward = gather."""
    import torch.nn.functional as F

    from jcfszxc_unet_b200 import ops
    from oracle import unet_oracle as O

    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(h * w + c)
    x = torch.randn(n, h, w, c, device=dev, generator=g).bfloat16()
    x[0, 0, :2, :] = 1.0                                   # ties inside a window
    x[0, -1, -1, 0] = float("nan")
    x[0, 1, 1, 1] = float("inf")
    ho, wo = h // 2, w // 2
    xn = x.float().permute(0, 3, 1, 2)
    p_ref, idx_ref = F.max_pool2d(xn, 2, 2, return_indices=True)
    # forward with codes
    y = torch.empty(n, ho, wo, c, device=dev, dtype=torch.bfloat16)
    code = torch.full((n, ho, wo, c), 255, device=dev, dtype=torch.uint8)
    ops.maxpool_fwd_codes(x, y, code)
    assert torch.equal(y.float().permute(0, 3, 1, 2).nan_to_num(7.0), p_ref.nan_to_num(7.0))
    assert int(code.max()) <= 3
    # the codes name the element ATen's indices name
    cq = code.permute(0, 3, 1, 2).long()
    hh = torch.arange(ho, device=dev).view(1, 1, ho, 1) * 2 + (cq >> 1)
    ww = torch.arange(wo, device=dev).view(1, 1, 1, wo) * 2 + (cq & 1)
    assert torch.equal(hh * w + ww, idx_ref)
    # unpool of fresh values z through both index forms, into a channel slice of a wider buffer
    z = torch.randn(n, ho, wo, c, device=dev, generator=g).bfloat16()
    zn = z.float().permute(0, 3, 1, 2).contiguous()
    ref = F.max_unpool2d(zn, idx_ref, 2, 2)
    assert np.array_equal(O.max_unpool2x2_numpy(zn.cpu().numpy(), idx_ref.cpu().numpy()), ref.cpu().numpy())   # the oracle
    for where in (code, idx_ref.contiguous()):
        buf = torch.full((n, h, w, pad[0] + c + pad[1]), -3.0, device=dev, dtype=torch.bfloat16)
        out = buf[..., pad[0]:pad[0] + c]
        ops.max_unpool(z, where, out)
        assert torch.equal(out.float().permute(0, 3, 1, 2), ref)
        assert (buf[..., :pad[0]] == -3.0).all() and (buf[..., pad[0] + c:] == -3.0).all()
        # backward: gather, plain and accumulating
        dy = torch.randn(n, h, w, c, device=dev, generator=g).bfloat16()
        dref = dy.float().permute(0, 3, 1, 2).flatten(2).gather(2, idx_ref.flatten(2)).view(n, c, ho, wo)
        dx = torch.empty(n, ho, wo, c, device=dev, dtype=torch.bfloat16)
        ops.max_unpool_bwd(dy, where, dx)
        assert torch.equal(dx.float().permute(0, 3, 1, 2), dref)
        base = torch.randn(n, ho, wo, c, device=dev, generator=g).bfloat16()
        dx2 = base.clone()
        ops.max_unpool_bwd(dy, where, dx2, accumulate=True)
        assert torch.equal(dx2, (base.float() + dref.permute(0, 2, 3, 1)).bfloat16())


@pytest.mark.parametrize("n,hs,ws,hd,wd,oy,ox,c", [(2, 6, 8, 7, 9, 0, 0, 16), (1, 4, 4, 7, 6, 1, 1, 8), (2, 7, 9, 6, 8, 0, 0, 24), (1, 5, 5, 5, 5, 0, 0, 8),
                                                   (1, 7, 9, 5, 6, -1, -2, 8)])
def test_shift_copy_is_f_pad(n, hs, ws, hd, wd, oy, ox, c):
    """unetk_shift_copy == F.pad (zero border, unet_parts.py:64-67) and, with negated offsets, its backward (a crop)."""
    from jcfszxc_unet_b200 import ops

    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(hs * wd + c)
    src = torch.randn(n, hs, ws, c, device=dev, generator=g).bfloat16()
    buf = torch.full((n, hd, wd, c + 16), 9.0, device=dev, dtype=torch.bfloat16)
    dst = buf[..., 8:8 + c]
    ops.shift_copy(dst, src, oy, ox)
    if hd >= hs:   # pad
        ref = F.pad(src.permute(0, 3, 1, 2), [ox, wd - ws - ox, oy, hd - hs - oy]).permute(0, 2, 3, 1)
    else:          # crop (negative pad)
        ref = src[:, -oy:-oy + hd, -ox:-ox + wd] if (oy or ox) else src[:, :hd, :wd]
    assert torch.equal(dst, ref)
    assert (buf[..., :8] == 9.0).all() and (buf[..., 8 + c:] == 9.0).all()


@pytest.mark.parametrize("n,h,w,c", [(1, 7, 9, 8), (2, 5, 6, 16), (1, 6, 5, 8)])
def test_maxpool_odd_sizes_floor(n, h, w, c):
    """MaxPool2d(2) floors odd sizes; the last row / column belongs to no window: zero gradient (written, not skipped)."""
    from jcfszxc_unet_b200 import ops

    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(h * w)
    x = torch.randn(n, h, w, c, device=dev, generator=g).bfloat16()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    yr = F.max_pool2d(xr, 2)
    gy = torch.randn(yr.shape, device=dev, generator=g).bfloat16().float()
    yr.backward(gy)
    y = torch.empty(n, h // 2, w // 2, c, device=dev, dtype=torch.bfloat16)
    ops.maxpool_fwd(x, y)
    assert torch.equal(y.float().permute(0, 3, 1, 2), yr.detach())
    dx = torch.full((n, h, w, c), float("nan"), device=dev, dtype=torch.bfloat16)
    ops.maxpool_bwd(x, gy.permute(0, 2, 3, 1).contiguous().bfloat16(), dx)
    assert torch.equal(dx.float().permute(0, 3, 1, 2), xr.grad)


def test_abi_is_reentrant_across_threads():
    """the C ABI is called from the main thread (forward) and from autograd worker threads (backward).
    (1) forward on the main thread, backward explicitly on another Python thread: same gradients as in one thread;
    (2) unetk_last_error is per thread: a failing call on a worker thread neither clobbers nor reads the main thread's
    message; (3) kernels that need > 48 KB of shared memory launch from a thread that never called them before."""
    import threading

    from jcfszxc_unet_b200 import _lib, ops
    from UNetFamily.utils.unet_parts import DoubleConv

    lib = _lib.load()
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    dc = DoubleConv(64, 64).to(dev).train()
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn(2, 64, 16, 160, device=dev, generator=g)
    gy = torch.randn(2, 64, 16, 160, device=dev, generator=g)

    def grads(threaded):
        dc.zero_grad(set_to_none=True)
        y = dc(x)
        if threaded:
            err = []

            def work():
                try:
                    torch.cuda.set_device(dev)
                    (y.float() * gy).sum().backward()
                except Exception as e:  # pragma: no cover
                    err.append(e)

            t = threading.Thread(target=work)
            t.start()
            t.join()
            assert not err, err
        else:
            (y.float() * gy).sum().backward()
        torch.cuda.synchronize()
        return [p.grad.clone() for p in dc.parameters()]

    a, b = grads(False), grads(True)
    assert all(torch.equal(p, q) for p, q in zip(a, b))

    # per-thread error text
    bad = torch.zeros(1, 4, 4, 12, device=dev, dtype=torch.bfloat16)          # C = 12: not a multiple of 8
    out = torch.zeros(1, 2, 2, 12, device=dev, dtype=torch.bfloat16)
    with pytest.raises(_lib.UnetkError, match="multiple of 8"):
        ops.maxpool_fwd(bad, out)
    main_msg = lib.unetk_last_error()
    seen = {}

    def worker():
        torch.cuda.set_device(dev)
        seen["before"] = lib.unetk_last_error()
        rc = lib.unetk_shift_copy(out.data_ptr(), 12, 2, 2, bad.data_ptr(), 12, 4, 4, 0, 0, 1, 12, None)
        seen["rc"], seen["after"] = rc, lib.unetk_last_error()
        # first use of a big-shared-memory kernel from this thread
        xw = torch.randn(1, 8, 128, 64, device=dev).bfloat16()
        w_ab, _ = ops.pack_weight(torch.randn(64, 64, 3, 3, device=dev) * 0.05, True, False)
        yw = torch.empty(1, 8, 128, 64, device=dev, dtype=torch.bfloat16)
        ops.conv_fwd(xw, w_ab, None, yw, 3)
        torch.cuda.synchronize()
        seen["conv_ok"] = bool(torch.isfinite(yw.float()).all())

    t = threading.Thread(target=worker)
    t.start()
    t.join()
    assert seen["before"] in (b"", None) and seen["rc"] != 0 and b"shift_copy" in seen["after"] and seen["conv_ok"]
    assert lib.unetk_last_error() == main_msg and b"multiple of 8" in main_msg
