"""Kernel-level parity of the tcgen05 tap-GEMM / weight-gradient kernels, called through the C ABI.

Checker: torch.nn.functional in fp32 (TF32 disabled) on the SAME bf16-rounded inputs, i.e. the
arithmetic the reference's nn.Conv2d / nn.ConvTranspose2d perform under bf16 autocast
(reference UNetFamily/utils/unet_parts.py:24-31,56-58,77) with fp32 accumulation.
Tolerance: one bf16 rounding of the output (2^-8 relative) plus accumulation-order noise.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _exact_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _ops():
    from jcfszxc_unet_b200 import ops

    return ops


def _nhwc_slice(n, h, w, c, dev, pad_lo=0, pad_hi=0, fill=None):
    """An NHWC bf16 tensor that is a channel slice of a wider buffer (ld = pad_lo + c + pad_hi)."""
    buf = torch.zeros(n, h, w, pad_lo + c + pad_hi, device=dev, dtype=torch.bfloat16)
    if fill is not None:
        buf.fill_(fill)
    return buf, buf[..., pad_lo:pad_lo + c]


def _close(got, ref, what):
    got, ref = got.float(), ref.float()
    scale = ref.abs().max().item() + 1e-6
    err = (got - ref).abs().max().item()
    assert err <= 1.2e-2 * scale, f"{what}: max err {err:.4g} vs scale {scale:.4g}"


CONV_SHAPES = [
    # n, h, w, cin, cout, x_pad(lo,hi), y_pad(lo,hi)
    (2, 32, 32, 64, 64, (0, 0), (0, 0)),       # BN=64, TW=32
    (1, 64, 64, 128, 256, (0, 0), (0, 0)),     # BN=256, multi k-chunk
    (2, 16, 16, 256, 128, (64, 0), (0, 128)),  # BN=128, sliced in/out (concat-buffer views)
    (1, 8, 256, 64, 64, (0, 0), (64, 0)),      # TW=128, two w tiles
    (1, 20, 24, 72, 72, (0, 0), (0, 0)),       # ragged: H,W not tile multiples; C not a multiple of 64
    (3, 5, 7, 8, 8, (0, 0), (0, 0)),           # tiny
    (1, 32, 32, 512, 320, (0, 0), (0, 0)),     # BN=256 with a masked N tail (320 = 256 + 64)
    # BN = 256 runs on CTA pairs (cta_group::2, two M tiles per M = 256 MMA): odd M-tile counts leave the second CTA of the
    # last pair with a tile past the end (all zero fill, nothing stored, nothing counted in the statistics)
    (1, 24, 16, 64, 256, (0, 0), (0, 0)),      # 3 M tiles
    (3, 8, 16, 256, 512, (0, 0), (0, 64)),     # 3 M tiles (one per image) x 2 N tiles, sliced output
    # W >= 128 and <= 128 output channels -> the halo kernel (conv3x3_halo.cu): two rows per tile, shifted descriptors
    (1, 5, 200, 64, 64, (0, 0), (0, 0)),       # odd H (last tile has one live row), ragged W, resident weights
    (2, 6, 128, 136, 72, (8, 0), (0, 56)),     # K tail chunk (136 = 2*64 + 8), 72 output channels, sliced in/out
    (1, 4, 384, 128, 128, (0, 0), (128, 0)),   # BN=128, three column tiles, weight ring (not resident)
    (1, 3, 256, 64, 128, (64, 0), (0, 0)),     # 64 -> 128 (dgrad of it is 128 -> 64)
    (1, 5, 128, 72, 96, (0, 0), (0, 32)),      # halo kernel on CTA pairs (BN = 128): three tiles, the last pair's second CTA idle
    # <= 64 output channels, K <= 128, W >= 128 -> the row-stacked kernel (conv3x3_rows.cu): 4 output rows per tile,
    # one input row multiplied by up to three filter rows in one N = 192 MMA
    (1, 9, 130, 64, 64, (0, 0), (0, 0)),       # H not a multiple of 4 (last tile: one live row), ragged W
    (2, 4, 128, 128, 64, (0, 0), (0, 64)),     # two k-chunks (3-slot row ring, single staging buffer), sliced output
    (1, 7, 300, 72, 40, (8, 0), (0, 0)),       # K tail chunk, 40 output channels (weight rows zero-filled), 3 column tiles
    (1, 2, 128, 8, 8, (0, 0), (0, 0)),         # smallest channel counts, H < 4
    (3, 16, 128, 64, 64, (0, 0), (0, 0)),      # several tiles per CTA: accumulator double-buffering and ring wrap-around
    # <= 32 output channels -> the row-stacked kernel with 32-wide accumulator blocks (N = 96), 64-byte staging rows
    (1, 9, 130, 96, 32, (0, 0), (0, 0)),       # K = 96: second chunk has two 16-channel steps
    (2, 4, 128, 192, 32, (0, 0), (0, 32)),     # UNet++ 192 -> 32: three k-chunks, sliced output
    (1, 6, 256, 32, 32, (32, 0), (0, 0)),      # 32 -> 32 (fwd and dgrad both take this path), sliced input
    (1, 5, 140, 160, 24, (0, 0), (0, 8)),      # 24 output channels (weight rows zero-filled), K tail, ragged W
    (3, 16, 128, 32, 32, (0, 0), (0, 0)),      # several tiles per CTA
    (1, 8, 128, 256, 16, (0, 0), (0, 0)),      # four k-chunks: 4-slot row ring with a single staging buffer
    # 129..192 output channels -> one 192-wide N tile; K not a multiple of 64 -> the last chunk issues fewer MMA steps
    (1, 24, 40, 192, 32, (0, 0), (0, 0)),      # UNet++ 192 -> 32 (its dgrad is 32 -> 192: N = 192, two 16-channel steps)
    (2, 12, 20, 160, 48, (32, 0), (0, 16)),    # 160 = 2.5 chunks; dgrad: N = 160 in a 192 tile, K = 48 (three steps)
    (1, 16, 16, 32, 176, (0, 0), (0, 0)),      # forward with N = 176 (masked tail of the 192 tile), K = 32
    # W >= 128, > 128 output channels, K <= 64 (the dgrad of UNet++'s concat-fed convs): column slices of <= 128 channels
    # through the halo / row-stacked kernels (conv_gemm_run, split_wide_thin)
    (1, 6, 128, 160, 32, (0, 0), (0, 0)),      # dgrad 32 -> 160 = 128 + 32
    (2, 4, 256, 192, 32, (32, 0), (0, 32)),    # dgrad 32 -> 192 = 128 + 64 into a sliced gradient buffer
    (1, 5, 140, 320, 64, (0, 0), (0, 0)),      # dgrad 64 -> 320 = 128 + 128 + 64, ragged W
    (1, 4, 128, 32, 176, (0, 0), (0, 16)),     # forward 32 -> 176 = 128 + 48 (no fused statistics)
]


@pytest.mark.parametrize("n,h,w,cin,cout,xpad,ypad", CONV_SHAPES)
def test_conv3x3_fwd(n, h, w, cin, cout, xpad, ypad):
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1234 + cin + cout)
    xbuf, x = _nhwc_slice(n, h, w, cin, dev, *xpad, fill=7.0)
    x.copy_(torch.randn(n, h, w, cin, device=dev, generator=g))
    wt = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * (1.0 / (3 * cin ** 0.5))
    bias = torch.randn(cout, device=dev, generator=g)
    ybuf, y = _nhwc_slice(n, h, w, cout, dev, *ypad, fill=-3.0)
    w_pack, _ = ops.pack_weight(wt, True, False)
    for b in (None, bias):
        ops.conv_fwd(x, w_pack, b, y, 3)
        torch.cuda.synchronize()
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.bfloat16().float(), b, padding=1).permute(0, 2, 3, 1)
        _close(y, ref, f"conv3x3_fwd bias={b is not None}")
    # neighbouring channels of the wider buffer must be untouched
    if ypad != (0, 0):
        lo, hi = ypad
        assert (ybuf[..., :lo] == -3.0).all() and (ybuf[..., lo + cout:] == -3.0).all()


@pytest.mark.parametrize("n,h,w,cin,cout,xpad,ypad", CONV_SHAPES)
def test_conv3x3_fwd_fused_bn_statistics(n, h, w, cin, cout, xpad, ypad):
    """The conv epilogue's per-channel (sum, sum of squares) must equal what unetk_bn_stats computes from the
    stored bf16 output (same values, different summation order), including ragged tiles and channel tails."""
    from jcfszxc_unet_b200 import _lib

    ops = _ops()
    dev = torch.device("cuda:0")
    lib = _lib.load()
    g = torch.Generator(device=dev).manual_seed(4321 + cin + cout)
    _, x = _nhwc_slice(n, h, w, cin, dev, *xpad)
    x.copy_(torch.randn(n, h, w, cin, device=dev, generator=g))
    wt = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * (1.0 / (3 * cin ** 0.5))
    bias = torch.randn(cout, device=dev, generator=g)
    _, y = _nhwc_slice(n, h, w, cout, dev, *ypad)
    w_pack, _ = ops.pack_weight(wt, True, False)
    partial = torch.empty(max(lib.unetk_conv_stats_partial_floats(cout), lib.unetk_chan_partial_floats(n * h * w, cout), 4096), device=dev)
    sums = torch.zeros(2 * cout, dtype=torch.float64, device=dev)
    xp, xld = ops.nhwc(x)
    yp, yld = ops.nhwc(y)
    stream = torch.cuda.current_stream().cuda_stream
    _lib.call("unetk_conv3x3_fwd_bnstats", xp, xld, w_pack.data_ptr(), bias.data_ptr(), yp, yld, partial.data_ptr(),
              sums.data_ptr(), n, h, w, cin, cout, stream)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.bfloat16().float(), bias, padding=1).permute(0, 2, 3, 1)
    _close(y, ref, "conv3x3_fwd_bnstats output")
    yf = y.float().double()
    s_ref = torch.cat([yf.sum(dim=(0, 1, 2)), (yf * yf).sum(dim=(0, 1, 2))])
    assert torch.allclose(sums, s_ref, rtol=1e-4, atol=1e-3 * float(s_ref.abs().max())), (sums - s_ref).abs().max()
    sums2 = torch.zeros_like(sums)
    ops.bn_stats(y, partial, sums2)
    assert torch.allclose(sums, sums2, rtol=1e-4, atol=1e-3 * float(s_ref.abs().max()))


@pytest.mark.parametrize("n,h,w,cin,cout,xpad,ypad", CONV_SHAPES)
def test_conv3x3_dgrad(n, h, w, cin, cout, xpad, ypad):
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(99 + cin + cout)
    _, dy = _nhwc_slice(n, h, w, cout, dev, *ypad)
    dy.copy_(torch.randn(n, h, w, cout, device=dev, generator=g))
    wt = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * (1.0 / (3 * cout ** 0.5))
    _, dx = _nhwc_slice(n, h, w, cin, dev, *xpad)
    _, w_pack_t = ops.pack_weight(wt, False, True)
    ops.conv_dgrad(dy, w_pack_t, dx, 3)
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), wt.bfloat16().float(), padding=1).permute(0, 2, 3, 1)
    _close(dx, ref, "conv3x3_dgrad")


@pytest.mark.parametrize("n,h,w,parts,cout", [(1, 6, 128, [32, 32, 64], 32), (2, 5, 40, [64, 128], 64), (1, 4, 256, [32, 32, 32, 32, 64], 32),
                                                 (1, 3, 130, [24, 40], 16)])
def test_conv3x3_dgrad_column_slices(n, h, w, parts, cout):
    """unetk_conv3x3_dgrad_cols: every member of a concat input gets its own columns of the dgrad, written (or bf16
    reduce-added) straight into that member's gradient view: equal to the matching slice of the whole dgrad."""
    ops = _ops()
    dev = torch.device("cuda:0")
    cin = sum(parts)
    g = torch.Generator(device=dev).manual_seed(cin + cout + w)
    dy = torch.randn(n, h, w, cout, device=dev, generator=g).to(torch.bfloat16)
    wt = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * (1.0 / (3 * cout ** 0.5))
    _, w_pack_t = ops.pack_weight(wt, False, True)
    whole = torch.empty(n, h, w, cin, device=dev, dtype=torch.bfloat16)
    ops.conv_dgrad(dy, w_pack_t, whole, 3)
    ref = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), wt.bfloat16().float(), padding=1).permute(0, 2, 3, 1)
    _close(whole, ref, "conv3x3_dgrad")
    c0 = 0
    for k, c in enumerate(parts):
        buf = torch.zeros(n, h, w, c + 24, device=dev, dtype=torch.bfloat16)     # a slice of a wider gradient buffer
        view = buf[..., 8:8 + c]
        ops.conv_dgrad_cols(dy, w_pack_t, c0, view, accumulate=False)
        _close(view, ref[..., c0:c0 + c], f"dgrad_cols part {k}")
        assert float(buf[..., :8].abs().max()) == 0 and float(buf[..., 8 + c:].abs().max()) == 0
        base = torch.randn(n, h, w, c, device=dev, generator=g).to(torch.bfloat16)
        view.copy_(base)
        ops.conv_dgrad_cols(dy, w_pack_t, c0, view, accumulate=True)
        _close(view, base.float() + ref[..., c0:c0 + c], f"dgrad_cols accumulate part {k}")
        c0 += c


WGRAD_EXTRA = [
    # Cout <= 64 < Cin: the weight-gradient GEMM runs with swapped operands and reversed taps (wgrad3x3_run, flip)
    (1, 16, 32, 128, 64, (0, 0), (0, 0)),
    (2, 9, 20, 136, 40, (8, 0), (0, 24)),
    (1, 4, 128, 256, 64, (0, 0), (64, 0)),
    (1, 3, 70, 96, 32, (0, 0), (0, 0)),      # UNet++'s 96 -> 32: the 5 + 4 tap split with half-empty N atoms, ragged W
]


@pytest.mark.parametrize("n,h,w,cin,cout,xpad,ypad", CONV_SHAPES + WGRAD_EXTRA)
def test_conv3x3_wgrad(n, h, w, cin, cout, xpad, ypad):
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(7 + cin + cout)
    _, x = _nhwc_slice(n, h, w, cin, dev, *xpad)
    x.copy_(torch.randn(n, h, w, cin, device=dev, generator=g))
    _, dy = _nhwc_slice(n, h, w, cout, dev, *ypad)
    dy.copy_(torch.randn(n, h, w, cout, device=dev, generator=g))
    dw = torch.full((cout, cin, 3, 3), float("nan"), device=dev)
    ops.conv_wgrad(x, dy, dw, 3)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (cout, cin, 3, 3),
                                      dy.float().permute(0, 3, 1, 2), padding=1)
    scale = ref.abs().max().item()
    err = (dw - ref).abs().max().item()
    assert err <= 2e-3 * scale, f"conv3x3_wgrad: err {err:.4g} scale {scale:.4g}"
    # accumulate mode adds on top (shared-weight recurrences, unet_parts.py:125-132)
    ops.conv_wgrad(x, dy, dw, 3, accumulate=True)
    torch.cuda.synchronize()
    err2 = (dw - 2 * ref).abs().max().item()
    assert err2 <= 4e-3 * scale


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 128, 64), (1, 32, 32, 1024, 512), (1, 8, 8, 64, 128)])
def test_conv1x1(n, h, w, cin, cout):
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn(n, h, w, cin, device=dev, generator=g).bfloat16()
    wt = torch.randn(cout, cin, 1, 1, device=dev, generator=g) / cin ** 0.5
    bias = torch.randn(cout, device=dev, generator=g)
    y = torch.empty(n, h, w, cout, device=dev, dtype=torch.bfloat16)
    w_pack, w_pack_t = ops.pack_weight(wt)
    ops.conv_fwd(x, w_pack, bias, y, 1)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.bfloat16().float(), bias).permute(0, 2, 3, 1)
    _close(y, ref, "conv1x1_fwd")
    dy = torch.randn(n, h, w, cout, device=dev, generator=g).bfloat16()
    dx = torch.empty_like(x)
    ops.conv_dgrad(dy, w_pack_t, dx, 1)
    ref = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), wt.bfloat16().float()).permute(0, 2, 3, 1)
    _close(dx, ref, "conv1x1_dgrad")
    dw = torch.empty(cout, cin, 1, 1, device=dev)
    ops.conv_wgrad(x, dy, dw, 1)
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (cout, cin, 1, 1), dy.float().permute(0, 3, 1, 2))
    assert (dw - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("n,h,w,cin,cout,ypad", [(2, 16, 16, 128, 64, (64, 0)), (1, 32, 32, 1024, 512, (0, 0)),
                                                 (1, 6, 10, 64, 64, (0, 0)), (1, 64, 64, 256, 128, (128, 0))])
def test_convT2x2(n, h, w, cin, cout, ypad):
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(11)
    x = torch.randn(n, h, w, cin, device=dev, generator=g).bfloat16()
    wt = torch.randn(cin, cout, 2, 2, device=dev, generator=g) / cin ** 0.5
    bias = torch.randn(cout, device=dev, generator=g)
    ybuf, y = _nhwc_slice(n, 2 * h, 2 * w, cout, dev, *ypad, fill=5.0)
    w_dgrad, w_fwd = ops.pack_weight(wt)  # [4,Cin,Cout] (dgrad), [4,Cout,Cin] (fwd)
    ops.convT_fwd(x, w_fwd, bias, y)
    ref = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), wt.bfloat16().float(), bias, stride=2).permute(0, 2, 3, 1)
    _close(y, ref, "convT_fwd")
    if ypad[0]:
        assert (ybuf[..., :ypad[0]] == 5.0).all()
    _, dy = _nhwc_slice(n, 2 * h, 2 * w, cout, dev, *ypad)
    dy.copy_(torch.randn(n, 2 * h, 2 * w, cout, device=dev, generator=g))
    dx = torch.empty_like(x)
    ops.convT_dgrad(dy, w_dgrad, dx)
    ref = F.conv2d(dy.float().permute(0, 3, 1, 2), wt.bfloat16().float(), stride=2).permute(0, 2, 3, 1)
    _close(dx, ref, "convT_dgrad")
    dw = torch.empty(cin, cout, 2, 2, device=dev)
    ops.convT_wgrad(x, dy, dw)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(False)
    wr = wt.clone().requires_grad_(True)
    F.conv_transpose2d(xr, wr, None, stride=2).backward(dy.float().permute(0, 3, 1, 2))
    assert (dw - wr.grad).abs().max().item() <= 2e-3 * wr.grad.abs().max().item()


@pytest.mark.parametrize("n,h,w,cin,cout,xpad,ypad", CONV_SHAPES[:4] + CONV_SHAPES[7:9])
def test_conv3x3_dgrad_colsum(n, h, w, cin, cout, xpad, ypad):
    """dgrad whose epilogue also returns the per-channel sums of dx (the ConvTranspose bias gradient when dx is the
    gradient of a concat buffer): same dx as the plain dgrad, sums equal to the column sums of the stored bf16 dx."""
    from jcfszxc_unet_b200 import _lib

    ops = _ops()
    dev = torch.device("cuda:0")
    lib = _lib.load()
    g = torch.Generator(device=dev).manual_seed(7 + cin + cout)
    _, dy = _nhwc_slice(n, h, w, cout, dev, *ypad)
    dy.copy_(torch.randn(n, h, w, cout, device=dev, generator=g))
    wt = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * (1.0 / (3 * cout ** 0.5))
    _, dx = _nhwc_slice(n, h, w, cin, dev, *xpad)
    _, dx_ref = _nhwc_slice(n, h, w, cin, dev, *xpad)
    _, w_pack_t = ops.pack_weight(wt, False, True)
    partial = torch.empty(max(lib.unetk_conv_stats_partial_floats(cin), 4096), device=dev)
    sums = torch.zeros(2 * cin, dtype=torch.float64, device=dev)
    ops.conv_dgrad_colsum(dy, w_pack_t, dx, partial, sums)
    ops.conv_dgrad(dy, w_pack_t, dx_ref, 3)
    torch.cuda.synchronize()
    assert torch.equal(dx, dx_ref)
    col = dx.float().double().sum(dim=(0, 1, 2))
    assert torch.allclose(sums[:cin], col, rtol=1e-5, atol=1e-4 * float(col.abs().max() + 1))
    out = torch.full((cin // 2,), 3.0, device=dev)
    ops.sums_to_f32(sums, cin // 2, cin // 2, out, accumulate=True)
    torch.cuda.synchronize()
    assert torch.allclose(out, 3.0 + col[cin // 2:].float(), rtol=1e-5, atol=1e-4 * float(col.abs().max() + 1))


# ---- fp32 mode: six-term bf16 split on the tensor core (csrc/f32path.cu) against float64 ------------------------
@pytest.mark.parametrize("n,h,w,cin,cout", [(1, 32, 32, 64, 64), (2, 16, 24, 128, 256), (1, 20, 12, 72, 40),
                                             (1, 8, 256, 64, 64), (1, 16, 16, 1024, 512)])
def test_f32_conv3x3(n, h, w, cin, cout):
    from jcfszxc_unet_b200 import f32

    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(31 + cin + cout)
    x = torch.randn(n, h, w, cin, device=dev, generator=g)
    wt = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * (1.0 / (3 * cin ** 0.5))
    bias = torch.randn(cout, device=dev, generator=g)
    y = f32.conv_f32(x, wt, bias)
    torch.cuda.synchronize()
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), wt.double(), bias.double(), padding=1).permute(0, 2, 3, 1)
    err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
    ref32 = F.conv2d(x.permute(0, 3, 1, 2), wt, bias, padding=1).permute(0, 2, 3, 1)
    err32 = (ref32.double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"f32 conv3x3 {cin}->{cout}: ours vs fp64 {err:.3g}; torch fp32 (cuDNN) vs fp64 {err32:.3g}")
    # the tensor core's fp32 accumulator truncates: ~3e-6 relative at K = 9*6*64, 3.6e-5 at K = 9*6*1024 (torch fp32: <1e-6); whole-network logits stay within 2.2e-5 (tests/test_gpu_unet.py), inside the 1e-4 budget
    assert err <= 5e-5, err


@pytest.mark.parametrize("n,h,w,cin,cout", [(1, 16, 16, 128, 64), (2, 8, 8, 1024, 512)])
def test_f32_convT2x2(n, h, w, cin, cout):
    from jcfszxc_unet_b200 import f32

    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(77 + cin + cout)
    x = torch.randn(n, h, w, cin, device=dev, generator=g)
    wt = torch.randn(cin, cout, 2, 2, device=dev, generator=g) * (1.0 / cin ** 0.5)
    bias = torch.randn(cout, device=dev, generator=g)
    y = f32.conv_f32(x, wt, bias, transposed=True)
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(x.double().permute(0, 3, 1, 2), wt.double(), bias.double(), stride=2).permute(0, 2, 3, 1)
    err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"f32 convT2x2 {cin}->{cout}: ours vs fp64 {err:.3g}")
    assert err <= 5e-5, err


def test_f32_split_is_exact_to_24_bits():
    from jcfszxc_unet_b200 import f32

    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn(1, 8, 8, 16, device=dev, generator=g) * torch.logspace(-6, 6, 16, device=dev)
    s = f32.split_activation(x).float()
    c = 16
    hi, mid, lo = s[..., 0:c], s[..., 3 * c:4 * c], s[..., 5 * c:6 * c]
    assert torch.equal(s[..., c:2 * c], hi) and torch.equal(s[..., 2 * c:3 * c], hi) and torch.equal(s[..., 4 * c:5 * c], mid)
    rec = hi.double() + mid.double() + lo.double()
    assert ((rec - x.double()).abs() <= x.abs().double() * 2.0 ** -23).all()


def test_sm_limit_keeps_results_and_scratch_sizes_valid():
    """unetk_set_sm_limit: persistent kernels launched by this thread use fewer SMs (a data-parallel trainer leaves some to
    NCCL).  Results stay correct, scratch sizes queried BEFORE the limit was set stay sufficient (the queries answer for
    every limit down to device_sms - 16), and the limit is clamped to that range."""
    from jcfszxc_unet_b200 import _lib

    ops = _ops()
    lib = _lib.load()
    dev = torch.device("cuda:0")
    real = lib.unetk_device_sms()
    assert real > 32
    n, h, w, cin, cout = 4, 64, 128, 128, 64
    g = torch.Generator(device=dev).manual_seed(77)
    x = torch.randn(n, h, w, cin, device=dev, generator=g).bfloat16()
    dy = torch.randn(n, h, w, cout, device=dev, generator=g).bfloat16()
    wt = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * 0.03
    w_ab, _ = ops.pack_weight(wt, True, False)
    ws = torch.empty(lib.unetk_conv_wgrad_workspace(n, h, w, cin, cout, 9), dtype=torch.uint8, device=dev)   # sized at full width
    partial = torch.empty(max(lib.unetk_conv_stats_partial_floats(cout), 4096), device=dev)
    ref_y = F.conv2d(x.float().permute(0, 3, 1, 2), wt.bfloat16().float(), None, padding=1).permute(0, 2, 3, 1)
    ref_dw = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (cout, cin, 3, 3), dy.float().permute(0, 3, 1, 2), padding=1)
    try:
        for limit in (0, real - 4, real - 16, real - 13, 7):
            prev = lib.unetk_set_sm_limit(limit)
            assert prev >= 0
            y = torch.empty(n, h, w, cout, device=dev, dtype=torch.bfloat16)
            sums = torch.zeros(2 * cout, dtype=torch.float64, device=dev)
            ops.conv_fwd_stats(x, w_ab, None, y, partial, sums, 3, 1)
            _close(y, ref_y, f"conv fwd under sm limit {limit}")
            assert torch.allclose(sums[:cout], y.double().sum(dim=(0, 1, 2)), rtol=1e-4, atol=1e-2)
            dw = torch.full((cout, cin, 3, 3), float("nan"), device=dev)
            ops.conv_wgrad(x, dy, dw, 3, ws=ws)
            torch.cuda.synchronize()
            assert (dw - ref_dw).abs().max().item() <= 2e-3 * ref_dw.abs().max().item()
        assert lib.unetk_set_sm_limit(0) == real - 16          # the request for 7 SMs was clamped to device_sms - 16
    finally:
        lib.unetk_set_sm_limit(0)


UPCONV_SHAPES = [
    # n, h, w (LOW resolution), cin, cout, ypad(lo,hi)
    (2, 8, 8, 64, 32, (0, 0)),        # the block test's shape; BN = 64, four phases
    (1, 32, 32, 1024, 512, (0, 0)),   # Up5 of the models: BN = 256, 16 k-chunks per tap
    (2, 16, 16, 128, 64, (64, 0)),    # output = upper slice of a concat buffer
    (1, 5, 7, 16, 8, (0, 0)),         # odd sizes: every border of the 2x2 windows, ragged tiles
    (1, 3, 130, 64, 64, (0, 64)),     # TW = 128, two column tiles (the second one ragged), sliced output
    (1, 64, 64, 256, 128, (0, 0)),    # BN = 128
    (3, 2, 2, 24, 40, (0, 0)),        # deepest level of a 32 x 32 image; channel counts that are not multiples of 64
    (1, 4, 128, 256, 128, (0, 0)),    # low-resolution W >= 128: the four phases on the halo kernel (conv3x3_halo.cu, TAPS = 4), BN = 128, weight ring
    (2, 3, 256, 128, 64, (64, 0)),    # the same with BN = 64, resident weights over two k-chunks, odd H, sliced output
    (1, 6, 20, 72, 40, (0, 0)),       # halo weight-gradient kernel (W >= 16), Cout <= 64 mode: ragged W, ragged channels
    (1, 16, 16, 192, 136, (0, 8)),    # the same, Cout > 64 mode: two m tiles (the second half empty), ragged second n tile
]


@pytest.mark.parametrize("n,h,w,cin,cout,ypad", UPCONV_SHAPES)
def test_upconv3x3_subpixel(n, h, w, cin, cout, ypad):
    """up_conv's nn.Upsample(scale_factor=2) -> nn.Conv2d(k=3,p=1) (unet_parts.py:103-104) in sub-pixel form on the
    low-resolution tensor: packs bit-exact against the oracle's restatement of the identity, forward (+ fused BatchNorm
    sums), input gradient (plain and accumulating) and the folded 3x3 weight gradient against torch on the up-sampled
    tensor."""
    from jcfszxc_unet_b200 import _lib
    from oracle import numpy_ops as NO

    ops = _ops()
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(3 + cin + cout + h)
    x = torch.randn(n, h, w, cin, device=dev, generator=g).bfloat16()
    wt = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * (1.0 / (3 * cin ** 0.5))
    bias = torch.randn(cout, device=dev, generator=g)
    w_fwd, w_dg = ops.pack_upconv_weight(wt)
    # packs: the fp32 sums of the taps that share a low-resolution pixel, rounded to bf16 once
    wq = NO.subpixel_weights(wt.double().cpu().numpy())            # [qy,qx,u,v,Cout,Cin] in float64
    wq32 = torch.zeros(2, 2, 2, 2, cout, cin, device=dev)
    for qy in range(2):
        for qx in range(2):
            for u in range(2):
                for v in range(2):
                    acc = torch.zeros(cout, cin, device=dev)
                    for kh in NO._subpixel_rows(qy, u):             # same order as the kernel: kh outer, kw inner
                        for kw in NO._subpixel_rows(qx, v):
                            acc = acc + wt[:, :, kh, kw]
                    wq32[qy, qx, u, v] = acc
    assert (wq32.double().cpu() - torch.from_numpy(wq)).abs().max().item() <= 1e-6
    exp_fwd = wq32.permute(2, 3, 0, 1, 4, 5).reshape(4, 4, cout, cin).bfloat16()          # [t4][q][Cout][Cin]
    exp_dg = wq32.reshape(16, cout, cin).transpose(1, 2).contiguous().bfloat16()          # [q*4+t4][Cin][Cout]
    assert torch.equal(w_fwd, exp_fwd) and torch.equal(w_dg, exp_dg)

    xr = x.float().permute(0, 3, 1, 2)
    up = F.interpolate(xr, scale_factor=2, mode="nearest")
    ybuf, y = _nhwc_slice(n, 2 * h, 2 * w, cout, dev, *ypad, fill=5.0)
    partial = torch.empty(max(lib.unetk_conv_stats_partial_floats(cout), 4096), device=dev)
    sums = torch.zeros(2 * cout, dtype=torch.float64, device=dev)
    ops.upconv_fwd(x, w_fwd, bias, y, partial, sums)
    # like-for-like reference: the four phase convs with the SAME bf16 sub-filters, fp32 accumulation
    refq = torch.zeros(n, cout, 2 * h, 2 * w, device=dev)
    xp = F.pad(xr, (1, 1, 1, 1))
    wb = exp_fwd.float().reshape(2, 2, 2, 2, cout, cin)   # [u][v][qy][qx]
    for qy in range(2):
        for qx in range(2):
            for u in range(2):
                for v in range(2):
                    patch = xp[:, :, qy + u:qy + u + h, qx + v:qx + v + w]
                    refq[:, :, qy::2, qx::2] += torch.einsum("nchw,oc->nohw", patch, wb[u, v, qy, qx])
    refq += bias.view(1, -1, 1, 1)
    _close(y, refq.permute(0, 2, 3, 1), "upconv_fwd vs the same bf16 sub-filters")
    # and the reference's own formulation (bf16 3x3 weights on the up-sampled tensor): one extra weight rounding apart
    ref33 = F.conv2d(up, wt.bfloat16().float(), bias, padding=1).permute(0, 2, 3, 1)
    _close(y, ref33, "upconv_fwd vs conv3x3(nearest2x(x))")
    if ypad[0]:
        assert (ybuf[..., :ypad[0]] == 5.0).all()
    if ypad[1]:
        assert (ybuf[..., ypad[0] + cout:] == 5.0).all()
    yd = y.float().double()
    s_ref = torch.stack([yd.sum(dim=(0, 1, 2)), (yd * yd).sum(dim=(0, 1, 2))]).reshape(-1)
    assert (sums - s_ref).abs().max().item() <= 1e-3 * max(1.0, s_ref.abs().max().item())
    y2buf, y2 = _nhwc_slice(n, 2 * h, 2 * w, cout, dev, *ypad)
    ops.upconv_fwd(x, w_fwd, bias, y2)
    assert torch.equal(y2, y)

    # eval-mode fold: relu(conv * scale + shift)
    scale = torch.rand(cout, device=dev, generator=g) + 0.5
    shift = torch.randn(cout, device=dev, generator=g)
    y3 = torch.empty(n, 2 * h, 2 * w, cout, device=dev, dtype=torch.bfloat16)
    ops.upconv_fwd_affine(x, w_fwd, scale, shift, True, y3)
    ref3 = torch.relu((refq - bias.view(1, -1, 1, 1)) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)).permute(0, 2, 3, 1)
    _close(y3, ref3, "upconv_fwd_affine")

    _, dy = _nhwc_slice(n, 2 * h, 2 * w, cout, dev, *ypad)
    dy.copy_(torch.randn(n, 2 * h, 2 * w, cout, device=dev, generator=g))
    dyr = dy.float().permute(0, 3, 1, 2)
    upr = up.clone().requires_grad_(True)
    wr = wt.bfloat16().float().requires_grad_(True)
    F.conv2d(upr, wr, None, padding=1).backward(dyr)
    dx_ref = (upr.grad[:, :, 0::2, 0::2] + upr.grad[:, :, 0::2, 1::2] + upr.grad[:, :, 1::2, 0::2] + upr.grad[:, :, 1::2, 1::2])
    dx = torch.full_like(x, 7.0)
    ops.upconv_dgrad(dy, w_dg, dx)
    _close(dx, dx_ref.permute(0, 2, 3, 1), "upconv_dgrad")
    base = torch.randn(n, h, w, cin, device=dev, generator=g).bfloat16()
    dx2 = base.clone()
    ops.upconv_dgrad(dy, w_dg, dx2, accumulate=True)
    _close(dx2, (base.float() + dx.float()), "upconv_dgrad accumulate")

    # weight gradient: mathematically the same sum as the 3x3 conv's on the up-sampled tensor (operands are bf16-exact)
    dw = torch.full((cout, cin, 3, 3), 9.0, device=dev)
    ops.upconv_wgrad(x, dy, dw)
    dw_ref = torch.nn.grad.conv2d_weight(up.double(), (cout, cin, 3, 3), dyr.double(), padding=1).float()
    assert (dw - dw_ref).abs().max().item() <= 2e-3 * dw_ref.abs().max().item()
    dw2 = dw.clone()
    ops.upconv_wgrad(x, dy, dw2, accumulate=True)
    assert (dw2 - 2 * dw).abs().max().item() <= 1e-5 * dw.abs().max().item()


def test_conv3x3_wgrad_cta_pairs_subprocess():
    """The rows2 weight gradient launched as clusters of two CTAs with TMA-multicast operands (UNETK_WGRAD3_CLUSTER=1, read
    once per process => a child process): same results as the default launch on the three shapes that take the 5 + 4 tap
    split (cluster barrier, multicast loads, multicast commit onto both CTAs' empty barriers)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import torch, sys
sys.path.insert(0, %r)
from jcfszxc_unet_b200 import ops
dev = torch.device("cuda:0")
for (n, h, w, cin, cout) in [(1, 4, 128, 256, 64), (1, 3, 70, 96, 32), (2, 64, 512, 128, 64)]:
    g = torch.Generator(device=dev).manual_seed(cin + cout)
    x = torch.randn(n, h, w, cin, device=dev, generator=g).bfloat16()
    dy = torch.randn(n, h, w, cout, device=dev, generator=g).bfloat16()
    dw = torch.full((cout, cin, 3, 3), float("nan"), device=dev)
    ops.conv_wgrad(x, dy, dw, 3)
    ref = torch.nn.grad.conv2d_weight(x.double().permute(0, 3, 1, 2), (cout, cin, 3, 3), dy.double().permute(0, 3, 1, 2), padding=1)
    err = (dw.double() - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item(), (n, h, w, cin, cout, err)
print("pairs ok")
''' % root
    env = dict(os.environ, UNETK_WGRAD3_CLUSTER="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "pairs ok" in r.stdout, r.stdout + r.stderr
