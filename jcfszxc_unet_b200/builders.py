"""Execution plans of the U-Net variants (AttentionUNet, R2UNet, R2AttentionUNet, ResUNet, NestedUNet):
each reference `forward` (cited per builder) wired once into the fused ops of engine.py over static NHWC
buffers.  Like the vanilla plan, every `torch.cat` of the reference is a pair of channel slices of one
buffer, the 2x2 max-pool is fused into the producing BatchNorm pass where the producer allows it, and
tensors with several consumers get their gradients accumulated in the consumers' epilogues.
"""
from __future__ import annotations

import os

from .engine import (Act, AddN, AttentionGate, BNAct, ConvBNReLU, ConvT2x2, Head, Image, MaxPool2x2, Plan,
                     Upsample2x, _require)


def _hw(P: Plan, x):
    return (P.H, P.W) if isinstance(x, Image) else (x.H, x.W)


# ------------------------------------------------------------------------------------------------ blocks
def _producer(P: Plan, y: Act):
    """The unit that was emitted last, if it is the conv+BN(+ReLU) producing `y` (then the head is y's only consumer
    and can take over that unit's BatchNorm passes, engine.Head)."""
    op = P.ops[-1] if P.ops else None
    return op if isinstance(op, ConvBNReLU) and op.out is y else None


def emit_conv_pair(P: Plan, x, seq, out: Act | None = None, pooled: Act | None = None) -> Act:
    """seq = [conv3x3, BN, ReLU, conv3x3, BN, ReLU]: conv_block (unet_parts.py:82-96), DoubleConv (:17-34),
    NestedUNet's DoubleConv (UNetPP.py:15-28)."""
    h, w = _hw(P, x)
    mid = P.act(h, w, seq[0].out_channels)
    ConvBNReLU(P, x, seq[0], seq[1], mid)
    out = out if out is not None else P.act(h, w, seq[3].out_channels)
    ConvBNReLU(P, mid, seq[3], seq[4], out, pooled)
    return out


def emit_up_conv(P: Plan, x: Act, mod, out: Act | None = None) -> Act:
    """up_conv.forward (unet_parts.py:99-111): nearest 2x -> conv3x3(bias) -> BN -> ReLU.  Default: ONE unit in sub-pixel
    form on the low-resolution tensor (ConvBNReLU(up=True): 2.25x fewer FLOPs, no up-sampled tensor);
    UNETK_UPCONV_SUBPIXEL=0 materialises the up-sampled tensor and runs the 3x3 conv on it."""
    conv = mod.up[1]
    if os.environ.get("UNETK_UPCONV_SUBPIXEL", "1") != "0" and conv.in_channels % 8 == 0 and conv.out_channels % 8 == 0:
        out = out if out is not None else P.act(2 * x.H, 2 * x.W, conv.out_channels)
        ConvBNReLU(P, x, conv, mod.up[2], out, up=True)
        return out
    up = P.act(2 * x.H, 2 * x.W, x.C)
    Upsample2x(P, x, up, "nearest")
    out = out if out is not None else P.act(up.H, up.W, mod.up[1].out_channels)
    ConvBNReLU(P, up, mod.up[1], mod.up[2], out)
    return out


def emit_recurrent(P: Plan, x: Act, mod, out: Act | None = None, res: Act | None = None) -> Act:
    """Recurrent_block.forward (unet_parts.py:124-132): x1 = f(x); then t times x1 = f(x + x1), with ONE shared
    f = conv3x3(bias)+BN+ReLU.  The `x + x1` adds ride on the BatchNorm pass of the f that produced x1; `res`
    (optional) is added to the final f (the `x + x1` of RRCNN_block, :146)."""
    conv, bn = mod.conv[0], mod.conv[1]
    cur = x
    for i in range(mod.t + 1):
        last = i == mod.t
        dst = (out if out is not None else P.act(x.H, x.W, x.C)) if last else P.act(x.H, x.W, x.C)
        ConvBNReLU(P, cur, conv, bn, dst, res=res if last else x)
        cur = dst
    return cur


def emit_rrcnn(P: Plan, x, mod, out: Act | None = None) -> Act:
    """RRCNN_block.forward (unet_parts.py:143-146): x = Conv_1x1(x); return x + RCNN(x)."""
    h, w = _hw(P, x)
    c = mod.Conv_1x1.out_channels
    x0 = P.act(h, w, c)
    ConvBNReLU(P, x, mod.Conv_1x1, None, x0, relu=False)
    r1 = emit_recurrent(P, x0, mod.RCNN[0])
    return emit_recurrent(P, r1, mod.RCNN[1], out=out, res=x0)


def emit_residual_conv(P: Plan, x: Act, mod, out: Act | None = None) -> Act:
    """ResidualConv.forward (unet_parts.py:454-475): conv_block(x) + conv_skip(x) with
    conv_block = BN -> ReLU -> conv3x3(stride) -> BN -> ReLU -> conv3x3 and conv_skip = conv3x3(stride) -> BN."""
    cb, cs = mod.conv_block, mod.conv_skip
    stride = cb[2].stride[0]
    _require(cb[2].padding == (1, 1) and cs[0].padding == (1, 1), "ResidualConv: only padding=1 is on this path")
    h, w = x.H // stride, x.W // stride
    cout = cb[2].out_channels
    t = P.act(x.H, x.W, x.C)
    BNAct(P, x, cb[0], t, relu=True)
    u = P.act(h, w, cout)
    ConvBNReLU(P, t, cb[2], cb[3], u)
    out = out if out is not None else P.act(h, w, cout)
    v = Act(P.act(h, w, cout, grad=False).t, out.g)   # d(out)/d(v) = 1: v's gradient IS out's
    ConvBNReLU(P, u, cb[5], None, v, relu=False)
    ConvBNReLU(P, x, cs[0], cs[1], out, relu=False, res=v)
    return out


def _check16(H, W, who):
    _require(H % 16 == 0 and W % 16 == 0 and H >= 16 and W >= 16,
             f"{who} plan needs H, W divisible by 16 (got {H}x{W})")


# ------------------------------------------------------------------------------------------------ models
def build_attention_unet_plan(model, N, H, W, device, training, grad_views=None, with_grad=None) -> Plan:
    """AttentionUNet.forward (UNetFamily/AttentionUNet.py:46-84)."""
    _check16(H, W, "AttentionUNet")
    P = Plan(device, N, H, W, training, with_grad)
    enc = [model.Conv1, model.Conv2, model.Conv3, model.Conv4, model.Conv5]
    C = [m.conv[3].out_channels for m in enc]
    xs, x = [], P.image
    for i, m in enumerate(enc):
        h, w = H >> i, W >> i
        out = P.act(h, w, C[i])
        pooled = P.act(h >> 1, w >> 1, C[i]) if i < 4 else None
        emit_conv_pair(P, x, m.conv, out, pooled)
        xs.append(out)
        x = pooled
    y = xs[4]
    for lvl, (up, att, upc) in zip((3, 2, 1, 0), ((model.Up5, model.Att5, model.Up_conv5), (model.Up4, model.Att4, model.Up_conv4),
                                                  (model.Up3, model.Att3, model.Up_conv3), (model.Up2, model.Att2, model.Up_conv2))):
        cat = P.act(H >> lvl, W >> lvl, 2 * C[lvl])          # cat((x_gated, d), dim=1), AttentionUNet.py:66
        d = cat.slice(C[lvl], C[lvl])
        emit_up_conv(P, y, up, d)
        AttentionGate(P, d, xs[lvl], att, cat.slice(0, C[lvl]))
        y = emit_conv_pair(P, cat, upc.conv)
    P.head = Head(P, y, model.Conv_1x1, fuse=_producer(P, y))
    return P.finalize(grad_views)


def _build_r2(model, N, H, W, device, training, grad_views, with_grad, gated: bool) -> Plan:
    _check16(H, W, type(model).__name__)
    P = Plan(device, N, H, W, training, with_grad)
    enc = [model.RRCNN1, model.RRCNN2, model.RRCNN3, model.RRCNN4, model.RRCNN5]
    C = [m.Conv_1x1.out_channels for m in enc]
    cats = [P.act(H >> i, W >> i, 2 * C[i]) for i in range(4)]
    xs, x = [], P.image
    for i, m in enumerate(enc):
        h, w = H >> i, W >> i
        if i < 4:
            # without a gate the skip is the lower half of the concat buffer; with one, the gate writes that half
            out = P.act(h, w, C[i]) if gated else cats[i].slice(0, C[i])
        else:
            out = P.act(h, w, C[i])
        emit_rrcnn(P, x, m, out)
        xs.append(out)
        if i < 4:
            x = P.act(h >> 1, w >> 1, C[i])
            MaxPool2x2(P, out, x)
    y = xs[4]
    ups = (model.Up5, model.Up4, model.Up3, model.Up2)
    blocks = (model.Up_RRCNN5, model.Up_RRCNN4, model.Up_RRCNN3, model.Up_RRCNN2)
    atts = (model.Att5, model.Att4, model.Att3, model.Att2) if gated else (None,) * 4
    for lvl, up, att, blk in zip((3, 2, 1, 0), ups, atts, blocks):
        d = cats[lvl].slice(C[lvl], C[lvl])
        emit_up_conv(P, y, up, d)
        if gated:
            AttentionGate(P, d, xs[lvl], att, cats[lvl].slice(0, C[lvl]))
        y = emit_rrcnn(P, cats[lvl], blk)
    P.head = Head(P, y, model.Conv_1x1)
    return P.finalize(grad_views)


def build_r2unet_plan(model, N, H, W, device, training, grad_views=None, with_grad=None) -> Plan:
    """R2UNet.forward (UNetFamily/R2UNet.py:45-79)."""
    return _build_r2(model, N, H, W, device, training, grad_views, with_grad, gated=False)


def build_r2attention_unet_plan(model, N, H, W, device, training, grad_views=None, with_grad=None) -> Plan:
    """R2AttentionUNet.forward (UNetFamily/R2AttentionUNet.py:48-91): R2UNet with attention-gated skips."""
    return _build_r2(model, N, H, W, device, training, grad_views, with_grad, gated=True)


def build_resunet_plan(model, N, H, W, device, training, grad_views=None, with_grad=None) -> Plan:
    """ResUNet.forward (UNetFamily/ResUNet.py:52-76).  Concat order is [up-sampled, skip] (:61,66,71)."""
    _require(H % 8 == 0 and W % 8 == 0 and H >= 8 and W >= 8, f"ResUNet plan needs H, W divisible by 8 (got {H}x{W})")
    P = Plan(device, N, H, W, training, with_grad)
    il, isk = model.input_layer, model.input_skip
    c1 = il[0].out_channels
    rcs = (model.residual_conv_1, model.residual_conv_2, model.bridge)
    C = [c1] + [m.conv_block[2].out_channels for m in rcs]          # 64, 128, 256, 512
    ups = (model.upsample_1, model.upsample_2, model.upsample_3)
    cu = [u.upsample.out_channels for u in ups]                       # 512, 256, 128
    # cat buffers at levels 2, 1, 0: [up | skip]
    cats = {2: P.act(H >> 2, W >> 2, cu[0] + C[2]), 1: P.act(H >> 1, W >> 1, cu[1] + C[1]), 0: P.act(H, W, cu[2] + C[0])}
    skip = {2: cats[2].slice(cu[0], C[2]), 1: cats[1].slice(cu[1], C[1]), 0: cats[0].slice(cu[2], C[0])}
    # x1 = input_layer(x) + input_skip(x)
    m = P.act(H, W, c1)
    ConvBNReLU(P, P.image, il[0], il[1], m)
    x1 = skip[0]
    ya = Act(P.act(H, W, c1, grad=False).t, x1.g)
    ConvBNReLU(P, m, il[3], None, ya, relu=False)
    za = Act(P.act(H, W, c1, grad=False).t, x1.g)
    ConvBNReLU(P, P.image, isk[0], None, za, relu=False)
    AddN(P, [ya, za], x1)
    x2 = emit_residual_conv(P, x1, rcs[0], skip[1])
    x3 = emit_residual_conv(P, x2, rcs[1], skip[2])
    x4 = emit_residual_conv(P, x3, rcs[2])
    y = x4
    for lvl, up, rc in zip((2, 1, 0), ups, (model.up_residual_conv1, model.up_residual_conv2, model.up_residual_conv3)):
        ConvT2x2(P, y, up.upsample, cats[lvl].slice(0, cu[2 - lvl]))
        y = emit_residual_conv(P, cats[lvl], rc)
    P.head = Head(P, y, model.output_layer[0], post_sigmoid=True)
    return P.finalize(grad_views)


def build_nested_unet_plan(model, N, H, W, device, training, grad_views=None, with_grad=None) -> Plan:
    """NestedUNet.forward (UNetFamily/UNetPP.py:73-107); deepsupervision=False is the reference's fixed setting (:37),
    True adds the heads final1..final3 on X[0][1..3] (:93-100) next to final4 on X[0][4].

    Node X[i][j] is produced straight into the concat buffer of the first node that reads it (conv{i}_{j+1}) and
    copied into the later ones (the reference re-copies all of them in every torch.cat); the bilinear up-sampling
    writes into the last slice of its consumer's concat buffer."""
    _check16(H, W, "NestedUNet")
    P = Plan(device, N, H, W, training, with_grad)
    nb = [model.conv0_0.conv[0].out_channels, model.conv1_0.conv[0].out_channels, model.conv2_0.conv[0].out_channels,
          model.conv3_0.conv[0].out_channels, model.conv4_0.conv[0].out_channels]
    maxj = [4, 3, 2, 1, 0]
    cat = {}
    for i in range(4):
        for j in range(1, maxj[i] + 1):
            cat[i, j] = P.act(H >> i, W >> i, nb[i] * j + nb[i + 1])
    X = {}
    # UNETK_SCATTER_DGRAD=0: every consumer writes the gradient of its whole concat buffer and one add pass per copy
    # brings the members' slices home (the first version: 2.6 ms of adds per step at B = 16, 512^2)
    scatter = os.environ.get("UNETK_SCATTER_DGRAD", "1") != "0"

    def node_out(i, j):
        """where conv{i}_{j} writes: slice j of the concat buffer of conv{i}_{j+1}, or its own buffer"""
        if j < maxj[i]:
            return cat[i, j + 1].slice(nb[i] * j, nb[i])
        return P.act(H >> i, W >> i, nb[i])

    def publish(i, j):
        """copy X[i][j] into the concat buffers of conv{i}_{j+2}, ... (UNetPP.py:80,86,88,93,95,97)"""
        for jj in range(j + 2, maxj[i] + 1):
            AddN(P, [X[i, j]], cat[i, jj].slice(nb[i] * j, nb[i])).scattered = scatter

    def up_into(i, j):
        """self.up(X[i+1][j-1]) -> last slice of conv{i}_{j}'s concat buffer"""
        Upsample2x(P, X[i + 1, j - 1], cat[i, j].slice(nb[i] * j, nb[i + 1]), "bilinear")

    pooled = {}

    def backbone(i, x):
        out = node_out(i, 0)
        pooled[i] = P.act(H >> (i + 1), W >> (i + 1), nb[i]) if i < 4 else None
        emit_conv_pair(P, x, getattr(model, f"conv{i}_0").conv, out, pooled[i])
        X[i, 0] = out
        publish(i, 0)

    def nested(i, j):
        up_into(i, j)
        out = node_out(i, j)
        first = len(P.ops)
        emit_conv_pair(P, cat[i, j], getattr(model, f"conv{i}_{j}").conv, out)
        if scatter and P.with_grad:
            # the conv that reads the concat sends every member's gradient columns to that member's own gradient
            # (X[i][jj].g, wherever its home slice is) and the up-sampled part to its slice of this buffer
            conv1 = P.ops[first]
            assert conv1.x is cat[i, j]
            conv1.x_parts = [(nb[i] * jj, nb[i], X[i, jj]) for jj in range(j)]
            conv1.x_parts.append((nb[i] * j, nb[i + 1], cat[i, j].slice(nb[i] * j, nb[i + 1])))
        X[i, j] = out
        publish(i, j)

    # same evaluation order as UNetPP.py:74-99
    backbone(0, P.image)
    backbone(1, pooled[0])
    nested(0, 1)
    backbone(2, pooled[1])
    nested(1, 1)
    nested(0, 2)
    backbone(3, pooled[2])
    nested(2, 1)
    nested(1, 2)
    nested(0, 3)
    backbone(4, pooled[3])
    nested(3, 1)
    nested(2, 2)
    nested(1, 3)
    nested(0, 4)
    if getattr(model, "deepsupervision", False):
        last = _producer(P, X[0, 4])
        # X[0][1..3] also feed the later nodes: their heads keep a private gradient buffer that is added to the node's
        # gradient (Head(own_grad=True)); X[0][4] feeds only final4, whose head takes over the node's BatchNorm passes
        P.heads = [Head(P, X[0, k], getattr(model, f"final{k}"), post_sigmoid=True, own_grad=True) for k in (1, 2, 3)]
        P.heads.append(Head(P, X[0, 4], model.final4, post_sigmoid=True, fuse=last))
        P.head = P.heads[-1]
    else:
        P.head = Head(P, X[0, 4], model.final, post_sigmoid=True, fuse=_producer(P, X[0, 4]))
    return P.finalize(grad_views)
