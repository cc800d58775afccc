"""Drop-in building blocks of the U-Net family, backed by the B200-native kernels.

Same class names, constructor signatures, sub-module attribute names and therefore the same
`state_dict` keys and default initialisation as the reference's UNetFamily/utils/unet_parts.py
(DoubleConv :17-34, Down :37-47, Up :50-70, OutConv :73-79, conv_block :82-96, up_conv :99-111,
Recurrent_block :114-132, RRCNN_block :135-146, Attention_block :149-176, ResidualConv :454-475,
Upsample :478-487), so checkpoints and pickled models interchange.  The torch.nn sub-modules here are PARAMETER CONTAINERS only: no torch kernel runs in
forward().  A model (UNetFamily.UNet.UNet) executes as one fused plan; a block used on its own runs a
small plan of the same ops (jcfszxc_unet_b200.blocks).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from jcfszxc_unet_b200 import blocks as _blocks
from jcfszxc_unet_b200 import builders as _builders
from jcfszxc_unet_b200 import engine as _engine


class DoubleConv(nn.Module):
    """[conv3x3 (no bias) -> BatchNorm -> ReLU] twice."""

    def __init__(self, in_channels, out_channels, mid_channels=None):
        super().__init__()
        mid = mid_channels if mid_channels else out_channels
        layers = []
        for cin, cout in ((in_channels, mid), (mid, out_channels)):
            layers += [nn.Conv2d(cin, cout, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(cout),
                       nn.ReLU(inplace=True)]
        self.double_conv = nn.Sequential(*layers)

    def forward(self, x):
        return _blocks.run_double_conv(self, x)


class Down(nn.Module):
    """2x2 max-pool, then DoubleConv."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels))

    def forward(self, x):
        return _blocks.run_down(self, x)


class Up(nn.Module):
    """ConvTranspose2d(k=2, s=2) on x1, concat [x2, up(x1)] on channels, DoubleConv."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.up = nn.ConvTranspose2d(in_channels, in_channels // 2, kernel_size=2, stride=2)
        self.conv = DoubleConv(in_channels, out_channels)

    def forward(self, x1, x2):
        return _blocks.run_up(self, x1, x2)


class OutConv(nn.Module):
    """1x1 convolution to the class logits."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1)

    def forward(self, x):
        return _blocks.run_out_conv(self, x)


class conv_block(nn.Module):
    """[conv3x3 (bias) -> BatchNorm -> ReLU] twice (AttentionUNet encoder/decoder)."""

    def __init__(self, ch_in, ch_out):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(ch_in, ch_out, kernel_size=3, stride=1, padding=1, bias=True), nn.BatchNorm2d(ch_out),
            nn.ReLU(inplace=True),
            nn.Conv2d(ch_out, ch_out, kernel_size=3, stride=1, padding=1, bias=True), nn.BatchNorm2d(ch_out),
            nn.ReLU(inplace=True))

    def forward(self, x):
        return _blocks.run_emit(self, "conv_block", [x], lambda P, a: _builders.emit_conv_pair(P, a[0], self.conv))


class up_conv(nn.Module):
    """nearest 2x up-sampling -> conv3x3 (bias) -> BatchNorm -> ReLU."""

    def __init__(self, ch_in, ch_out):
        super().__init__()
        self.up = nn.Sequential(
            nn.Upsample(scale_factor=2),
            nn.Conv2d(ch_in, ch_out, kernel_size=3, stride=1, padding=1, bias=True), nn.BatchNorm2d(ch_out),
            nn.ReLU(inplace=True))

    def forward(self, x):
        return _blocks.run_emit(self, "up_conv", [x], lambda P, a: _builders.emit_up_conv(P, a[0], self))


class Recurrent_block(nn.Module):
    """x1 = f(x); t times x1 = f(x + x1), one shared f = conv3x3 (bias) -> BatchNorm -> ReLU."""

    def __init__(self, ch_out, t=2):
        super().__init__()
        self.t = t
        self.ch_out = ch_out
        self.conv = nn.Sequential(
            nn.Conv2d(ch_out, ch_out, kernel_size=3, stride=1, padding=1, bias=True), nn.BatchNorm2d(ch_out),
            nn.ReLU(inplace=True))

    def forward(self, x):
        return _blocks.run_emit(self, ("recurrent", self.t), [x], lambda P, a: _builders.emit_recurrent(P, a[0], self))


class RRCNN_block(nn.Module):
    """x = Conv_1x1(x); return x + RCNN(x) with RCNN = two Recurrent_blocks."""

    def __init__(self, ch_in, ch_out, t=2):
        super().__init__()
        self.RCNN = nn.Sequential(Recurrent_block(ch_out, t=t), Recurrent_block(ch_out, t=t))
        self.Conv_1x1 = nn.Conv2d(ch_in, ch_out, kernel_size=1, stride=1, padding=0)

    def forward(self, x):
        return _blocks.run_emit(self, ("rrcnn", self.RCNN[0].t), [x], lambda P, a: _builders.emit_rrcnn(P, a[0], self))


class Attention_block(nn.Module):
    """out = x * sigmoid(BN(psi(relu(BN(W_g g) + BN(W_x x)))))."""

    def __init__(self, F_g, F_l, F_int):
        super().__init__()
        self.W_g = nn.Sequential(nn.Conv2d(F_g, F_int, kernel_size=1, stride=1, padding=0, bias=True),
                                 nn.BatchNorm2d(F_int))
        self.W_x = nn.Sequential(nn.Conv2d(F_l, F_int, kernel_size=1, stride=1, padding=0, bias=True),
                                 nn.BatchNorm2d(F_int))
        self.psi = nn.Sequential(nn.Conv2d(F_int, 1, kernel_size=1, stride=1, padding=0, bias=True), nn.BatchNorm2d(1),
                                 nn.Sigmoid())
        self.relu = nn.ReLU(inplace=True)

    def forward(self, g, x):
        def emit(P, a):
            out = P.act(a[1].H, a[1].W, a[1].C)
            _engine.AttentionGate(P, a[0], a[1], self, out)
            return out

        return _blocks.run_emit(self, "attention", [g, x], emit)


class ResidualConv(nn.Module):
    """(BN -> ReLU -> conv3x3(stride) -> BN -> ReLU -> conv3x3) + (conv3x3(stride) -> BN)."""

    def __init__(self, input_dim, output_dim, stride, padding):
        super().__init__()
        self.conv_block = nn.Sequential(
            nn.BatchNorm2d(input_dim), nn.ReLU(),
            nn.Conv2d(input_dim, output_dim, kernel_size=3, stride=stride, padding=padding),
            nn.BatchNorm2d(output_dim), nn.ReLU(),
            nn.Conv2d(output_dim, output_dim, kernel_size=3, padding=1))
        self.conv_skip = nn.Sequential(
            nn.Conv2d(input_dim, output_dim, kernel_size=3, stride=stride, padding=1), nn.BatchNorm2d(output_dim))

    def forward(self, x):
        return _blocks.run_emit(self, "residual_conv", [x], lambda P, a: _builders.emit_residual_conv(P, a[0], self))


class Upsample(nn.Module):
    """ConvTranspose2d wrapper (kernel 2, stride 2 on this path)."""

    def __init__(self, input_dim, output_dim, kernel, stride):
        super().__init__()
        self.upsample = nn.ConvTranspose2d(input_dim, output_dim, kernel_size=kernel, stride=stride)

    def forward(self, x):
        def emit(P, a):
            out = P.act(2 * a[0].H, 2 * a[0].W, self.upsample.out_channels)
            _engine.ConvT2x2(P, a[0], self.upsample, out)
            return out

        return _blocks.run_emit(self, "upsample", [x], emit)


__all__ = ["DoubleConv", "Down", "Up", "OutConv", "conv_block", "up_conv", "Recurrent_block", "RRCNN_block",
           "Attention_block", "ResidualConv", "Upsample", "torch", "nn"]
