"""Drop-in building blocks of the U-Net family, backed by the B200-native kernels.

Same class names, constructor signatures, sub-module attribute names and therefore the same
`state_dict` keys and default initialisation as the reference's UNetFamily/utils/unet_parts.py
(DoubleConv :17-34, Down :37-47, Up :50-70, OutConv :73-79), so checkpoints and pickled models
interchange.  The torch.nn sub-modules here are PARAMETER CONTAINERS only: no torch kernel runs in
forward().  A model (UNetFamily.UNet.UNet) executes as one fused plan; a block used on its own runs a
small plan of the same ops (jcfszxc_unet_b200.blocks).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from jcfszxc_unet_b200 import blocks as _blocks


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


__all__ = ["DoubleConv", "Down", "Up", "OutConv", "torch", "nn"]
