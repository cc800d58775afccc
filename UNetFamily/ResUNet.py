"""ResUNet — drop-in for the reference's UNetFamily/ResUNet.py:15-76 (class path, ctor signature, 145 state_dict
keys and default init identical).  The output is post-sigmoid like the reference's (output_layer ends in
nn.Sigmoid, :47-50).  One fused plan (jcfszxc_unet_b200.builders.build_resunet_plan): stride-2 convolutions on
the tcgen05 tap-GEMM through strided TMA boxes, pre-activation BatchNorm as a streaming pass, residual adds on
the BatchNorm pass of the skip branch.
"""
from __future__ import annotations

import torch.nn as nn

from jcfszxc_unet_b200 import bridge as _bridge
from jcfszxc_unet_b200 import builders as _builders
from UNetFamily.utils.unet_parts import ResidualConv, Upsample


class ResUNet(nn.Module):
    def __init__(self, channel=3, out_channels=1):
        super().__init__()
        self.n_channels = channel
        self.n_classes = out_channels
        self.bilinear = False
        self.input_layer = nn.Sequential(
            nn.Conv2d(channel, 64, kernel_size=3, padding=1), nn.BatchNorm2d(64), nn.ReLU(),
            nn.Conv2d(64, 64, kernel_size=3, padding=1))
        self.input_skip = nn.Sequential(nn.Conv2d(channel, 64, kernel_size=3, padding=1))
        self.residual_conv_1 = ResidualConv(64, 128, 2, 1)
        self.residual_conv_2 = ResidualConv(128, 256, 2, 1)
        self.bridge = ResidualConv(256, 512, 2, 1)
        self.upsample_1 = Upsample(512, 512, 2, 2)
        self.up_residual_conv1 = ResidualConv(512 + 256, 256, 1, 1)
        self.upsample_2 = Upsample(256, 256, 2, 2)
        self.up_residual_conv2 = ResidualConv(128 + 256, 128, 1, 1)
        self.upsample_3 = Upsample(128, 128, 2, 2)
        self.up_residual_conv3 = ResidualConv(128 + 64, 64, 1, 1)
        self.output_layer = nn.Sequential(nn.Conv2d(64, out_channels, 1, 1), nn.Sigmoid())

    def forward(self, x):
        """[N, channel, H, W] -> fp32 probabilities [N, out_channels, H, W] (post-sigmoid, as the reference)."""
        return _bridge.run_model(self, _builders.build_resunet_plan, x)
