"""Vanilla U-Net — drop-in for the reference's UNetFamily/UNet.py:14-55 (class path, ctor signature,
`.n_channels/.n_classes`, state_dict keys and default init identical), executed as ONE fused plan of
hand-written sm_100a kernels (jcfszxc_unet_b200.engine.build_unet_plan).
"""
from __future__ import annotations

import torch.nn as nn

from jcfszxc_unet_b200 import bridge as _bridge
from jcfszxc_unet_b200 import engine as _engine
from UNetFamily.utils.unet_parts import DoubleConv, Down, OutConv, Up

_WIDTHS = (64, 128, 256, 512, 1024)


class UNet(nn.Module):
    def __init__(self, n_channels=3, n_classes=1, bilinear=False):
        super().__init__()
        if bilinear:
            # BASELINE.json's north_star names a `bilinear` argument; the reference removed it
            # (UNet.py:30).  Accepted for signature compatibility, only the transposed-conv path exists.
            raise NotImplementedError("bilinear up-sampling was removed from the reference UNet (UNet.py:30)")
        self.n_channels = n_channels
        self.n_classes = n_classes
        w = _WIDTHS
        self.inc = DoubleConv(n_channels, w[0])
        for i in range(1, 5):
            setattr(self, f"down{i}", Down(w[i - 1], w[i]))
        for i in range(1, 5):
            setattr(self, f"up{i}", Up(w[5 - i], w[4 - i]))
        self.outc = OutConv(w[0], n_classes)

    def forward(self, x):
        """[N, n_channels, H, W] float image (any strides) -> fp32 logits [N, n_classes, H, W]."""
        return _bridge.run_model(self, _engine.build_unet_plan, x)
