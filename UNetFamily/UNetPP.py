"""UNet++ (NestedUNet) — drop-in for the reference's UNetFamily/UNetPP.py:15-107 (class paths
`UNetFamily.UNetPP.NestedUNet` / `.DoubleConv`, ctor signature, 212 state_dict keys and default init identical;
deepsupervision defaults to the reference's hard-coded False; True is a superset keyword).  The output is post-sigmoid (:105-106).  One fused plan
(jcfszxc_unet_b200.builders.build_nested_unet_plan): every node is produced straight into the concat buffer of
its first consumer and the bilinear up-sampling writes into the consumer's concat slice.
"""
from __future__ import annotations

import torch.nn as nn

from jcfszxc_unet_b200 import blocks as _blocks
from jcfszxc_unet_b200 import bridge as _bridge
from jcfszxc_unet_b200 import builders as _builders


class DoubleConv(nn.Module):
    """[conv3x3 (bias) -> BatchNorm -> ReLU] twice (UNetPP.py:15-28; not unet_parts.DoubleConv)."""

    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(in_ch, out_ch, 3, padding=1), nn.BatchNorm2d(out_ch), nn.ReLU(inplace=True),
            nn.Conv2d(out_ch, out_ch, 3, padding=1), nn.BatchNorm2d(out_ch), nn.ReLU(inplace=True))

    def forward(self, input):
        return _blocks.run_emit(self, "pp_double_conv", [input], lambda P, a: _builders.emit_conv_pair(P, a[0], self.conv))


class NestedUNet(nn.Module):
    def __init__(self, in_channel=3, out_channel=1, deepsupervision=False):
        """deepsupervision is a SUPERSET keyword: the reference hard-codes `self.deepsupervision = False` (UNetPP.py:38) but
        carries the deep-supervision branch (:65-69, 93-102) that BASELINE.json's configs[4] names.  True builds the four
        1x1 heads final1..final4 (same registration order, hence the same initialisation draws as the reference class
        with the attribute forced on) and forward returns [output1..output4], each post-sigmoid."""
        super().__init__()
        self.n_channels = in_channel
        self.n_classes = out_channel
        self.deepsupervision = bool(deepsupervision)
        nb = [32, 64, 128, 256, 512]
        self.pool = nn.MaxPool2d(2, 2)
        self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        # registration (= initialisation) order of UNetPP.py:46-65
        self.conv0_0 = DoubleConv(in_channel, nb[0])
        for i in range(1, 5):
            setattr(self, f"conv{i}_0", DoubleConv(nb[i - 1], nb[i]))
        for j in range(1, 5):
            for i in range(0, 5 - j):
                setattr(self, f"conv{i}_{j}", DoubleConv(nb[i] * j + nb[i + 1], nb[i]))
        self.sigmoid = nn.Sigmoid()
        if self.deepsupervision:
            for k in range(1, 5):
                setattr(self, f"final{k}", nn.Conv2d(nb[0], out_channel, kernel_size=1))
        else:
            self.final = nn.Conv2d(nb[0], out_channel, kernel_size=1)

    def forward(self, input):
        """[N, in_channel, H, W] -> fp32 probabilities [N, out_channel, H, W] (post-sigmoid, as the reference); with
        deepsupervision a list of four such tensors (UNetPP.py:93-102)."""
        return _bridge.run_model(self, _builders.build_nested_unet_plan, input)
