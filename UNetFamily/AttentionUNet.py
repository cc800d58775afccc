"""Attention U-Net — drop-in for the reference's UNetFamily/AttentionUNet.py:15-84 (class path, ctor signature,
`.n_channels/.n_classes`, 240 state_dict keys and default init identical), executed as ONE fused plan of
hand-written sm_100a kernels (jcfszxc_unet_b200.builders.build_attention_unet_plan): tensor-core convs with
the BatchNorm statistics in their epilogue and the attention gates as fused kernels (csrc/gate.cu).
"""
from __future__ import annotations

import torch.nn as nn

from jcfszxc_unet_b200 import bridge as _bridge
from jcfszxc_unet_b200 import builders as _builders
from UNetFamily.utils.unet_parts import Attention_block, conv_block, up_conv


class AttentionUNet(nn.Module):
    def __init__(self, img_ch=3, output_ch=1):
        super().__init__()
        self.n_channels = img_ch
        self.n_classes = output_ch
        self.Maxpool = nn.MaxPool2d(kernel_size=2, stride=2)
        w = (64, 128, 256, 512, 1024)
        self.Conv1 = conv_block(ch_in=img_ch, ch_out=w[0])
        for i in range(1, 5):
            setattr(self, f"Conv{i + 1}", conv_block(ch_in=w[i - 1], ch_out=w[i]))
        for i in (5, 4, 3, 2):   # same registration (= initialisation) order as AttentionUNet.py:28-42
            c = w[i - 2]
            setattr(self, f"Up{i}", up_conv(ch_in=2 * c, ch_out=c))
            setattr(self, f"Att{i}", Attention_block(F_g=c, F_l=c, F_int=c // 2))
            setattr(self, f"Up_conv{i}", conv_block(ch_in=2 * c, ch_out=c))
        self.Conv_1x1 = nn.Conv2d(w[0], output_ch, kernel_size=1, stride=1, padding=0)

    def forward(self, x):
        """[N, img_ch, H, W] float image (any strides) -> fp32 logits [N, output_ch, H, W]."""
        return _bridge.run_model(self, _builders.build_attention_unet_plan, x)
