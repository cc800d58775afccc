"""R2 Attention U-Net — drop-in for the reference's UNetFamily/R2AttentionUNet.py:14-91: R2UNet's recurrent
residual blocks with AttentionUNet's gated skips (jcfszxc_unet_b200.builders.build_r2attention_unet_plan).
"""
from __future__ import annotations

import torch.nn as nn

from jcfszxc_unet_b200 import bridge as _bridge
from jcfszxc_unet_b200 import builders as _builders
from UNetFamily.utils.unet_parts import Attention_block, RRCNN_block, up_conv


class R2AttentionUNet(nn.Module):
    def __init__(self, img_ch=3, output_ch=1, t=2):
        super().__init__()
        self.n_channels = img_ch
        self.n_classes = output_ch
        self.Maxpool = nn.MaxPool2d(kernel_size=2, stride=2)
        self.Upsample = nn.Upsample(scale_factor=2)
        w = (64, 128, 256, 512, 1024)
        self.RRCNN1 = RRCNN_block(ch_in=img_ch, ch_out=w[0], t=t)
        for i in range(1, 5):
            setattr(self, f"RRCNN{i + 1}", RRCNN_block(ch_in=w[i - 1], ch_out=w[i], t=t))
        for i in (5, 4, 3, 2):   # registration order of R2AttentionUNet.py:30-44
            c = w[i - 2]
            setattr(self, f"Up{i}", up_conv(ch_in=2 * c, ch_out=c))
            setattr(self, f"Att{i}", Attention_block(F_g=c, F_l=c, F_int=c // 2))
            setattr(self, f"Up_RRCNN{i}", RRCNN_block(ch_in=2 * c, ch_out=c, t=t))
        self.Conv_1x1 = nn.Conv2d(w[0], output_ch, kernel_size=1, stride=1, padding=0)

    def forward(self, x):
        return _bridge.run_model(self, _builders.build_r2attention_unet_plan, x)
