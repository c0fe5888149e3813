"""DySample with the reference's constructor signature and state_dict layout (DySample_md.py:20-81); the offset
projection + point-sampling upsample run in libkmunet.so."""
import torch
import torch.nn as nn

from .. import ops


class DySample(nn.Module):
    def __init__(self, in_channels, scale=2, style='lp', groups=4, dyscope=False):
        super().__init__()
        self.scale, self.style, self.groups = scale, style, groups
        assert style in ['lp', 'pl']
        if style == 'pl':
            assert in_channels >= scale ** 2 and in_channels % scale ** 2 == 0
        assert in_channels >= groups and in_channels % groups == 0
        if style == 'pl':
            in_channels = in_channels // scale ** 2
            out_channels = 2 * groups
        else:
            out_channels = 2 * groups * scale ** 2
        self.offset = nn.Conv2d(in_channels, out_channels, 1)
        nn.init.normal_(self.offset.weight, 0, 0.001)
        nn.init.constant_(self.offset.bias, 0)
        if dyscope:
            self.scope = nn.Conv2d(in_channels, out_channels, 1, bias=False)
            nn.init.constant_(self.scope.weight, 0.)
        self.register_buffer('init_pos', self._init_pos())

    def _init_pos(self):
        s = self.scale
        h = (torch.arange(s, dtype=torch.float32) - (s - 1) / 2) / s
        pos = torch.empty(2, self.groups, s, s)
        pos[0] = h.view(1, 1, s)      # x start depends on the sub-pixel column j
        pos[1] = h.view(1, s, 1)      # y start depends on the sub-pixel row i
        return pos.reshape(1, -1, 1, 1)

    def sample(self, x, offset):
        return ops.dysample_sample(x, offset, self.scale, self.groups)

    def forward(self, x):
        # KM-UNet builds DySample(in_channels, scale=2, style='lp') without dyscope (KM_UNetV3_SH.py:411,425,434); the other styles
        # of the reference class (DySample_md.py:63-76) are never executed by it and are not provided
        if self.style != 'lp' or hasattr(self, 'scope'):
            raise NotImplementedError("km_unet_b200.DySample implements style='lp' without dyscope (the configuration KM-UNet uses)")
        return ops.dysample(x, self.offset.weight, self.offset.bias, self.init_pos, self.scale, self.groups)
