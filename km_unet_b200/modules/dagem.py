"""DAGEM with the reference's constructor signature and state_dict layout (DAGEM_md.py:7-111).

The edge / vertex gating MLPs, their five train-mode BatchNorms and the final 1x1 fusion run in libkmunet.so
(ops.dagem_gate), and so does the deformable 3x3 convolution (ops.deformconv3x3, csrc/deform.cu): torchvision's CUDA op
launches on the legacy default stream, so it is silently dropped from a CUDA-graph capture, and its backward scatters with
atomics.  `deform_conv` stays a torchvision DeformConv2d module for its parameters / state_dict keys; it is only CALLED for
maps larger than the kernel takes (H*W > 4096).  The 3x3 offset convolution (DAGEM_md.py:45,98) is a plain cuDNN convolution.
"""
import torch.nn as nn
from torchvision.ops import DeformConv2d

from .. import ops
from .vim import count_batch


class DAGEM(nn.Module):
    def __init__(self, sync_bn=False, input_channels=256):
        super().__init__()
        self.input_channels = c = input_channels
        self.edge_aggregation_func = nn.Sequential(nn.Linear(4, 1), nn.BatchNorm1d(1), nn.ReLU(inplace=True))
        self.vertex_update_func = nn.Sequential(nn.Linear(2 * c, c // 2), nn.BatchNorm1d(c // 2), nn.ReLU(inplace=True))
        self.edge_update_func = nn.Sequential(nn.Linear(2 * c, c // 2), nn.BatchNorm1d(c // 2), nn.ReLU(inplace=True))
        self.update_edge_reduce_func = nn.Sequential(nn.Linear(4, 1), nn.BatchNorm1d(1), nn.ReLU(inplace=True))
        self.offset_conv = nn.Conv2d(c, 18, kernel_size=3, padding=1)
        self.deform_conv = DeformConv2d(c, c, kernel_size=3, padding=1)
        self.final_aggregation_layer = nn.Sequential(nn.Conv2d(c + c // 2, c, kernel_size=1, stride=1, padding=0, bias=False),
                                                     nn.BatchNorm2d(c), nn.ReLU(inplace=True))

    def _bns(self):
        return [self.edge_aggregation_func[1], self.edge_update_func[1], self.vertex_update_func[1],
                self.update_edge_reduce_func[1], self.final_aggregation_layer[1]]

    def forward(self, input):
        x = input
        offset = self.offset_conv(x)
        if ops.deformconv3x3_supported(x, self.deform_conv.weight):
            deformed = ops.deformconv3x3(x, offset, self.deform_conv.weight, self.deform_conv.bias) + x
        else:
            deformed = self.deform_conv(x, offset) + x
        bns = self._bns()
        training = self.training or any(bn.running_mean is None for bn in bns)
        lin = (self.edge_aggregation_func[0].weight, self.edge_aggregation_func[0].bias,
               self.vertex_update_func[0].weight, self.vertex_update_func[0].bias,
               self.edge_update_func[0].weight, self.edge_update_func[0].bias,
               self.update_edge_reduce_func[0].weight, self.update_edge_reduce_func[0].bias,
               self.final_aggregation_layer[0].weight)
        out = ops.dagem_gate(x, deformed, lin, [(bn.weight, bn.bias, bn.running_mean, bn.running_var) for bn in bns], training,
                             bns[0].momentum, bns[0].eps)
        if self.training:
            for bn in bns:
                count_batch(bn)
        return out
