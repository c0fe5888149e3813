"""Host-side mirror of the callers of the hot path: KM_UNetV3 (SH and LAPS variants) and the glue blocks around the
CUDA-backed operators, with the reference's module tree so a reference state_dict loads unchanged.

Reference: KM_UNetV3_SH.py:21-94 (StableHybridKANConv), :97-151 (EnhancedViMBlock), :154-212 (DirectionViM), :215-263
(DirectionAttention), :266-284 (TripleNorm), :287-332 (MultiScaleFusion / ChannelAttention), :336-368
(LocalContrastAttention), :371-517 (KM_UNetV3); KM_UNetV3_LAPS.py:411-437,483 (plain bilinear upsampling, no bridge);
WPL/iwp.py:9-132 (Haar wavelet pooling).  This file exists because the reference tree cannot travel to the GPU box: the
benchmark of the full training step (BASELINE configs[2..4]) needs the caller.  With the reference tree present,
`km_unet_b200.enable_dropin()` + the reference's own KM_UNetV3_SH.py gives the same network.

Differences that are deliberate: no self-imposed fp16 autocast (the reference decorates forward with
torch.cuda.amp.autocast; here the KAN convolutions pick their precision from km_unet_b200.config and the rest runs in
fp32), the dead parameters of the reference (StableHybridKANConv.branches / .attn, DirectionViM.dt_proj,
IWP.high_freq_conv) are kept for checkpoint compatibility but never executed, and the wavelet pooling does not rebuild
numpy matrices and upload them on every call.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import config, ops
from .dagem import DAGEM
from .dysample import DySample
from .kan import KANConv2d
from .vim import EfficientViMBlock, batch_counters, conv1x1, conv_same


def group_norm(m, x):
    """nn.GroupNorm forward through kmu_groupnorm_fwd on the GPU (CPU oracle runs keep the torch module)."""
    if ops.groupnorm_supported(x):
        return ops.groupnorm(x, m.weight, m.bias, m.num_groups, m.eps)
    return m(x)


def global_mean(x):
    """Mean over (H, W) of (B,C,H,W) -> (B,C), written as sum * (1 / HW): autograd's backward of `sum` is an expanded VIEW of
    the pooled gradient, whereas the backward of `mean` materialises a full-size tensor (one extra write + read of the feature
    map per call; 28 such poolings per KM_UNetV3 training step)."""
    return x.sum(dim=(2, 3)) * (1.0 / (x.shape[2] * x.shape[3]))


def _run(seq, x):
    """nn.Sequential forward with the plain convolutions / GroupNorms routed through the CUDA kernels where they apply."""
    for m in seq:
        x = conv_same(m, x) if type(m) is nn.Conv2d else (group_norm(m, x) if type(m) is nn.GroupNorm else m(x))
    return x


_SIDE, _LEVEL = {}, {}


def _side_streams(device, n, level):
    """n side streams tied to (current stream of `device`, nesting level) -- cached: the same streams every step, as graph capture
    wants; a nested parallel region on the same stream gets its own set so it does not queue behind an outer region's branches."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, level)
    have = _SIDE.setdefault(key, [])
    while len(have) < n:
        have.append(torch.cuda.Stream(device))
    return have[:n]


def parallel(fns, inputs, device):
    """Run independent callables concurrently: fns[0] on the current stream, the others on side streams (fork after `inputs` are
    ready, join before anyone consumes the results).  Inside a CUDA-graph capture these become parallel branches of the graph; the
    backward nodes run on the streams their forward ran on.  Falls back to a plain loop on CPU / when config.parallel_branches is off."""
    if not (device.type == "cuda" and config.parallel_branches) or len(fns) < 2:
        return [f() for f in fns]
    cur = torch.cuda.current_stream(device)
    outs = [None] * len(fns)
    lkey = (device.index, cur.cuda_stream)
    level = _LEVEL.get(lkey, 0)
    sides = _side_streams(device, len(fns) - 1, level)
    for i, s in enumerate(sides):
        s.wait_stream(cur)
        for t in inputs:
            t.record_stream(s)
        with torch.cuda.stream(s):
            outs[i + 1] = fns[i + 1]()
    _LEVEL[lkey] = level + 1
    try:
        outs[0] = fns[0]()
    finally:
        _LEVEL[lkey] = level
    for i, s in enumerate(sides):
        cur.wait_stream(s)
        outs[i + 1].record_stream(cur)
    return outs


class DropPath(nn.Module):
    """timm 0.9.16 DropPath: per-sample bernoulli(keep) / keep in training, identity otherwise."""

    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


class StableHybridKANConv(nn.Module):
    """ReLU(residual(GN4(x)) + KANConv2d(GN4(x)))."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1):
        super().__init__()
        self.branches = nn.ModuleDict({'plain': KANConv2d(in_channels, out_channels, kernel_size, padding=padding)})  # dead
        self.kanconv2d = nn.Sequential(KANConv2d(in_channels, out_channels, kernel_size, padding=padding))
        self.attn = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(in_channels, 1, 1), nn.Softmax(dim=1))          # dead
        self.pre_norm = nn.GroupNorm(4, in_channels)
        self.post_act = nn.ReLU(inplace=True)
        self.residual = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else nn.Identity()
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out')
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x):
        x = group_norm(self.pre_norm, x)
        identity = x if isinstance(self.residual, nn.Identity) else conv1x1(x, self.residual.weight, self.residual.bias)
        return self.post_act(identity + self.kanconv2d(x))


class DirectionAttention(nn.Module):
    """dwconv3x3(sigmoid(q k) v) scaled by a squeeze-excite weight of the globally pooled input."""

    def __init__(self, dim, mode):
        super().__init__()
        self.mode = mode
        self.qkv = nn.Conv2d(dim, dim * 3, 1)
        self.conv = nn.Conv2d(dim, dim, 3, padding=1, groups=dim)
        self.fc = nn.Sequential(nn.Linear(dim, dim // 4), nn.GELU(), nn.Linear(dim // 4, dim), nn.Sigmoid())

    def forward(self, x):
        # the three pooling modes of the reference (mean over W then H, over H then W, or both) are the same global mean
        weight = self.fc(global_mean(x))
        attn = ops.qkv_gate(conv1x1(x, self.qkv.weight, self.qkv.bias))          # sigmoid(q k) v
        return ops.dwconv3x3(attn, self.conv.weight, self.conv.bias, scale=weight)   # conv(attn) * weight[:, :, None, None]


class DirectionViM(nn.Module):
    def __init__(self, dim, mode='height', state_dim=64):
        super().__init__()
        self.mode, self.state_dim = mode, state_dim
        self.dt_proj = nn.Linear(dim, state_dim)                                                                     # dead
        self.vit_mamba = EfficientViMBlock(dim=dim, mlp_ratio=4, ssd_expand=1, state_dim=64)
        if mode == 'height':
            self.proj = nn.Conv2d(dim, dim, (3, 1), padding=(1, 0))
        elif mode == 'width':
            self.proj = nn.Conv2d(dim, dim, (1, 3), padding=(0, 1))
        else:
            self.proj = nn.Conv2d(dim, dim, 1)
        self.attn = DirectionAttention(dim, mode)

    def forward(self, x):
        return self.attn(self.vit_mamba(conv_same(self.proj, x)))


class TripleNorm(nn.Module):
    """(GN1(x; h) + GN1(x; w) + LayerNorm_C(x)) / 3 -- GroupNorm(1) statistics do not see the H/W permutation."""

    def __init__(self, dim):
        super().__init__()
        self.norm_h = nn.GroupNorm(num_groups=1, num_channels=dim)
        self.norm_w = nn.GroupNorm(num_groups=1, num_channels=dim)
        self.norm_c = nn.LayerNorm(dim)

    def forward(self, x):
        if ops.triplenorm_supported(x.shape[1]):
            return ops.triplenorm(x, self.norm_h.weight, self.norm_h.bias, self.norm_w.weight, self.norm_w.bias, self.norm_c.weight,
                                  self.norm_c.bias, self.norm_h.eps, self.norm_c.eps)
        c = self.norm_c(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
        return (self.norm_h(x) + self.norm_w(x) + c) / 3


class EnhancedViMBlock(nn.Module):
    def __init__(self, dim, expansion=4, state_dim=64, drop_path=0.1):
        super().__init__()
        self.dim, self.state_dim = dim, state_dim
        self.height_block = DirectionViM(dim, mode='height', state_dim=state_dim)
        self.width_block = DirectionViM(dim, mode='width', state_dim=state_dim)
        self.channel_block = DirectionViM(dim, mode='channel', state_dim=state_dim)
        self.fusion_gate = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(dim * 3, dim // 4, 1), nn.GELU(),
                                         nn.Conv2d(dim // 4, 3, 1), nn.Softmax(dim=1))
        self.ffn = nn.Sequential(nn.Conv2d(dim, dim * expansion, 1), nn.GELU(), nn.Conv2d(dim * expansion, dim, 1))
        self.norm = TripleNorm(dim)
        self.drop_path = DropPath(drop_path) if drop_path > 0 else nn.Identity()

    def forward(self, x):
        # the three direction branches are independent until the fusion gate
        feats = parallel([lambda: self.height_block(x), lambda: self.width_block(x), lambda: self.channel_block(x)], [x], x.device)
        # fusion_gate[0] is a global average pool: pool the three branches separately instead of concatenating the feature maps
        g = _run(self.fusion_gate[1:], torch.cat([global_mean(f) for f in feats], dim=1)[:, :, None, None])
        if ops.combine3_supported(x):
            # x + DropPath(g0 f0 + g1 f1 + g2 f2) in one pass: the per-sample DropPath factor is folded into the gate weights
            coef = g.reshape(g.shape[0], 3)
            if isinstance(self.drop_path, DropPath):
                coef = self.drop_path(coef)            # same bernoulli(keep) / keep per sample, applied to the (B,3) coefficients
            x = ops.combine3(x, feats[0], feats[1], feats[2], coef)
        else:
            x = x + self.drop_path(g[:, 0:1] * feats[0] + g[:, 1:2] * feats[1] + g[:, 2:3] * feats[2])
        h = F.gelu(conv1x1(self.norm(x), self.ffn[0].weight, self.ffn[0].bias))
        y = conv1x1(h, self.ffn[2].weight, self.ffn[2].bias)
        if isinstance(self.drop_path, DropPath) and self.training and self.drop_path.drop_prob > 0.0:
            # x + DropPath(y) in one pass: the per-sample factor (same draw as DropPath(y)) as the multiplier of an addcmul
            return torch.addcmul(x, y, self.drop_path(x.new_ones((x.shape[0], 1, 1, 1))))
        return x + y


class ChannelAttention(nn.Module):
    def __init__(self, channel, reduction=8):
        super().__init__()
        self.gap = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(nn.Linear(channel, channel // reduction), nn.SiLU(), nn.Linear(channel // reduction, channel),
                                nn.Sigmoid())

    def forward(self, x):
        return x * self.fc(global_mean(x))[:, :, None, None]


class MultiScaleFusion(nn.Module):
    def __init__(self, channels, reduction=4):
        super().__init__()
        out = channels[-1]
        self.blocks = nn.ModuleList([nn.Sequential(nn.Conv2d(c, out, s, padding=s // 2, stride=1), nn.GroupNorm(1, out), nn.SiLU())
                                     for c, s in zip(channels, [3, 5, 7])])
        self.fusion = nn.Sequential(nn.Conv2d(out * 3, out, 1), nn.Conv2d(out, out, 3, padding=1), ChannelAttention(out, reduction))

    def forward(self, features):
        branches = parallel([(lambda blk=blk, f=f: _run(blk, f)) for blk, f in zip(self.blocks, features)], list(features), features[0].device)
        return _run(self.fusion, torch.cat(branches, dim=1))


class LocalContrastAttention(nn.Module):
    def __init__(self, in_channels, reduction_ratio=4):
        super().__init__()
        self.reduction_ratio = reduction_ratio
        self.fc = nn.Sequential(nn.Linear(in_channels // reduction_ratio, 64), nn.ReLU(), nn.Linear(64, in_channels), nn.Sigmoid())

    def forward(self, x):
        avg = global_mean(x)
        g = self.fc(avg.view(avg.size(0), -1, self.reduction_ratio).mean(-1))[:, :, None, None]
        return torch.addcmul(g, x, 1 - g)           # x (1 - g) + g in one pass


class _NoParams(nn.Module):
    """Placeholder for the reference's parameter-free DWT_2D child (keeps the module tree / state_dict identical)."""


class IntelligentWaveletPoolingModule(nn.Module):
    """Haar 2x2 analysis + 1x1 fusion (WPL/iwp.py:116-132).  With s = 1/sqrt(2) taps the four sub-bands of a 2x2 block
    [[a,b],[c,d]] are LL=(a+b+c+d)/2, LH=(a-b+c-d)/2, HL=(a+b-c-d)/2, HH=(a-b-c+d)/2; the reference's get_matrix leaves
    the LAST high-pass row and column zero (:79-82) and its Softmax2d over a one-channel map is identically 1, so the output
    is fusion_conv(cat[LL, mean_c(LH, HL, HH)])."""

    def __init__(self, in_channels, wavename='haar'):
        super().__init__()
        if wavename != 'haar':
            raise NotImplementedError("only the Haar wavelet (the one KM-UNet uses) is implemented")
        self.dwt = _NoParams()
        self.high_freq_conv = nn.Conv2d(3 * in_channels, 1, kernel_size=(1, 1))                                      # inert
        self.softmax = nn.Softmax2d()
        self.fusion_conv = nn.Conv2d(in_channels + 1, in_channels, kernel_size=(1, 1))

    def forward(self, x):
        B, C, H, W = x.shape
        if H != W or H % 2:
            raise RuntimeError("IntelligentWaveletPoolingModule: square, even-sized maps only (as in KM-UNet)")
        if ops.iwp_supported(C, H, W):
            return ops.iwp(x, self.fusion_conv.weight, self.fusion_conv.bias)
        return conv1x1(torch.cat(haar_pool(x), dim=1), self.fusion_conv.weight, self.fusion_conv.bias)


def haar_pool(x):
    """(LL, mean over channels of the three high-pass bands) of WPL/iwp.py in strided-slice form."""
    B, C, H, W = x.shape
    a, b = x[:, :, 0::2, 0::2], x[:, :, 0::2, 1::2]
    c, d = x[:, :, 1::2, 0::2], x[:, :, 1::2, 1::2]
    ll = (a + b + c + d) * 0.5
    lh = (a - b + c - d) * 0.5
    hl = (a + b - c - d) * 0.5
    hh = (a - b - c + d) * 0.5
    row = torch.ones(H // 2, 1, device=x.device, dtype=x.dtype)
    row[-1] = 0
    col = torch.ones(1, W // 2, device=x.device, dtype=x.dtype)
    col[:, -1] = 0
    high = (lh * col + hl * row + hh * (row * col)).sum(dim=1, keepdim=True) / (3 * C)
    return ll, high


class KM_UNetV3(nn.Module):
    """variant='SH': DAGEM bridge + DySample decoders (KM_UNetV3_SH.py:371-517); variant='LAPS': no bridge, bilinear
    align_corners upsampling (KM_UNetV3_LAPS.py:411-437,483)."""

    def __init__(self, num_classes=3, embed_dims=(16, 32, 64), variant='SH'):
        super().__init__()
        if variant not in ('SH', 'LAPS'):
            raise ValueError(variant)
        self.variant = variant
        e0, e1, e2 = embed_dims
        self.conv_f = nn.Conv2d(5, 16, kernel_size=3, padding=1, stride=1)
        self.lca1, self.lca2, self.lca3 = (LocalContrastAttention(e) for e in (e0, e1, e2))
        self.enc1 = nn.Sequential(StableHybridKANConv(16, e0), EnhancedViMBlock(e0, state_dim=16), IntelligentWaveletPoolingModule(e0))
        self.enc2 = nn.Sequential(StableHybridKANConv(e0, e1), EnhancedViMBlock(e1, state_dim=16), IntelligentWaveletPoolingModule(e1))
        self.enc3 = nn.Sequential(StableHybridKANConv(e1, e2), EnhancedViMBlock(e2, state_dim=16), IntelligentWaveletPoolingModule(e2))

        def up():
            return DySample(e2, scale=2, style='lp') if variant == 'SH' else nn.Upsample(scale_factor=2, mode='bilinear',
                                                                                         align_corners=True)
        if variant == 'SH':
            self.bridge_attention = DAGEM(sync_bn=False, input_channels=e2)
        self.dec1 = nn.Sequential(up(), StableHybridKANConv(e2, e1))
        self.attention1 = nn.Sequential(MultiScaleFusion([e0, e1, e1]))
        self.attention2 = nn.Sequential(MultiScaleFusion([e0, e1, e1]))
        self.dec2 = nn.Sequential(up(), nn.Conv2d(e1 * 2, e1, kernel_size=3, padding=1, stride=1), EnhancedViMBlock(e1, state_dim=16))
        self.dec3 = nn.Sequential(up(), nn.Conv2d(e1 * 2, e0, 3, padding=1), EnhancedViMBlock(e0),
                                  nn.Conv2d(e0, num_classes, 3, padding=1))
        self.output_norm = nn.GroupNorm(1, num_classes)
        self.activation = nn.Sigmoid()

    @staticmethod
    def _skips(e1, e2, size):
        if e1.is_cuda:
            r1, r2 = ops.resize_bilinear_ac(e1, size), ops.resize_bilinear_ac(e2, size)
        else:                           # CPU oracle runs (oracle/model.py swaps the CUDA entry points, not this torch call)
            r1 = F.interpolate(e1, size=size, mode='bilinear', align_corners=True)
            r2 = F.interpolate(e2, size=size, mode='bilinear', align_corners=True)
        return [r1, r2, r2]            # the reference feeds e2 twice (KM_UNetV3_SH.py:495)

    def forward(self, x):
        with batch_counters():          # the 65 BatchNorm counters of a training step: one multi-tensor add at the end
            return self._forward(x)

    def _forward(self, x):
        x = self.conv_f(x)
        e1 = self.lca1(self.enc1(x))
        e2 = self.lca2(self.enc2(e1))
        e3 = self.lca3(self.enc3(e2))
        if self.variant == 'SH':
            e3 = self.bridge_attention(e3)
        # the multi-scale skip fusions depend on e1 / e2 only: they run beside the decoder stage they are concatenated with
        s1 = (2 * e3.shape[2], 2 * e3.shape[3])
        d1, a1 = parallel([lambda: self.dec1(e3), lambda: self.attention1(self._skips(e1, e2, s1))], [e1, e2, e3], x.device)
        d1 = torch.cat([d1, a1], dim=1)
        s2 = (2 * d1.shape[2], 2 * d1.shape[3])
        d2, a2 = parallel([lambda: _run(self.dec2, d1), lambda: self.attention2(self._skips(e1, e2, s2))], [e1, e2, d1], x.device)
        d2 = torch.cat([d2, a2], dim=1)
        return self.activation(group_norm(self.output_norm, _run(self.dec3, d2)))


def KM_UNetV3_SH(num_classes=3, embed_dims=(16, 32, 64)):
    return KM_UNetV3(num_classes, embed_dims, 'SH')


def KM_UNetV3_LAPS(num_classes=3, embed_dims=(16, 32, 64)):
    return KM_UNetV3(num_classes, embed_dims, 'LAPS')
