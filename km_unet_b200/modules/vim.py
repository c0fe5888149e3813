"""EfficientViM building blocks with the reference's names, constructor signatures and state_dict layout
(vim_block_init/vim_utils_init.py:34-130, vim_block_init/efficient_vim_init.py:14-97).

LayerNorm1D, the HSM-SSD mixer, the depthwise 3x3 convolutions and every BatchNorm of the block (with its ReLU /
sigmoid-layer-scale epilogue) run in libkmunet.so; the two 1x1 FFN convolutions are cuDNN library GEMMs.
"""
import torch
import torch.nn as nn

from .. import ops


class LayerNorm1D(nn.Module):
    """Channel LayerNorm of (B,C,L)."""

    def __init__(self, num_channels, eps=1e-5, affine=True):
        super().__init__()
        self.num_channels, self.eps, self.affine = num_channels, eps, affine
        if affine:
            self.weight = nn.Parameter(torch.ones(1, num_channels, 1))
            self.bias = nn.Parameter(torch.zeros(1, num_channels, 1))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)

    def forward(self, x):
        if self.affine:
            return ops.layernorm1d(x, self.weight, self.bias, self.eps)
        one = torch.ones(1, self.num_channels, 1, device=x.device)
        return ops.layernorm1d(x, one, torch.zeros_like(one), self.eps)


class LayerNorm2D(nn.Module):
    """Channel LayerNorm of (B,C,H,W) (unused by KM-UNet; kept for import parity)."""

    def __init__(self, num_channels, eps=1e-5, affine=True):
        super().__init__()
        self.num_channels, self.eps, self.affine = num_channels, eps, affine
        if affine:
            self.weight = nn.Parameter(torch.ones(1, num_channels, 1, 1))
            self.bias = nn.Parameter(torch.zeros(1, num_channels, 1, 1))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)

    def forward(self, x):
        B, C, H, W = x.shape
        w = self.weight if self.affine else torch.ones(1, C, 1, device=x.device)
        b = self.bias if self.affine else torch.zeros(1, C, 1, device=x.device)
        return ops.layernorm1d(x.reshape(B, C, H * W), w.reshape(1, C, 1), b.reshape(1, C, 1), self.eps).reshape(B, C, H, W)


class _ConvLayer(nn.Module):
    conv_cls = None

    def _finish(self, out_dim, norm, act_layer, bn_weight_init):
        self.norm = norm(num_features=out_dim) if norm else None
        self.act = act_layer() if act_layer else None
        if self.norm:
            nn.init.constant_(self.norm.weight, bn_weight_init)
            nn.init.constant_(self.norm.bias, 0)

    def forward(self, x):
        x = self.conv(x)
        if self.norm:
            x = self.norm(x)
        if self.act:
            x = self.act(x)
        return x


def conv1x1(x, weight, bias=None):
    """Pointwise convolution on NCHW through the streaming CUDA kernel (cuDNN's implicit-GEMM path transposes to NHWC and
    back around every such call); channel pairs its weight-gradient kernel does not cover stay a library conv2d."""
    if ops.pwconv_supported(weight.shape[1], weight.shape[0]):
        from .. import config
        return ops.pwconv(x, weight, bias, config.precision_code(config.conv_precision))
    return torch.nn.functional.conv2d(x, weight.reshape(weight.shape[0], weight.shape[1], 1, 1), bias)


def conv_same(conv, x):
    """An nn.Conv2d with stride 1 and 'same' zero padding through the streaming CUDA kernels when they cover it (1x1 and
    <= 9-tap dense kernels with power-of-two input channels); anything else stays the library convolution."""
    kh, kw = conv.kernel_size
    plain = (conv.stride == (1, 1) and conv.dilation == (1, 1) and conv.groups == 1 and conv.padding == (kh // 2, kw // 2)
             and conv.padding_mode == "zeros")
    if plain and kh == kw == 1:
        return conv1x1(x, conv.weight, conv.bias)
    # 3x3 and larger stay on cuDNN: the per-tap weight-gradient passes re-read dy once per tap, which only pays for <= 3 taps
    if plain and kh * kw <= 3 and ops.smallconv_supported(conv.in_channels, conv.out_channels, kh, kw):
        return ops.smallconv(x, conv.weight, conv.bias)
    return conv(x)


_PENDING_COUNTS = None       # list while a whole-model forward collects its BatchNorm counters (batch_counters below)


def count_batch(bn):
    """`bn.num_batches_tracked += 1` as nn.BatchNorm2d.forward does in training -- immediately, or deferred to one multi-tensor
    add at the end of the enclosing model forward (65 counters per KM_UNetV3 step, each its own 1-thread kernel otherwise)."""
    if bn.training and bn.num_batches_tracked is not None:
        if _PENDING_COUNTS is not None:
            _PENDING_COUNTS.append(bn.num_batches_tracked)
        else:
            bn.num_batches_tracked += 1


class batch_counters:
    """Context of one model forward: BatchNorm counters touched inside are incremented together on exit (only if the forward
    completed).  Nested use joins the outer context."""

    def __enter__(self):
        global _PENDING_COUNTS
        self.outer = _PENDING_COUNTS is not None
        if not self.outer:
            _PENDING_COUNTS = []
        return self

    def __exit__(self, exc_type, exc, tb):
        global _PENDING_COUNTS
        if self.outer:
            return False
        pending, _PENDING_COUNTS = _PENDING_COUNTS, None
        if exc_type is None and pending:
            by_dev = {}
            for t in pending:
                by_dev.setdefault(t.device, []).append(t)
            for ts in by_dev.values():
                torch._foreach_add_(ts, 1)
        return False


def _bn2d(bn, x, relu=False, res=None, alpha=None):
    """BatchNorm2d module `bn` applied through the fused CUDA op (keeps the module's parameters, buffers and counters)."""
    training = bn.training or bn.running_mean is None
    y = ops.bnmix(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, training, bn.momentum, bn.eps, relu, res, alpha)
    count_batch(bn)
    return y


class ConvLayer2D(_ConvLayer):
    def __init__(self, in_dim, out_dim, kernel_size=3, stride=1, padding=0, dilation=1, groups=1, norm=nn.BatchNorm2d,
                 act_layer=nn.ReLU, bn_weight_init=1):
        super().__init__()
        self.conv = nn.Conv2d(in_dim, out_dim, (kernel_size, kernel_size), (stride, stride), (padding, padding),
                              (dilation, dilation), groups, bias=False)
        self._finish(out_dim, norm, act_layer, bn_weight_init)
        self._dw3 = (kernel_size == 3 and stride == 1 and padding == 1 and dilation == 1 and groups == in_dim == out_dim)
        self._pw = (kernel_size == 1 and stride == 1 and padding == 0 and groups == 1)

    def conv_out(self, x):
        if self._dw3:
            return ops.dwconv3x3(x, self.conv.weight)
        return conv1x1(x, self.conv.weight) if self._pw else self.conv(x)

    def forward(self, x, res=None, alpha=None):
        """conv -> norm -> act; with res/alpha the block's layer-scale mix is fused into the normalisation."""
        fusable = isinstance(self.norm, nn.BatchNorm2d) and (self.act is None or isinstance(self.act, nn.ReLU)) \
            and self.norm.affine and self.norm.momentum is not None
        if (fusable and self._dw3 and res is x and self.act is None and x.is_cuda and self.conv.bias is None
                and getattr(ops, "dwconv_bnmix_enabled", True)):
            # the depthwise branches of EfficientViMBlock: conv, BatchNorm and the layer-scale mix with the conv's own input as one
            # autograd node (the residual gradient is added inside the convolution's dx kernel)
            bn = self.norm
            y = ops.dwconv_bnmix(x, self.conv.weight, bn.weight, bn.bias, alpha, bn.running_mean, bn.running_var,
                                 bn.training or bn.running_mean is None, bn.momentum, bn.eps)
            count_batch(bn)
            return y
        x = self.conv_out(x)
        if fusable:
            return _bn2d(self.norm, x, relu=self.act is not None, res=res, alpha=alpha)
        if self.norm:
            x = self.norm(x)
        if self.act:
            x = self.act(x)
        if res is not None:
            a = torch.sigmoid(alpha).view(1, -1, 1, 1)
            x = (1 - a) * res + a * x
        return x


class ConvLayer1D(_ConvLayer):
    def __init__(self, in_dim, out_dim, kernel_size=3, stride=1, padding=0, dilation=1, groups=1, norm=nn.BatchNorm1d,
                 act_layer=nn.ReLU, bn_weight_init=1):
        super().__init__()
        self.conv = nn.Conv1d(in_dim, out_dim, kernel_size, stride, padding, dilation, groups, bias=False)
        self._finish(out_dim, norm, act_layer, bn_weight_init)


class FFN(nn.Module):
    def __init__(self, in_dim, dim):
        super().__init__()
        self.fc1 = ConvLayer2D(in_dim, dim, 1)
        self.fc2 = ConvLayer2D(dim, in_dim, 1, act_layer=None, bn_weight_init=0)

    def forward(self, x, res=None, alpha=None):
        return self.fc2(self.fc1(x), res, alpha)


class HSMSSD(nn.Module):
    """Hidden-state mixer.  Parameters live in the same child modules as the reference (BCdt_proj.conv, dw.conv,
    hz_proj.conv, out_proj.conv, A, D) so checkpoints load unchanged; forward() hands their weights to one fused op."""

    def __init__(self, d_model, ssd_expand=1, A_init_range=(1, 16), state_dim=64):
        super().__init__()
        self.ssd_expand = ssd_expand
        self.d_inner = int(ssd_expand * d_model)
        self.state_dim = state_dim
        if self.d_inner != d_model:
            raise NotImplementedError("km_unet_b200.HSMSSD implements ssd_expand == 1 (the only value KM-UNet uses)")
        conv_dim = 3 * state_dim
        self.BCdt_proj = ConvLayer1D(d_model, conv_dim, 1, norm=None, act_layer=None)
        self.dw = ConvLayer2D(conv_dim, conv_dim, 3, 1, 1, groups=conv_dim, norm=None, act_layer=None, bn_weight_init=0)
        self.hz_proj = ConvLayer1D(d_model, 2 * self.d_inner, 1, norm=None, act_layer=None)
        self.out_proj = ConvLayer1D(self.d_inner, d_model, 1, norm=None, act_layer=None, bn_weight_init=0)
        self.A = nn.Parameter(torch.empty(state_dim, dtype=torch.float32).uniform_(*A_init_range))
        self.act = nn.SiLU()
        self.D = nn.Parameter(torch.ones(1))
        self.D._no_weight_decay = True

    def forward(self, x):
        from .. import config
        return ops.hsmssd(x, self.BCdt_proj.conv.weight, self.dw.conv.weight, self.hz_proj.conv.weight,
                          self.out_proj.conv.weight, self.A, self.D, self.state_dim, config.precision_code(config.hsm_precision))


class EfficientViMBlock(nn.Module):
    def __init__(self, dim, mlp_ratio=4., ssd_expand=1, state_dim=64):
        super().__init__()
        self.dim, self.mlp_ratio = dim, mlp_ratio
        self.mixer = HSMSSD(d_model=dim, ssd_expand=ssd_expand, state_dim=state_dim)
        self.norm = LayerNorm1D(dim)
        self.dwconv1 = ConvLayer2D(dim, dim, 3, padding=1, groups=dim, bn_weight_init=0, act_layer=None)
        self.dwconv2 = ConvLayer2D(dim, dim, 3, padding=1, groups=dim, bn_weight_init=0, act_layer=None)
        self.ffn = FFN(in_dim=dim, dim=int(dim * mlp_ratio))
        self.alpha = nn.Parameter(1e-4 * torch.ones(4, dim), requires_grad=True)

    def forward(self, x):
        # x <- (1 - sigmoid(alpha_i)) x + sigmoid(alpha_i) f_i(x): the mix of the three conv branches is the epilogue of
        # their BatchNorm kernel; the mixer's is one lerp
        # one unbind instead of four selects: its backward is ONE stack of the four row gradients (a select's backward is a
        # zero-fill + a copy, and the four results are then summed: 11 tiny launches per block and step)
        a0, a1, a2, a3 = self.alpha.unbind(0)
        x = self.dwconv1(x, res=x, alpha=a0)
        mixed, _ = self.mixer(self.norm(x.flatten(2)))
        from .. import config
        if config.fused_lerp and ops.lerpmix_supported(x):
            x = ops.lerpmix(x, mixed, a1)
        else:
            x = torch.lerp(x, mixed, torch.sigmoid(a1).view(1, -1, 1, 1))
        x = self.dwconv2(x, res=x, alpha=a2)
        return self.ffn(x, res=x, alpha=a3)
