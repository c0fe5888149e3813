"""Full-model CPU oracle (TEST INFRASTRUCTURE -- see oracle/__init__.py).

`cpu_ops()` is a context manager that swaps the CUDA entry points of km_unet_b200.ops for the oracle restatements of
this package, so the host-side model mirror (km_unet_b200/modules/km_unet.py, whose glue is plain torch) runs end to end
on the CPU with the reference arithmetic.  Used by tests/ (parity of the CUDA model against it), by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.  Never used by the product path: outside
this context manager the operators raise on CPU tensors.
"""
import contextlib

from . import dagem as _dagem
from . import dysample as _dysample
from . import hsmssd as _hsmssd
from . import kan as _kan


def _kanconv2d(x, base_weight, spline_weight, spline_scaler, grid, kernel_size, stride=1, padding=0, grid_size=5,
               spline_order=3, precision=0, grid_meta=None):
    return _kan.kanconv2d(x, base_weight, spline_weight, spline_scaler, grid, kernel_size, stride, padding, spline_order)


def _kanlinear(x, base_weight, spline_weight, spline_scaler, grid, grid_size=5, spline_order=3, precision=0, grid_meta=None):
    return _kan.kan_linear(x, base_weight, spline_weight, spline_scaler, grid, spline_order)


def _hsm(x, w_bcdt, w_dw, w_hz, w_out, A, D, state_dim=64, precision=0):
    y, h = _hsmssd.hsmssd(x, w_bcdt, w_dw, w_hz, w_out, A, D, state_dim)
    B, C, L = x.shape
    H = int(round(L ** 0.5))
    return y.reshape(B, C, H, H), h


def _dys(x, w_offset, b_offset, init_pos, scale=2, groups=4):
    return _dysample.dysample_lp(x, w_offset, b_offset, init_pos)


def _bnmix(x, weight, bias, running_mean, running_var, training, momentum=0.1, eps=1e-5, relu=False, res=None, alpha=None):
    """vim_utils_init.py:83-89 (BatchNorm2d inside ConvLayer2D) + efficient_vim_init.py:82-96 (sigmoid layer-scale mix)."""
    import torch
    import torch.nn.functional as F
    y = F.batch_norm(x, running_mean, running_var, weight, bias, training, momentum, eps)
    if relu:
        y = F.relu(y)
    if res is not None:
        a = torch.sigmoid(alpha).view(1, -1, 1, 1)
        y = (1 - a) * res + a * y
    return y


def _dwconv3x3(x, weight, bias=None, scale=None):
    import torch.nn.functional as F
    y = F.conv2d(x, weight, bias, stride=1, padding=1, groups=x.shape[1])
    return y if scale is None else y * scale[:, :, None, None]      # DirectionAttention: self.conv(attn) * weight (KM_UNetV3_SH.py:130-151)


def _combine3(x, f0, f1, f2, coef):
    c = coef.reshape(coef.shape[0], 3, 1, 1, 1)                      # KM_UNetV3_SH.py:361-364: x + drop_path(sum_i g_i f_i)
    return x + c[:, 0] * f0 + c[:, 1] * f1 + c[:, 2] * f2


def _pwconv(x, weight, bias=None, precision=0):
    import torch.nn.functional as F
    return F.conv2d(x, weight.reshape(weight.shape[0], weight.shape[1], 1, 1), bias)


def _triplenorm(x, gh, bh, gw, bw, gc, bc, eps_gn=1e-5, eps_ln=1e-5):
    """KM_UNetV3_SH.py:277-284, literally (incl. the permutes)."""
    import torch.nn.functional as F
    h = F.group_norm(x.permute(0, 1, 3, 2), 1, gh, bh, eps_gn).permute(0, 1, 3, 2)
    w = F.group_norm(x, 1, gw, bw, eps_gn)
    c = F.layer_norm(x.permute(0, 2, 3, 1), (x.shape[1],), gc, bc, eps_ln).permute(0, 3, 1, 2)
    return (h + w + c) / 3


def _qkv_gate(qkv):
    import torch
    q, k, v = qkv.chunk(3, dim=1)
    return torch.sigmoid(q * k) * v


def _smallconv(x, weight, bias=None):
    import torch.nn.functional as F
    return F.conv2d(x, weight, bias, padding=(weight.shape[2] // 2, weight.shape[3] // 2))


def _iwp(x, fusion_weight, fusion_bias):
    """WPL/iwp.py:58-132 with its banded Haar matrices written out (last high-pass row / column zero)."""
    import torch
    import torch.nn.functional as F
    B, C, H, W = x.shape
    s = 2 ** -0.5

    def mats(n):
        lo = torch.zeros(n // 2, n, dtype=x.dtype)
        hi = torch.zeros(n // 2, n, dtype=x.dtype)
        for i in range(n // 2):
            lo[i, 2 * i], lo[i, 2 * i + 1] = s, s
            if i < n // 2 - 1:
                hi[i, 2 * i], hi[i, 2 * i + 1] = s, -s
        return lo, hi
    lo_h, hi_h = mats(H)
    lo_w, hi_w = mats(W)
    L, Hh = lo_h @ x, hi_h @ x
    LL, LH, HL, HH = L @ lo_w.t(), L @ hi_w.t(), Hh @ lo_w.t(), Hh @ hi_w.t()
    high = torch.cat([LH, HL, HH], dim=1).mean(dim=1, keepdim=True)
    return F.conv2d(torch.cat([LL, high], dim=1), fusion_weight.reshape(C, C + 1, 1, 1), fusion_bias)


def _gate(x, deformed, linears, bns, training, momentum=0.1, eps=1e-5):
    return _dagem.dagem_gate(x, deformed, linears, bns, training)


@contextlib.contextmanager
def cpu_ops():
    from km_unet_b200 import ops
    saved = {n: getattr(ops, n) for n in ("kanconv2d", "kanlinear", "layernorm1d", "hsmssd", "dysample", "dagem_gate", "bnmix",
                                          "dwconv3x3", "pwconv", "triplenorm", "qkv_gate", "smallconv", "iwp", "combine3",
                                          "combine3_supported")}
    ops.kanconv2d, ops.kanlinear, ops.layernorm1d = _kanconv2d, _kanlinear, _hsmssd.layernorm1d
    ops.hsmssd, ops.dysample, ops.dagem_gate = _hsm, _dys, _gate
    ops.bnmix, ops.dwconv3x3, ops.pwconv = _bnmix, _dwconv3x3, _pwconv
    ops.triplenorm, ops.qkv_gate, ops.smallconv, ops.iwp = _triplenorm, _qkv_gate, _smallconv, _iwp
    ops.combine3, ops.combine3_supported = _combine3, (lambda x: True)
    try:
        yield
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)
