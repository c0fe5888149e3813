"""Oracle restatement of HSMSSD / LayerNorm1D / EfficientViMBlock (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Follows the reference:
    HSMSSD.forward            vim_block_init/efficient_vim_init.py:33-61   (hsmssd)
    LayerNorm1D.forward       vim_block_init/vim_utils_init.py:50-59       (layernorm1d)
    ConvLayer2D / FFN         vim_block_init/vim_utils_init.py:83-89,128-130
    EfficientViMBlock.forward vim_block_init/efficient_vim_init.py:81-97   (vim_block)
Parameters are passed as a dict keyed like the reference state_dict (prefix stripped), e.g.
"mixer.BCdt_proj.conv.weight".  Pure torch-CPU arithmetic, fp32 or fp64.
"""
import math

import torch
import torch.nn.functional as F


def layernorm1d(x, weight, bias, eps=1e-5):
    """Per-position normalisation over the channel axis of (B,C,L); biased variance."""
    mu = x.mean(dim=1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * weight.reshape(1, -1, 1) + bias.reshape(1, -1, 1)


def hsmssd(x, w_bcdt, w_dw, w_hz, w_out, A, D, state_dim=64):
    """x (B,C,L), L = H*H.  Returns (y (B,C,H,H), h (B,C,N)).

    w_bcdt (3N,C[,1]); w_dw (3N,1,3,3); w_hz (2C,C[,1]); w_out (C,C[,1]); A (N,), D (1,).
    """
    Bsz, C, L = x.shape
    H = int(math.sqrt(L))
    N = state_dim
    wp = w_bcdt.reshape(3 * N, C)
    q = torch.einsum("nc,bcl->bnl", wp, x).reshape(Bsz, 3 * N, H, H)
    p = F.conv2d(q, w_dw.reshape(3 * N, 1, 3, 3), padding=1, groups=3 * N).reshape(Bsz, 3 * N, L)
    Bm, Cm, dt = p[:, :N], p[:, N:2 * N], p[:, 2 * N:]
    a = torch.softmax(dt + A.reshape(1, N, 1), dim=-1)          # softmax along L
    hs = torch.einsum("bcl,bnl->bcn", x, a * Bm)
    hz = torch.einsum("dc,bcn->bdn", w_hz.reshape(2 * C, C), hs)
    hh, z = hz[:, :C], hz[:, C:]
    gated = hh * F.silu(z) + hh * D
    ho = torch.einsum("dc,bcn->bdn", w_out.reshape(C, C), gated)
    y = torch.einsum("bcn,bnl->bcl", ho, Cm).reshape(Bsz, C, H, H)
    return y, ho


def hsmssd_grads(x, dy, w_bcdt, w_dw, w_hz, w_out, A, D, state_dim=64):
    """Closed-form backward of hsmssd w.r.t. y only (the block discards h).  Returns a dict of gradients.
    This is the three-sweep structure the CUDA backward implements."""
    Bsz, C, L = x.shape
    H = int(math.sqrt(L))
    N = state_dim
    wp = w_bcdt.reshape(3 * N, C)
    wd = w_dw.reshape(3 * N, 1, 3, 3)
    whz = w_hz.reshape(2 * C, C)
    wo = w_out.reshape(C, C)
    dy = dy.reshape(Bsz, C, L)
    q = torch.einsum("nc,bcl->bnl", wp, x)
    p = F.conv2d(q.reshape(Bsz, 3 * N, H, H), wd, padding=1, groups=3 * N).reshape(Bsz, 3 * N, L)
    Bm, Cm, dt = p[:, :N], p[:, N:2 * N], p[:, 2 * N:]
    a = torch.softmax(dt, dim=-1)
    hs = torch.einsum("bcl,bnl->bcn", x, a * Bm)
    hz = torch.einsum("dc,bcn->bdn", whz, hs)
    hh, z = hz[:, :C], hz[:, C:]
    sz = torch.sigmoid(z)
    silu_z = z * sz
    v = hh * (silu_z + D)
    ho = torch.einsum("dc,bcn->bdn", wo, v)
    # sweep 1 -- reductions over L
    dho = torch.einsum("bcl,bnl->bcn", dy, Cm)
    dCm = torch.einsum("bcn,bcl->bnl", ho, dy)
    dv = torch.einsum("dc,bdn->bcn", wo, dho)
    d_wo = torch.einsum("bdn,bcn->dc", dho, v)
    dhh = dv * (silu_z + D)
    dz = dv * hh * (sz * (1 + z * (1 - sz)))
    d_D = (dv * hh).sum().reshape(1)
    dhz = torch.cat([dhh, dz], dim=1)
    dhs = torch.einsum("dc,bdn->bcn", whz, dhz)
    d_whz = torch.einsum("bdn,bcn->dc", dhz, hs)
    # sweep 2 -- softmax backward needs sum_L(dA * A)
    dx = torch.einsum("bcn,bnl->bcl", dhs, a * Bm)
    dG = torch.einsum("bcn,bcl->bnl", dhs, x)
    dBm = dG * a
    dA = dG * Bm
    ddt = a * (dA - (dA * a).sum(-1, keepdim=True))
    # sweep 3 -- transpose depthwise conv + projection
    dP = torch.cat([dBm, dCm, ddt], dim=1).reshape(Bsz, 3 * N, H, H)
    dQ = F.conv_transpose2d(dP, wd, padding=1, groups=3 * N).reshape(Bsz, 3 * N, L)
    qimg = F.pad(q.reshape(Bsz, 3 * N, H, H), (1, 1, 1, 1))
    d_wd = torch.stack([(qimg[:, :, i:i + H, j:j + H] * dP).sum(dim=(0, 2, 3)) for i in range(3) for j in range(3)],
                       dim=-1).reshape(3 * N, 1, 3, 3)
    dx = dx + torch.einsum("nc,bnl->bcl", wp, dQ)
    d_wp = torch.einsum("bnl,bcl->nc", dQ, x)
    return {"x": dx, "BCdt_proj": d_wp.reshape(w_bcdt.shape), "dw": d_wd, "hz_proj": d_whz.reshape(w_hz.shape),
            "out_proj": d_wo.reshape(w_out.shape), "A": torch.zeros_like(A), "D": d_D}


def batchnorm2d(x, weight, bias, running_mean, running_var, training, eps=1e-5, momentum=0.1):
    """Returns (y, new_running_mean, new_running_var); batch statistics in training mode."""
    if training:
        n = x.numel() // x.shape[1]
        mu = x.mean(dim=(0, 2, 3))
        var = ((x - mu.reshape(1, -1, 1, 1)) ** 2).mean(dim=(0, 2, 3))
        new_rm = (1 - momentum) * running_mean + momentum * mu
        new_rv = (1 - momentum) * running_var + momentum * var * n / max(n - 1, 1)
    else:
        mu, var, new_rm, new_rv = running_mean, running_var, running_mean, running_var
    y = (x - mu.reshape(1, -1, 1, 1)) / torch.sqrt(var.reshape(1, -1, 1, 1) + eps)
    return y * weight.reshape(1, -1, 1, 1) + bias.reshape(1, -1, 1, 1), new_rm, new_rv


def _conv_bn(x, P, prefix, training, groups=1, padding=0, relu=False, stats=None):
    y = F.conv2d(x, P[prefix + ".conv.weight"], padding=padding, groups=groups)
    y, rm, rv = batchnorm2d(y, P[prefix + ".norm.weight"], P[prefix + ".norm.bias"],
                            P[prefix + ".norm.running_mean"], P[prefix + ".norm.running_var"], training)
    if stats is not None:
        stats[prefix + ".norm.running_mean"] = rm
        stats[prefix + ".norm.running_var"] = rv
    return F.relu(y) if relu else y


def vim_block(x, P, training=True, state_dim=64, return_stats=False):
    """EfficientViMBlock forward.  x (B,C,H,W) with H == W; P = state_dict of the block."""
    Bsz, C, H, W = x.shape
    stats = {}
    al = torch.sigmoid(P["alpha"]).reshape(4, C, 1, 1)
    x = (1 - al[0]) * x + al[0] * _conv_bn(x, P, "dwconv1", training, groups=C, padding=1, stats=stats)
    xn = layernorm1d(x.reshape(Bsz, C, H * W), P["norm.weight"], P["norm.bias"])
    y, _ = hsmssd(xn, P["mixer.BCdt_proj.conv.weight"], P["mixer.dw.conv.weight"], P["mixer.hz_proj.conv.weight"],
                  P["mixer.out_proj.conv.weight"], P["mixer.A"], P["mixer.D"], state_dim)
    x = (1 - al[1]) * x + al[1] * y
    x = (1 - al[2]) * x + al[2] * _conv_bn(x, P, "dwconv2", training, groups=C, padding=1, stats=stats)
    f = _conv_bn(x, P, "ffn.fc1", training, relu=True, stats=stats)
    f = _conv_bn(f, P, "ffn.fc2", training, stats=stats)
    x = (1 - al[3]) * x + al[3] * f
    return (x, stats) if return_stats else x
