"""HybridLoss of the training scripts (train_shanghai.py:298-326, train_LAPS.py:347-375), caller-side glue in torch.

loss = alpha * (0.55 * MSE + 0.45 * mean((pred-target)^2 * exp(2 target))) + (1 - alpha) * (1 - SSIM(pred_n, target_n))
with pred_n / target_n min-max normalised by their own detached global extrema (+1e-8).
The SSIM term lives in third-party torchmetrics==1.5.2 in the reference (not under the reference tree, no reference test
pins it): it is restated here with torchmetrics' documented defaults -- 11x11 Gaussian window, sigma 1.5, k1 0.01,
k2 0.03, data_range 1, reflect padding then crop, mean over pixels, channels and batch ("parity unpinned").
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _gaussian_taps(size, sigma, device, dtype):
    d = torch.arange((1 - size) / 2, (1 + size) / 2, 1, device=device, dtype=dtype)
    g = torch.exp(-(d / sigma) ** 2 / 2)
    return g / g.sum()


def ssim(pred, target, data_range=1.0, size=11, sigma=1.5, k1=0.01, k2=0.03):
    C = pred.shape[1]
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    pad = (size - 1) // 2
    g = _gaussian_taps(size, sigma, pred.device, pred.dtype)       # the 2-D window is the outer product: filter separably
    p = F.pad(pred, (pad, pad, pad, pad), mode="reflect")
    t = F.pad(target, (pad, pad, pad, pad), mode="reflect")
    stack = torch.cat([p, t, p * p, t * t, p * t])
    n, _, hp, wp = stack.shape
    flat = stack.reshape(n * C, 1, hp, wp)
    out = F.conv2d(F.conv2d(flat, g.view(1, 1, size, 1)), g.view(1, 1, 1, size)).reshape(n, C, hp - 2 * pad, wp - 2 * pad)
    mu_p, mu_t, pp, tt, pt = out.split(pred.shape[0])
    s_p, s_t, s_pt = pp - mu_p * mu_p, tt - mu_t * mu_t, pt - mu_p * mu_t
    m = ((2 * mu_p * mu_t + c1) * (2 * s_pt + c2)) / ((mu_p * mu_p + mu_t * mu_t + c1) * (s_p + s_t + c2))
    m = m[..., pad:-pad, pad:-pad]
    return m.reshape(m.shape[0], -1).mean(-1).mean()


class HybridLoss(nn.Module):
    def __init__(self, alpha=0.7):
        super().__init__()
        self.alpha = alpha

    def forward(self, pred, target):
        mse = F.mse_loss(pred, target)
        weighted = ((pred - target).pow(2) * torch.exp(target * 2)).mean()
        t_min, t_max = target.min().detach(), target.max().detach()
        p_min, p_max = pred.min().detach(), pred.max().detach()
        t_n = (target - t_min) / (t_max - t_min + 1e-8)
        p_n = (pred - p_min) / (p_max - p_min + 1e-8)
        return self.alpha * (0.55 * mse + 0.45 * weighted) + (1 - self.alpha) * (1 - ssim(p_n, t_n))
