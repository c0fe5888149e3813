"""CPU restatement of HybridLoss (test infrastructure; see oracle/__init__.py).

Follows train_shanghai.py:298-326 (identical in train_LAPS.py:347-375):
    loss = alpha*(0.55*MSE + 0.45*mean((p-t)^2 * exp(2t))) + (1-alpha)*(1 - SSIM(p_n, t_n)),   alpha = 0.7
with p_n, t_n min-max normalised by their own DETACHED global extrema (+1e-8).

The SSIM term is third-party: torchmetrics==1.5.2 (requirements.txt:82; call sites train_shanghai.py:21,302,323), absent
from the reference tree and from this image, and no reference test pins it -> "parity unpinned" for this term.  Its
published algorithm (StructuralSimilarityIndexMeasure defaults): 11x11 Gaussian window, sigma 1.5, k1 = 0.01, k2 = 0.03,
data_range = 1; inputs reflect-padded by 5, filtered with a depthwise conv, the padded border cropped again, mean over
(C, H, W) per sample and then over the batch.  It is written here literally (pad -> conv2d -> crop), NOT in the banded-GEMM
form of km_unet_b200/loss.py, so the two are independent statements of the same definition.
"""
import torch
import torch.nn.functional as F


def gaussian_window(size=11, sigma=1.5, dtype=torch.float32):
    d = torch.arange((1 - size) / 2, (1 + size) / 2, 1, dtype=dtype)
    g = torch.exp(-((d / sigma) ** 2) / 2)
    g = g / g.sum()
    return torch.outer(g, g)


def ssim(pred, target, data_range=1.0, size=11, sigma=1.5, k1=0.01, k2=0.03):
    B, C, H, W = pred.shape
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    pad = (size - 1) // 2
    win = gaussian_window(size, sigma, pred.dtype).to(pred.device).expand(C, 1, size, size)
    p = F.pad(pred, (pad, pad, pad, pad), mode="reflect")
    t = F.pad(target, (pad, pad, pad, pad), mode="reflect")
    maps = F.conv2d(torch.cat([p, t, p * p, t * t, p * t]), win, groups=C)      # 5B images, depthwise
    mu_p, mu_t, pp, tt, pt = maps.split(B)
    s_p, s_t, s_pt = pp - mu_p * mu_p, tt - mu_t * mu_t, pt - mu_p * mu_t
    m = ((2 * mu_p * mu_t + c1) * (2 * s_pt + c2)) / ((mu_p * mu_p + mu_t * mu_t + c1) * (s_p + s_t + c2))
    m = m[..., pad:-pad, pad:-pad]
    return m.reshape(B, -1).mean(-1).mean()


def hybrid_loss(pred, target, alpha=0.7):
    mse = F.mse_loss(pred, target)
    weighted = ((pred - target).pow(2) * torch.exp(target * 2)).mean()
    t_min, t_max = target.min().detach(), target.max().detach()
    p_min, p_max = pred.min().detach(), pred.max().detach()
    t_n = (target - t_min) / (t_max - t_min + 1e-8)
    p_n = (pred - p_min) / (p_max - p_min + 1e-8)
    return alpha * (0.55 * mse + 0.45 * weighted) + (1 - alpha) * (1 - ssim(p_n, t_n))
