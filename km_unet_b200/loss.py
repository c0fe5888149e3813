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


_BANDS = {}


def _band(n, size, sigma, device, dtype):
    """(n, n - size + 1) banded matrix whose column j holds the Gaussian taps at rows j .. j + size - 1."""
    key = (n, size, sigma, str(device), dtype)
    if key not in _BANDS:
        g = _gaussian_taps(size, sigma, device, dtype)
        m = torch.zeros(n, n - size + 1, device=device, dtype=dtype)
        idx = torch.arange(n - size + 1, device=device)
        for k in range(size):
            m[idx + k, idx] = g[k]
        _BANDS[key] = m
    return _BANDS[key]


def ssim(pred, target, data_range=1.0, size=11, sigma=1.5, k1=0.01, k2=0.03):
    """torchmetrics pads by reflection, filters, then crops the padded border again: what survives are exactly the
    windows that lie inside the image, i.e. a VALID 11x11 Gaussian filter.  The window is separable, so the filter is two
    small dense products with constant banded matrices (x @ G_w, G_h^T @ .), which run as batched GEMMs."""
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    H, W = pred.shape[-2:]
    gw = _band(W, size, sigma, pred.device, pred.dtype)
    gh = _band(H, size, sigma, pred.device, pred.dtype).t()
    stack = torch.stack([pred, target, pred * pred, target * target, pred * target])
    out = torch.matmul(gh, torch.matmul(stack, gw))
    mu_p, mu_t, pp, tt, pt = out.unbind(0)
    s_p, s_t, s_pt = pp - mu_p * mu_p, tt - mu_t * mu_t, pt - mu_p * mu_t
    m = ((2 * mu_p * mu_t + c1) * (2 * s_pt + c2)) / ((mu_p * mu_p + mu_t * mu_t + c1) * (s_p + s_t + c2))
    return m.reshape(m.shape[0], -1).mean(-1).mean()


class _HybridLossFn(torch.autograd.Function):
    """The same loss as four streaming CUDA passes (csrc/loss.cu) around the two banded GEMMs of the SSIM filter; analytic backward."""

    @staticmethod
    def forward(ctx, pred, target, alpha, size, sigma, k1, k2):
        import ctypes as C
        from . import _lib
        from .ops import _call, _workspace, check, ptr, stream_ptr
        lib = _lib.lib()
        pred = pred.detach().to(torch.float32).contiguous()
        target = target.detach().to(torch.float32).contiguous()
        n = pred.numel()
        H, W = pred.shape[-2:]
        dev = pred.device
        scal = torch.empty(8, dtype=torch.float32, device=dev)
        ws = _workspace(lib.kmu_hybridloss_workspace_bytes(), dev)
        key = tuple(pred.shape)
        check(_call("kmu_hybridloss_stats", key, lib.kmu_hybridloss_stats, ptr(pred), ptr(target), n, ptr(scal), ws.data_ptr(), ws.numel(),
                    stream_ptr()), "kmu_hybridloss_stats")
        stack = torch.empty((5,) + tuple(pred.shape), dtype=torch.float32, device=dev)
        check(_call("kmu_hybridloss_stack", key, lib.kmu_hybridloss_stack, ptr(pred), ptr(target), ptr(scal), ptr(stack), n, stream_ptr()),
              "kmu_hybridloss_stack")
        gw = _band(W, size, sigma, dev, torch.float32)
        gh = _band(H, size, sigma, dev, torch.float32).t()
        filt = torch.matmul(gh, torch.matmul(stack, gw)).contiguous()          # (5, ..., H - size + 1, W - size + 1)
        nv = filt.numel() // 5
        gm = torch.empty((3,) + tuple(filt.shape[1:]), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        check(_call("kmu_hybridloss_ssim", key, lib.kmu_hybridloss_ssim, ptr(filt), ptr(scal), ptr(gm), ptr(loss), nv, n, float(alpha),
                    float(k1) ** 2, float(k2) ** 2, ws.data_ptr(), ws.numel(), stream_ptr()), "kmu_hybridloss_ssim")
        ctx.save_for_backward(pred, target, scal, gm, gw, gh)
        ctx.alpha = float(alpha)
        return loss

    @staticmethod
    def backward(ctx, g):
        from . import _lib
        from .ops import _call, check, ptr, stream_ptr
        lib = _lib.lib()
        pred, target, scal, gm, gw, gh = ctx.saved_tensors
        dstack = torch.matmul(gh.t(), torch.matmul(gm, gw.t())).contiguous()   # transposed filter: (3, ..., H, W)
        dp = torch.empty_like(pred)
        g = g.detach().to(torch.float32).contiguous()
        check(_call("kmu_hybridloss_bwd", tuple(pred.shape), lib.kmu_hybridloss_bwd, ptr(pred), ptr(target), ptr(scal), ptr(dstack), ptr(g),
                    ptr(dp), pred.numel(), ctx.alpha, stream_ptr()), "kmu_hybridloss_bwd")
        return dp, None, None, None, None, None, None


class HybridLoss(nn.Module):
    def __init__(self, alpha=0.7):
        super().__init__()
        self.alpha = alpha

    def forward(self, pred, target):
        if pred.is_cuda and pred.numel() % 4 == 0 and not target.requires_grad and min(pred.shape[-2:]) >= 11:
            return _HybridLossFn.apply(pred, target, self.alpha, 11, 1.5, 0.01, 0.03)
        return self.forward_torch(pred, target)

    def forward_torch(self, pred, target):
        """The formula in plain torch (CPU oracle runs, and the reference the CUDA path is tested against)."""
        mse = F.mse_loss(pred, target)
        weighted = ((pred - target).pow(2) * torch.exp(target * 2)).mean()
        t_min, t_max = target.min().detach(), target.max().detach()
        p_min, p_max = pred.min().detach(), pred.max().detach()
        t_n = (target - t_min) / (t_max - t_min + 1e-8)
        p_n = (pred - p_min) / (p_max - p_min + 1e-8)
        return self.alpha * (0.55 * mse + 0.45 * weighted) + (1 - self.alpha) * (1 - ssim(p_n, t_n))
