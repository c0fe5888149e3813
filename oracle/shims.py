"""Import shims that let the UNMODIFIED reference modules import in this container.

Test infrastructure (see oracle/__init__.py).  The reference imports a few third-party
names that are not installed here; none of them carries hot-path arithmetic except
DropPath (train mode only), whose published timm-0.9.16 behaviour is restated below.

    timm.models.layers.{trunc_normal_, DropPath}   KM_UNetV3_SH.py:7
    timm.layers.{trunc_normal_, SqueezeExcite}     vim_block_init/efficient_vim_init.py:7, vim_utils_init.py:3
    timm.models.register_model                     vim_block_init/efficient_vim_init.py:8
    fvcore.nn.flop_count                           vim_block_init/efficient_vim_init.py:9
    pywt.Wavelet('haar')                           WPL/iwp.py:5,50-52
    lpips.LPIPS                                    metrics.py (only the evaluator's LPIPS column)
"""
import sys
import types

import torch
import torch.nn as nn


class _DropPath(nn.Module):
    """Stochastic depth, per sample: keep with prob (1-p) and rescale by 1/(1-p)."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = float(drop_prob)
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if not self.training or self.drop_prob == 0.0:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        if self.scale_by_keep and keep > 0.0:
            mask.div_(keep)
        return x * mask


class _Unused(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()


class _HaarWavelet:
    def __init__(self, name):
        if name != "haar":
            raise ValueError("shim only provides the 'haar' taps the reference uses")
        s = 2.0 ** -0.5
        self.rec_lo = [s, s]
        self.rec_hi = [s, -s]
        self.dec_lo = [s, s]
        self.dec_hi = [-s, s]


class _ZeroLPIPS(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, a, b, *args, **kwargs):
        return torch.zeros(a.shape[0], 1, 1, 1)


def _new_module(name):
    mod = types.ModuleType(name)
    mod.__dict__["__shim__"] = True
    sys.modules[name] = mod
    return mod


def install():
    """Idempotently register the stand-in modules in sys.modules (real packages win)."""
    def missing(name):
        if name in sys.modules:
            return False
        try:
            __import__(name)
            return False
        except Exception:
            return True

    if missing("timm"):
        timm = _new_module("timm")
        layers = _new_module("timm.layers")
        models = _new_module("timm.models")
        mlayers = _new_module("timm.models.layers")
        timm.layers, timm.models, models.layers = layers, models, mlayers
        for m in (layers, mlayers):
            m.trunc_normal_ = nn.init.trunc_normal_
            m.DropPath = _DropPath
            m.SqueezeExcite = _Unused
        models.register_model = lambda fn: fn
    if missing("fvcore"):
        fv = _new_module("fvcore")
        fvnn = _new_module("fvcore.nn")
        fv.nn = fvnn
        fvnn.flop_count = lambda *a, **k: ({}, {})
    if missing("pywt"):
        pywt = _new_module("pywt")
        pywt.Wavelet = _HaarWavelet
    if missing("lpips"):
        lp = _new_module("lpips")
        lp.LPIPS = _ZeroLPIPS
