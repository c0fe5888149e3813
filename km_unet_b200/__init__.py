"""km_unet_b200 -- B200-native (sm_100a) hot path of KM-UNet behind the reference's module API.

    from km_unet_b200 import KANConv2d, KANLinear, EfficientViMBlock, HSMSSD, DySample   # CUDA-backed nn.Modules
    km_unet_b200.enable_dropin()   # then the reference's `from convKAN.KANConv2Dlayers import *` etc. resolve to these

The arithmetic lives in libkmunet.so (km_unet_b200/csrc, C ABI in include/kmunet.h).  No CPU fallback.
"""
import os
import sys

from . import config  # noqa: F401

DROPIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")


def enable_dropin():
    """Put the drop-in module tree (convKAN/, vim_block_init/, DySample_md.py, DAGEM_md.py) first on sys.path so the
    reference's KM_UNetV3_SH.py / KM_UNetV3_LAPS.py / train_shanghai.py import the CUDA-backed operators unchanged."""
    if DROPIN_DIR in sys.path:
        sys.path.remove(DROPIN_DIR)
    sys.path.insert(0, DROPIN_DIR)
    return DROPIN_DIR


def __getattr__(name):
    if name in ("KANConv2d", "KANLinear", "KAN_Convolutional_Layer", "EfficientViMBlock", "HSMSSD", "LayerNorm1D", "DySample",
                "DAGEM", "KM_UNetV3", "KM_UNetV3_SH", "KM_UNetV3_LAPS", "StableHybridKANConv", "EnhancedViMBlock",
                "IntelligentWaveletPoolingModule", "modules"):
        if name == "modules":
            import importlib
            return importlib.import_module(".modules", __name__)
        from . import modules
        return getattr(modules, name)
    if name == "FusedAdamW":
        from .optim import FusedAdamW
        return FusedAdamW
    raise AttributeError(name)
