from .dagem import DAGEM
from .dysample import DySample
from .km_unet import (KM_UNetV3, KM_UNetV3_LAPS, KM_UNetV3_SH, EnhancedViMBlock, IntelligentWaveletPoolingModule,
                      StableHybridKANConv)
from .kan import KAN_Convolutional_Layer, KANConv2d, KANLinear
from .vim import FFN, ConvLayer1D, ConvLayer2D, EfficientViMBlock, HSMSSD, LayerNorm1D, LayerNorm2D

__all__ = ["KM_UNetV3", "KM_UNetV3_SH", "KM_UNetV3_LAPS", "StableHybridKANConv", "EnhancedViMBlock",
           "IntelligentWaveletPoolingModule", "DAGEM", "DySample", "KANConv2d", "KANLinear", "KAN_Convolutional_Layer", "EfficientViMBlock", "HSMSSD", "LayerNorm1D",
           "LayerNorm2D", "ConvLayer1D", "ConvLayer2D", "FFN"]
