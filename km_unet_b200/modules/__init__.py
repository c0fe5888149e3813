from .dagem import DAGEM
from .dysample import DySample
from .kan import KAN_Convolutional_Layer, KANConv2d, KANLinear
from .vim import FFN, ConvLayer1D, ConvLayer2D, EfficientViMBlock, HSMSSD, LayerNorm1D, LayerNorm2D

__all__ = ["DAGEM", "DySample", "KANConv2d", "KANLinear", "KAN_Convolutional_Layer", "EfficientViMBlock", "HSMSSD", "LayerNorm1D",
           "LayerNorm2D", "ConvLayer1D", "ConvLayer2D", "FFN"]
