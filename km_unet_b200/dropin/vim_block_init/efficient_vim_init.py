"""Drop-in for the reference's vim_block_init/efficient_vim_init.py (`from vim_block_init.efficient_vim_init import
EfficientViMBlock` in KM_UNetV3_*.py).  The classifier backbone in that file is never built by KM-UNet and is not
provided."""
from km_unet_b200.modules.vim import HSMSSD, EfficientViMBlock  # noqa: F401
from .vim_utils_init import FFN, ConvLayer1D, ConvLayer2D, LayerNorm1D, LayerNorm2D  # noqa: F401
