"""Drop-in for the reference's vim_block_init/vim_utils_init.py."""
from km_unet_b200.modules.vim import FFN, ConvLayer1D, ConvLayer2D, LayerNorm1D, LayerNorm2D  # noqa: F401
