"""Drop-in for the reference's convKAN/KANConv2Dlayers.py (`from convKAN.KANConv2Dlayers import *` in KM_UNetV3_*.py)."""
import torch.nn as nn  # noqa: F401
import torch.nn.functional as F  # noqa: F401

from .KANlayers import *  # noqa: F401,F403
from km_unet_b200.modules.kan import KAN_Convolutional_Layer, KANConv2d, KANLinear  # noqa: F401
