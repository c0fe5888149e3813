"""Drop-in for the reference's convKAN/KANlayers.py: same import path, CUDA-backed KANLinear."""
import math  # noqa: F401  (the reference's star-import exposes these names to KANConv2Dlayers)

import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import torch.nn.functional as F  # noqa: F401

from km_unet_b200.modules.kan import KANLinear  # noqa: F401
