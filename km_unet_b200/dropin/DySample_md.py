"""Drop-in for the reference's DySample_md.py (`from DySample_md import DySample` in KM_UNetV3_SH.py)."""
from km_unet_b200.modules.dysample import DySample  # noqa: F401
