"""Drop-in for the reference's DAGEM_md.py (`from DAGEM_md import DAGEM` in KM_UNetV3_SH.py)."""
from km_unet_b200.modules.dagem import DAGEM  # noqa: F401
