"""Process-wide knobs of the drop-in modules."""
import os

# "fp32": CUDA-core fp32 kernels (1e-4 parity gate).  "bf16": tcgen05 tensor-core kernels where the shape allows
# (2e-2 parity gate), fp32 family otherwise.  Both run on the GPU; there is no CPU path.
kan_precision = os.environ.get("KMU_KAN_PRECISION", "fp32")
# same switch for the HSM-SSD projection (BCdt_proj + depthwise 3x3 of the forward as one tcgen05 convolution)
hsm_precision = os.environ.get("KMU_HSM_PRECISION", "fp32")
# same switch for the pointwise (1x1) convolutions of the callers / EfficientViMBlock FFN
conv_precision = os.environ.get("KMU_CONV_PRECISION", "fp32")
# backward of the pointwise convolutions: "split" = input-gradient kernel + weight-gradient kernel of conv_precision's family;
# "fused" = one persistent TMA -> tcgen05 kernel that reads x and dy once and writes dx, dW, db (bf16 operands, 2e-2 gate) where
# the shape allows, the split kernels otherwise
conv_bwd = os.environ.get("KMU_CONV_BWD", "split")
# forward of the pointwise convolutions: "simt" = fp32 streaming kernel (or the tcgen05 kernel of conv_precision = "bf16");
# "tma" = persistent TMA -> tcgen05 pipeline (bf16 operands, 2e-2 gate) where the shape allows
conv_fwd = os.environ.get("KMU_CONV_FWD", "simt")


def precision_code(name=None):
    from ._lib import KMU_PREC_BF16, KMU_PREC_FP32
    name = kan_precision if name is None else name
    if name not in ("fp32", "bf16"):
        raise ValueError(f"unknown KAN precision {name!r} (expected 'fp32' or 'bf16')")
    return KMU_PREC_BF16 if name == "bf16" else KMU_PREC_FP32
# pairs the TMA pipelines do not take (Cin = 256: the tile ring does not fit): "simt" = fp32 streaming kernels (default: these few
# layers were the largest remaining source of end-to-end error when run with plain bf16 operands, tools/bf16_ablation.py),
# "tc" = per-tile tcgen05 kernels of pwconv_tc.cu (bf16 operands)
conv_wide = os.environ.get("KMU_CONV_WIDE", "simt")
# EfficientViMBlock's mixer layer-scale (torch.lerp with a broadcast weight) through kmu_lerpmix (one pass per direction)
fused_lerp = os.environ.get("KMU_FUSED_LERP", "1") == "1"
# EnhancedViMBlock's three direction branches (height / width / channel: independent until the fusion gate) on three CUDA streams:
# inside the captured step graph they become parallel branches, so the many small latency-bound kernels of one branch fill the
# drain / fill bubbles of the others (the backward nodes run on the streams their forward ran on)
parallel_branches = os.environ.get("KMU_PARALLEL_BRANCHES", "1") == "1"
