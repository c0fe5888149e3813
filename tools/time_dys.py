"""Device time of DySample forward / backward at 16..128 px (CUDA events around the C-ABI calls): python tools/time_dys.py"""
import sys
import torch
sys.path.insert(0, ".")
import km_unet_b200 as K
from km_unet_b200 import ops
for S in (16, 32, 64, 128):
    m = K.DySample(64).cuda()
    x = torch.randn(32 if S < 128 else 8, 64, S, S, device="cuda", requires_grad=True)
    for _ in range(3):
        m(x).sum().backward()
    ops.profile_start()
    for _ in range(10):
        m(x).sum().backward()
    pr = ops.profile_stop()
    B = x.shape[0]
    t = {k[0]: v["ms"] / v["calls"] for k, v in pr.items()}
    print(S, {k: round(v * 1000, 1) for k, v in t.items()}, "fwd GB/s %.0f bwd GB/s %.0f" % (20.0 * B * 64 * S * S / t["kmu_dysample_fwd"] / 1e6, 24.0 * B * 64 * S * S / t["kmu_dysample_bwd"] / 1e6))
