"""Tiny driver for ncu captures of the KANConv2d kernels: python tools/prof_kan.py [fwd|fwdbwd] [B] [precision]."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import km_unet_b200 as K

mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
K.config.kan_precision = sys.argv[3] if len(sys.argv) > 3 else "bf16"
torch.manual_seed(0)
m = K.KANConv2d(64, 64, 3, padding=1).cuda()
x = torch.randn(B, 64, 128, 128, device="cuda", requires_grad=True)
g = torch.randn(B, 64, 128, 128, device="cuda")
for _ in range(3):
    y = m(x)
    if mode == "fwdbwd":
        y.backward(g)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
