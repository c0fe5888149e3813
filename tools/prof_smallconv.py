"""ncu driver: forward + backward of the 1x3 / 3x1 projections (ops.smallconv) at the model's shapes: python tools/prof_smallconv.py"""
import sys
import torch
sys.path.insert(0, ".")
from km_unet_b200 import ops
for (C, S, kh, kw) in [(16, 128, 3, 1), (16, 128, 1, 3), (32, 64, 3, 1), (64, 32, 1, 3)]:
    x = torch.randn(32, C, S, S, device="cuda", requires_grad=True)
    w = torch.randn(C, C, kh, kw, device="cuda", requires_grad=True)
    b = torch.randn(C, device="cuda", requires_grad=True)
    for _ in range(2):
        y = ops.smallconv(x, w, b)
        y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("ok")
