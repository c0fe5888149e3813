"""ncu driver: HSMSSD fwd+bwd at one shape.  python tools/prof_hsm_big.py [C] [S] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import km_unet_b200 as K

C = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
torch.manual_seed(0)
m = K.HSMSSD(C).cuda()
x = torch.randn(B, C, S * S, device="cuda", requires_grad=True)
for _ in range(2):
    y, _ = m(x)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("ok")
