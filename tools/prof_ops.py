"""Tiny driver for ncu launch lists of the non-KAN ops: python tools/prof_ops.py [hsm|dys|vim|dagem|all] [B]."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import km_unet_b200 as K

what = sys.argv[1] if len(sys.argv) > 1 else "all"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.manual_seed(0)
dev = "cuda"


def run(m, x, tup=False):
    for _ in range(reps):
        y = m(x)
        y = y[0] if tup else y
        y.backward(torch.ones_like(y))
    torch.cuda.synchronize()


if what in ("hsm", "all"):
    for C, S in [(16, 128), (32, 64), (64, 32)]:
        run(K.HSMSSD(C).to(dev), torch.randn(B, C, S * S, device=dev, requires_grad=True), tup=True)
if what in ("vim", "all"):
    for C, S in [(16, 128), (32, 64), (64, 32)]:
        run(K.EfficientViMBlock(C).to(dev), torch.randn(B, C, S, S, device=dev, requires_grad=True))
if what in ("dys", "all"):
    for S in [16, 32, 64]:
        run(K.DySample(64).to(dev), torch.randn(B, 64, S, S, device=dev, requires_grad=True))
if what in ("dagem", "all"):
    run(K.DAGEM(input_channels=64).to(dev), torch.randn(B, 64, 16, 16, device=dev, requires_grad=True))
if what in ("pw", "all"):
    from km_unet_b200 import ops
    for cin, cout, S in [(16, 64, 128), (64, 16, 128), (16, 48, 128), (32, 128, 64), (64, 256, 32)]:
        w = torch.randn(cout, cin, 1, 1, device=dev, requires_grad=True)
        bb = torch.randn(cout, device=dev, requires_grad=True)
        x = torch.randn(B, cin, S, S, device=dev, requires_grad=True)
        for _ in range(reps):
            y = ops.pwconv(x, w, bb, K.config.precision_code(K.config.conv_precision))
            y.backward(torch.ones_like(y))
        torch.cuda.synchronize()
if what in ("shell", "all"):
    from km_unet_b200 import ops
    for C, S in [(16, 128), (64, 128), (32, 64)]:
        x = torch.randn(B, C, S, S, device=dev, requires_grad=True)
        r = torch.randn(B, C, S, S, device=dev, requires_grad=True)
        w = torch.ones(C, device=dev, requires_grad=True)
        bb = torch.zeros(C, device=dev, requires_grad=True)
        al = torch.zeros(C, device=dev, requires_grad=True)
        rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
        wd = torch.randn(C, 1, 3, 3, device=dev, requires_grad=True)
        for _ in range(reps):
            y = ops.bnmix(ops.dwconv3x3(x, wd), w, bb, rm, rv, True, 0.1, 1e-5, False, r, al)
            y.backward(torch.ones_like(y))
        torch.cuda.synchronize()
print("ok")
