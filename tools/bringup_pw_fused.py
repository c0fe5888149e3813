"""Bring-up aid for the fused pointwise-convolution backward: errors vs fp64 and timings vs the split kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import km_unet_b200 as K
from km_unet_b200 import ops


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


for B, Cin, Cout, H, W in [(2, 16, 64, 32, 32), (3, 64, 16, 16, 20), (2, 32, 128, 24, 24), (2, 128, 32, 16, 16), (2, 32, 32, 13, 12)]:
    torch.manual_seed(1)
    x, w, bv, g = torch.randn(B, Cin, H, W), torch.randn(Cout, Cin, 1, 1) / Cin ** 0.5, torch.randn(Cout), torch.randn(B, Cout, H, W)
    xd, wd, bd = x.double().requires_grad_(True), w.double().requires_grad_(True), bv.double().requires_grad_(True)
    F.conv2d(xd, wd, bd).backward(g.double())
    for mode in ("split", "fused"):
        K.config.conv_bwd = mode
        xc, wc, bc = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True), bv.cuda().requires_grad_(True)
        ops.pwconv(xc, wc, bc).backward(g.cuda())
        torch.cuda.synchronize()
        print((B, Cin, Cout, H, W), mode, "dx %.2e dw %.2e db %.2e" % (rel(xc.grad, xd.grad), rel(wc.grad, wd.grad), rel(bc.grad, bd.grad)), flush=True)

for B, Cin, Cout, S in [(32, 16, 64, 128), (32, 64, 16, 128), (32, 16, 48, 128), (32, 32, 128, 64), (32, 128, 32, 64), (32, 32, 96, 64), (32, 64, 64, 32)]:
    x = torch.randn(B, Cin, S, S, device="cuda", requires_grad=True)
    w = torch.randn(Cout, Cin, 1, 1, device="cuda", requires_grad=True)
    bv = torch.randn(Cout, device="cuda", requires_grad=True)
    g = torch.randn(B, Cout, S, S, device="cuda")
    for mode in ("split", "fused"):
        K.config.conv_bwd = mode
        y = ops.pwconv(x, w, bv)
        for _ in range(3):
            y.backward(g, retain_graph=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            y.backward(g, retain_graph=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gb = 4.0 * (2 * Cin + Cout) * B * S * S / 1e9
        print((B, Cin, Cout, S), mode, "bwd %.3f ms  (%.0f GB/s of the x + dy + dx minimum)" % (ms, gb / ms * 1e3), flush=True)
print("ok")
