"""Bring-up aid for the tcgen05 HSM-SSD backward: per-tensor errors vs the fp64 oracle and timings.  python tools/bringup_hsm_bwd.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import km_unet_b200 as K
from km_unet_b200 import HSMSSD
from oracle import hsmssd as O


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


for B, C, H in [(2, 16, 40), (1, 32, 33), (2, 64, 32), (3, 64, 17)]:
    torch.manual_seed(C + H)
    m = HSMSSD(d_model=C)
    sd = m.state_dict()
    x = torch.randn(B, C, H * H)
    xd = x.double().requires_grad_(True)
    W = [sd[k].double().requires_grad_(True) for k in ("BCdt_proj.conv.weight", "dw.conv.weight", "hz_proj.conv.weight", "out_proj.conv.weight")]
    want, _ = O.hsmssd(xd, W[0], W[1], W[2], W[3], sd["A"].double(), sd["D"].double())
    gout = torch.randn(want.shape)
    want.backward(gout.double())
    m = m.cuda()
    for prec in ("fp32", "bf16"):
        K.config.hsm_precision = prec
        m.zero_grad()
        xc = x.cuda().requires_grad_(True)
        y, _ = m(xc)
        y.backward(gout.cuda().reshape(y.shape))
        torch.cuda.synchronize()
        print((B, C, H), prec, "y %.2e dx %.2e dWp %.2e dWd %.2e dWhz %.2e dWo %.2e" % (
            rel(y.reshape(want.shape), want), rel(xc.grad, xd.grad), rel(m.BCdt_proj.conv.weight.grad, W[0].grad),
            rel(m.dw.conv.weight.grad, W[1].grad), rel(m.hz_proj.conv.weight.grad, W[2].grad), rel(m.out_proj.conv.weight.grad, W[3].grad)), flush=True)

# timings at the model's shapes
for B, C, H in [(32, 16, 128), (32, 32, 64), (32, 64, 32)]:
    m = HSMSSD(d_model=C).cuda()
    x = torch.randn(B, C, H * H, device="cuda", requires_grad=True)
    for prec in ("fp32", "bf16"):
        K.config.hsm_precision = prec
        for _ in range(3):
            y, _ = m(x)
            y.backward(torch.ones_like(y))
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf = tb = 0.0
        for _ in range(10):
            e[0].record()
            y, _ = m(x)
            e[1].record()
            y.backward(torch.ones_like(y))
            e[2].record()
            torch.cuda.synchronize()
            tf += e[0].elapsed_time(e[1])
            tb += e[1].elapsed_time(e[2])
        print((B, C, H), prec, "fwd %.3f ms  bwd %.3f ms" % (tf / 10, tb / 10), flush=True)
print("ok")
