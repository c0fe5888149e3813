"""Bring-up of the tcgen05 KANConv2d backward: errors of dX / dW against the fp64 oracle for a few shapes, with the
MN-major descriptor convention both ways (kmu_debug_flags bit 0).  Scratch tool, not a test."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import km_unet_b200 as K
from km_unet_b200 import _lib
from oracle import kan as O


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def run(B, cin, cout, H, W, flags):
    torch.manual_seed(B + cin + cout + H)
    m = K.KANConv2d(cin, cout, 3, padding=1)
    m.kanlayer.precision = "bf16"
    kl = m.kanlayer
    x = torch.randn(B, cin, H, W) * 1.1
    g = torch.randn(B, cout, H, W)
    xd = x.double().requires_grad_(True)
    ps = [p.detach().double().requires_grad_(True) for p in (kl.base_weight, kl.spline_weight, kl.spline_scaler)]
    y = O.kanconv2d(xd, ps[0], ps[1], ps[2], kl.grid, 3, 1, 1)
    y.backward(g.double())
    _lib.lib().kmu_debug_flags(flags)
    m = m.cuda()
    xc = x.cuda().requires_grad_(True)
    yc = m(xc)
    yc.backward(g.cuda())
    torch.cuda.synchronize()
    kl = m.kanlayer
    return dict(y=rel(yc, y), dx=rel(xc.grad, xd.grad), dbw=rel(kl.base_weight.grad, ps[0].grad),
                dsw=rel(kl.spline_weight.grad, ps[1].grad), dsc=rel(kl.spline_scaler.grad, ps[2].grad))


if __name__ == "__main__":
    shapes = [(1, 16, 16, 16, 8), (2, 16, 16, 32, 32), (1, 16, 32, 17, 13), (2, 32, 64, 32, 32), (1, 64, 32, 32, 32),
              (1, 64, 64, 48, 40)]
    for flags in (0, 1):
        for s in shapes:
            try:
                r = run(*s, flags)
                print("flags", flags, s, {k: "%.2e" % v for k, v in r.items()}, flush=True)
            except Exception as e:  # noqa: BLE001
                print("flags", flags, s, "FAILED", repr(e)[:300], flush=True)
                raise
