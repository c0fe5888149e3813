"""One small forward + backward per tcgen05 / TMA kernel family, for compute-sanitizer (memcheck / racecheck).

    compute-sanitizer --tool memcheck python tools/sanitize_cases.py
"""
import sys

import torch

sys.path.insert(0, ".")
import km_unet_b200 as K
from km_unet_b200 import ops

torch.manual_seed(0)
dev = "cuda"
K.config.kan_precision = K.config.hsm_precision = "bf16"
K.config.conv_fwd, K.config.conv_bwd = "tma", "fused"


def run(name, f):
    f()
    torch.cuda.synchronize()
    print("ok", name, flush=True)


def kan():
    m = K.KANConv2d(16, 32, 3, padding=1).to(dev)
    x = torch.randn(1, 16, 32, 32, device=dev, requires_grad=True)
    m(x).sum().backward()


def kan64():
    m = K.KANConv2d(64, 64, 3, padding=1).to(dev)
    x = torch.randn(1, 64, 32, 32, device=dev, requires_grad=True)
    m(x).sum().backward()


def hsm(C, S):
    def f():
        m = K.HSMSSD(C).to(dev)
        x = torch.randn(2, C, S * S, device=dev, requires_grad=True)
        y, _ = m(x)
        y.sum().backward()
    return f


def pw(cin, cout):
    def f():
        w = torch.randn(cout, cin, 1, 1, device=dev, requires_grad=True)
        b = torch.randn(cout, device=dev, requires_grad=True)
        x = torch.randn(2, cin, 32, 32, device=dev, requires_grad=True)
        ops.pwconv(x, w, b).sum().backward()
    return f


def dys():
    m = K.DySample(64).to(dev)
    x = torch.randn(2, 64, 16, 16, device=dev, requires_grad=True)
    m(x).sum().backward()


def dagem():
    m = K.DAGEM(input_channels=64).to(dev)
    x = torch.randn(2, 64, 16, 16, device=dev, requires_grad=True)
    m(x).sum().backward()


run("kan_tc 16->32", kan)
run("kan_tc 64->64", kan64)
run("hsm fused C=16", hsm(16, 32))
run("hsm fused C=32", hsm(32, 16))
run("hsm fused C=64", hsm(64, 16))
run("pw tma/fused 16->64", pw(16, 64))
run("pw tma/fused 64->16", pw(64, 16))
run("pw tma/fused 128->32", pw(128, 32))
run("dysample fused", dys)
run("dagem + deformconv", dagem)
print("all ok")
