"""Debug aid: repeat the 128x128 DySample backward on the GPU and list where dX / dW differ from the fp64 oracle between runs
(how the derivative jumps at integer sample coordinates were told apart from a real defect; see tools/dys_flip_check.py)."""
import sys
import torch
sys.path.insert(0, ".")
from km_unet_b200 import DySample
from oracle import dysample as O
B, C, H, W, std = 1, 64, 128, 128, 0.1
torch.manual_seed(H * W)
m = DySample(C)
with torch.no_grad():
    m.offset.weight.normal_(0, std)
    m.offset.bias.uniform_(-0.3, 0.3)
x = torch.randn(B, C, H, W)
xd = x.double().requires_grad_(True)
wd = m.offset.weight.detach().double().requires_grad_(True)
bd = m.offset.bias.detach().double().requires_grad_(True)
want = O.dysample_lp(xd, wd, bd, m.init_pos.double())
gout = torch.randn(want.shape)
want.backward(gout.double())
m = m.cuda()
rel = lambda a, b: ((a.double().cpu() - b).abs().max() / b.abs().max()).item()
for it in range(12):
    xc = x.cuda().requires_grad_(True)
    m.zero_grad()
    y = m(xc)
    y.backward(gout.cuda())
    e = (xc.grad.double().cpu() - xd.grad).abs()
    idx = (e == e.max()).nonzero()[0].tolist()
    print(it, "y %.1e dx %.1e dw %.1e db %.1e" % (rel(y, want), rel(xc.grad, xd.grad), rel(m.offset.weight.grad, wd.grad), rel(m.offset.bias.grad, bd.grad)), idx, float(xc.grad[tuple(idx)]), float(xd.grad[tuple(idx)]))
