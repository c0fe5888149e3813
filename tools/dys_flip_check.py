"""Are the isolated dX / dW differences of the 128x128 DySample case derivative jumps at integer sample coordinates?"""
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
off = torch.nn.functional.conv2d(x.double(), wd.detach(), bd.detach()) * 0.25 + m.init_pos.double()      # (1, 32, H, W): x offsets then y offsets
gx = torch.arange(W, dtype=torch.float64).view(1, 1, 1, W) + off[:, :16]
gy = torch.arange(H, dtype=torch.float64).view(1, 1, H, 1) + off[:, 16:]
dist = torch.minimum((gx - gx.round()).abs(), (gy - gy.round()).abs())          # distance of each sample to the nearest grid line
near = (dist < 2e-5)
print("samples", dist.numel(), "within 2e-5 of a grid line:", int(near.sum()))
mc = m.cuda()
xc = x.cuda().requires_grad_(True)
mc(xc).backward(gout.cuda())
err = (xc.grad.double().cpu() - xd.grad).abs().amax(dim=1)[0]                    # (H, W) max over channels
bad = err > 1e-4 * xd.grad.abs().max()
print("pixels with a dX difference:", int(bad.sum()), "of", bad.numel())
near_px = near.any(dim=1)[0]                                                       # input pixels owning a near-grid sample
# a flipped sample of pixel (h, w) changes doffset at (h, w) -> dX at (h, w) (all channels, through the 1x1 conv) and the 4 taps around the sample
dil = torch.nn.functional.max_pool2d(near_px[None, None].double(), 5, 1, 2)[0, 0] > 0
print("bad pixels explained by a near-grid sample within 2 pixels:", int((bad & dil).sum()), "unexplained:", int((bad & ~dil).sum()))
