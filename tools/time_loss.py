"""Device time of HybridLoss forward + backward at the bench shape (32,20,128,128): python tools/time_loss.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from km_unet_b200.loss import HybridLoss

torch.backends.cuda.matmul.allow_tf32 = True
crit = HybridLoss()
pred = torch.rand(32, 20, 128, 128, device="cuda", requires_grad=True)
tgt = torch.rand(32, 20, 128, 128, device="cuda")
g = torch.cuda.CUDAGraph()
for _ in range(3):
    crit(pred, tgt).backward()
torch.cuda.synchronize()
pred.grad = None
with torch.cuda.graph(g):
    loss = crit(pred, tgt)
    loss.backward()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g.replay()
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    g.replay()
e1.record()
torch.cuda.synchronize()
print("HybridLoss fwd+bwd (graph replay): %.3f ms" % (e0.elapsed_time(e1) / 20), float(loss))
