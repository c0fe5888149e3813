"""torch.profiler breakdown of one full-model training step: python tools/prof_model.py [B] [size]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import km_unet_b200 as K
from km_unet_b200.loss import HybridLoss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
K.config.kan_precision = "bf16"
torch.manual_seed(0)
m = K.KM_UNetV3_SH(num_classes=20).cuda().train()
crit = HybridLoss()
x = torch.rand(B, 5, S, S, device="cuda")
t = torch.rand(B, 20, S, S, device="cuda")
crit(m(x[:2]), t[:2]).backward()
live = [p for p in m.parameters() if p.grad is not None]
opt = torch.optim.AdamW(live, lr=1e-3, weight_decay=0.05, fused=True)


def step():
    opt.zero_grad(set_to_none=True)
    loss = crit(m(x), t)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile, record_function
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=False) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
rows = [(e.key, e.self_device_time_total / 2000.0, e.count // 2) for e in prof.key_averages() if e.self_device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"GPU ms per step {tot:.2f}")
for k, ms, n in rows[:70]:
    print(f"{ms:8.3f} ms {n:5d}x  {k[:110]}")
# module-level attribution with hooks
import collections
tot = collections.defaultdict(float)
evs = []


def hook_pair(name):
    def pre(mod, inp):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        mod._t0 = e

    def post(mod, inp, out):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        evs.append((name, mod._t0, e))
    return pre, post


for name, mod in m.named_modules():
    depth = name.count(".")
    if name and depth <= 1:
        pre, post = hook_pair(name)
        mod.register_forward_pre_hook(pre)
        mod.register_forward_hook(post)
with torch.no_grad():
    m(x)
evs.clear()
a = torch.cuda.Event(enable_timing=True)
b = torch.cuda.Event(enable_timing=True)
a.record()
with torch.no_grad():
    m(x)
b.record()
torch.cuda.synchronize()
print("forward total ms", a.elapsed_time(b))
for name, e0, e1 in evs:
    print(f"{name:30s} {e0.elapsed_time(e1):8.3f} ms")
