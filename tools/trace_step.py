"""GPU timeline of one replay of the captured training step (torch.profiler / CUPTI): every kernel with start, duration and stream,
written to gpurun_out/trace_step.csv -- where the step is idle, what runs concurrently, which chains are serial.

    python tools/trace_step.py [--batch 32] [--variant SH]
"""
import argparse
import csv
import os
import sys

import torch

sys.path.insert(0, ".")
import km_unet_b200 as K  # noqa: E402
from km_unet_b200.loss import HybridLoss  # noqa: E402
from km_unet_b200.train import GraphedTrainStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--variant", default="SH")
a = ap.parse_args()
K.config.kan_precision = K.config.hsm_precision = "bf16"
K.config.conv_bwd, K.config.conv_fwd = "fused", "tma"
torch.backends.cudnn.benchmark = True
torch.backends.cuda.matmul.allow_tf32 = True
classes = 20 if a.variant == "SH" else 3
torch.manual_seed(1234)
model = K.KM_UNetV3(num_classes=classes, variant=a.variant).cuda().train()
crit = HybridLoss()
x = torch.rand(a.batch, 5, 128, 128, device="cuda")
t = torch.rand(a.batch, classes, 128, 128, device="cuda")
crit(model(x[:2]), t[:2]).backward()
live = [p for p in model.parameters() if p.grad is not None]
for p in model.parameters():
    p.grad = None
opt = K.FusedAdamW(live, lr=1e-3, weight_decay=0.05)
step = GraphedTrainStep(model, crit, opt, x, t, world=1, warmup=3)
for _ in range(5):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    step()
    torch.cuda.synchronize()
rows = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        rows.append((ev.time_range.start, ev.time_range.end - ev.time_range.start, getattr(ev, "device_resource_id", -1) if hasattr(ev, "device_resource_id") else -1, ev.name))
rows.sort()
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/trace_step.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["start_us", "dur_us", "stream", "name"])
    t0 = rows[0][0] if rows else 0
    for s, d, st, n in rows:
        w.writerow([f"{s - t0:.3f}", f"{d:.3f}", st, n[:120]])
print("kernels:", len(rows))
try:
    prof.export_chrome_trace("gpurun_out/trace_step.json")
except Exception as e:  # noqa: BLE001
    print("chrome trace export failed:", e)
step.close()
