"""Host-side cost of replaying the captured training step: how long `cudaGraphLaunch` keeps the launching thread busy per step,
next to the device time of the step -- with N ranks on one box the N launching threads share the host's cores, so a launch
cost close to the step time would make the step host-bound at N = 8 while it is not at N = 1.

    python tools/graph_launch_cost.py [--batch 32] [--steps 20] [--busy K]

--busy K starts K busy-spinning Python processes first (a stand-in for the other ranks' host threads).
"""
import argparse
import multiprocessing as mp
import os
import sys
import time

import torch

sys.path.insert(0, ".")
import km_unet_b200 as K  # noqa: E402
from km_unet_b200.loss import HybridLoss  # noqa: E402
from km_unet_b200.train import GraphedTrainStep  # noqa: E402


def _spin(stop):
    while not stop.value:
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--busy", type=int, default=0)
    a = ap.parse_args()
    K.config.kan_precision = K.config.hsm_precision = "bf16"
    K.config.conv_bwd, K.config.conv_fwd = "fused", "tma"
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.manual_seed(1234)
    model = K.KM_UNetV3(num_classes=20, variant="SH").cuda().train()
    crit = HybridLoss()
    x = torch.rand(a.batch, 5, 128, 128, device="cuda")
    t = torch.rand(a.batch, 20, 128, 128, device="cuda")
    crit(model(x[:2]), t[:2]).backward()
    live = [p for p in model.parameters() if p.grad is not None]
    for p in model.parameters():
        p.grad = None
    opt = torch.optim.AdamW(live, lr=1e-3, weight_decay=0.05, fused=True, capturable=True)
    step = GraphedTrainStep(model, crit, opt, x, t, world=1, warmup=3)
    print("host cores:", os.cpu_count(), " affinity:", len(os.sched_getaffinity(0)), " torch threads:", torch.get_num_threads())
    stop = mp.Value("i", 0)
    procs = [mp.Process(target=_spin, args=(stop,), daemon=True) for _ in range(a.busy)]
    for p in procs:
        p.start()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host = []
    e0.record()
    w0 = time.perf_counter()
    for _ in range(a.steps):
        h0 = time.perf_counter()
        step()
        host.append(time.perf_counter() - h0)
    w_issue = time.perf_counter() - w0
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / a.steps
    # the same with a synchronise after every step: launch latency no longer hidden behind the previous step
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    for _ in range(a.steps):
        step()
        torch.cuda.synchronize()
    sync_ms = (time.perf_counter() - w0) / a.steps * 1e3
    stop.value = 1
    host.sort()
    print(f"busy={a.busy}  device {dev_ms:.2f} ms/step   host launch median {host[len(host) // 2] * 1e3:.2f} ms  max {host[-1] * 1e3:.2f} ms  "
          f"(issue loop {w_issue / a.steps * 1e3:.2f} ms/step)   step+sync {sync_ms:.2f} ms")
    step.close()


if __name__ == "__main__":
    main()
