"""Stage-by-stage bring-up of the captured-NCCL training step (run under torchrun with a short timeout)."""
import faulthandler
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(int(os.environ.get("DUMP_AFTER", "60")), exit=True)
rank = int(os.environ["RANK"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)


def say(*a):
    print(f"[r{rank} {time.time() % 1000:.1f}]", *a, flush=True)


mode = sys.argv[1] if len(sys.argv) > 1 else "tiny"
if mode == "tiny":
    net = torch.nn.Sequential(torch.nn.Linear(64, 64), torch.nn.ReLU(), torch.nn.Linear(64, 8)).to(dev)
    from km_unet_b200.ddp import BucketedGradAllReduce
    red = BucketedGradAllReduce(list(net.parameters()), bucket_bytes=1024)
    x = torch.randn(16, 64, device=dev)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            net.zero_grad(set_to_none=True)
            net(x).square().mean().backward()
            red.finish()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    say("eager ok; capturing", len(red.buckets), "buckets")
    g = torch.cuda.CUDAGraph()
    net.zero_grad(set_to_none=True)
    with torch.cuda.graph(g, capture_error_mode=os.environ.get("CAPMODE", "global")):
        net.zero_grad(set_to_none=True)
        net(x).square().mean().backward()
        red.finish()
    say("captured")
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    say("replayed; grad sum", float(net[0].weight.grad.sum()))
    del g
else:
    import km_unet_b200 as K
    from km_unet_b200.loss import HybridLoss
    from km_unet_b200.train import GraphedTrainStep
    K.config.kan_precision = K.config.hsm_precision = "bf16"
    K.config.conv_bwd, K.config.conv_fwd = "fused", "tma"
    torch.manual_seed(0)
    model = K.KM_UNetV3(num_classes=4, variant="SH").to(dev).train()
    x = torch.rand(2, 5, 64, 64, device=dev)
    t = torch.rand(2, 4, 64, 64, device=dev)
    crit = HybridLoss()
    crit(model(x), t).backward()
    live = [p for p in model.parameters() if p.grad is not None]
    for p in model.parameters():
        p.grad = None
    opt = torch.optim.AdamW(live, lr=1e-3, fused=True, capturable=True)
    say("building graphed step", mode)
    step = GraphedTrainStep(model, crit, opt, x, t, world=2, warmup=3, comm=mode)
    say("built")
    for _ in range(3):
        loss = step()
    torch.cuda.synchronize()
    say("replayed, loss", float(loss))
    step.close()
import gc
gc.collect()
torch.cuda.synchronize()
dist.barrier()
say("destroying")
dist.destroy_process_group()
say("done")
