"""Run N eager training steps of the bench model and print the loss / the first non-finite tensor (debug aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import km_unet_b200 as K
from km_unet_b200.loss import HybridLoss

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
K.config.kan_precision = K.config.hsm_precision = "bf16"
K.config.conv_bwd = os.environ.get("KMU_CONV_BWD", "fused")
K.config.conv_fwd = os.environ.get("KMU_CONV_FWD", "tma")
torch.backends.cudnn.benchmark = True
torch.manual_seed(1234)
m = K.KM_UNetV3_SH(num_classes=20).cuda().train()
crit = HybridLoss()
g = torch.Generator().manual_seed(20240518)
data = torch.rand(32, 25, 128, 128, generator=g).cuda()
x, t = data[:, :5].contiguous(), data[:, 5:].contiguous()
crit(m(x[:2]), t[:2]).backward()
live = [p for p in m.parameters() if p.grad is not None]
opt = torch.optim.AdamW(live, lr=1e-3, weight_decay=0.05, fused=True)
for i in range(steps):
    opt.zero_grad(set_to_none=True)
    loss = crit(m(x), t)
    loss.backward()
    bad = [n for n, p in m.named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
    print(i, float(loss), "non-finite grads: %d %s" % (len(bad), bad[:3]), flush=True)
    if bad or not torch.isfinite(loss):
        break
    opt.step()
print("done")

if os.environ.get("KMU_CHECK_GRAPH"):
    from km_unet_b200.train import GraphedTrainStep
    opt2 = torch.optim.AdamW(live, lr=1e-3, weight_decay=0.05, fused=True, capturable=True)
    gs = GraphedTrainStep(m, crit, opt2, x, t, world=1, warmup=3)
    for i in range(30):
        loss = gs()
        torch.cuda.synchronize()
        bad = [n for n, p in m.named_parameters() if not torch.isfinite(p).all()]
        print("graph", i, float(loss), "non-finite params: %d %s" % (len(bad), bad[:3]), flush=True)
        if bad:
            break
