"""ncu driver: a few eager full-model training steps (same model / loss / optimizer as bench.py, no CUDA graph so that every
kernel is a separate launch).  python tools/prof_model_step.py [steps] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import km_unet_b200 as K
from km_unet_b200.loss import HybridLoss

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
K.config.kan_precision = "bf16"
K.config.hsm_precision = "bf16"
torch.manual_seed(1234)
m = K.KM_UNetV3_SH(num_classes=20).cuda().train()
crit = HybridLoss()
g = torch.Generator().manual_seed(20240518)
data = torch.rand(B, 25, 128, 128, generator=g).cuda()
x, t = data[:, :5].contiguous(), data[:, 5:].contiguous()
crit(m(x[:2]), t[:2]).backward()
live = [p for p in m.parameters() if p.grad is not None]
opt = torch.optim.AdamW(live, lr=1e-3, weight_decay=0.05, fused=True)
for _ in range(steps):
    opt.zero_grad(set_to_none=True)
    loss = crit(m(x), t)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("ok", float(loss))
