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
K.config.conv_bwd = "fused"
K.config.conv_fwd = "tma"
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.benchmark = os.environ.get("KMU_CUDNN_BENCHMARK", "1") == "1"   # train_shanghai.py:331
torch.manual_seed(1234)
m = K.KM_UNetV3_SH(num_classes=20).cuda().train()
crit = HybridLoss()
g = torch.Generator().manual_seed(20240518)
data = torch.rand(B, 25, 128, 128, generator=g).cuda()
x, t = data[:, :5].contiguous(), data[:, 5:].contiguous()
crit(m(x[:2]), t[:2]).backward()
live = [p for p in m.parameters() if p.grad is not None]
opt = K.FusedAdamW(live, lr=1e-3, weight_decay=0.05)          # as bench.py
def step():
    opt.zero_grad(set_to_none=True)
    loss = crit(m(x), t)
    loss.backward()
    opt.step()
    return loss


for i in range(steps):
    if i == steps - 1 and os.environ.get("KMU_PROFILE_LAST_STEP"):   # ncu --profile-from-start off: only the last step is profiled
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    loss = step()
torch.cuda.synchronize()
if os.environ.get("KMU_PROFILE_LAST_STEP"):
    torch.cuda.profiler.stop()
if os.environ.get("KMU_ATEN_PROFILE"):
    # ATen-level attribution of the non-libkmunet part of the step (torch.profiler, device time per op and input shape)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        step()
        torch.cuda.synchronize()
    with open(os.environ["KMU_ATEN_PROFILE"], "w") as f:
        f.write(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=90, max_name_column_width=60,
                                                                    max_shapes_column_width=70))
print("ok", float(loss))
