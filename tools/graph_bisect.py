"""Find the first sub-module of KM_UNetV3 whose output differs between an eager forward and a CUDA-graph replay of it.

    python tools/graph_bisect.py [SH|LAPS] [fp32|bf16] [--bwd]
"""
import sys

import torch

sys.path.insert(0, ".")
import km_unet_b200 as K  # noqa: E402
from km_unet_b200.modules import km_unet as MM  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "SH"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
classes = 20 if variant == "SH" else 3
if prec == "bf16":
    K.config.kan_precision = K.config.hsm_precision = "bf16"
    K.config.conv_bwd, K.config.conv_fwd = "fused", "tma"
    torch.backends.cuda.matmul.allow_tf32 = True
torch.manual_seed(1234)
model = K.KM_UNetV3(num_classes=classes, variant=variant).cuda().train()
for m in model.modules():
    if isinstance(m, MM.DropPath):
        m.drop_prob = 0.0
x = torch.rand(2, 5, 128, 128, device="cuda")
state0 = {k: v.clone() for k, v in model.state_dict().items()}


def restore():
    with torch.no_grad():
        for k, v in model.state_dict().items():
            if not torch.equal(v, state0[k]):
                v.copy_(state0[k])


names = {m: n for n, m in model.named_modules()}
rec = {}


def hook(mod, inp, out):
    if torch.is_tensor(out):
        rec[names[mod]] = out
    elif isinstance(out, (tuple, list)) and torch.is_tensor(out[0]):
        rec[names[mod]] = out[0]


hs = [m.register_forward_hook(hook) for m in model.modules()]
with torch.no_grad():
    model(x)                                            # lazy init
    restore()
    model(x)
    eager = {k: v.clone() for k, v in rec.items()}
    order = list(rec.keys())
    restore()
    rec.clear()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        model(x)
    torch.cuda.current_stream().wait_stream(s)
    restore()
    rec.clear()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        model(x)
    graph_rec = dict(rec)
    restore()
    g.replay()
    torch.cuda.synchronize()
bad = 0
for k in order:
    a, b = eager[k].float(), graph_rec[k].float()
    err = ((a - b).abs().max() / a.abs().max().clamp_min(1e-30)).item()
    if err > 1e-6:
        print(f"DIFF {err:.3e} {k} ({type(dict(model.named_modules())[k]).__name__})")
        bad += 1
        if bad > 12:
            break
print("modules compared", len(order), "diffs", bad)
