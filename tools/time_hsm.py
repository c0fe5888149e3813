"""HSMSSD fwd / bwd device time (CUDA events around the C-ABI calls) and parity vs the fp64 oracle at the model's shapes.

    [KMU_HSM_FUSED=0] python tools/time_hsm.py [fp32|bf16]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import km_unet_b200 as K
from km_unet_b200 import ops
from oracle import hsmssd as OH

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
K.config.hsm_precision = prec
out = {}
for (B, C, S) in [(2, 16, 32), (32, 16, 128), (32, 32, 64), (32, 64, 32), (8, 16, 256)]:
    torch.manual_seed(0)
    m = K.HSMSSD(C).cuda()
    x = (torch.randn(B, C, S * S, device="cuda") * 1.0).requires_grad_(True)
    gout = torch.randn(B, C, S, S, device="cuda")
    err = {}
    if B * S * S <= 64 * 1024:
        sd = {k: v.detach().double().cpu() for k, v in m.state_dict().items()}
        xd = x.detach().double().cpu().requires_grad_(True)
        ps = [sd["BCdt_proj.conv.weight"].requires_grad_(True), sd["dw.conv.weight"].requires_grad_(True),
              sd["hz_proj.conv.weight"].requires_grad_(True), sd["out_proj.conv.weight"].requires_grad_(True)]
        D = sd["D"].requires_grad_(True)
        want, _ = OH.hsmssd(xd, ps[0], ps[1], ps[2], ps[3], sd["A"], D)
        want.backward(gout.double().cpu())
        y, _ = m(x)
        y.backward(gout)
        rel = lambda a, b: ((a.double().cpu() - b).abs().max() / b.abs().max()).item()
        err = {"y": rel(y, want), "dx": rel(x.grad, xd.grad), "dWp": rel(m.BCdt_proj.conv.weight.grad, ps[0].grad),
               "dWd": rel(m.dw.conv.weight.grad, ps[1].grad), "dWhz": rel(m.hz_proj.conv.weight.grad, ps[2].grad),
               "dWo": rel(m.out_proj.conv.weight.grad, ps[3].grad), "dD": rel(m.D.grad, D.grad)}
    for _ in range(3):
        x.grad = None
        y, _ = m(x)
        y.backward(gout)
    ops.profile_start()
    for _ in range(10):
        x.grad = None
        y, _ = m(x)
        y.backward(gout)
    prof = ops.profile_stop()
    t = {name: e["ms"] / e["calls"] for (name, key), e in prof.items()}
    out[f"{B}x{C}x{S * S}"] = {"fwd_ms": t.get("kmu_hsmssd_fwd"), "bwd_ms": t.get("kmu_hsmssd_bwd"), "err": err}
    print(B, C, S * S, {k: round(v, 4) for k, v in t.items() if "hsm" in k}, {k: float("%.2e" % v) for k, v in err.items()}, flush=True)
print(json.dumps(out))
