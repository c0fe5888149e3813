"""Which tensor-core family contributes what to the end-to-end error of the bf16 precision class (vs the fp64 reference fixture)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import train_fixture as TF  # noqa: E402
import km_unet_b200 as K  # noqa: E402
from km_unet_b200.loss import HybridLoss  # noqa: E402
from km_unet_b200.modules import km_unet as MM  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
CFG = {
    "fp32": {},
    "kan": {"kan_precision": "bf16"},
    "hsm": {"hsm_precision": "bf16"},
    "conv_fwd_tma": {"conv_fwd": "tma"},
    "conv_bwd_fused": {"conv_bwd": "fused"},
    "tf32_matmul": {"tf32": True},
    "all": {"kan_precision": "bf16", "hsm_precision": "bf16", "conv_fwd": "tma", "conv_bwd": "fused", "tf32": True},
}
for tag in sys.argv[1:] or ["sh", "laps"]:
    variant, classes = TF.VARIANTS[tag]
    z = np.load(os.path.join(ROOT, "tests", "golden", f"km_unetv3_{tag}_train_128.npz"))
    for name, cfg in CFG.items():
        K.config.kan_precision = K.config.hsm_precision = K.config.conv_precision = "fp32"
        K.config.conv_bwd, K.config.conv_fwd = "split", "simt"
        torch.backends.cuda.matmul.allow_tf32 = bool(cfg.get("tf32"))
        for k, v in cfg.items():
            if k != "tf32":
                setattr(K.config, k, v)
        torch.manual_seed(TF.SEED_WEIGHTS)
        model = K.KM_UNetV3(num_classes=classes, variant=variant)
        TF.perturb_(model)
        model = model.cuda().train()
        x, t = TF.make_batch(classes)
        with TF.DropPathReplayer(MM.DropPath, list(z["masks"])):
            out = model(x.cuda())
            loss = HybridLoss()(out, t.cuda())
            loss.backward()
        grads = {k: p.grad for k, p in model.named_parameters()}
        e = TF.grad_errors(grads, z)
        want = z["out0"].astype(np.float64)
        oe = np.abs(out.detach().double().cpu().numpy() - want)
        print(f"{tag:5s} {name:15s} out max {oe.max() / np.abs(want).max():.2e} mean {oe.mean() / np.abs(want).mean():.2e}  grad L2 {TF.grad_global_l2(grads, z):.2e} "
              f"median {np.median([v[0] for v in e.values()]):.2e} max {max(v[0] for v in e.values()):.2e}", flush=True)
