"""The UNMODIFIED reference model file running on the drop-in operators: `KM_UNetV3_SH.py` /
`KM_UNetV3_LAPS.py` are imported from the byte-for-byte mirror oracle/_ref/ (recipe: oracle/make_ref.py) with
km_unet_b200.enable_dropin() in front, so their `from convKAN.KANConv2Dlayers import *`, `from vim_block_init.efficient_vim_init
import EfficientViMBlock`, `from DAGEM_md import DAGEM`, `from DySample_md import DySample` resolve to the CUDA-backed modules
while every other line of the model (StableHybridKANConv, EnhancedViMBlock, DirectionViM, TripleNorm, MultiScaleFusion, IWP with
its numpy matrices, ...) is the reference's own torch code on the GPU.  fp16 autocast is neutralised on both sides (section 8c) -- except
in the last test, which runs the file exactly as shipped: its `@autocast()` forwards (KM_UNetV3_SH.py:71,306,327,465) inside the fp16
autocast + GradScaler step of train_shanghai.py:159-181.

Checked against (i) the fp64 fixture of the same reference file on its own operators (tests/golden/km_unetv3_*_train_128.npz)
and (ii) the same reference file on its own operators run live on the same GPU in fp32.  Also K5: the reference's
StableHybridKANConv over the drop-in KANConv2d vs over its own.
"""
import json
import os

import numpy as np
import pytest
import torch

import train_fixture as TF
from conftest import GOLDEN, ROOT, rel_err
from oracle import ref_loader, shims

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="no reference tree and no oracle/_ref mirror")]


def _step(cls, classes, masks, loss_fn):
    torch.manual_seed(TF.SEED_WEIGHTS)
    model = cls(num_classes=classes)
    TF.perturb_(model)
    model = model.cuda().train()
    x, t = TF.make_batch(classes)
    x, t = x.cuda(), t.cuda()
    with TF.DropPathReplayer(shims._DropPath, masks):          # both imports take DropPath from the timm stand-in (KM_UNetV3_SH.py:7)
        out = model(x)
        loss = loss_fn(out, t)
        loss.backward()
    torch.cuda.synchronize()
    return model, out.detach(), loss.detach()


@pytest.mark.parametrize("tag", ["sh", "laps"])
def test_unmodified_reference_model_file_trains_on_the_dropin_operators(tag):
    from oracle import loss as OL
    variant, classes = TF.VARIANTS[tag]
    z = np.load(os.path.join(GOLDEN, f"km_unetv3_{tag}_train_128.npz"))
    masks = list(z["masks"])
    D = ref_loader.load_models(dropin=True, autocast=False)
    R = ref_loader.load_models(dropin=False, autocast=False)
    dcls = D.KM_UNetV3_SH if variant == "SH" else D.KM_UNetV3_LAPS
    rcls = R.KM_UNetV3_SH if variant == "SH" else R.KM_UNetV3_LAPS
    assert dcls.__module__ == rcls.__module__ and D.KANConv2d.__module__.startswith("km_unet_b200") \
        and R.KANConv2d.__module__.startswith("convKAN")
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        ours, out_o, loss_o = _step(dcls, classes, masks, OL.hybrid_loss)
        TF.assert_same_state({k: v for k, v in ours.state_dict().items() if "running_" not in k and "num_batches" not in k},
                             {"cs/" + k: z["cs/" + k] for k in ours.state_dict()})
        ref, out_r, loss_r = _step(rcls, classes, masks, OL.hybrid_loss)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    g_o = {k: p.grad for k, p in ours.named_parameters()}
    g_r = {k: p.grad for k, p in ref.named_parameters()}
    e_o, e_r = TF.grad_errors(g_o, z), TF.grad_errors(g_r, z)
    med_o, med_r = float(np.median([v[0] for v in e_o.values()])), float(np.median([v[0] for v in e_r.values()]))
    want = z["out0"].astype(np.float64)
    rep = {"dropin_vs_fp64": {"out": rel_err(out_o, torch.from_numpy(want)), "loss": abs(loss_o.item() - float(z["loss"])) / float(z["loss"]),
                              "grad_l2": TF.grad_global_l2(g_o, z), "grad_median": med_o, "grad_max": max(v[0] for v in e_o.values())},
           "reference_gpu_fp32_vs_fp64": {"out": rel_err(out_r, torch.from_numpy(want)), "loss": abs(loss_r.item() - float(z["loss"])) / float(z["loss"]),
                                          "grad_l2": TF.grad_global_l2(g_r, z), "grad_median": med_r, "grad_max": max(v[0] for v in e_r.values())},
           "dropin_vs_reference_gpu": {"out": rel_err(out_o, out_r), "loss": abs(loss_o.item() - loss_r.item()) / abs(loss_r.item())}}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"parity_{tag}_reference_dropin.json"), "w") as f:
        json.dump(rep, f, indent=1)
    d = rep["dropin_vs_fp64"]
    r = rep["reference_gpu_fp32_vs_fp64"]
    assert d["out"] <= 1e-4 and d["loss"] <= 1e-4, rep
    # gradients: the drop-in must be as close to fp64 as the reference's own fp32 GPU run is (see test_gpu_model_train.py on gates)
    assert d["grad_l2"] <= 1e-4 + 8 * max(r["grad_l2"], float(z["ref32/grad_l2"])), rep
    assert d["grad_median"] <= 1e-4 + 4 * max(r["grad_median"], 1e-4), rep
    ref32 = {k[7:]: float(z[k][0]) for k in z.files if k.startswith("ref32g/")}
    bad = {k: v[0] for k, v in e_o.items() if not v[0] <= 1e-4 + 10 * max(ref32[k], e_r[k][0], med_r)}
    assert len(bad) <= 0.02 * len(e_o) and max(v[0] for v in e_o.values()) <= 1e-1, bad     # see test_gpu_model_train.py on the gates


@pytest.mark.parametrize("cin,cout", [(16, 16), (16, 32), (64, 32)])
def test_k5_reference_stable_hybrid_kanconv_on_dropin_kanconv2d(cin, cout):
    """KM_UNetV3_SH.py:72-94: ReLU(residual(GN4(x)) + KANConv2d(GN4(x))) -- the reference's class over our KANConv2d (GPU) vs
    over its own KANConv2d (CPU fp64)."""
    D = ref_loader.load_models(dropin=True, autocast=False)
    R = ref_loader.load_models(dropin=False, autocast=False)
    torch.manual_seed(5)
    ref = R.sh_module.StableHybridKANConv(cin, cout)
    ours = D.sh_module.StableHybridKANConv(cin, cout)
    with torch.no_grad():
        ref.pre_norm.weight.uniform_(0.5, 1.5)
        ref.pre_norm.bias.uniform_(-0.3, 0.3)
    ours.load_state_dict(ref.state_dict())
    ref = ref.double()
    ours = ours.cuda()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, cin, 24, 24, generator=g) * 1.5 + 0.3
    gout = torch.randn(2, cout, 24, 24, generator=g)
    xr = x.double().requires_grad_(True)
    yr = ref(xr)
    yr.backward(gout.double())
    xo = x.cuda().requires_grad_(True)
    yo = ours(xo)
    yo.backward(gout.cuda())
    assert rel_err(yo, yr) <= 1e-4
    assert rel_err(xo.grad, xr.grad) <= 1e-4
    live = {k: p.grad for k, p in ref.named_parameters() if p.grad is not None}
    got = dict(ours.named_parameters())
    assert any("kanlayer.spline_weight" in k for k in live)
    for k, want in live.items():
        assert got[k].grad is not None, k
        assert rel_err(got[k].grad, want) <= 1e-4, k


def _amp_step(cls, classes, masks, loss_fn):
    """One iteration of train_shanghai.py:159-181: autocast forward + loss, scaled backward, scaler.step(AdamW), scaler.update()."""
    torch.manual_seed(TF.SEED_WEIGHTS)
    model = cls(num_classes=classes)
    TF.perturb_(model)
    model = model.cuda().train()
    before = {k: p.detach().clone() for k, p in model.named_parameters()}
    x, t = TF.make_batch(classes)
    x, t = x.cuda(), t.cuda()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.05)          # train_shanghai.py:342
    scaler = torch.amp.GradScaler("cuda")
    skipped = 0
    for _ in range(6):                  # GradScaler starts at 65536: it may skip (and halve) a few times before the first real step
        scale = scaler.get_scale()
        with TF.DropPathReplayer(shims._DropPath, masks):
            opt.zero_grad()
            with torch.autocast("cuda", dtype=torch.float16):
                out = model(x)
                loss = loss_fn(out, t)
            scaler.scale(loss).backward()
            scaler.step(opt)                                                        # unscales .grad in place, skips on inf / nan
            scaler.update()
        if scaler.get_scale() >= scale:
            break
        skipped += 1
    torch.cuda.synchronize()
    moved = sum(int(not torch.equal(p.detach(), before[k])) for k, p in model.named_parameters())
    return model, out.detach().float(), loss.detach().float(), moved, skipped


@pytest.mark.parametrize("tag", ["sh", "laps"])
def test_reference_model_file_as_shipped_fp16_autocast_and_gradscaler_on_the_dropin(tag):
    """What a user of the reference gets after switching: nothing neutralised.  The drop-in operators compute in fp32 / bf16-tcgen05
    inside the autocast region (custom_fwd casts their inputs); everything else is the reference's fp16 torch code.  Compared with
    the same file on its own operators under the same autocast (fp16 everywhere) and with the fp64 fixture."""
    from oracle import loss as OL
    variant, classes = TF.VARIANTS[tag]
    z = np.load(os.path.join(GOLDEN, f"km_unetv3_{tag}_train_128.npz"))
    masks = list(z["masks"])
    D = ref_loader.load_models(dropin=True, autocast=True)
    R = ref_loader.load_models(dropin=False, autocast=True)
    dcls = D.KM_UNetV3_SH if variant == "SH" else D.KM_UNetV3_LAPS
    rcls = R.KM_UNetV3_SH if variant == "SH" else R.KM_UNetV3_LAPS
    ours, out_o, loss_o, moved_o, skipped_o = _amp_step(dcls, classes, masks, OL.hybrid_loss)
    ref, out_r, loss_r, moved_r, skipped_r = _amp_step(rcls, classes, masks, OL.hybrid_loss)
    assert out_o.shape == out_r.shape and bool(torch.isfinite(out_o).all()) and bool(torch.isfinite(loss_o))
    # a real step was taken (GradScaler found no inf / nan after at most as many halvings as the reference's own operators need + 1)
    # and it moved every live parameter
    live = [k for k, p in ours.named_parameters() if p.grad is not None and bool(p.grad.any())]
    assert skipped_o <= skipped_r + 1 and skipped_o < 6 and moved_o >= 0.98 * len(live), (skipped_o, skipped_r, moved_o, moved_r, len(live))
    want = torch.from_numpy(z["out0"].astype(np.float64))
    g_o = {k: p.grad.float() for k, p in ours.named_parameters() if p.grad is not None}
    g_r = {k: p.grad.float() for k, p in ref.named_parameters() if p.grad is not None}
    assert all(bool(torch.isfinite(v).all()) for v in g_o.values())
    rep = {"dropin_amp_vs_fp64": {"out": rel_err(out_o, want), "loss": abs(loss_o.item() - float(z["loss"])) / float(z["loss"]),
                                  "grad_l2": TF.grad_global_l2(g_o, z)},
           "reference_amp_vs_fp64": {"out": rel_err(out_r, want), "loss": abs(loss_r.item() - float(z["loss"])) / float(z["loss"]),
                                     "grad_l2": TF.grad_global_l2(g_r, z)},
           "dropin_amp_vs_reference_amp": {"out": rel_err(out_o, out_r), "loss": abs(loss_o.item() - loss_r.item()) / abs(loss_r.item())},
           "gradscaler_skips": {"dropin": skipped_o, "reference": skipped_r}}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"parity_{tag}_reference_dropin_amp.json"), "w") as f:
        json.dump(rep, f, indent=1)
    d, r = rep["dropin_amp_vs_fp64"], rep["reference_amp_vs_fp64"]
    # fp16 autocast is the reference's own precision class here: the drop-in run (fp32 operators inside fp16 glue) must be at least as
    # close to fp64 as the reference's fp16 run is, up to the 2e-2 class gate
    assert d["out"] <= max(2e-2, 2 * r["out"]) and d["loss"] <= max(2e-2, 2 * r["loss"]), rep
    assert d["grad_l2"] <= max(5e-2, 2 * r["grad_l2"]), rep
