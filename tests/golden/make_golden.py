"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules (/root/reference) on CPU.

    python tests/golden/make_golden.py

Each fixture stores: the module's state_dict ("sd/<key>"), the seeded input(s), the forward output(s)
and the autograd gradients w.r.t. the input and every parameter ("grad/<key>") for the upstream
gradient "gout".  Everything is fp32 computed by the reference in fp32 on CPU (torch 2.11), except the
"*_f64" fixtures which run the same module after .double() to give a tighter pin for closed forms.
The reference tree cannot travel to the GPU box, these small vectors can.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_loader  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def dump(name, module, inputs, train=True, double=False, out_index=None, extra=None):
    if double:
        module = module.double()
        inputs = [i.double() for i in inputs]
    module.train(train)
    sd_before = {k: v.clone() for k, v in module.state_dict().items()}
    inputs = [i.clone().requires_grad_(True) for i in inputs]
    out = module(*inputs)
    outs = out if isinstance(out, (tuple, list)) else (out,)
    main = outs[0] if out_index is None else outs[out_index]
    g = torch.Generator().manual_seed(99)
    gout = torch.randn(main.shape, generator=g, dtype=torch.float32).to(main.dtype)
    main.backward(gout)
    rec = {}
    for k, v in sd_before.items():
        rec["sd/" + k] = _np(v)
    for k, v in module.state_dict().items():
        if "running_" in k or "num_batches" in k:
            rec["sd_after/" + k] = _np(v)
    for i, t in enumerate(inputs):
        rec[f"in{i}"] = _np(t)
        rec[f"grad_in{i}"] = _np(t.grad)
    for i, t in enumerate(outs):
        rec[f"out{i}"] = _np(t)
    rec["gout"] = _np(gout)
    for k, p in module.named_parameters():
        if p.grad is not None:
            rec["grad/" + k] = _np(p.grad)
    if extra:
        rec.update(extra)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB, out {tuple(main.shape)}")


def main():
    R = ref_loader.load(with_models=True)
    torch.manual_seed(1234)
    g = torch.Generator().manual_seed(20240518)

    # ---- KANLinear / KANConv2d (K1-K4) ------------------------------------------------------
    lin = R.KANLinear(6, 5)
    x = torch.randn(7, 6, generator=g) * 1.3          # some samples leave [-2.2, 2.2): zero-basis branch
    x[0, 0], x[0, 1], x[0, 2] = -2.2, 2.2, 0.0        # knot / range edges
    dump("kanlinear_6_5", lin, [x])
    dump("kanlinear_6_5_f64", R.KANLinear(6, 5), [x], double=True)
    conv = R.KANConv2d(4, 8, 3, padding=1)
    dump("kanconv2d_4_8_k3p1", conv, [torch.randn(2, 4, 8, 8, generator=g)])
    conv = R.KANConv2d(3, 5, 3, stride=2, padding=0)
    dump("kanconv2d_3_5_k3s2p0", conv, [torch.randn(2, 3, 9, 7, generator=g)])
    conv = R.KANConv2d(16, 16, 3, padding=1)
    dump("kanconv2d_16_16_k3p1", conv, [torch.randn(1, 16, 16, 16, generator=g) * 0.8])
    with torch.no_grad():
        bases0 = lin.float().b_splines(torch.zeros(1, 6))
    np.savez_compressed(os.path.join(HERE, "kan_phi0.npz"), bases0=_np(bases0), grid=_np(lin.grid))

    # ---- HSMSSD / LayerNorm1D / EfficientViMBlock (S1-S3) -----------------------------------
    mixer = R.HSMSSD(d_model=16)
    dump("hsmssd_16_L64", mixer, [torch.randn(2, 16, 64, generator=g)])
    dump("hsmssd_16_L64_f64", R.HSMSSD(d_model=16), [torch.randn(2, 16, 64, generator=g)], double=True)
    mixer = R.HSMSSD(d_model=32)
    dump("hsmssd_32_L144", mixer, [torch.randn(1, 32, 144, generator=g)])
    ln = R.LayerNorm1D(16)
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.uniform_(-0.3, 0.3)
    dump("layernorm1d_16", ln, [torch.randn(2, 16, 40, generator=g)])
    for tag, train in (("train", True), ("eval", False)):
        blk = R.EfficientViMBlock(dim=16)
        with torch.no_grad():                         # default init zeroes three BN gammas: make every branch live
            for m in blk.modules():
                if isinstance(m, torch.nn.BatchNorm2d):
                    m.weight.uniform_(0.5, 1.5)
                    m.bias.uniform_(-0.2, 0.2)
                    m.running_mean.uniform_(-0.1, 0.1)
                    m.running_var.uniform_(0.5, 1.5)
            blk.alpha.uniform_(-1.0, 1.0)
        dump(f"vimblock_16_{tag}", blk, [torch.randn(2, 16, 8, 8, generator=g)], train=train)
    dump("vimblock_16_init", R.EfficientViMBlock(dim=16), [torch.randn(2, 16, 8, 8, generator=g)])

    # ---- DySample (D1-D2) -------------------------------------------------------------------
    dy = R.DySample(8, scale=2, style="lp", groups=4)
    with torch.no_grad():
        dy.offset.weight.normal_(0, 0.3)              # init std 0.001 barely moves the samples: stress the gather
        dy.offset.bias.uniform_(-0.5, 0.5)
    dump("dysample_8_g4", dy, [torch.randn(2, 8, 6, 5, generator=g)])
    dump("dysample_64_init", R.DySample(64), [torch.randn(1, 64, 4, 4, generator=g)])

    # ---- DAGEM (G1) -------------------------------------------------------------------------
    for tag, train in (("train", True), ("eval", False)):
        dg = R.DAGEM(input_channels=8)
        with torch.no_grad():
            dg.offset_conv.weight.mul_(3.0)
            for m in dg.modules():
                if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                    m.running_mean.uniform_(-0.1, 0.1)
                    m.running_var.uniform_(0.5, 1.5)
        dump(f"dagem_8_{tag}", dg, [torch.randn(2, 8, 6, 6, generator=g)], train=train)

    # ---- whole model, config 1 scaled down (end-to-end forward pin for the drop-in) -----------
    torch.manual_seed(1234)
    model = R.KM_UNetV3_SH(num_classes=4).eval()
    xin = torch.rand(1, 5, 32, 32, generator=g)
    with torch.no_grad():
        out = model(xin)
    rec = {"in0": _np(xin), "out0": _np(out)}
    for k, v in model.state_dict().items():
        rec["sd/" + k] = _np(v)
    np.savez_compressed(os.path.join(HERE, "km_unetv3_sh_eval_32.npz"), **rec)
    print("km_unetv3_sh_eval_32:", os.path.getsize(os.path.join(HERE, "km_unetv3_sh_eval_32.npz")) / 1024, "KiB")


if __name__ == "__main__":
    main()
