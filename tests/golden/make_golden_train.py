"""Generate the END-TO-END train-mode fixtures by running the UNMODIFIED reference model on CPU.

    python tests/golden/make_golden_train.py      # needs /root/reference or oracle/_ref

km_unetv3_sh_train_128.npz   : KM_UNetV3_SH.KM_UNetV3(num_classes=20)   (KM_UNetV3_SH.py:371-517)
km_unetv3_laps_train_128.npz : KM_UNetV3_LAPS.KM_UNetV3(num_classes=3)  (KM_UNetV3_LAPS.py:366-511)
each in train() mode at B = 2, 5 -> classes frames of 128 x 128, one training step's forward + HybridLoss + backward
(train_shanghai.py:159-181 without the fp16 autocast / GradScaler; loss = oracle/loss.py restating :298-326).

The reference runs in **fp64** (model.double()): the fixture is then exact to ~1e-13 and all of a test's error budget belongs
to the implementation under test.  The same step in the reference's fp32 is run as well and its own deviation from the fp64
result is stored (`ref32/...`) for context.  Stored: DropPath masks in call order (input / target are rebuilt from the seed), output, loss, the gradient of
every live parameter (tests/train_fixture.py: whole or sampled + sum / norm), BatchNorm running statistics after the step, and
checksums of the state_dict the step started from (weights are rebuilt from the seed by the tests).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import train_fixture as TF  # noqa: E402
from oracle import loss as OL  # noqa: E402
from oracle import ref_loader, shims  # noqa: E402


def run(cls, classes, dtype, masks=None):
    torch.manual_seed(TF.SEED_WEIGHTS)
    model = cls(num_classes=classes)
    TF.perturb_(model)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.to(dtype).train()
    x, t = TF.make_batch(classes)
    x, t = x.to(dtype), t.to(dtype)
    torch.manual_seed(TF.SEED_DROPPATH)
    if masks is None:
        with TF.DropPathRecorder(shims._DropPath) as rec:
            out = model(x)
        masks = rec.masks
    else:
        with TF.DropPathReplayer(shims._DropPath, masks):
            out = model(x)
    loss = OL.hybrid_loss(out, t)
    loss.backward()
    return model, sd0, x, t, masks, out, loss


def main():
    R = ref_loader.load_models(dropin=False, autocast=False)
    for tag, (variant, classes) in TF.VARIANTS.items():
        cls = R.KM_UNetV3_SH if variant == "SH" else R.KM_UNetV3_LAPS
        model, sd0, x, t, masks, out, loss = run(cls, classes, torch.float64)
        m32, _, _, _, _, out32, loss32 = run(cls, classes, torch.float32, masks)
        rec = {"out0": out.detach().float().numpy(), "loss": np.array(loss.item()),
               "masks": np.stack(masks), "ref32/out_err": np.array(((out32.double() - out).abs().max() / out.abs().max()).item()),
               "ref32/loss_err": np.array(abs(loss32.item() - loss.item()) / abs(loss.item()))}
        for k, v in TF.checksums(sd0).items():
            rec["cs/" + k] = v
        nlive, nelem = 0, 0
        for k, p in model.named_parameters():
            if p.grad is None:
                continue
            nlive += 1
            nelem += p.numel()
            rec["grad/" + k], rec["gstat/" + k] = TF.compress(p.grad)
        rec["gfloor"] = np.array(TF.grad_floor([np.abs(v).max() for k, v in rec.items() if k.startswith("grad/")]))

        class _Z(dict):                                              # the fixture so far, in the np.load interface grad_errors expects
            files = property(lambda self: list(self))
        e32 = TF.grad_errors({k: p.grad for k, p in m32.named_parameters()}, _Z(rec))
        worst32 = max(v[0] for v in e32.values())
        rec["ref32/grad_err"] = np.array(worst32)
        thr = lambda a: np.stack([(np.clip(a, 0, 1) * 90).astype(np.uint16) >= t for t in (20, 30, 35, 40)])   # metrics.py:45-47,106-107
        rec["ref32/mask_flips"] = np.array(int((thr(out32.detach().numpy()) != thr(out.detach().numpy())).sum()))
        rec["ref32/grad_l2"] = np.array(TF.grad_global_l2({k: p.grad for k, p in m32.named_parameters()}, _Z(rec)))
        for k, v in e32.items():                                     # the reference's own fp32 error per tensor: context for the gates
            rec["ref32g/" + k] = np.array(v)
        for k, v in model.state_dict().items():
            if "running_" in k:
                rec["sd_after/" + k] = v.double().numpy()
        path = os.path.join(HERE, f"km_unetv3_{tag}_train_128.npz")
        np.savez_compressed(path, **rec)
        print(f"{tag}: {os.path.getsize(path) / 1e6:.2f} MB; {nlive} live parameters ({nelem} elements), {len(masks)} DropPath masks, "
              f"loss {loss.item():.6f}; the reference's own fp32 vs fp64: out {rec['ref32/out_err']:.2e} ({int(rec['ref32/mask_flips'])} thresholded-mask cells flip) loss {rec['ref32/loss_err']:.2e} "
              f"grad max {worst32:.2e} median {np.median([v[0] for v in e32.values()]):.2e} global L2 {rec['ref32/grad_l2']:.2e}")


if __name__ == "__main__":
    main()
