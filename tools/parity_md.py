"""gpurun_out/parity_*.json (written by tests/test_gpu_model_train.py and tests/test_gpu_reference_dropin.py) -> profiles/rNN_parity.md"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "gpurun_out")


def load(name):
    with open(os.path.join(src, name)) as f:
        return json.load(f)


print("# End-to-end train-mode parity on B200 (round 2)\n")
print("One training step (forward + HybridLoss + backward), B = 2, 128x128, weights from seed 1234 + `tests/train_fixture.perturb_`, DropPath masks "
      "replayed, against the UNMODIFIED reference run in fp64 on CPU (`tests/golden/km_unetv3_{sh,laps}_train_128.npz`).  Errors are "
      "max|got - want| / max(max|want|, floor) per tensor (floor = 0.1 x median gradient scale, for gradients that are mathematically zero); "
      "`ref32` = the reference's own fp32 CPU run against its fp64 run.  Written by `tests/test_gpu_model_train.py` / `tests/test_gpu_reference_dropin.py`.\n")
print("| run | output | loss | BN running stats | grad global L2 (ref32) | grad median (ref32) | grad max | mask cells flipped | CSI/POD/FAR/HSS max diff |")
print("|---|---|---|---|---|---|---|---|---|")
for tag in ("sh", "laps"):
    for cls, label in (("fp32", "mirror, fp32 class"), ("bf16_graphed", "mirror, bench configuration (bf16 / split-bf16 / TF32), CUDA graph replay")):
        r = load(f"parity_{tag}_{cls}.json")
        print(f"| {tag.upper()} {label} | {r['out_err']:.2e} | {r['loss_err']:.2e} | {r['running_stat_err']:.2e} | {r['grad_l2']:.2e} ({r['grad_l2_ref32']:.1e}) | "
              f"{r['grad_err_median']:.2e} ({r['grad_err_median_ref32']:.1e}) | {r['grad_err_max']:.2e} | {r['mask_flips']} / {r['mask_cells']} | {r['score_max_abs_diff']:.1e} |")
print()
print("Graph replay vs the same step run eagerly (bench configuration): output, loss and BatchNorm running statistics are bit-identical; gradients:\n")
print("| model | replay vs eager, global L2 | eager vs eager (same step twice), global L2 | worst tensor (floor-ed scale) |")
print("|---|---|---|---|")
for tag in ("sh", "laps"):
    r = load(f"parity_{tag}_bf16_graphed.json")
    print(f"| {tag.upper()} | {r['replay_vs_eager_grad_l2']:.2e} | {r.get('eager_vs_eager_grad_l2', float('nan')):.2e} | {r['replay_vs_eager_grad']:.2e} |")
print("\n(two eager runs differ because DySample's dX scatter, ATen's bilinear-upsample backward and cuDNN's wgrad use fp32 atomics; in the bf16 class a 1e-7 "
      "difference can flip the bf16 rounding of an operand, and several gradients of this network are cancellations 1e3 deep.)\n")
print("The reference's UNMODIFIED model file on the drop-in operators (`enable_dropin()`), fp32, same fixture; `reference GPU` = the same file on its own operators on the same B200:\n")
print("| model | run | output | loss | grad global L2 | grad median | grad max |")
print("|---|---|---|---|---|---|---|")
for tag in ("sh", "laps"):
    r = load(f"parity_{tag}_reference_dropin.json")
    for k, label in (("dropin_vs_fp64", "reference file + drop-in operators"), ("reference_gpu_fp32_vs_fp64", "reference file + its own operators (GPU fp32)")):
        d = r[k]
        print(f"| {tag.upper()} | {label} | {d['out']:.2e} | {d['loss']:.2e} | {d['grad_l2']:.2e} | {d['grad_median']:.2e} | {d['grad_max']:.2e} |")
print("\nWorst gradients of the bench configuration (SH):\n")
r = load("parity_sh_bf16_graphed.json")
print("| parameter | err | norm err | ref32 err |")
print("|---|---|---|---|")
for w in r["worst"][:8]:
    print(f"| `{w['param']}` | {w['err']:.2e} | {w['norm_err']:.2e} | {w['ref32_err']:.1e} |")
print("\nThe reference's model file AS SHIPPED (its `@autocast()` forwards active) inside the fp16 autocast + GradScaler + AdamW iteration of "
      "`train_shanghai.py:159-181`, on the drop-in operators and on its own; both against the fp64 fixture:\n")
print("| model | run | output | loss | grad global L2 | GradScaler skips before the first step |")
print("|---|---|---|---|---|---|")
for tag in ("sh", "laps"):
    r = load(f"parity_{tag}_reference_dropin_amp.json")
    for k, label, who in (("dropin_amp_vs_fp64", "reference file + drop-in operators, fp16 autocast", "dropin"),
                          ("reference_amp_vs_fp64", "reference file + its own operators, fp16 autocast", "reference")):
        d = r[k]
        print(f"| {tag.upper()} | {label} | {d['out']:.2e} | {d['loss']:.2e} | {d['grad_l2']:.2e} | {r['gradscaler_skips'][who]} |")
