"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of tools/prof_model_step.py into a markdown table of
the LAST training step (launches between the last two fused-AdamW groups): python tools/launches_md.py launches.csv > out.md"""
import collections
import csv
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", "")) / 1000.0) for r in csv.DictReader(lines)]
adam = [i for i, (k, _) in enumerate(rows) if "FusedOptimizer" in k or "multi_tensor_apply_kernel" in k]
groups, prev = [], None
for i in adam:                      # consecutive AdamW launches form one group
    if prev is None or i - prev > 50:
        groups.append([i, i])
    groups[-1][1] = i
    prev = i
if len(groups) >= 2:
    lo, hi = groups[-2][1] + 1, groups[-1][1] + 1
else:
    lo, hi = 0, len(rows)
step = rows[lo:hi]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in step:
    key = k.split("(")[0].replace("void ", "")[:100]
    agg[key][0] += 1
    agg[key][1] += v
tot = sum(v for _, v in step)
ours = sum(v[1] for k, v in agg.items() if any(ns in k for ns in ("kan::", "hsm::", "dys", "dagem", "shell::", "pw::", "sc::", "glue::", "tc::", "tcb::", "fused::", "pwtc::", "iwp", "fz::", "deform::", "loss::",
                                                                    "kmu::")))
print("# Kernel launches of one full-model training step (KM_UNetV3_SH, B=32, 128x128, tensor-core precision class as in bench.py)\n")
print(f"Source: `{sys.argv[1].split('/')[-1]}` -- `ncu --metrics gpu__time_duration.sum --clock-control none` over `tools/prof_model_step.py`")
print("(eager execution, one launch per kernel; ncu serialises launches and runs them cold-cache: compare SHARES, not absolutes).")
print(f"\nLast step: {len(step)} launches, {tot / 1000:.2f} ms of kernel time under ncu; libkmunet kernels = {100 * ours / tot:.1f} % of it.\n")
print("| kernel | launches | total [us] | share |")
print("|---|---|---|---|")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"| `{k}` | {c} | {v:.1f} | {100 * v / tot:.1f} % |")
