"""Summarise gpurun_out/trace_step.csv (tools/trace_step.py: torch.profiler / CUPTI timeline of graph replays) as markdown:
how much of the step the GPU is idle, how many kernels run concurrently, and each kernel family's share of the WALL time of the step
(time a kernel runs alone + 1/k of the time it shares the GPU with k - 1 others).

    python tools/trace_md.py gpurun_out/trace_step.csv > profiles/r02_timeline_model_step.md
"""
import collections
import csv
import gzip
import re
import sys

path = sys.argv[1]
op = gzip.open if path.endswith(".gz") else open
with op(path, "rt") as f:
    rows = sorted((float(r["start_us"]), float(r["dur_us"]), r["stream"], r["name"]) for r in csv.DictReader(f))
step = rows[len(rows) // 2:]                       # the trace holds two replays: take the second
t0, t1 = step[0][0], max(s + d for s, d, _, _ in step)
pts = []
for i, (s, d, _, _) in enumerate(step):
    pts.append((s, 1, i))
    pts.append((s + d, -1, i))
pts.sort()
active, last = set(), pts[0][0]
level = collections.Counter()
wall = collections.Counter()
alone = collections.Counter()
for t, dl, i in pts:
    dt = t - last
    if dt > 0:
        level[len(active)] += dt
        for j in active:
            wall[j] += dt / len(active)
        if len(active) == 1:
            alone[next(iter(active))] += dt
    last = t
    if dl == 1:
        active.add(i)
    else:
        active.discard(i)


def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*", "", n)
    n = re.sub(r"^kmu::", "", n)
    return n[:100]


fam = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for i, (s, d, _, n) in enumerate(step):
    e = fam[short(n)]
    e[0] += 1
    e[1] += d
    e[2] += alone[i]
    e[3] += wall[i]
span = t1 - t0
ours = sum(v[3] for k, v in fam.items() if not re.match(r"^(at::|cudnn|cutlass|sm\d|wgrad|implicit|precomputed|detail::|nccl|void|cub|dgrad|conv)", k))
print("# GPU timeline of one replay of the captured training step (KM_UNetV3_SH, B = 32, 128 x 128, bench configuration)\n")
print("Source: `tools/trace_step.py` (torch.profiler / CUPTI activity records of a graph replay, not ncu: kernels run warm, concurrently, as")
print("in the benchmark) summarised by `tools/trace_md.py`; raw records in `r02_timeline_model_step.csv.gz`.\n")
print(f"* kernels in the step: {len(step)}; span {span / 1e3:.2f} ms; sum of kernel durations {sum(d for _, d, _, _ in step) / 1e3:.2f} ms "
      f"on {len(set(st for _, _, st, _ in step))} streams (the graph's parallel branches)")
print(f"* GPU idle (no kernel running): {level[0] / 1e3:.2f} ms = {100 * level[0] / span:.1f} % of the step -- the graph leaves no launch gaps")
print("* time with exactly k kernels running: " + ", ".join(f"k={k}: {v / 1e3:.2f} ms" for k, v in sorted(level.items()) if k > 0 and v > 5))
print(f"* libkmunet kernels' share of the wall time: {100 * ours / span:.1f} %\n")
print("Wall share = time the kernel runs alone + 1/k of the time it shares the GPU with k - 1 other kernels; the column sums to the step.\n")
print("| kernel | launches | sum of durations [us] | alone [us] | wall share [us] | % of step |")
print("|---|---|---|---|---|---|")
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][3])[:40]:
    print(f"| `{k}` | {v[0]} | {v[1]:.0f} | {v[2]:.0f} | {v[3]:.0f} | {100 * v[3] / span:.1f} |")
