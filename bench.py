#!/usr/bin/env python
"""bench.py -- headline measurement of the KM-UNet hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload model|laps|infer|kan]
                    [--batch B] [--precision fp32|bf16]

Default workload = BASELINE.json configs[2], the configuration the headline metric "KM_UNetV3_SH train samples/sec at
1/2/4/8 B200" is quoted on: KM_UNetV3_SH(num_classes=20), 5 -> 20 frames of 128x128, B = 32 per GPU, synthetic U[0,1)
frames, random-init weights, train mode.  A "step" = forward + HybridLoss + backward + gradient all-reduce (N > 1) +
AdamW update of one batch.  With N > 1 (torchrun, one rank per GPU) every rank processes its own batch shard (weak
scaling); the only collective is the bucketed NCCL all-reduce of the 5.1 MB of live gradients, launched from grad-ready
hooks so it overlaps the rest of backward.  `--workload laps` = configs[3] (KM_UNetV3_LAPS, 5 -> 3 frames of 256x256),
`--workload infer` = configs[4] (SH eval, 256x256, B = 64), `--workload kan` = configs[1] alone (the KANConv2d 64->64
microbench the "KANConv2D % tensor-core peak" half of the metric is quoted on; it is also attached to every default
line as `kan_microbench`).

One JSON line on stdout (rank 0).  `value` = whole-job samples/s with the batch resident in HBM; `e2e` = the same
metric through the public module API with the batch in pinned host memory (H2D of the 25-frame batch and D2H of the
loss inside the timed region); `roofline` = the C-ABI entry point with the largest share of the step (CUDA events on
the launching stream around every libkmunet call during a profiling pass of the same workload) against the roofline
that bounds it (`roofline.kan` = the KANConv2d microbench fractions of the tensor peak); `cpu_baseline` = the UNMODIFIED
reference model (byte-for-byte mirror oracle/_ref/, recipe oracle/make_ref.py; kind "reference") on the box's host cores,
bounded sample -- or, when the mirror is absent, the CPU oracle port (oracle/model.py; kind "port");
`gpu_eager_reference` = the same unmodified reference run eagerly on the B200 under fp16 autocast + GradScaler exactly as
train_shanghai.py:159-181 does, at the largest batch that fits; `gpu_eager_reference_dropin` = that same unmodified model file and
iteration with its operator imports resolved to the libkmunet drop-in modules (what switching alone buys, no graph, no mirror).
`--impl reference` times the reference's CPU path alone; that arm never imports km_unet_b200.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# stdout carries exactly ONE JSON line.  Libraries that write to file descriptor 1 behind Python's back (NCCL prints its version
# banner there) are sent to stderr: fd 1 is re-pointed at fd 2 and the JSON line goes to a private duplicate of the real stdout.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "samples/s"
WORKLOADS = {
    # name: (metric, description, variant, classes, frames_in, size, default batch, train)
    "model": ("KM_UNetV3_SH train samples/sec", "BASELINE configs[2]: KM_UNetV3_SH(num_classes=20) full training step "
              "(fwd + HybridLoss + bwd + AdamW), 5->20 frames 128x128", "SH", 20, 5, 128, 32, True),
    "laps": ("KM_UNetV3_LAPS train samples/sec", "BASELINE configs[3]: KM_UNetV3_LAPS(num_classes=3) full training step, "
             "5->3 frames 256x256", "LAPS", 3, 5, 256, 32, True),
    "infer": ("KM_UNetV3_SH inference samples/sec", "BASELINE configs[4]: KM_UNetV3_SH(num_classes=20) eval forward, "
              "256x256", "SH", 20, 5, 256, 64, False),
}
KAN = dict(CIN=64, COUT=64, KS=3, S=128)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "hbm_gbs": p["hbm_gbs"], "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------- algorithmic work
def op_work(name, key):
    """(bound, algorithmic units per call, unit) of one C-ABI entry point -- SURVEY section 8d / DESIGN section 4."""
    if name.startswith("kmu_kanconv2d"):
        B, Cin, H, W, Cout = key
        f = 2.0 * B * H * W * Cout * Cin * 81
        return ("tensor", f if name.endswith("fwd") else 2 * f, "flop")
    if name.startswith("kmu_hsmssd"):
        B, C, L = key
        return ("hbm", (8.0 if name.endswith("fwd") else 16.0) * B * C * L, "byte")
    if name.startswith("kmu_hybridloss"):
        n = 1
        for v in key:
            n *= v                           # stats: read p, t; stack: + write 5 maps; ssim: 5 + 3 filtered maps; bwd: p, t, 3 maps, dp
        per = {"kmu_hybridloss_stats": 8.0, "kmu_hybridloss_stack": 28.0, "kmu_hybridloss_ssim": 32.0 * (118.0 / 128.0) ** 2,
               "kmu_hybridloss_bwd": 24.0}[name]
        return ("hbm", per * n, "byte")
    if name.startswith("kmu_groupnorm"):
        B, C, HW, G = key                    # statistics pass + apply pass: read x twice, write y
        return ("hbm", 12.0 * B * C * HW, "byte")
    if name.startswith("kmu_resize_bilinear"):
        B, C, H, W, OH, OW = key
        return ("hbm", 4.0 * B * C * (H * W + OH * OW), "byte")
    if name.startswith("kmu_lerpmix"):
        B, C, HW = key                       # fwd: read x, m, write y; bwd: read x, m, dy, write dx, dm
        return ("hbm", (12.0 if name.endswith("fwd") else 20.0) * B * C * HW, "byte")
    if name.startswith("kmu_combine3"):
        B, n = key                           # fwd: read x, f0..f2, write out; bwd: read dy, f0..f2, write df0..df2
        return ("hbm", (20.0 if name.endswith("fwd") else 28.0) * B * n, "byte")
    if name.startswith("kmu_layernorm1d"):
        B, C, L = key
        return ("hbm", (8.0 if name.endswith("fwd") else 12.0) * B * C * L, "byte")
    if name.startswith("kmu_dysample"):
        B, C, H, W = key
        return ("hbm", (20.0 if name.endswith("fwd") else 24.0) * B * C * H * W, "byte")
    if name.startswith("kmu_deformconv3x3"):
        B, C, H, W = key                     # fwd: read x, offset, write out; bwd: + dout, dx, doffset (latency-bound at the bridge size)
        return ("hbm", (4.0 * (2 * C + 18) if name.endswith("fwd") else 4.0 * (4 * C + 36)) * B * H * W, "byte")
    if name.startswith("kmu_dagem"):
        B, C, H, W = key
        return ("hbm", (12.0 if name.endswith("fwd") else 20.0) * B * C * H * W, "byte")
    if name.startswith("kmu_bnmix"):
        B, C, HW = key                       # fwd: read x (+res) twice (statistics, apply), write y; bwd: x, dy (+res) twice, dx (+dres)
        return ("hbm", (12.0 if name.endswith("fwd") else 20.0) * B * C * HW, "byte")
    if name.startswith("kmu_dwconv3x3"):
        B, C, H, W = key                     # fwd: read x, write y; bwd: read dy (dx), read x and dy (dw), write dx
        return ("hbm", (8.0 if name.endswith("fwd") else 16.0) * B * C * H * W, "byte")
    if name == "kmu_pwconv_tma_fwd":
        B, Cin, Cout, HW = key
        return ("hbm", 4.0 * (Cin + Cout) * B * HW, "byte")
    if name == "kmu_pwconv_fused_bwd":
        B, Cin, Cout, HW = key               # one kernel: read x and dy once, write dx
        return ("hbm", 4.0 * (2 * Cin + Cout) * B * HW, "byte")
    if name.startswith("kmu_pwconv"):
        B, Cin, Cout, HW = key               # fwd: read x, write y; bwd: read dy (dx) + x and dy (dw), write dx
        return ("hbm", (4.0 * (Cin + Cout) if name.endswith("fwd") else 4.0 * (2 * Cin + 2 * Cout)) * B * HW, "byte")
    if name.startswith("kmu_smallconv"):
        B, Cin, Cout, H, W, kh, kw = key
        return ("hbm", (4.0 * (Cin + Cout) if name.endswith("fwd") else 4.0 * (2 * Cin + 2 * Cout)) * B * H * W, "byte")
    if name.startswith("kmu_iwp"):
        B, C, H, W = key                     # fwd: read x, write x/4; bwd: read x and dout, write dx
        return ("hbm", (5.0 if name.endswith("fwd") else 9.0) * B * C * H * W, "byte")
    if name.startswith("kmu_triplenorm"):
        B, C, HW = key                       # fwd: x twice (statistics, apply) + y; bwd: x, dy twice + dx
        return ("hbm", (12.0 if name.endswith("fwd") else 20.0) * B * C * HW, "byte")
    if name.startswith("kmu_qkv_gate"):
        B, C, HW = key                       # fwd: 3C in, C out; bwd: 3C + C in, 3C out
        return ("hbm", (16.0 if name.endswith("fwd") else 28.0) * B * C * HW, "byte")
    return ("hbm", 0.0, "byte")


# ----------------------------------------------------------------------------------------------------- CPU oracle arm
def cpu_oracle_model_step_factory(workload, batch):
    """Full-model CPU oracle (oracle/model.py: the model mirror's torch glue + the op restatements of oracle/), the same
    training step as the GPU arm.  The checker, timed here as the CPU baseline only."""
    import torch
    import km_unet_b200 as K
    from km_unet_b200.loss import HybridLoss
    from oracle import model as OM
    _, _, variant, classes, fin, size, _, train = WORKLOADS[workload]
    torch.manual_seed(1234)
    model = K.KM_UNetV3(num_classes=classes, variant=variant)
    model.train(train)
    g = torch.Generator().manual_seed(20240518)
    data = torch.rand(batch, fin + classes, size, size, generator=g)
    x, target = data[:, :fin].contiguous(), data[:, fin:].contiguous()
    crit = HybridLoss()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.05)

    def step():
        with OM.cpu_ops():
            if not train:
                with torch.no_grad():
                    return float(model(x).mean())
            opt.zero_grad(set_to_none=True)
            loss = crit(model(x), target)
            loss.backward()
            opt.step()
            return float(loss.detach())
    return step


def cpu_oracle_kan_step_factory(batch):
    import torch
    from oracle import kan as OK
    CIN, COUT, KS, S = KAN["CIN"], KAN["COUT"], KAN["KS"], KAN["S"]
    torch.manual_seed(1234)
    grid = OK.make_grid(CIN * KS * KS)
    bw = (torch.rand(COUT, CIN * KS * KS) - 0.5) * 0.083
    sw = torch.randn(COUT, CIN * KS * KS, 8) * 0.009
    sc = (torch.rand(COUT, CIN * KS * KS) - 0.5) * 0.083
    params = [t.requires_grad_(True) for t in (bw, sw, sc)]
    g = torch.Generator().manual_seed(20240518)
    x = torch.randn(batch, CIN, S, S, generator=g).requires_grad_(True)
    gout = torch.randn(batch, COUT, S, S, generator=g)

    def step():
        for t in params + [x]:
            t.grad = None
        y = OK.kanconv2d(x, params[0], params[1], params[2], grid, KS, 1, 1)
        y.backward(gout)
        return float(params[0].grad[0, 0])
    return step


def reference_mirror_available():
    from oracle import ref_loader
    return ref_loader.available()


def reference_model_step_factory(workload, batch, device="cpu", fp16_autocast=False, dropin=False):
    """The UNMODIFIED reference (oracle/_ref mirror or /root/reference; imported through oracle/ref_loader.py with the timm /
    fvcore / pywt stand-ins of oracle/shims.py) running train_shanghai.py's train() body (:159-181): zero_grad, forward, loss,
    backward, AdamW(lr 1e-3, wd 0.05) (:342).  HybridLoss = oracle/loss.py (train_shanghai.py:298-326; torchmetrics is absent).
    Nothing of km_unet_b200 is imported -- unless dropin=True: then the same unmodified model file resolves its operator imports to
    the drop-in modules (km_unet_b200.enable_dropin()), which is what a user of the reference gets by switching and nothing else."""
    import warnings
    import torch
    from oracle import loss as OL
    from oracle import ref_loader
    _, _, variant, classes, fin, size, _, train = WORKLOADS[workload]
    R = ref_loader.load_models(dropin=dropin, autocast=True)           # decorators as shipped (inert on CPU tensors)
    torch.manual_seed(1234)
    model = (R.KM_UNetV3_SH if variant == "SH" else R.KM_UNetV3_LAPS)(num_classes=classes).to(device)
    model.train(train)
    g = torch.Generator().manual_seed(20240518)
    data = torch.rand(batch, fin + classes, size, size, generator=g).to(device)
    x, target = data[:, :fin].contiguous(), data[:, fin:].contiguous()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.05)
    scaler = torch.amp.GradScaler("cuda") if fp16_autocast else None

    def step():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if not train:
                with torch.no_grad():
                    return float(model(x).float().mean())
            opt.zero_grad()
            if fp16_autocast:
                with torch.autocast("cuda", dtype=torch.float16):
                    loss = OL.hybrid_loss(model(x), target)
                scaler.scale(loss).backward()
                scaler.step(opt)
                scaler.update()
            else:
                loss = OL.hybrid_loss(model(x), target)
                loss.backward()
                opt.step()
            return loss
    return step


def reference_kan_step_factory(batch):
    import torch
    from oracle import ref_loader
    R = ref_loader.load(with_models=False)
    CIN, COUT, KS, S = KAN["CIN"], KAN["COUT"], KAN["KS"], KAN["S"]
    torch.manual_seed(1234)
    layer = R.KANConv2d(CIN, COUT, KS, padding=1)
    g = torch.Generator().manual_seed(20240518)
    x = torch.randn(batch, CIN, S, S, generator=g).requires_grad_(True)
    gout = torch.randn(batch, COUT, S, S, generator=g)

    def step():
        layer.zero_grad()
        x.grad = None
        layer(x).backward(gout)
        return float(layer.kanlayer.base_weight.grad[0, 0])
    return step


def time_cpu_oracle(workload, batch, steps, warmup):
    """-> (samples/s, ms per step, threads, kind): the reference itself when its mirror is present, else the oracle port."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    if reference_mirror_available():
        kind = "reference"
        step = reference_kan_step_factory(batch) if workload == "kan" else reference_model_step_factory(workload, batch)
    else:
        kind = "port"
        step = cpu_oracle_kan_step_factory(batch) if workload == "kan" else cpu_oracle_model_step_factory(workload, batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, torch.get_num_threads(), kind


def cpu_sample_text(steps, warmup, sb, kind):
    what = ("the UNMODIFIED reference model (oracle/_ref mirror) + oracle/loss.py on torch CPU fp32" if kind == "reference"
            else "oracle/model.py (CPU port) on torch CPU fp32")
    return f"{steps} steps x {sb} sample(s) of the workload after {warmup} warm-up, {what}"


def gpu_eager_reference(workload, dev, steps=3, warmup=2, dropin=False):
    """The unmodified reference on the B200, eager, fp16 autocast + GradScaler as train_shanghai.py:159-181 -- 'the meaningful GPU
    baseline' of SURVEY section 8d.  Largest batch of 32 / 16 / 8 / 4 that fits (its B-spline temporaries are GBs each)."""
    import torch
    if not reference_mirror_available():
        return {"unavailable": "no oracle/_ref mirror on this box"}
    torch.backends.cudnn.benchmark = True                               # train_shanghai.py:330
    last = None
    for B in (32, 16, 8, 4):
        try:
            step = reference_model_step_factory(workload, B, device=dev, fp16_autocast=WORKLOADS[workload][7], dropin=dropin)
            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                res = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            peak = torch.cuda.max_memory_allocated(dev) / 2 ** 30
            return {"value": B / (ms / 1e3), "unit": UNIT, "batch": B, "ms_per_step": ms, "steps": steps, "warmup": warmup,
                    "precision": ("fp16 autocast + GradScaler (train_shanghai.py:172-181), cudnn.benchmark" if WORKLOADS[workload][7] else
                                  "fp16 autocast from the model's own @autocast() decorators (KM_UNetV3_SH.py:465), no_grad, cudnn.benchmark"),
                    "mode": "eager",
                    "loss": float(res) if not isinstance(res, float) else res, "peak_mem_gib": peak,
                    "what": ("unmodified reference model file (oracle/_ref) on the libkmunet drop-in operators (enable_dropin()), the rest stock "
                             "PyTorch, + oracle/loss.py, 1 x B200" if dropin else
                             "unmodified reference model (oracle/_ref) + oracle/loss.py, stock PyTorch kernels, 1 x B200")}
        except torch.cuda.OutOfMemoryError as e:
            last = str(e).split("\n")[0]
            step = None
            import gc
            gc.collect()
            torch.cuda.empty_cache()
    return {"unavailable": f"out of memory down to B=4: {last}"}


def cpu_sample_batch(workload):
    return 1 if workload in ("kan", "laps", "infer") else 2


def metric_of(workload):
    return "train samples/sec (KANConv2d 64->64 3x3 128x128 fwd+bwd)" if workload == "kan" else WORKLOADS[workload][0]


def describe(workload):
    return ("BASELINE configs[1]: KANConv2d 64->64 3x3 grid5 order3, 128x128, fwd+bwd" if workload == "kan"
            else WORKLOADS[workload][1])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sb = cpu_sample_batch(args.workload)
    steps, warmup = min(args.steps, 3), min(args.warmup, 1)
    value, ms, cores, kind = time_cpu_oracle(args.workload, sb, steps, warmup)
    sample = cpu_sample_text(steps, warmup, sb, kind)
    line = {
        "impl": "reference", "metric": metric_of(args.workload), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": describe(args.workload) + f"; CPU sample of {sb} sample(s) per step", "batch_per_step": sb},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------------------- our arm
class Harness:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device: km_unet_b200 has no CPU fallback")
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.args = args

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world == 1:
            return ms
        t = self.torch.tensor([ms], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, step, steps, warmup):
        torch = self.torch
        for _ in range(warmup):
            step()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        self.barrier()
        self.per_rank_ms = self.gather(ms)
        return self.max_over_ranks(ms)

    def gather(self, obj):
        """Every rank's value of `obj`, in rank order (diagnostics only: which rank is the slow one, and at which clock)."""
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out


def kan_microbench(h, B, steps, warmup, precision, with_total=False):
    """BASELINE configs[1]: per-kernel-family device time of one KANConv2d(64->64, 3x3) at B x 64 x 128 x 128 and the
    fraction of the measured bf16 tensor peak (dense FLOP count 2*M*Cout*Cin*81 per GEMM family)."""
    torch = h.torch
    import km_unet_b200 as K
    CIN, COUT, KS, S = KAN["CIN"], KAN["COUT"], KAN["KS"], KAN["S"]
    old = K.config.kan_precision
    K.config.kan_precision = precision
    torch.manual_seed(1234)
    layer = K.KANConv2d(CIN, COUT, KS, padding=1).to(h.dev)
    params = list(layer.parameters())
    g = torch.Generator().manual_seed(20240518 + h.rank)
    x_dev = torch.randn(B, CIN, S, S, generator=g).to(h.dev).requires_grad_(True)
    gout = torch.randn(B, COUT, S, S, generator=g).to(h.dev)
    flops = 2.0 * B * S * S * COUT * CIN * KS * KS * 9

    def time_family(which):
        ts = []
        xin = x_dev if which != "dw" else x_dev.detach()
        for p in params:
            p.requires_grad_(which != "dx")
        for i in range(warmup + steps):
            for p in params:
                p.grad = None
            x_dev.grad = None
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if which == "fwd":
                a.record()
                y = layer(xin)
                b.record()
            else:
                y = layer(xin)
                a.record()
                y.backward(gout)
                b.record()
            torch.cuda.synchronize()
            if i >= warmup:
                ts.append(a.elapsed_time(b))
        for p in params:
            p.requires_grad_(True)
        return sum(ts) / len(ts)

    fam_ms = {"kanconv2d_fwd": time_family("fwd"), "kanconv2d_bwd_dx": time_family("dx"), "kanconv2d_bwd_dw": time_family("dw")}
    peaks = load_peaks()
    fam = {k: {"ms": ms, "achieved": flops / ms / 1e9, "frac": flops / ms / 1e9 / peaks["bf16_tflops"]} for k, ms in fam_ms.items()}
    out = {"workload": describe("kan"), "batch": B, "precision": precision, "flops_per_family": flops, "unit": "TFLOP/s",
           "peak": peaks["bf16_tflops"], "peak_source": peaks["source"] + " bf16 burst", "families": fam}
    if with_total:
        def step():
            for p in params:
                p.grad = None
            x_dev.grad = None
            layer(x_dev).backward(gout)
        ms = h.timed(step, steps, warmup) / steps
        out["fwd_bwd_ms"] = ms
        out["tflops_fwd_bwd"] = 3 * flops / ms / 1e9
    K.config.kan_precision = old
    return out


def run_kan(h, args):
    from km_unet_b200 import _lib
    B = args.batch or 32
    launches0 = _lib.launch_count()
    mb = kan_microbench(h, B, args.steps, args.warmup, args.precision, with_total=True)
    launches = _lib.launch_count() - launches0
    if h.rank != 0:
        return
    fam = mb["families"]
    dom = max(fam, key=lambda k: fam[k]["ms"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath) and args.precision == "bf16" and B == 32:
        with open(tpath) as f:
            traffic = json.load(f).get(dom)
    cpu = None
    if h.world == 1 and not args.no_cpu_baseline:
        v, ms, cores, kind = time_cpu_oracle("kan", 1, 3, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "3 steps x 1 image of the workload (B=1, 64x128x128) after 1 warm-up, " +
                         ("the reference's KANConv2d (oracle/_ref)" if kind == "reference" else "oracle/kan.py") + " on torch CPU fp32"}
    value = h.world * B / (mb["fwd_bwd_ms"] / 1e3)
    line = {"metric": metric_of("kan"), "value": value, "unit": UNIT, "n_gpus": h.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": mb["fwd_bwd_ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": describe("kan"), "batch_per_gpu": B, "global_batch": B * h.world, "precision": args.precision,
                       "parallelism": f"dp{h.world}", "l2": "inputs (x, dy: 2 x %.0f MB) exceed the 126 MB L2" % (B * 64 * 128 * 128 * 4 / 1e6)},
            "e2e": None, "gpu_launches": int(launches // (3 * (args.steps + args.warmup) + args.steps + args.warmup)),
            "roofline": {"kernel": dom, "bound": "tensor", "achieved": fam[dom]["achieved"], "peak": mb["peak"], "unit": "TFLOP/s",
                         "frac": fam[dom]["frac"], "traffic": traffic, "peak_source": mb["peak_source"], "families": fam},
            "cpu_baseline": cpu, "tflops_fwd_bwd": mb["tflops_fwd_bwd"]}
    emit(line)


def run_model(h, args):
    torch, dist = h.torch, h.dist
    import km_unet_b200 as K
    from km_unet_b200 import _lib, ops
    from km_unet_b200.ddp import BucketedGradAllReduce, broadcast_parameters
    from km_unet_b200.loss import HybridLoss
    K.config.kan_precision = args.precision
    K.config.hsm_precision = args.precision               # HSM-SSD BCdt projection on tcgen05 as well
    # pointwise convolutions: forward on the fp32 streaming kernels, backward as ONE fused TMA -> tcgen05 kernel (dx, dW, db from a
    # single pass over x and dy, pwconv_bwd_tc.cu) when the tensor-core precision class is selected
    K.config.conv_bwd = "fused" if args.precision == "bf16" else "split"
    K.config.conv_fwd = "tma" if args.precision == "bf16" else "simt"
    torch.backends.cudnn.benchmark = True                 # as the reference's training script does (train_shanghai.py:331)
    # the loss' SSIM filter (two banded GEMMs) and the small nn.Linear layers: TF32 tensor-core GEMMs in the tensor-core precision
    # class -- the reference computes all of them in fp16 (its loss sits inside autocast, train_shanghai.py:172-174)
    torch.backends.cuda.matmul.allow_tf32 = args.precision == "bf16"
    _lib.lib()                                            # fail loudly if the extension is missing
    metric, desc, variant, classes, fin, size, default_b, train = WORKLOADS[args.workload]
    B = args.batch or default_b
    dev, world, rank = h.dev, h.world, h.rank

    torch.manual_seed(1234)
    model = K.KM_UNetV3(num_classes=classes, variant=variant).to(dev)
    model.train(train)
    broadcast_parameters(model)
    crit = HybridLoss()
    g = torch.Generator().manual_seed(20240518 + rank)
    batch_host = torch.rand(B, fin + classes, size, size, generator=g).pin_memory()
    batch_dev = batch_host.to(dev)
    x_dev, t_dev = batch_dev[:, :fin].contiguous(), batch_dev[:, fin:].contiguous()
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    reducer = opt = None
    if train:
        # parameters that never receive a gradient (the reference's dead branches) stay out of the optimizer and the all-reduce
        crit(model(x_dev[:2]), t_dev[:2]).backward()
        live = [p for p in model.parameters() if p.grad is not None]
        for p in model.parameters():
            p.grad = None
        # train_shanghai.py:342; one launch per step (csrc/optim.cu) instead of torch's 19 multi_tensor_apply launches
        opt = (K.FusedAdamW(live, lr=1e-3, weight_decay=0.05) if args.optimizer == "kmu" else
               torch.optim.AdamW(live, lr=1e-3, weight_decay=0.05, fused=True, capturable=args.graph))
        reducer = BucketedGradAllReduce(live, bucket_bytes=2 << 20) if (world > 1 and not args.graph) else None
    graphed, graphed_launches = None, 0
    if train and args.graph:
        from km_unet_b200.train import GraphedTrainStep
        l0 = _lib.launch_count()
        graphed = GraphedTrainStep(model, crit, opt, x_dev, t_dev, world=world, warmup=3, comm=args.comm)
        graphed_launches = (_lib.launch_count() - l0) // 4          # 3 eager warm-up steps + the captured one

    def train_step(x, t):
        if graphed is not None:
            return graphed(None if x is x_dev else x, None if t is t_dev else t)
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x), t)
        loss.backward()
        if reducer is not None:
            reducer.finish()
        opt.step()
        return loss

    def infer_step(x):
        with torch.no_grad():
            return model(x).mean()

    def step_resident():
        return train_step(x_dev, t_dev) if train else infer_step(x_dev)

    # e2e: the batch lives in pinned host memory; every step copies it host->device (copy stream, double buffered so step
    # i+1's copy overlaps step i's kernels) and reads the step's loss back to the host before the step counts as done.
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [torch.empty_like(batch_dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    state = {"i": 0, "primed": False}

    def prefetch(slot):
        with torch.cuda.stream(copy_stream), torch.no_grad():
            copy_stream.wait_event(consumed[slot])
            bufs[slot].copy_(batch_host, non_blocking=True)
            ready[slot].record(copy_stream)

    def step_e2e():
        slot = state["i"] & 1
        if not state["primed"]:
            for e in consumed:
                e.record()
            prefetch(slot)
            state["primed"] = True
        prefetch(slot ^ 1)
        torch.cuda.current_stream().wait_event(ready[slot])
        x, t = bufs[slot][:, :fin].contiguous(), bufs[slot][:, fin:].contiguous()
        res = train_step(x, t) if train else infer_step(x)
        consumed[slot].record()
        loss_host.copy_(res.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        state["i"] += 1

    sampler = ClockSampler(h.local)            # every rank samples its own GPU; rank 0's goes into `clocks`, the rest into per_rank
    sampler.start()
    launches0 = _lib.launch_count()
    total_ms = h.timed(step_resident, args.steps, args.warmup)
    per_rank_ms = [m / args.steps for m in h.per_rank_ms]
    launches = (_lib.launch_count() - launches0) // (args.steps + args.warmup)
    if graphed is not None:
        launches = graphed_launches            # replays do not pass through the host-side counter: count of the captured step
    clocks = sampler.stop()
    if world > 1:
        every = h.gather({"sm_mhz": clocks["sm_mhz"], "reasons": clocks["reasons"]})
        clocks["per_rank"] = [dict(c, ms_per_step=m) for c, m in zip(every, per_rank_ms)]
    value = world * B * args.steps / (total_ms / 1e3)
    last_loss = float(step_resident().detach().float().cpu())     # sanity: a diverged / corrupted run must not pass as a number
    e2e_ms = h.timed(step_e2e, args.steps, args.warmup)
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)
    used_graph = graphed is not None
    if graphed is not None and graphed.reducer is not None:
        graphed.reducer.remove()       # the captured graph keeps its all-reduces; the eager passes below must not launch more

    # N > 1: what the gradient exchange costs on the critical path = the same step captured WITHOUT the all-reduce, timed the same way
    comm = None
    if world > 1 and train and graphed is not None:
        nocomm = GraphedTrainStep(model, crit, opt, x_dev, t_dev, world=1, warmup=1)
        nocomm_ms = h.timed(lambda: nocomm(), args.steps, args.warmup)
        comm = {"mode": args.comm, "graph_launches_per_step": graphed.graph_launches_per_step,
                "buckets": len(graphed.reducer.buckets), "bytes_per_step": sum(f.numel() * 4 for _, f in graphed.reducer.buckets),
                "ms_per_step_without_allreduce": nocomm_ms / args.steps,
                "per_rank_ms_without_allreduce": [m / args.steps for m in h.per_rank_ms],
                "exposed_comm_ms": (total_ms - nocomm_ms) / args.steps}
        nocomm.close()
        del nocomm

    # op-level profile of the same step: CUDA events on the launching stream around every libkmunet call
    prof_steps = 3

    def eager_step():
        if not train:
            return infer_step(x_dev)
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x_dev), t_dev)
        loss.backward()
        opt.step()
        return loss

    # one stream for the profiling pass: with the parallel branches on, an op's event pair would also time whatever the other
    # branches run concurrently
    branches, K.config.parallel_branches = K.config.parallel_branches, False
    eager_step()
    ops.profile_start()
    for _ in range(prof_steps):
        eager_step()
    prof = ops.profile_stop()
    K.config.parallel_branches = branches
    peaks = load_peaks()
    table = []
    for (name, key), e in prof.items():
        bound, work, unit = op_work(name, key)
        ms = e["ms"] / e["calls"]
        peak = peaks["hbm_gbs"] if bound == "hbm" else peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
        ach = work / ms / (1e6 if bound == "hbm" else 1e9)
        table.append({"op": name, "shape": list(key), "calls_per_step": e["calls"] / prof_steps, "ms_per_call": ms,
                      "ms_per_step": e["ms"] / prof_steps, "bound": bound, "achieved": ach,
                      "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "frac": ach / peak})
    table.sort(key=lambda r: -r["ms_per_step"])
    ours_ms = sum(r["ms_per_step"] for r in table)
    dom = table[0]
    dom_key = f"{dom['op']} {tuple(dom['shape'])}"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(dom_key)         # DRAM bytes per call from the committed `ncu --set full` capture
    roofline = {"kernel": dom_key, "bound": dom["bound"], "achieved": dom["achieved"],
                "peak": peaks["hbm_gbs"] if dom["bound"] == "hbm" else (peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]),
                "unit": dom["unit"], "frac": dom["frac"], "traffic": traffic,
                "peak_source": peaks["source"] + (" HBM copy" if dom["bound"] == "hbm" else " bf16 sustained"),
                "share_of_step": dom["ms_per_step"] / (total_ms / args.steps),
                "libkmunet_ms_per_step": ours_ms, "library_and_glue_ms_per_step": total_ms / args.steps - ours_ms, "ops": table[:12]}

    if rank == 0 and os.environ.get("KMU_BENCH_OPS_DUMP"):   # full per-entry-point table for profiles/
        with open(os.environ["KMU_BENCH_OPS_DUMP"], "w") as f:
            json.dump({"ms_per_step": total_ms / args.steps, "libkmunet_ms_per_step": ours_ms, "ops": table}, f, indent=1)

    # the headline numbers are in hand: a failure in one of the side legs below must not cost the line
    def side_leg(fn):
        try:
            return fn()
        except Exception as e:                                          # noqa: BLE001 -- reported in the line, not swallowed
            return {"unavailable": f"{type(e).__name__}: {str(e).splitlines()[0] if str(e) else ''}"[:300]}

    kan = kan_microbench(h, 32, 5, 3, args.precision) if not args.no_kan_microbench else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        def cpu_leg():
            sb = cpu_sample_batch(args.workload)
            v, ms, cores, kind = time_cpu_oracle(args.workload, sb, 2, 1)
            return {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": cpu_sample_text(2, 1, sb, kind)}
        cpu = side_leg(cpu_leg)
    if graphed is not None:
        graphed.close()                # a graph holding NCCL kernels must be gone before destroy_process_group()
    gpu_ref = gpu_ref_dropin = None
    if rank == 0 and world == 1 and not args.no_gpu_reference:
        graphed = None
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats(dev)
        gpu_ref = side_leg(lambda: gpu_eager_reference(args.workload, dev))
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats(dev)
        gpu_ref_dropin = side_leg(lambda: gpu_eager_reference(args.workload, dev, dropin=True))
    if kan is not None:
        roofline["kan"] = {k: {"ms": v["ms"], "achieved": v["achieved"], "frac": v["frac"]} for k, v in kan["families"].items()}
        roofline["kan"].update({"unit": "TFLOP/s", "peak": kan["peak"], "peak_source": kan["peak_source"], "workload": kan["workload"],
                                "batch": kan["batch"]})

    if rank == 0:
        nlive = sum(p.numel() for p in opt.param_groups[0]["params"]) if train else 0
        line = {
            "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": desc, "batch_per_gpu": B, "global_batch": B * world, "frames": f"{fin}->{classes}",
                       "size": size, "precision": (f"KANConv2d, the HSM-SSD projection (forward, dgrad, wgrad) and the 1x1-convolution backward {args.precision} (tcgen05), torch GEMMs (SSIM filter of the loss, nn.Linear) TF32, everything else fp32"
                                     if args.precision == "bf16" else "fp32 everywhere"),
                       "parallelism": f"dp{world}", "cuda_graph": used_graph, "parallel_graph_branches": bool(K.config.parallel_branches), "optimizer": ("AdamW(lr 1e-3, wd 0.05), " + ("libkmunet kmu_adamw_step" if args.optimizer == "kmu" else "torch fused")) if train else None,
                       "l2": "activations per step (hundreds of %.0f MB tensors) exceed the 126 MB L2" % (B * 16 * size * size * 4 / 1e6),
                       "grad_allreduce": ((f"bucketed NCCL all-reduce of {nlive * 4 / 1e6:.1f} MB launched from grad-ready hooks, captured inside the step graph "
                                           "(overlaps the rest of backward)" if args.comm == "captured" else
                                           f"bucketed NCCL all-reduce of {nlive * 4 / 1e6:.1f} MB between two graphs") if world > 1 else "none (1 GPU)")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": batch_host.numel() * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches), "loss_after_timed_steps": last_loss, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "kan_microbench": kan,
            "gpu_eager_reference": gpu_ref,
            "gpu_eager_reference_dropin": gpu_ref_dropin,
        }
        if comm is not None:
            line["comm"] = comm
        emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="model", choices=["model", "laps", "infer", "kan"])
    ap.add_argument("--batch", type=int, default=0, help="samples per GPU per step (0 = the workload's default)")
    ap.add_argument("--precision", default=os.environ.get("KMU_KAN_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--optimizer", default="kmu", choices=["kmu", "torch"], help="AdamW update: libkmunet's one-launch kernel or torch's fused one")
    ap.add_argument("--no-kan-microbench", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the eager fp16-autocast run of the unmodified reference on the GPU")
    ap.add_argument("--comm", default="captured", choices=["captured", "split"], help="N > 1: NCCL all-reduce inside the step graph, "
                    "overlapped with backward (default), or between two graphs (round-1 scheme)")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="run the training step eagerly instead of as CUDA graphs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)
    h = Harness(args)
    if args.workload == "kan":
        run_kan(h, args)
    else:
        run_model(h, args)
    if h.world > 1:
        # every rank is done and the line is out: leave without tearing NCCL down.  destroy_process_group() blocks for ever when a CUDA
        # graph that recorded NCCL kernels is still alive anywhere (measured: 900 s hang), and nothing after this point matters.
        h.barrier()
        _REAL_STDOUT.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
