#!/usr/bin/env python
"""bench.py -- headline measurement of the KM-UNet hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--precision fp32|bf16]

Workload (BASELINE.json configs[1], the configuration the "KANConv2D % tensor-core peak" metric is quoted on):
one KANConv2d(64 -> 64, 3x3, padding 1, grid 5, cubic) forward+backward on B x 64 x 128 x 128 fp32 inputs
(synthetic randn, seeded), B = 32 per GPU.  A "step" = forward + backward (dX, dWbase, dWspline, dWscaler) of one batch.
With N > 1 (torchrun, one rank per GPU) every rank processes its own batch shard (weak scaling) and the weight
gradients are averaged with a bucketed NCCL all-reduce launched from grad-ready hooks.

One JSON line on stdout (rank 0).  `value` = whole-job samples/s with inputs resident in HBM; `e2e` = same metric
through the public module API with the batch in pinned host memory (H2D of x and D2H of the weight gradients inside
the timed region); `roofline` = the dominant kernel family vs the measured bf16 tensor peak; `cpu_baseline` = the
oracle port of the same layer on the box's host cores (bounded sample).
`--impl reference` times that CPU oracle port alone (the reference is pure PyTorch: its own CPU path is the same
arithmetic; the reference tree itself cannot travel to the GPU box).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CIN, COUT, KS, S = 64, 64, 3, 128
METRIC = "train samples/sec (KANConv2d 64->64 3x3 128x128 fwd+bwd)"
UNIT = "samples/s"


def flops_fwd(batch):
    return 2.0 * batch * S * S * COUT * CIN * KS * KS * 9      # 2*M*Cout*Cin*81


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


# ----------------------------------------------------------------------------------------------------- CPU oracle arm
def cpu_oracle_step_factory(batch):
    """The oracle port (oracle/kan.py, torch CPU fp32, autograd backward) of the same layer: the checker, timed here as
    the CPU baseline only."""
    import torch
    from oracle import kan as OK
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


def time_cpu_oracle(batch, steps, warmup):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_oracle_step_factory(batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample_b = 1
    value, ms, cores = time_cpu_oracle(sample_b, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: KANConv2d 64->64 3x3 grid5 order3, 128x128, fwd+bwd; CPU sample of "
                               f"{sample_b} image per step", "batch_per_step": sample_b},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {sample_b} image(s) of the workload, oracle/kan.py on torch CPU fp32"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: km_unet_b200 has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import km_unet_b200 as K
    from km_unet_b200 import _lib
    from km_unet_b200.ddp import BucketedGradAllReduce, broadcast_parameters
    K.config.kan_precision = args.precision
    _lib.lib()                                            # fail loudly if the extension is missing

    B = args.batch
    torch.manual_seed(1234)
    layer = K.KANConv2d(CIN, COUT, KS, padding=1).to(dev)
    broadcast_parameters(layer)
    params = list(layer.parameters())
    reducer = BucketedGradAllReduce(params, bucket_bytes=1 << 20) if world > 1 else None
    g = torch.Generator().manual_seed(20240518 + rank)
    x_host = torch.randn(B, CIN, S, S, generator=g).pin_memory()
    gout = torch.randn(B, COUT, S, S, generator=g).to(dev)
    x_dev = x_host.to(dev).requires_grad_(True)
    grads_host = [torch.empty(p.shape, dtype=p.dtype).pin_memory() for p in params]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_resident():
        for p in params:
            p.grad = None
        x_dev.grad = None
        y = layer(x_dev)
        y.backward(gout)
        if reducer is not None:
            reducer.finish()

    # e2e: the batch lives in pinned host memory; every step copies it host->device (on a copy stream, double buffered so
    # step i+1's copy overlaps step i's kernels -- the loop a user of the module API writes) and reads the step's result
    # (the weight gradients) back to the host before the step counts as done.
    copy_stream = torch.cuda.Stream(device=dev)
    x_bufs = [torch.empty_like(x_dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    state = {"i": 0, "primed": False}

    def prefetch(slot):
        with torch.cuda.stream(copy_stream), torch.no_grad():
            copy_stream.wait_event(consumed[slot])
            x_bufs[slot].copy_(x_host, non_blocking=True)        # H2D from pinned memory, every step
            ready[slot].record(copy_stream)

    def step_e2e():
        slot = state["i"] & 1
        if not state["primed"]:
            for e in consumed:
                e.record()
            prefetch(slot)
            state["primed"] = True
        prefetch(slot ^ 1)                                       # next step's input, overlapped with this step's kernels
        torch.cuda.current_stream().wait_event(ready[slot])
        for p in params:
            p.grad = None
        xin = x_bufs[slot].requires_grad_(True)
        xin.grad = None
        y = layer(xin)
        y.backward(gout)
        consumed[slot].record()
        if reducer is not None:
            reducer.finish()
        for h, p in zip(grads_host, params):
            h.copy_(p.grad, non_blocking=True)            # D2H of the step's result (the weight gradients)
        torch.cuda.current_stream().synchronize()         # the host must hold the result before the next step
        state["i"] += 1

    def timed(step, steps, warmup):
        for _ in range(warmup):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        barrier()
        return max_over_ranks(ms)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = _lib.launch_count()
    total_ms = timed(step_resident, args.steps, args.warmup)
    launches = (_lib.launch_count() - launches0) * args.steps // (args.steps + args.warmup)
    clocks = sampler.stop() if sampler else None
    value = world * B * args.steps / (total_ms / 1e3)
    e2e_ms = timed(step_e2e, args.steps, args.warmup)
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)

    # per-kernel-family device time: CUDA events on the launching (current) stream around each family's launches.
    # backward-input alone = parameters frozen (the C call gets d_base_weight = NULL), backward-weights alone = detached input.
    def time_family(which):
        ts = []
        xin = x_dev if which != "dw" else x_dev.detach()
        for p in params:
            p.requires_grad_(which != "dx")
        for i in range(args.warmup + args.steps):
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
            if i >= args.warmup:
                ts.append(a.elapsed_time(b))
        for p in params:
            p.requires_grad_(True)
        return sum(ts) / len(ts)

    fam_ms = {"kanconv2d_fwd": time_family("fwd"), "kanconv2d_bwd_dx": time_family("dx"), "kanconv2d_bwd_dw": time_family("dw")}
    peaks = load_peaks()
    peak_tf = peaks["bf16_tflops"]
    fam = {k: (flops_fwd(B), ms) for k, ms in fam_ms.items()}        # each family is one 2*M*Cout*Cin*81 GEMM
    dom = max(fam, key=lambda k: fam[k][1])
    roof_all = {k: {"ms": ms, "achieved": fl / ms / 1e9, "frac": fl / ms / 1e9 / peak_tf} for k, (fl, ms) in fam.items()}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath) and args.precision == "bf16" and B == 32:
        with open(tpath) as f:
            traffic = json.load(f).get(dom)
    roofline = {"kernel": dom, "bound": "tensor", "achieved": roof_all[dom]["achieved"], "peak": peak_tf, "unit": "TFLOP/s",
                "frac": roof_all[dom]["frac"], "traffic": traffic, "peak_source": peaks["source"] + " bf16 burst",
                "algorithmic_flops_per_launch": fam[dom][0], "families": roof_all}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, ms, cores = time_cpu_oracle(1, 3, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "3 steps x 1 image of the workload (B=1, 64x128x128) after 1 warm-up, oracle/kan.py on torch CPU fp32"}

    if rank == 0:
        h2d = x_host.numel() * 4
        d2h = sum(p.numel() * 4 for p in params)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: KANConv2d 64->64 3x3 grid5 order3, 128x128, fwd+bwd",
                       "batch_per_gpu": B, "global_batch": B * world, "precision": args.precision,
                       "parallelism": f"dp{world}", "l2": "inputs (x, dy: 2 x %.0f MB) exceed the 126 MB L2" % (B * CIN * S * S * 4 / 1e6),
                       "grad_allreduce": "bucketed NCCL, hooks" if world > 1 else "none (1 GPU)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "tflops_fwd_bwd": 3 * flops_fwd(B) * world / (total_ms / args.steps) / 1e9,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--precision", default=os.environ.get("KMU_KAN_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
