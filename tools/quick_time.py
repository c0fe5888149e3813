"""Scratch timing of every op at BASELINE config shapes (CUDA events, on the current stream).  Not the bench."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import km_unet_b200 as K


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = "cuda"
    res = {}
    B = int(os.environ.get("QT_B", "8"))
    for (cin, cout, S) in [(64, 64, 128), (16, 16, 128), (16, 32, 64), (32, 64, 32), (64, 32, 32)]:
        m = K.KANConv2d(cin, cout, 3, padding=1).to(dev)
        x = torch.randn(B, cin, S, S, device=dev, requires_grad=True)
        y = m(x)
        g = torch.randn_like(y)
        f = timeit(lambda: m(x))
        def fb():
            yy = m(x)
            yy.backward(g)
        t = timeit(fb)
        gf = 2 * B * S * S * cout * cin * 81 / 1e9
        res[f"kan_{cin}_{cout}_{S}_B{B}"] = {"fwd_ms": f, "fwdbwd_ms": t, "fwd_TF": gf / f, "fwdbwd_TF": 3 * gf / t}
        print(res, flush=True)
    if os.environ.get("QT_ONLY") == "kan":
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(res, open("gpurun_out/quick_time.json", "w"), indent=1)
        return
    for (C, S) in [(16, 128), (32, 64), (64, 32)]:
        m = K.HSMSSD(C).to(dev)
        x = torch.randn(B, C, S * S, device=dev, requires_grad=True)
        y, _ = m(x)
        g = torch.randn_like(y)
        f = timeit(lambda: m(x))
        def fb():
            yy, _ = m(x)
            yy.backward(g)
        t = timeit(fb)
        res[f"hsm_{C}_{S}_B{B}"] = {"fwd_ms": f, "fwdbwd_ms": t, "fwd_GBs_alg": 8 * B * C * S * S / f / 1e6}
        print(res, flush=True)
    for S in [16, 32, 64, 128]:
        m = K.DySample(64).to(dev)
        x = torch.randn(B, 64, S, S, device=dev, requires_grad=True)
        y = m(x)
        g = torch.randn_like(y)
        f = timeit(lambda: m(x))
        def fb():
            yy = m(x)
            yy.backward(g)
        t = timeit(fb)
        res[f"dys_{S}_B{B}"] = {"fwd_ms": f, "fwdbwd_ms": t, "fwd_GBs_alg": 20 * B * 64 * S * S / f / 1e6}
        print(res, flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/quick_time.json", "w"), indent=1)


if __name__ == "__main__":
    main()
