"""Library reference point: cuDNN's weight gradient of the 1x3 / 3x1 convolutions at the model's shapes (what sc3_wgrad_strip_kernel replaced)."""
import torch
torch.backends.cudnn.benchmark = True
for (C, S, kh, kw) in [(16, 128, 3, 1), (16, 128, 1, 3), (32, 64, 3, 1), (64, 32, 1, 3)]:
    x = torch.randn(32, C, S, S, device="cuda")
    dy = torch.randn(32, C, S, S, device="cuda")
    w = torch.randn(C, C, kh, kw, device="cuda")
    f = lambda: torch.ops.aten.convolution_backward(dy, x, w, [C], [1, 1], [kh // 2, kw // 2], [1, 1], False, [0, 0], 1, [False, True, True])
    for _ in range(5):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        f()
    b.record()
    torch.cuda.synchronize()
    print(C, S, kh, kw, "aten wgrad us", a.elapsed_time(b) / 20 * 1000)
