"""GPU parity of the tcgen05 (bf16 operands, fp32 TMEM accumulation) KANConv2d forward against the fp64 oracle.
Gate: 2e-2 relative (north_star, "under TF32/bf16 tensor-core math")."""
import ctypes

import pytest
import torch

from conftest import Golden, rel_err

pytestmark = pytest.mark.gpu
TOL_BF16 = 2e-2


def _oracle(m, x, k=3, s=1, p=1):
    from oracle import kan as O
    kl = m.kanlayer
    return O.kanconv2d(x.double(), kl.base_weight.detach().double(), kl.spline_weight.detach().double(),
                       kl.spline_scaler.detach().double(), kl.grid, k, s, p)


def _takes_tensor_path(m, x):
    from km_unet_b200 import _lib
    uniform, t0, h = m.kanlayer._grid_meta()
    B, C, H, W = x.shape
    d = _lib.KanDesc(B, C, H, W, m.out_channels, m.kernel_size, m.stride, m.padding, 5, 3, _lib.KMU_PREC_BF16, 1,
                     1 if uniform else 0, t0, h)
    return _lib.lib().kmu_kanconv2d_path(ctypes.byref(d)) == 1


@pytest.mark.parametrize("B,cin,cout,H,W", [
    (1, 16, 16, 16, 8),       # a single strip, a single M tile
    (2, 16, 16, 32, 32),      # enc1.0 channels
    (1, 16, 32, 17, 13),      # ragged: strips and tiles overhang the image on both axes
    (2, 32, 64, 32, 32),      # enc3.0
    (1, 64, 32, 32, 32),      # dec1.1
    (1, 64, 64, 128, 128),    # config-2 microbench image (TT=1 at this batch)
    (5, 16, 16, 128, 128),    # enough strips for the 32-row strip kernel (TT=2)
    (10, 16, 16, 128, 128),   # enough strips for the 64-row strip kernel (TT=4)
])
def test_kanconv2d_tc_forward_vs_oracle(B, cin, cout, H, W):
    from km_unet_b200 import KANConv2d
    torch.manual_seed(B + cin + cout + H)
    m = KANConv2d(cin, cout, 3, padding=1)
    m.kanlayer.precision = "bf16"
    x = torch.randn(B, cin, H, W) * 1.1
    want = _oracle(m, x)
    m = m.cuda()
    xc = x.cuda()
    assert _takes_tensor_path(m, xc)
    y = m(xc)
    torch.cuda.synchronize()
    err = rel_err(y, want)
    assert err < TOL_BF16, err
    # the same layer in the fp32 family agrees to 1e-4: the bf16 error is operand rounding, not a layout bug
    m.kanlayer.precision = "fp32"
    y32 = m(xc)
    assert rel_err(y32, want) < 1e-4
    assert rel_err(y, y32) < TOL_BF16


def test_kanconv2d_tc_golden_16_16():
    from km_unet_b200 import KANConv2d
    g = Golden("kanconv2d_16_16_k3p1")
    m = KANConv2d(16, 16, 3, padding=1)
    m.load_state_dict(g.sd())
    m.kanlayer.precision = "bf16"
    m = m.cuda()
    x = g.t("in0", "cuda").requires_grad_(True)
    assert _takes_tensor_path(m, x)
    y = m(x)
    assert rel_err(y, g.t("out0")) < TOL_BF16
    y.backward(g.t("gout", "cuda"))                      # backward = the tcgen05 dX / dW kernels
    assert rel_err(x.grad, g.t("grad_in0")) < TOL_BF16
    want = g.grads()
    for n, p in m.named_parameters():
        assert rel_err(p.grad, want[n]) < TOL_BF16, n


@pytest.mark.parametrize("B,cin,cout,H,W", [
    (1, 16, 16, 16, 8),       # one tile
    (2, 16, 16, 32, 32),      # enc1.0 channels: 8 tap-row slots per A descriptor
    (1, 16, 32, 17, 13),      # ragged tiles on both axes
    (2, 32, 64, 32, 32),      # enc3.0: two accumulator sets (tap rows 2,1 | 0,pad), 8-channel blocks
    (1, 64, 32, 32, 32),      # dec1.1: four 16-channel blocks
    (3, 64, 64, 48, 40),      # microbench channels, more tiles than CTAs per block
])
def test_kanconv2d_tc_backward_vs_oracle(B, cin, cout, H, W):
    """dX, d base_weight, d spline_weight, d spline_scaler of the tensor-core family against fp64 autograd of the oracle."""
    from km_unet_b200 import KANConv2d
    from oracle import kan as O
    torch.manual_seed(B + cin + cout + H)
    m = KANConv2d(cin, cout, 3, padding=1)
    m.kanlayer.precision = "bf16"
    kl = m.kanlayer
    x = torch.randn(B, cin, H, W) * 1.1
    g = torch.randn(B, cout, H, W)
    xd = x.double().requires_grad_(True)
    ps = [p.detach().double().requires_grad_(True) for p in (kl.base_weight, kl.spline_weight, kl.spline_scaler)]
    O.kanconv2d(xd, ps[0], ps[1], ps[2], kl.grid, 3, 1, 1).backward(g.double())
    m = m.cuda()
    xc = x.cuda().requires_grad_(True)
    assert _takes_tensor_path(m, xc)
    m(xc).backward(g.cuda())
    torch.cuda.synchronize()
    assert rel_err(xc.grad, xd.grad) < TOL_BF16
    assert rel_err(m.kanlayer.base_weight.grad, ps[0].grad) < TOL_BF16
    assert rel_err(m.kanlayer.spline_weight.grad, ps[1].grad) < TOL_BF16
    assert rel_err(m.kanlayer.spline_scaler.grad, ps[2].grad) < TOL_BF16


def test_kanconv2d_tc_backward_is_deterministic_and_skips_unneeded_grads():
    from km_unet_b200 import KANConv2d
    torch.manual_seed(3)
    m = KANConv2d(32, 32, 3, padding=1)
    m.kanlayer.precision = "bf16"
    m = m.cuda()
    x = torch.randn(2, 32, 40, 24, device="cuda", requires_grad=True)
    g = torch.randn(2, 32, 40, 24, device="cuda")
    outs = []
    for _ in range(2):
        x.grad = None
        m.zero_grad(set_to_none=True)
        m(x).backward(g)
        outs.append((x.grad.clone(), m.kanlayer.spline_weight.grad.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])   # fixed-order split reduction
    x2 = x.detach()                                       # no dX requested: only the dW kernel runs
    m.zero_grad(set_to_none=True)
    m(x2).backward(g)
    assert torch.equal(m.kanlayer.spline_weight.grad, outs[0][1])


def test_kanconv2d_tc_backward_full_size_property():
    """At the microbench size (64->64, 128x128) the oracle is too slow; check linearity of the backward in dY and that
    d base_weight of a constant-one upstream gradient equals the tap-shifted sums of SiLU(x) (bf16 tolerance)."""
    from km_unet_b200 import KANConv2d
    torch.manual_seed(5)
    m = KANConv2d(64, 64, 3, padding=1)
    m.kanlayer.precision = "bf16"
    m = m.cuda()
    x = torch.randn(2, 64, 128, 128, device="cuda", requires_grad=True)
    g1 = torch.randn(2, 64, 128, 128, device="cuda")
    g2 = torch.randn(2, 64, 128, 128, device="cuda")

    def grads(g):
        x.grad = None
        m.zero_grad(set_to_none=True)
        m(x).backward(g)
        return x.grad.clone(), m.kanlayer.base_weight.grad.clone()
    a, b, c = grads(g1), grads(g2), grads(g1 + g2)
    assert rel_err(a[0] + b[0], c[0]) < TOL_BF16 and rel_err(a[1] + b[1], c[1]) < TOL_BF16
    _, dbw = grads(torch.ones_like(g1))
    act = torch.nn.functional.pad(torch.nn.functional.silu(x.detach()), (1, 1, 1, 1))
    want = torch.stack([act[:, :, ki:ki + 128, kj:kj + 128].sum(dim=(0, 2, 3)) for ki in range(3) for kj in range(3)], dim=1)
    assert rel_err(dbw[7].view(64, 9), want) < TOL_BF16


def test_kanconv2d_tc_unit_impulse_isolates_each_tap():
    """Structure test: with one-hot weights the layer must return a SHIFTED copy of SiLU(x) for every tap, which
    pins the shifted-window descriptors (tap -> (ki,kj) offset) and the zero padding exactly."""
    from km_unet_b200 import KANConv2d
    torch.manual_seed(0)
    m = KANConv2d(16, 16, 3, padding=1)
    m.kanlayer.precision = "bf16"
    x = torch.randn(1, 16, 24, 16)
    m = m.cuda()
    xc = x.cuda()
    act = torch.nn.functional.silu(x)
    for tap in range(9):
        ki, kj = divmod(tap, 3)
        with torch.no_grad():
            m.kanlayer.base_weight.zero_()
            m.kanlayer.spline_weight.zero_()
            m.kanlayer.base_weight[3, 5 * 9 + tap] = 1.0            # out channel 3 <- SiLU(in channel 5) through this tap
        y = m(xc).cpu()
        want = torch.zeros(24 + 2, 16 + 2)
        want[1:-1, 1:-1] = act[0, 5]
        want = want[ki:ki + 24, kj:kj + 16]
        assert (y[0, 3] - want).abs().max() < 2e-2, tap
        assert y[0, :3].abs().max() == 0 and y[0, 4:].abs().max() == 0


def test_nonuniform_grid_falls_back_to_fp32_family_on_gpu():
    from km_unet_b200 import KANConv2d
    torch.manual_seed(1)
    m = KANConv2d(16, 16, 3, padding=1)
    m.kanlayer.precision = "bf16"
    with torch.no_grad():
        m.kanlayer.grid.copy_(m.kanlayer.grid + torch.cumsum(torch.rand_like(m.kanlayer.grid) * 0.05, dim=1))
    x = torch.randn(1, 16, 8, 8)
    want = _oracle(m, x)
    m = m.cuda()
    assert not _takes_tensor_path(m, x.cuda())
    assert rel_err(m(x.cuda()), want) < 1e-4
