"""GPU parity of the KANConv2d / KANLinear op (through the drop-in module -> ctypes -> C ABI) against the reference's
golden vectors and against the oracle on seeded inputs.  fp32 family gate: 1e-4 relative (north_star)."""
import pytest
import torch

from conftest import Golden, rel_err

pytestmark = pytest.mark.gpu

TOL_FP32 = 1e-4


def _load(mod, sd):
    mod.load_state_dict(sd)
    return mod.cuda()


def _check_grads(mod, prefix, want, tol):
    for name, p in mod.named_parameters():
        key = name if name in want else prefix + name
        assert p.grad is not None, name
        assert rel_err(p.grad, want[key]) < tol, (name, rel_err(p.grad, want[key]))


def test_kanlinear_golden():
    from km_unet_b200 import KANLinear
    g = Golden("kanlinear_6_5")
    m = _load(KANLinear(6, 5), g.sd())
    x = g.t("in0", "cuda").requires_grad_(True)
    y = m(x)
    assert rel_err(y, g.t("out0")) < TOL_FP32
    y.backward(g.t("gout", "cuda"))
    assert rel_err(x.grad, g.t("grad_in0")) < TOL_FP32
    _check_grads(m, "", g.grads(), TOL_FP32)


@pytest.mark.parametrize("name,cin,cout,k,s,p", [("kanconv2d_4_8_k3p1", 4, 8, 3, 1, 1), ("kanconv2d_3_5_k3s2p0", 3, 5, 3, 2, 0),
                                                 ("kanconv2d_16_16_k3p1", 16, 16, 3, 1, 1)])
def test_kanconv2d_golden(name, cin, cout, k, s, p):
    from km_unet_b200 import KANConv2d
    g = Golden(name)
    m = _load(KANConv2d(cin, cout, k, stride=s, padding=p), g.sd())
    x = g.t("in0", "cuda").requires_grad_(True)
    y = m(x)
    assert y.shape == g.t("out0").shape
    assert rel_err(y, g.t("out0")) < TOL_FP32
    y.backward(g.t("gout", "cuda"))
    assert rel_err(x.grad, g.t("grad_in0")) < TOL_FP32
    _check_grads(m, "", g.grads(), TOL_FP32)


@pytest.mark.parametrize("B,cin,cout,H,W,k,s,p", [
    (2, 16, 16, 24, 24, 3, 1, 1),     # enc1.0-like
    (1, 16, 32, 17, 13, 3, 1, 1),     # ragged spatial size, residual-branch shape of enc2.0
    (2, 32, 64, 8, 8, 3, 1, 1),       # enc3.0-like
    (1, 64, 64, 12, 12, 3, 1, 1),     # config-2 microbench channels
    (1, 5, 70, 9, 9, 3, 1, 1),        # Cout > 64 (two output tiles), odd Cin
    (3, 2, 3, 7, 6, 2, 1, 0),         # even kernel, no padding
    (2, 3, 4, 11, 10, 5, 2, 2),       # 5x5 stride 2
    (4, 1, 1, 1, 1, 1, 1, 0),         # degenerate 1x1
])
def test_kanconv2d_vs_oracle(B, cin, cout, H, W, k, s, p):
    from km_unet_b200 import KANConv2d
    from oracle import kan as O
    torch.manual_seed(B * 1000 + cin * 10 + cout)
    m = KANConv2d(cin, cout, k, stride=s, padding=p)
    kl = m.kanlayer
    x = torch.randn(B, cin, H, W) * 1.2
    want = O.kanconv2d(x.double(), kl.base_weight.detach().double(), kl.spline_weight.detach().double(),
                       kl.spline_scaler.detach().double(), kl.grid, k, s, p)
    gout = torch.randn(want.shape)
    dx, dwb, dws, dsc = O.kanconv2d_grads(x.double(), gout.double(), kl.base_weight.detach().double(),
                                          kl.spline_weight.detach().double(), kl.spline_scaler.detach().double(), kl.grid,
                                          k, s, p)
    m = m.cuda()
    xc = x.cuda().requires_grad_(True)
    y = m(xc)
    assert rel_err(y, want) < TOL_FP32
    y.backward(gout.cuda())
    assert rel_err(xc.grad, dx) < TOL_FP32
    assert rel_err(m.kanlayer.base_weight.grad, dwb) < TOL_FP32
    assert rel_err(m.kanlayer.spline_weight.grad, dws) < TOL_FP32
    assert rel_err(m.kanlayer.spline_scaler.grad, dsc) < TOL_FP32


def test_kanconv2d_nonuniform_per_feature_grid():
    """After KANLinear.update_grid every feature has its own non-uniform knot row: the fp32 family must follow it."""
    from km_unet_b200 import KANConv2d
    from oracle import kan as O
    torch.manual_seed(5)
    m = KANConv2d(3, 4, 3, padding=1)
    kl = m.kanlayer
    with torch.no_grad():
        kl.grid.copy_(kl.grid + torch.cumsum(torch.rand_like(kl.grid) * 0.1, dim=1))
    x = torch.randn(2, 3, 6, 6)
    want = O.kanconv2d(x.double(), kl.base_weight.detach().double(), kl.spline_weight.detach().double(),
                       kl.spline_scaler.detach().double(), kl.grid.double(), 3, 1, 1)
    y = m.cuda()(x.cuda())
    assert rel_err(y, want) < TOL_FP32


def test_out_of_range_inputs_have_zero_basis_but_live_silu():
    from km_unet_b200 import KANLinear
    m = KANLinear(4, 3).cuda()
    x = torch.tensor([[5.0, -7.0, 2.2, -2.2000005]]).cuda().repeat(2, 1)
    y = m(x)
    want = torch.nn.functional.silu(x) @ m.base_weight.t()      # basis is zero outside [-2.2, 2.2)
    assert rel_err(y, want) < 1e-5


def test_no_cpu_fallback():
    from km_unet_b200 import KANConv2d
    with pytest.raises(RuntimeError):
        KANConv2d(2, 2, 3, padding=1)(torch.randn(1, 2, 4, 4))


def test_kanconv2d_property_linearity_in_weights_full_size():
    """Size-independent property at a config-3 layer size (enc1.0: 16->16 @128x128): the layer is linear in its
    weights, so f(x; 2W) == 2 f(x; W) and the weight gradient does not depend on W."""
    from km_unet_b200 import KANConv2d
    torch.manual_seed(11)
    m = KANConv2d(16, 16, 3, padding=1).cuda()
    x = torch.randn(2, 16, 128, 128, device="cuda")
    y1 = m(x)
    g1 = torch.autograd.grad(y1.sum(), m.kanlayer.base_weight)[0]
    with torch.no_grad():
        m.kanlayer.base_weight.mul_(2.0)
        m.kanlayer.spline_weight.mul_(2.0)
    y2 = m(x)
    g2 = torch.autograd.grad(y2.sum(), m.kanlayer.base_weight)[0]
    assert rel_err(y2, 2 * y1) < 1e-5
    assert rel_err(g2, g1) < 1e-5
