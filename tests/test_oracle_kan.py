"""Oracle (oracle/kan.py) vs golden vectors from the reference, and vs the live reference when present."""
import pytest
import torch

from conftest import Golden, rel_err
from oracle import kan, ref_loader

TOL32 = 2e-6   # fp32 vs fp32, different summation order
TOL64 = 1e-12


def _params(g):
    sd = g.sd()
    return sd["base_weight"], sd["spline_weight"], sd["spline_scaler"], sd["grid"]


@pytest.mark.parametrize("name,tol", [("kanlinear_6_5", TOL32), ("kanlinear_6_5_f64", TOL64)])
def test_kanlinear_forward_and_grads(name, tol):
    g = Golden(name)
    wb, ws, sc, grid = _params(g)
    x, gout = g.t("in0"), g.t("gout")
    assert rel_err(kan.kan_linear(x, wb, ws, sc, grid), g.t("out0")) < tol
    dx, dwb, dws, dsc = kan.kan_linear_grads(x, gout, wb, ws, sc, grid)
    gr = g.grads()
    assert rel_err(dx, g.t("grad_in0")) < 10 * tol
    assert rel_err(dwb, gr["base_weight"]) < 10 * tol
    assert rel_err(dws, gr["spline_weight"]) < 10 * tol
    assert rel_err(dsc, gr["spline_scaler"]) < 10 * tol


@pytest.mark.parametrize("name,k,s,p", [("kanconv2d_4_8_k3p1", 3, 1, 1), ("kanconv2d_3_5_k3s2p0", 3, 2, 0),
                                        ("kanconv2d_16_16_k3p1", 3, 1, 1)])
def test_kanconv2d_forward_and_grads(name, k, s, p):
    g = Golden(name)
    sd = g.sd()
    wb, ws, sc, grid = (sd["kanlayer." + n] for n in ("base_weight", "spline_weight", "spline_scaler", "grid"))
    x, gout = g.t("in0"), g.t("gout")
    assert rel_err(kan.kanconv2d(x, wb, ws, sc, grid, k, s, p), g.t("out0")) < TOL32
    dx, dwb, dws, dsc = kan.kanconv2d_grads(x, gout, wb, ws, sc, grid, k, s, p)
    gr = g.grads()
    assert rel_err(dx, g.t("grad_in0")) < 2e-5
    assert rel_err(dwb, gr["kanlayer.base_weight"]) < 2e-5
    assert rel_err(dws, gr["kanlayer.spline_weight"]) < 2e-5
    assert rel_err(dsc, gr["kanlayer.spline_scaler"]) < 2e-5
    # implicit-GEMM formulation (what the tensor-core kernel computes): conv over the Phi image padded with Phi(0)
    assert rel_err(kan.kanconv2d_as_phi_conv(x, wb, ws, sc, grid, k, s, p), g.t("out0")) < TOL32


def test_phi_of_zero_padding_value():
    """Border patches see x = 0 whose spline part is NOT zero: [0,0,1/48,23/48,23/48,1/48,0,0]."""
    g = Golden("kan_phi0")
    grid = g.t("grid")
    got = kan.bspline_basis(torch.zeros(1, grid.shape[0]), grid)
    assert torch.allclose(got, g.t("bases0"), atol=1e-7)
    want = torch.tensor([0, 0, 1 / 48, 23 / 48, 23 / 48, 1 / 48, 0, 0])
    assert torch.allclose(got[0, 0], want, atol=2e-7)


def test_uniform_closed_form_matches_cox_de_boor():
    grid = kan.make_grid(3)
    x = torch.linspace(-2.6, 2.6, 1041).reshape(-1, 1).repeat(1, 3)
    ref = kan.bspline_basis(x, grid)
    got = kan.uniform_cubic_basis(x, t0=-2.2, h=0.4)
    assert (ref - got).abs().max() < 1e-6
    assert ref[(x < -2.2) | (x >= 2.2)].abs().max() == 0


def test_make_grid_matches_reference_buffer():
    g = Golden("kanlinear_6_5")
    assert torch.equal(kan.make_grid(6), g.sd()["grid"])


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
def test_live_reference_kanconv2d_microbench_shape():
    R = ref_loader.load()
    torch.manual_seed(7)
    m = R.KANConv2d(64, 64, 3, padding=1)
    x = torch.randn(1, 64, 12, 12, requires_grad=True)
    y = m(x)
    gout = torch.randn_like(y)
    y.backward(gout)
    k = m.kanlayer
    args = (k.base_weight.detach(), k.spline_weight.detach(), k.spline_scaler.detach(), k.grid)
    assert rel_err(kan.kanconv2d(x.detach(), *args, 3, 1, 1), y) < TOL32
    dx, dwb, dws, dsc = kan.kanconv2d_grads(x.detach(), gout, *args, 3, 1, 1)
    assert rel_err(dx, x.grad) < 2e-5
    assert rel_err(dwb, k.base_weight.grad) < 2e-5
    assert rel_err(dws, k.spline_weight.grad) < 2e-5
    assert rel_err(dsc, k.spline_scaler.grad) < 2e-5
