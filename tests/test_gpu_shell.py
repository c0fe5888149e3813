"""GPU parity of the EfficientViMBlock shell ops (kmu_bnmix, kmu_dwconv3x3) against torch fp64 autograd of the same
arithmetic (vim_utils_init.py:83-89 BatchNorm2d inside ConvLayer2D; efficient_vim_init.py:82-96 layer-scale mix)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _ref_bnmix(x, w, b, rm, rv, training, relu, res, alpha):
    y = F.batch_norm(x, rm, rv, w, b, training, 0.1, 1e-5)
    if relu:
        y = F.relu(y)
    if res is not None:
        a = torch.sigmoid(alpha).view(1, -1, 1, 1)
        y = (1 - a) * res + a * y
    return y


@pytest.mark.parametrize("B,C,H,W,training,relu,mix", [
    (4, 16, 32, 32, True, False, True), (2, 64, 16, 16, True, True, False), (3, 8, 5, 7, True, False, True),
    (2, 16, 8, 8, False, False, True), (2, 32, 12, 12, False, True, False), (32, 16, 128, 128, True, False, True)])
def test_bnmix_vs_torch(B, C, H, W, training, relu, mix):
    from km_unet_b200 import ops
    torch.manual_seed(C + H)
    x = torch.randn(B, C, H, W) * 1.7 + 0.4
    w, b = torch.rand(C) + 0.5, torch.randn(C) * 0.2
    rm, rv = torch.randn(C) * 0.1, torch.rand(C) + 0.5
    res = torch.randn(B, C, H, W) if mix else None
    alpha = torch.randn(C) if mix else None
    gout = torch.randn(B, C, H, W)
    dbl = lambda t: None if t is None else t.double().requires_grad_(True)
    xd, wd, bd, rd, ad = dbl(x), dbl(w), dbl(b), dbl(res), dbl(alpha)
    rm_ref, rv_ref = rm.double().clone(), rv.double().clone()
    want = _ref_bnmix(xd, wd, bd, rm_ref, rv_ref, training, relu, rd, ad)
    want.backward(gout.double())
    cu = lambda t: None if t is None else t.cuda().requires_grad_(True)
    xc, wc, bc, rc, ac = cu(x), cu(w), cu(b), cu(res), cu(alpha)
    rmc, rvc = rm.cuda(), rv.cuda()
    y = ops.bnmix(xc, wc, bc, rmc, rvc, training, 0.1, 1e-5, relu, rc, ac)
    assert rel_err(y, want) < TOL
    y.backward(gout.cuda())
    assert rel_err(xc.grad, xd.grad) < TOL
    assert rel_err(wc.grad, wd.grad) < TOL
    assert rel_err(bc.grad, bd.grad) < TOL
    if mix:
        assert rel_err(rc.grad, rd.grad) < TOL
        assert rel_err(ac.grad, ad.grad) < TOL
    assert rel_err(rmc, rm_ref) < TOL and rel_err(rvc, rv_ref) < TOL


def test_bnmix_large_channel_offset_does_not_cancel():
    """mean 100, std 0.1 -- E[x^2] - mean^2 in fp32 loses the variance entirely; the kernel shifts by a sample."""
    from km_unet_b200 import ops
    torch.manual_seed(3)
    B, C, H, W = 4, 16, 32, 32
    x = torch.randn(B, C, H, W) * 0.1 + 100.0 * (torch.arange(C).float().view(1, C, 1, 1) - 7.5)
    w, b = torch.rand(C) + 0.5, torch.randn(C) * 0.2
    rm, rv = torch.zeros(C), torch.ones(C)
    gout = torch.randn(B, C, H, W)
    xd, wd, bd = (t.double().requires_grad_(True) for t in (x, w, b))
    rm_ref, rv_ref = rm.double().clone(), rv.double().clone()
    want = _ref_bnmix(xd, wd, bd, rm_ref, rv_ref, True, False, None, None)
    want.backward(gout.double())
    xc, wc, bc = (t.cuda().requires_grad_(True) for t in (x, w, b))
    rmc, rvc = rm.cuda(), rv.cuda()
    y = ops.bnmix(xc, wc, bc, rmc, rvc, True, 0.1, 1e-5, False, None, None)
    y.backward(gout.cuda())
    # x itself carries ~1e-7 * 750 / 0.1 = 1e-3 of a standard deviation of fp32 representation error; the statistics must not add to it
    assert rel_err(rvc, rv_ref) < 1e-4 and rel_err(rmc, rm_ref) < 1e-6
    assert rel_err(y, want) < 2e-3
    assert rel_err(wc.grad, wd.grad) < 2e-3 and rel_err(bc.grad, bd.grad) < 1e-4


@pytest.mark.parametrize("B,C,H,W,bias", [(2, 16, 32, 32, False), (3, 8, 7, 5, True), (1, 64, 16, 16, True), (4, 16, 128, 128, False),
                                          (2, 4, 1, 1, True)])
def test_dwconv3x3_vs_torch(B, C, H, W, bias):
    from km_unet_b200 import ops
    torch.manual_seed(H * W + C)
    x = torch.randn(B, C, H, W)
    w = torch.randn(C, 1, 3, 3) * 0.4
    bv = torch.randn(C) if bias else None
    gout = torch.randn(B, C, H, W)
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    bd = bv.double().requires_grad_(True) if bias else None
    want = F.conv2d(xd, wd, bd, padding=1, groups=C)
    want.backward(gout.double())
    xc, wc = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    bc = bv.cuda().requires_grad_(True) if bias else None
    y = ops.dwconv3x3(xc, wc, bc)
    assert rel_err(y, want) < TOL
    y.backward(gout.cuda())
    assert rel_err(xc.grad, xd.grad) < TOL
    assert rel_err(wc.grad, wd.grad) < TOL
    if bias:
        assert rel_err(bc.grad, bd.grad) < TOL


def test_shell_ops_raise_on_cpu():
    from km_unet_b200 import ops
    with pytest.raises(RuntimeError):
        ops.dwconv3x3(torch.randn(1, 4, 4, 4), torch.randn(4, 1, 3, 3))
    with pytest.raises(RuntimeError):
        ops.bnmix(torch.randn(1, 4, 4, 4), torch.ones(4), torch.zeros(4), torch.zeros(4), torch.ones(4), True)


@pytest.mark.parametrize("B,Cin,Cout,H,W,bias", [(2, 16, 64, 32, 32, True), (3, 64, 16, 7, 9, False), (2, 16, 48, 16, 16, True),
                                                (1, 256, 64, 8, 8, True), (2, 64, 256, 12, 12, False), (2, 128, 32, 5, 5, True),
                                                (32, 16, 64, 128, 128, False), (2, 32, 32, 64, 64, True)])
def test_pwconv_vs_torch(B, Cin, Cout, H, W, bias):
    from km_unet_b200 import ops
    assert ops.pwconv_supported(Cin, Cout)
    torch.manual_seed(Cin + Cout)
    x = torch.randn(B, Cin, H, W)
    w = torch.randn(Cout, Cin, 1, 1) / Cin ** 0.5
    bv = torch.randn(Cout) if bias else None
    gout = torch.randn(B, Cout, H, W)
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    bd = bv.double().requires_grad_(True) if bias else None
    want = F.conv2d(xd, wd, bd)
    want.backward(gout.double())
    xc, wc = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    bc = bv.cuda().requires_grad_(True) if bias else None
    y = ops.pwconv(xc, wc, bc)
    assert rel_err(y, want) < TOL
    y.backward(gout.cuda())
    assert rel_err(xc.grad, xd.grad) < TOL
    assert rel_err(wc.grad, wd.grad) < TOL
    if bias:
        assert rel_err(bc.grad, bd.grad) < TOL


def test_pwconv_unsupported_pairs_are_reported():
    from km_unet_b200 import ops
    assert not ops.pwconv_supported(17, 16)      # IWP fusion conv (C + 1 inputs) stays a library conv
    assert not ops.pwconv_supported(96, 32)


@pytest.mark.parametrize("B,C,H,W", [(2, 16, 32, 32), (3, 32, 7, 9), (2, 64, 8, 8), (4, 16, 128, 128)])
def test_triplenorm_vs_reference_composition(B, C, H, W):
    """KM_UNetV3_SH.py:277-284 restated literally in torch fp64 (two GroupNorm(1), permuted LayerNorm, /3)."""
    from km_unet_b200 import ops
    from oracle.model import _triplenorm
    torch.manual_seed(C * H)
    x = torch.randn(B, C, H, W) * 1.3 + 0.2
    ps = [torch.rand(C) + 0.5 if i % 2 == 0 else torch.randn(C) * 0.2 for i in range(6)]
    gout = torch.randn(B, C, H, W)
    xd = x.double().requires_grad_(True)
    pd = [p.double().requires_grad_(True) for p in ps]
    want = _triplenorm(xd, *pd)
    want.backward(gout.double())
    xc = x.cuda().requires_grad_(True)
    pc = [p.cuda().requires_grad_(True) for p in ps]
    y = ops.triplenorm(xc, *pc)
    assert rel_err(y, want) < TOL
    y.backward(gout.cuda())
    assert rel_err(xc.grad, xd.grad) < TOL
    for a, b in zip(pc, pd):
        assert rel_err(a.grad, b.grad) < TOL


@pytest.mark.parametrize("B,C,H,W", [(2, 16, 16, 16), (3, 8, 5, 7), (4, 16, 128, 128)])
def test_qkv_gate_vs_torch(B, C, H, W):
    from km_unet_b200 import ops
    torch.manual_seed(C + W)
    qkv = torch.randn(B, 3 * C, H, W)
    gout = torch.randn(B, C, H, W)
    qd = qkv.double().requires_grad_(True)
    q, k, v = qd.chunk(3, dim=1)
    want = torch.sigmoid(q * k) * v
    want.backward(gout.double())
    qc = qkv.cuda().requires_grad_(True)
    y = ops.qkv_gate(qc)
    assert rel_err(y, want) < TOL
    y.backward(gout.cuda())
    assert rel_err(qc.grad, qd.grad) < TOL


@pytest.mark.parametrize("B,Cin,Cout,H,W,kh,kw,bias", [(2, 16, 16, 32, 32, 3, 1, True), (2, 16, 16, 32, 32, 1, 3, True),
                                                      (3, 32, 32, 7, 9, 3, 1, False), (2, 64, 32, 16, 16, 3, 3, True),
                                                      (2, 64, 16, 12, 20, 3, 3, True), (1, 16, 20, 9, 9, 3, 3, True),
                                                      (32, 16, 16, 128, 128, 1, 3, True), (2, 32, 32, 64, 64, 3, 3, False)])
def test_smallconv_vs_torch(B, Cin, Cout, H, W, kh, kw, bias):
    from km_unet_b200 import ops
    assert ops.smallconv_supported(Cin, Cout, kh, kw)
    torch.manual_seed(Cin + Cout + kh)
    x = torch.randn(B, Cin, H, W)
    w = torch.randn(Cout, Cin, kh, kw) / (Cin * kh * kw) ** 0.5
    bv = torch.randn(Cout) if bias else None
    gout = torch.randn(B, Cout, H, W)
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    bd = bv.double().requires_grad_(True) if bias else None
    want = F.conv2d(xd, wd, bd, padding=(kh // 2, kw // 2))
    want.backward(gout.double())
    xc, wc = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    bc = bv.cuda().requires_grad_(True) if bias else None
    y = ops.smallconv(xc, wc, bc)
    assert rel_err(y, want) < TOL
    y.backward(gout.cuda())
    assert rel_err(xc.grad, xd.grad) < TOL
    assert rel_err(wc.grad, wd.grad) < TOL
    if bias:
        assert rel_err(bc.grad, bd.grad) < TOL


def test_smallconv_unsupported_shapes_are_reported():
    from km_unet_b200 import ops
    assert not ops.smallconv_supported(5, 16, 3, 3)       # conv_f: 5 input frames
    assert not ops.smallconv_supported(32, 32, 5, 5)      # more than 9 taps


@pytest.mark.parametrize("B,C,S", [(2, 16, 32), (3, 32, 12), (2, 64, 8), (4, 16, 128), (1, 16, 2)])
def test_iwp_vs_matrix_form(B, C, S):
    """Against the banded-matrix definition of WPL/iwp.py:58-132 (oracle.model._iwp, fp64 autograd)."""
    from km_unet_b200 import ops
    from oracle.model import _iwp
    torch.manual_seed(C + S)
    x = torch.randn(B, C, S, S)
    w = torch.randn(C, C + 1, 1, 1) / C ** 0.5
    bv = torch.randn(C) * 0.3
    gout = torch.randn(B, C, S // 2, S // 2)
    xd, wd, bd = x.double().requires_grad_(True), w.double().requires_grad_(True), bv.double().requires_grad_(True)
    want = _iwp(xd, wd, bd)
    want.backward(gout.double())
    xc, wc, bc = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True), bv.cuda().requires_grad_(True)
    y = ops.iwp(xc, wc, bc)
    assert rel_err(y, want) < TOL
    y.backward(gout.cuda())
    assert rel_err(xc.grad, xd.grad) < TOL
    assert rel_err(wc.grad, wd.grad) < TOL
    assert rel_err(bc.grad, bd.grad) < TOL


@pytest.mark.parametrize("B,Cin,Cout,H,W,bias", [(2, 16, 64, 32, 32, True), (3, 64, 16, 7, 9, False), (2, 16, 48, 16, 16, True),
                                                (1, 256, 64, 8, 8, True), (2, 64, 256, 12, 12, False), (2, 128, 32, 5, 5, True),
                                                (8, 16, 64, 128, 128, True), (2, 32, 32, 64, 64, True), (2, 64, 192, 32, 32, True)])
def test_pwconv_tcgen05_within_2e2(B, Cin, Cout, H, W, bias):
    """bf16 tensor-core path of the pointwise convolution (forward, dgrad, MN-major wgrad with the ones-column bias gradient)."""
    from km_unet_b200 import ops
    torch.manual_seed(Cin + Cout + H)
    x = torch.randn(B, Cin, H, W)
    w = torch.randn(Cout, Cin, 1, 1) / Cin ** 0.5
    bv = torch.randn(Cout) if bias else None
    gout = torch.randn(B, Cout, H, W)
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    bd = bv.double().requires_grad_(True) if bias else None
    want = F.conv2d(xd, wd, bd)
    want.backward(gout.double())
    xc, wc = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    bc = bv.cuda().requires_grad_(True) if bias else None
    y = ops.pwconv(xc, wc, bc, ops.KMU_PREC_BF16)
    assert rel_err(y, want) < 2e-2
    y.backward(gout.cuda())
    assert rel_err(xc.grad, xd.grad) < 2e-2
    assert rel_err(wc.grad, wd.grad) < 2e-2
    if bias:
        assert rel_err(bc.grad, bd.grad) < 2e-2


@pytest.mark.parametrize("B,Cin,Cout,H,W,bias", [(2, 16, 64, 32, 32, True), (3, 64, 16, 16, 20, False), (2, 16, 48, 16, 16, True),
                                                (2, 32, 128, 24, 24, True), (2, 128, 32, 16, 16, True), (5, 16, 16, 12, 12, True),
                                                (8, 16, 64, 128, 128, True), (2, 32, 96, 64, 64, False), (40, 64, 64, 16, 8, True),
                                                (2, 32, 32, 13, 12, True)])
def test_pwconv_fused_backward_within_2e2(B, Cin, Cout, H, W, bias):
    """config.conv_bwd = "fused": dx, dW, db from ONE persistent TMA -> tcgen05 kernel (bf16 operands, fp32 accumulation).
    Covers ragged last tiles (HW not a multiple of 128), several tiles per CTA, both plane-buffer / stage plans."""
    import km_unet_b200 as K
    from km_unet_b200 import _lib, ops
    import ctypes as C
    assert _lib.lib().kmu_pwconv_fused_bwd_supported(C.byref(ops.PwDesc(B, Cin, Cout, H * W)))
    torch.manual_seed(Cin + Cout + H)
    x = torch.randn(B, Cin, H, W)
    w = torch.randn(Cout, Cin, 1, 1) / Cin ** 0.5
    bv = torch.randn(Cout) if bias else None
    gout = torch.randn(B, Cout, H, W)
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    bd = bv.double().requires_grad_(True) if bias else None
    want = F.conv2d(xd, wd, bd)
    want.backward(gout.double())
    xc, wc = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    bc = bv.cuda().requires_grad_(True) if bias else None
    old = K.config.conv_bwd
    K.config.conv_bwd = "fused"
    try:
        n0 = _lib.launch_count()
        y = ops.pwconv(xc, wc, bc)
        y.backward(gout.cuda())
    finally:
        K.config.conv_bwd = old
    assert rel_err(y, want) < 1e-4                   # the forward stays on the fp32 kernel
    assert rel_err(xc.grad, xd.grad) < 2e-2
    assert rel_err(wc.grad, wd.grad) < 2e-2
    if bias:
        assert rel_err(bc.grad, bd.grad) < 2e-2


@pytest.mark.parametrize("B,Cin,Cout,H,W,bias", [(2, 16, 64, 32, 32, True), (3, 64, 16, 16, 20, False), (2, 16, 48, 16, 16, True),
                                                (2, 32, 128, 24, 24, True), (2, 128, 32, 16, 16, True), (2, 64, 256, 16, 16, True),
                                                (8, 16, 64, 128, 128, True), (2, 32, 32, 13, 12, False), (2, 64, 192, 32, 32, True)])
def test_pwconv_tma_forward_within_2e2(B, Cin, Cout, H, W, bias):
    """config.conv_fwd = "tma": forward on the persistent TMA -> tcgen05 pipeline (bf16 operands, fp32 accumulation), incl. ragged
    last tiles and the widest accumulators (2 x 256 TMEM columns)."""
    import km_unet_b200 as K
    from km_unet_b200 import _lib, ops
    import ctypes as C
    assert _lib.lib().kmu_pwconv_tma_fwd_supported(C.byref(ops.PwDesc(B, Cin, Cout, H * W)))
    torch.manual_seed(Cin + Cout + H)
    x = torch.randn(B, Cin, H, W)
    w = torch.randn(Cout, Cin, 1, 1) / Cin ** 0.5
    bv = torch.randn(Cout) if bias else None
    want = F.conv2d(x.double(), w.double(), bv.double() if bias else None)
    old = K.config.conv_fwd
    K.config.conv_fwd = "tma"
    try:
        y = ops.pwconv(x.cuda(), w.cuda(), bv.cuda() if bias else None)
    finally:
        K.config.conv_fwd = old
    assert rel_err(y, want) < 2e-2


@pytest.mark.parametrize("B,C,H,W,bias", [(2, 16, 32, 32, True), (3, 32, 17, 20, True), (2, 64, 8, 8, False), (4, 16, 128, 128, True)])
def test_dwconv3x3_with_plane_scale_vs_torch(B, C, H, W, bias):
    """y = scale[b, c] * (dwconv3x3(x) + bias) in one kernel; gradients of x, weight, bias and scale (fp32, 1e-4)."""
    from km_unet_b200 import ops
    torch.manual_seed(H * W + C)
    x = torch.randn(B, C, H, W)
    w = torch.randn(C, 1, 3, 3) * 0.4
    bv = torch.randn(C) if bias else None
    sc = torch.rand(B, C) + 0.25
    gout = torch.randn(B, C, H, W)
    xd, wd, sd = x.double().requires_grad_(True), w.double().requires_grad_(True), sc.double().requires_grad_(True)
    bd = bv.double().requires_grad_(True) if bias else None
    want = F.conv2d(xd, wd, bd, padding=1, groups=C) * sd[:, :, None, None]
    want.backward(gout.double())
    xc, wc, scc = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True), sc.cuda().requires_grad_(True)
    bc = bv.cuda().requires_grad_(True) if bias else None
    y = ops.dwconv3x3(xc, wc, bc, scale=scc)
    assert rel_err(y, want) < TOL
    y.backward(gout.cuda())
    assert rel_err(xc.grad, xd.grad) < TOL
    assert rel_err(wc.grad, wd.grad) < TOL
    assert rel_err(scc.grad, sd.grad) < TOL
    if bias:
        assert rel_err(bc.grad, bd.grad) < TOL


@pytest.mark.parametrize("B,C,H,W", [(2, 16, 32, 32), (5, 32, 12, 10), (32, 16, 64, 64)])
def test_combine3_vs_torch(B, C, H, W):
    """x + sum_i coef[b, i] f_i (EnhancedViMBlock's gated fusion + DropPath + residual) and its gradients (fp32, 1e-4)."""
    from km_unet_b200 import ops
    torch.manual_seed(B + C + H)
    t = [torch.randn(B, C, H, W) for _ in range(4)]
    coef = torch.rand(B, 3)
    gout = torch.randn(B, C, H, W)
    td = [v.double().requires_grad_(True) for v in t]
    cd = coef.double().requires_grad_(True)
    want = td[0] + sum(cd[:, i].reshape(B, 1, 1, 1) * td[i + 1] for i in range(3))
    want.backward(gout.double())
    tc = [v.cuda().requires_grad_(True) for v in t]
    cc = coef.cuda().requires_grad_(True)
    y = ops.combine3(tc[0], tc[1], tc[2], tc[3], cc)
    assert rel_err(y, want) < TOL
    y.backward(gout.cuda())
    for a, b in zip(tc, td):
        assert rel_err(a.grad, b.grad) < TOL
    assert rel_err(cc.grad, cd.grad) < TOL


@pytest.mark.parametrize("B,C,H,W,OH,OW", [(2, 16, 64, 64, 32, 32), (3, 32, 32, 32, 64, 64), (2, 8, 17, 23, 9, 40), (1, 4, 8, 8, 1, 1)])
def test_resize_bilinear_align_corners_vs_torch(B, C, H, W, OH, OW):
    """F.interpolate(bilinear, align_corners=True) forward kernel + gradient (ATen's backward) vs torch on the CPU."""
    from km_unet_b200 import ops
    torch.manual_seed(H + OW)
    x = torch.randn(B, C, H, W)
    xd = x.double().requires_grad_(True)
    want = F.interpolate(xd, size=(OH, OW), mode="bilinear", align_corners=True)
    g = torch.randn(B, C, OH, OW)
    want.backward(g.double())
    xc = x.cuda().requires_grad_(True)
    y = ops.resize_bilinear_ac(xc, (OH, OW))
    assert rel_err(y, want) < 1e-5
    y.backward(g.cuda())
    assert rel_err(xc.grad, xd.grad) < 1e-5


@pytest.mark.parametrize("B,C,H,W,G", [(2, 16, 32, 32, 4), (3, 32, 12, 10, 1), (32, 20, 128, 128, 1), (4, 64, 16, 16, 4)])
def test_groupnorm_vs_torch(B, C, H, W, G):
    """kmu_groupnorm_fwd (split statistics) + the library backward on its mean / rstd vs torch.nn.functional.group_norm in fp64."""
    from km_unet_b200 import ops
    torch.manual_seed(B + C + H)
    x = torch.randn(B, C, H, W) * 1.7 + 0.6
    w, b = torch.randn(C), torch.randn(C)
    g = torch.randn(B, C, H, W)
    xd, wd, bd = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    want = F.group_norm(xd, G, wd, bd, 1e-5)
    want.backward(g.double())
    xc, wc, bc = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    y = ops.groupnorm(xc, wc, bc, G, 1e-5)
    assert rel_err(y, want) < TOL
    y.backward(g.cuda())
    assert rel_err(xc.grad, xd.grad) < TOL
    assert rel_err(wc.grad, wd.grad) < TOL
    assert rel_err(bc.grad, bd.grad) < TOL


@pytest.mark.parametrize("B,C,H,W", [(2, 16, 32, 32), (3, 32, 10, 12), (32, 64, 32, 32)])
def test_lerpmix_vs_torch(B, C, H, W):
    """(1 - sigmoid(alpha_c)) x + sigmoid(alpha_c) m and the gradients of x, m, alpha vs torch.lerp in fp64."""
    from km_unet_b200 import ops
    torch.manual_seed(B + C + H)
    x, m, al, g = torch.randn(B, C, H, W), torch.randn(B, C, H, W), torch.randn(C), torch.randn(B, C, H, W)
    xd, md, ad = x.double().requires_grad_(True), m.double().requires_grad_(True), al.double().requires_grad_(True)
    want = torch.lerp(xd, md, torch.sigmoid(ad).view(1, -1, 1, 1))
    want.backward(g.double())
    xc, mc, ac = x.cuda().requires_grad_(True), m.cuda().requires_grad_(True), al.cuda().requires_grad_(True)
    y = ops.lerpmix(xc, mc, ac)
    assert rel_err(y, want) < TOL
    y.backward(g.cuda())
    assert rel_err(xc.grad, xd.grad) < TOL
    assert rel_err(mc.grad, md.grad) < TOL
    assert rel_err(ac.grad, ad.grad) < TOL
