"""CPU checks of the host-side model mirror (km_unet_b200/modules/km_unet.py): the module tree must carry exactly the
reference's state_dict (keys + shapes, pinned by the golden full-model fixture generated from the unmodified reference),
and the torch-only glue blocks must match the reference modules when the reference tree is present."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, Golden, rel_err


def test_state_dict_layout_matches_reference_fixture():
    from km_unet_b200 import KM_UNetV3_SH
    g = Golden("km_unetv3_sh_eval_32")
    want = g.sd()
    m = KM_UNetV3_SH(num_classes=4)
    have = m.state_dict()
    assert set(have.keys()) == set(want.keys())
    for k, v in want.items():
        assert tuple(have[k].shape) == tuple(v.shape), k
    m.load_state_dict(want)       # strict


def test_laps_variant_has_no_bridge_and_no_dysample():
    from km_unet_b200 import KM_UNetV3_LAPS
    m = KM_UNetV3_LAPS(num_classes=3)
    keys = m.state_dict().keys()
    assert not any(k.startswith("bridge_attention") for k in keys)
    assert not any(k.startswith("dec1.0.") or k.startswith("dec2.0.") or k.startswith("dec3.0.") for k in keys)


@pytest.mark.parametrize("C,S", [(4, 8), (16, 6)])
def test_wavelet_pooling_matches_matrix_form(C, S):
    """Strided 2x2 form vs the banded-matrix definition of WPL/iwp.py:58-103 restated with numpy (last high-pass row /
    column zero), through the same fusion conv."""
    from km_unet_b200.modules.km_unet import IntelligentWaveletPoolingModule
    torch.manual_seed(C)
    m = IntelligentWaveletPoolingModule(C)
    x = torch.randn(2, C, S, S)
    s = 2 ** -0.5
    lo = np.zeros((S // 2, S))
    hi = np.zeros((S // 2, S))
    for i in range(S // 2):
        lo[i, 2 * i], lo[i, 2 * i + 1] = s, s
        if i < S // 2 - 1:
            hi[i, 2 * i], hi[i, 2 * i + 1] = s, -s
    lo, hi = torch.tensor(lo, dtype=torch.float32), torch.tensor(hi, dtype=torch.float32)
    L, Hh = lo @ x, hi @ x
    LL, LH, HL, HH = L @ lo.t(), L @ hi.t(), Hh @ lo.t(), Hh @ hi.t()
    high = torch.cat([LH, HL, HH], dim=1).mean(dim=1, keepdim=True)
    want = m.fusion_conv(torch.cat([LL, high], dim=1))
    from km_unet_b200.modules.km_unet import haar_pool
    ll, hi_mean = haar_pool(x)                      # the strided-slice restatement used for channel counts without a kernel
    assert rel_err(ll, LL) < 1e-6 and rel_err(hi_mean, high) < 1e-6
    from oracle import model as OM
    with OM.cpu_ops():                              # the module itself (CUDA op swapped for the matrix-form oracle)
        assert rel_err(m(x), want) < 1e-6


def _ref_available():
    from oracle import ref_loader
    return ref_loader.available()


@pytest.mark.skipif(not _ref_available(), reason="no reference tree and no oracle/_ref mirror")
@pytest.mark.parametrize("name", ["StableHybridKANConv", "StableHybridKANConv_residual", "TripleNorm", "DirectionAttention", "LocalContrastAttention",
                                  "MultiScaleFusion", "IntelligentWaveletPoolingModule"])
def test_glue_blocks_match_live_reference(name):
    from oracle import ref_loader
    R = ref_loader.load(with_models=True)
    import km_unet_b200.modules.km_unet as M
    torch.manual_seed(11)
    sh = R.sh_module
    if name == "TripleNorm":
        ref, ours, x = sh.TripleNorm(16), M.TripleNorm(16), torch.randn(2, 16, 8, 8)
    elif name == "DirectionAttention":
        ref, ours, x = sh.DirectionAttention(16, "height"), M.DirectionAttention(16, "height"), torch.randn(2, 16, 8, 8)
    elif name == "LocalContrastAttention":
        ref, ours, x = sh.LocalContrastAttention(16), M.LocalContrastAttention(16), torch.randn(2, 16, 8, 8)
    elif name == "MultiScaleFusion":
        ref, ours = sh.MultiScaleFusion([16, 32, 32]), M.MultiScaleFusion([16, 32, 32])
        x = [torch.randn(2, 16, 8, 8), torch.randn(2, 32, 8, 8), torch.randn(2, 32, 8, 8)]
    elif name == "IntelligentWaveletPoolingModule":
        ref, ours, x = sh.IntelligentWaveletPoolingModule(16), M.IntelligentWaveletPoolingModule(16), torch.randn(2, 16, 8, 8)
    elif name == "StableHybridKANConv":          # K5, identity residual (KM_UNetV3_SH.py:72-94); the KAN op runs as its CPU oracle
        ref, ours, x = sh.StableHybridKANConv(16, 16), M.StableHybridKANConv(16, 16), torch.randn(2, 16, 8, 8) * 1.5
    else:                                        # K5 with the 1x1 residual convolution
        ref, ours, x = sh.StableHybridKANConv(16, 32), M.StableHybridKANConv(16, 32), torch.randn(2, 16, 8, 8) * 1.5
    with torch.no_grad():
        for p in ref.parameters():
            p.add_(torch.randn_like(p) * 0.1)
    ours.load_state_dict(ref.state_dict())
    from oracle import model as OM
    with OM.cpu_ops():
        got = ours(x)
    assert rel_err(got, ref(x)) < 1e-5


def test_full_model_cpu_oracle_matches_reference_golden():
    """oracle/model.py (model mirror glue + the op restatements of oracle/) against the forward of the unmodified
    reference model: this pins the full-model oracle used for GPU parity, smoke() and the CPU baseline."""
    from km_unet_b200 import KM_UNetV3_SH
    from oracle import model as OM
    g = Golden("km_unetv3_sh_eval_32")
    m = KM_UNetV3_SH(num_classes=4)
    m.load_state_dict(g.sd())
    m.eval()
    with OM.cpu_ops(), torch.no_grad():
        y = m(g.t("in0"))
    assert rel_err(y, g.t("out0")) < 1e-5


@pytest.mark.parametrize("tag", ["sh", "laps"])
def test_full_model_train_step_cpu_oracle_matches_reference_fp64_fixture(tag):
    """TRAIN mode, end to end: the mirror's glue (DropPath folded into the combine3 coefficients, BatchNorm statistics, dead
    parameters) + the op restatements of oracle/ + oracle/loss.py, in fp64, against output / loss / every live gradient of the
    unmodified reference's own fp64 step (tests/golden/make_golden_train.py).  Exact to the fixture's fp32 storage."""
    import train_fixture as TF
    import km_unet_b200 as K
    from km_unet_b200.modules import km_unet as MM
    from oracle import loss as OL
    from oracle import model as OM
    variant, classes = TF.VARIANTS[tag]
    z = np.load(os.path.join(GOLDEN, f"km_unetv3_{tag}_train_128.npz"))
    torch.manual_seed(TF.SEED_WEIGHTS)
    m = K.KM_UNetV3(num_classes=classes, variant=variant)
    TF.perturb_(m)
    TF.assert_same_state(m.state_dict(), z)
    m = m.double().train()
    x, t = TF.make_batch(classes)
    with OM.cpu_ops(), TF.DropPathReplayer(MM.DropPath, list(z["masks"])):
        out = m(x.double())
        loss = OL.hybrid_loss(out, t.double())
        loss.backward()
    grads = {k: p.grad for k, p in m.named_parameters()}
    assert rel_err(out, torch.from_numpy(z["out0"])) < 1e-6
    assert abs(loss.item() - float(z["loss"])) < 1e-9
    errs = TF.grad_errors(grads, z)
    assert max(v[0] for v in errs.values()) < 1e-6, sorted(errs.items(), key=lambda kv: -kv[1][0])[:3]
    assert TF.grad_global_l2(grads, z) < 1e-6
    for k in z.files:
        if k.startswith("sd_after/"):
            assert rel_err(m.state_dict()[k[9:]], torch.from_numpy(z[k])) < 1e-6, k


def test_product_path_raises_on_cpu_outside_the_oracle_context():
    from km_unet_b200 import KM_UNetV3_SH
    m = KM_UNetV3_SH(num_classes=4).eval()
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 5, 32, 32))


def test_hybrid_loss_ssim_properties():
    from km_unet_b200.loss import HybridLoss, ssim
    torch.manual_seed(0)
    a = torch.rand(2, 3, 32, 32)
    assert abs(ssim(a, a).item() - 1.0) < 1e-6
    b = torch.rand(2, 3, 32, 32)
    assert ssim(a, b).item() < 0.2
    loss = HybridLoss()
    assert loss(a, a).item() < 1e-6
    p = b.clone().requires_grad_(True)
    loss(p, a).backward()
    assert torch.isfinite(p.grad).all()


def test_global_mean_gradient_is_an_expanded_view():
    """The pooled gradient must reach autograd's accumulation as a stride-0 view (what `sum` gives), not as a materialised
    full-size tensor (what `mean` gives): that is the point of km_unet.global_mean."""
    from km_unet_b200.modules.km_unet import global_mean
    x = torch.randn(2, 3, 4, 6, dtype=torch.float64, requires_grad=True)
    y = global_mean(x)
    assert rel_err(y, x.mean(dim=(2, 3))) < 1e-14
    g, = torch.autograd.grad(y.sum(), x)
    assert g.stride()[2:] == (0, 0)
    assert rel_err(g, torch.full_like(x, 1.0 / 24)) < 1e-14


def test_ffn_residual_addcmul_draws_the_same_droppath_mask():
    """EnhancedViMBlock folds `x + DropPath(y)` into one addcmul whose multiplier is DropPath applied to ones: same RNG consumption
    (B bernoulli draws) and the same per-sample factor as timm's DropPath(y)."""
    from km_unet_b200.modules.km_unet import DropPath
    dp = DropPath(0.3).train()
    x, y = torch.randn(5, 2, 3, 3), torch.randn(5, 2, 3, 3)
    torch.manual_seed(3)
    want = x + dp(y)
    after_want = torch.rand(1)
    torch.manual_seed(3)
    got = torch.addcmul(x, y, dp(x.new_ones((5, 1, 1, 1))))
    after_got = torch.rand(1)
    assert torch.equal(after_want, after_got)            # the generator advanced by the same amount
    assert rel_err(got, want) < 1e-6
