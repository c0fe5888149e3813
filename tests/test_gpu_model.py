"""End-to-end GPU parity of the model mirror (all CUDA-backed operators inside the full KM_UNetV3_SH graph) against the
golden forward of the UNMODIFIED reference model (tests/golden/km_unetv3_sh_eval_32.npz), plus training-step sanity."""
import numpy as np
import pytest
import torch

from conftest import Golden, rel_err

pytestmark = pytest.mark.gpu


def _thresholded(x):
    """metrics.py:45-47,106-107: uint16(clip(x,0,1) * 90) >= thr for thr in [20,30,35,40]."""
    q = (np.clip(x, 0, 1) * 90).astype(np.uint16)
    return np.stack([q >= t for t in (20, 30, 35, 40)])


def _scores(pred, truth):
    """CSI / POD / FAR / HSS of metrics.py:258-266 (incl. its non-standard HSS denominator) from binary masks."""
    out = []
    for p, t in zip(pred, truth):
        tp, fn = float((p & t).sum()), float((~p & t).sum())
        fp, tn = float((p & ~t).sum()), float((~p & ~t).sum())
        eps = 1e-6
        csi = tp / (tp + fn + fp + eps)
        pod = tp / (tp + fn + eps)
        far = fp / (tp + fp + eps)
        hss = 2 * (tp * tn - fn * fp) / ((tp + fn) * (fn + tn) + (tp + fp) * (fp + tn) + eps)
        out.append((csi, pod, far, hss))
    return out


def test_full_model_eval_matches_reference_golden():
    import km_unet_b200 as K
    g = Golden("km_unetv3_sh_eval_32")
    K.config.kan_precision = "fp32"
    m = K.KM_UNetV3_SH(num_classes=4)
    m.load_state_dict(g.sd())
    m = m.cuda().eval()
    with torch.no_grad():
        y = m(g.t("in0", "cuda"))
    want = g.t("out0")
    assert y.shape == want.shape
    assert rel_err(y, want) < 1e-4            # fp32 gate of north_star
    # thresholded cloud masks and the scores built from them must be identical on this fixed batch
    truth = _thresholded(np.random.RandomState(0).rand(*want.shape).astype(np.float32))
    a, b = _thresholded(y.cpu().numpy()), _thresholded(want.numpy())
    assert int((a != b).sum()) == 0
    assert _scores(a, truth) == _scores(b, truth)


def test_full_model_eval_bf16_tensor_core_path_within_2e2():
    import km_unet_b200 as K
    g = Golden("km_unetv3_sh_eval_32")
    m = K.KM_UNetV3_SH(num_classes=4)
    m.load_state_dict(g.sd())
    m = m.cuda().eval()
    K.config.kan_precision = K.config.hsm_precision = K.config.conv_precision = "bf16"
    K.config.conv_bwd, K.config.conv_fwd = "fused", "tma"
    try:
        with torch.no_grad():
            y = m(g.t("in0", "cuda"))
    finally:
        K.config.kan_precision = K.config.hsm_precision = K.config.conv_precision = "fp32"
        K.config.conv_bwd, K.config.conv_fwd = "split", "simt"
    assert rel_err(y, g.t("out0")) < 2e-2


@pytest.mark.parametrize("variant,classes", [("SH", 20), ("LAPS", 3)])
def test_training_step_runs_and_all_live_parameters_get_finite_grads(variant, classes):
    import km_unet_b200 as K
    torch.manual_seed(1234)
    m = K.KM_UNetV3(num_classes=classes, variant=variant).cuda().train()
    K.config.kan_precision = K.config.hsm_precision = K.config.conv_precision = "bf16"
    K.config.conv_bwd, K.config.conv_fwd = "fused", "tma"
    try:
        x = torch.rand(2, 5, 64, 64, device="cuda")
        t = torch.rand(2, classes, 64, 64, device="cuda")
        loss = torch.nn.functional.mse_loss(m(x), t)
        loss.backward()
    finally:
        K.config.kan_precision = K.config.hsm_precision = K.config.conv_precision = "fp32"
        K.config.conv_bwd, K.config.conv_fwd = "split", "simt"
    assert torch.isfinite(loss)
    dead = ("branches.", ".attn.1.", "dt_proj.", "high_freq_conv.")
    for k, p in m.named_parameters():
        if any(d in k for d in dead) and ".attn.qkv" not in k and ".attn.conv" not in k and ".attn.fc" not in k:
            continue
        assert p.grad is not None, k
        assert torch.isfinite(p.grad).all(), k
