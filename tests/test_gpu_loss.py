"""HybridLoss (train_shanghai.py:298-326) CUDA path vs the plain-torch formula in fp64."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(2, 20, 128, 128), (3, 3, 64, 40), (1, 4, 24, 32)])
def test_hybrid_loss_value_and_gradient(shape):
    from km_unet_b200.loss import HybridLoss
    torch.manual_seed(sum(shape))
    pred, tgt = torch.rand(shape), torch.rand(shape)
    crit = HybridLoss()
    pd = pred.double().requires_grad_(True)
    want = crit.forward_torch(pd, tgt.double())
    want.backward()
    pc = pred.cuda().requires_grad_(True)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False       # fp32 GEMMs for the SSIM filter: the tolerance below is an fp32 one
    try:
        got = crit(pc, tgt.cuda())
        (got * 3.0).backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert abs(float(got.detach()) - float(want.detach())) < 1e-5 * abs(float(want.detach()))
    assert rel_err(pc.grad, 3.0 * pd.grad) < 1e-4


def test_hybrid_loss_matches_torch_path_under_tf32():
    """With TF32 GEMMs (the bench setting) both formulations see the same filter precision: compare them with each other."""
    from km_unet_b200.loss import HybridLoss
    torch.manual_seed(5)
    pred, tgt = torch.rand(4, 20, 128, 128, device="cuda"), torch.rand(4, 20, 128, 128, device="cuda")
    crit = HybridLoss()
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = pred.clone().requires_grad_(True)
        b = pred.clone().requires_grad_(True)
        la, lb = crit(a, tgt), crit.forward_torch(b, tgt)
        la.backward()
        lb.backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert abs(float(la) - float(lb)) < 2e-3 * abs(float(lb))
    assert rel_err(a.grad, b.grad) < 2e-2
