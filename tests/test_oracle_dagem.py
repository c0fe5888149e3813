"""Oracle (oracle/dagem.py) vs golden vectors from the reference DAGEM, and deform_conv2d vs torchvision CPU."""
import pytest
import torch

from conftest import Golden, rel_err
from oracle import dagem as O


@pytest.mark.parametrize("name,train", [("dagem_8_train", True), ("dagem_8_eval", False)])
def test_dagem_forward(name, train):
    g = Golden(name)
    assert rel_err(O.dagem(g.t("in0"), g.sd(), training=train), g.t("out0")) < 5e-6


def test_dagem_grads_by_autograd_of_oracle():
    g = Golden("dagem_8_train")
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in g.sd().items()}
    x = g.t("in0").requires_grad_(True)
    O.dagem(x, sd, training=True).backward(g.t("gout"))
    assert rel_err(x.grad, g.t("grad_in0")) < 1e-4
    for k, v in g.grads().items():
        # Linear biases that feed a train-mode BatchNorm have a mathematically zero gradient (fp32 noise only)
        assert (sd[k].grad - v).abs().max() < 2e-4 * v.abs().max() + 1e-5, k


def test_deform_conv_restatement_vs_torchvision():
    tv = pytest.importorskip("torchvision.ops")
    torch.manual_seed(3)
    x = torch.randn(2, 6, 7, 5, dtype=torch.float64)
    off = torch.randn(2, 18, 7, 5, dtype=torch.float64) * 2.5     # large offsets: samples leave the image
    w = torch.randn(4, 6, 3, 3, dtype=torch.float64)
    b = torch.randn(4, dtype=torch.float64)
    assert rel_err(O.deform_conv2d(x, off, w, b, padding=1), tv.deform_conv2d(x, off, w, b, padding=1)) < 1e-12
