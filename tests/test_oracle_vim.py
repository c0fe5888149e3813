"""Oracle (oracle/hsmssd.py) vs golden vectors from the reference HSMSSD / LayerNorm1D / EfficientViMBlock."""
import pytest
import torch

from conftest import Golden, rel_err
from oracle import hsmssd as O


def _mixer_args(sd, prefix=""):
    return (sd[prefix + "BCdt_proj.conv.weight"], sd[prefix + "dw.conv.weight"], sd[prefix + "hz_proj.conv.weight"],
            sd[prefix + "out_proj.conv.weight"], sd[prefix + "A"], sd[prefix + "D"])


@pytest.mark.parametrize("name,tol", [("hsmssd_16_L64", 5e-6), ("hsmssd_32_L144", 5e-6), ("hsmssd_16_L64_f64", 1e-12)])
def test_hsmssd_forward_and_grads(name, tol):
    g = Golden(name)
    args = _mixer_args(g.sd())
    x = g.t("in0")
    y, h = O.hsmssd(x, *args)
    assert rel_err(y, g.t("out0")) < tol
    assert rel_err(h, g.t("out1")) < tol
    gr = O.hsmssd_grads(x, g.t("gout"), *args)
    want = g.grads()
    assert rel_err(gr["x"], g.t("grad_in0")) < 20 * tol
    for mine, theirs in (("BCdt_proj", "BCdt_proj.conv.weight"), ("dw", "dw.conv.weight"),
                         ("hz_proj", "hz_proj.conv.weight"), ("out_proj", "out_proj.conv.weight"), ("D", "D")):
        assert rel_err(gr[mine], want[theirs]) < 20 * tol, mine
    # the per-state parameter A is a dead shift under the over-L softmax: reference gradient is ~0
    assert want["A"].abs().max() < 1e-6 * max(1.0, want["D"].abs().max().item())


def test_layernorm1d():
    g = Golden("layernorm1d_16")
    sd = g.sd()
    assert rel_err(O.layernorm1d(g.t("in0"), sd["weight"], sd["bias"]), g.t("out0")) < 2e-6


@pytest.mark.parametrize("name,train", [("vimblock_16_train", True), ("vimblock_16_eval", False),
                                        ("vimblock_16_init", True)])
def test_vim_block_forward(name, train):
    g = Golden(name)
    y, stats = O.vim_block(g.t("in0"), g.sd(), training=train, return_stats=True)
    assert rel_err(y, g.t("out0")) < 5e-6
    if train:
        after = g.sd(after=True)
        for k, v in stats.items():
            assert rel_err(v, after[k]) < 5e-6, k


def test_vim_block_grads_by_autograd_of_oracle():
    g = Golden("vimblock_16_train")
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in g.sd().items()}
    x = g.t("in0").requires_grad_(True)
    O.vim_block(x, sd, training=True).backward(g.t("gout"))
    assert rel_err(x.grad, g.t("grad_in0")) < 5e-5
    for k, v in g.grads().items():
        if k == "mixer.A":
            continue
        assert rel_err(sd[k].grad, v) < 1e-4, k
