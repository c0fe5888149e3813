"""Oracle (oracle/dysample.py) vs golden vectors from the reference DySample."""
import pytest
import torch

from conftest import Golden, rel_err
from oracle import dysample as O


@pytest.mark.parametrize("name", ["dysample_8_g4", "dysample_64_init"])
def test_dysample_forward_and_grads(name):
    g = Golden(name)
    sd = g.sd()
    x = g.t("in0")
    assert torch.equal(O.init_pos(2, 4), sd["init_pos"])
    out = O.dysample_lp(x, sd["offset.weight"], sd["offset.bias"], sd["init_pos"])
    assert rel_err(out, g.t("out0")) < 2e-6
    # closed-form backward: sampler grads, then chain through the 1x1 offset conv
    off = torch.nn.functional.conv2d(x, sd["offset.weight"], sd["offset.bias"]) * 0.25 + sd["init_pos"]
    dx_s, doff = O.sample_grads(x, off, g.t("gout"))
    doff = doff * 0.25
    w = sd["offset.weight"].reshape(sd["offset.weight"].shape[0], -1)
    dx = dx_s + torch.einsum("oc,bohw->bchw", w, doff)
    dw = torch.einsum("bohw,bchw->oc", doff, x)
    db = doff.sum(dim=(0, 2, 3))
    want = g.grads()
    assert rel_err(dx, g.t("grad_in0")) < 1e-5
    assert rel_err(dw.reshape(sd["offset.weight"].shape), want["offset.weight"]) < 1e-5
    assert rel_err(db, want["offset.bias"]) < 1e-5


def test_border_clip_zeroes_offset_gradient():
    """init_pos = +-0.25 puts the j=0 sub-pixels of column 0 outside [0, W-1]: their offset gradient must be 0."""
    x = torch.randn(1, 4, 3, 3, dtype=torch.float64)
    off = O.init_pos(2, 4, torch.float64).expand(1, 32, 3, 3).clone()
    _, doff = O.sample_grads(x, off, torch.ones(1, 4, 6, 6, dtype=torch.float64))
    d = doff.reshape(2, 4, 2, 2, 3, 3)
    assert d[0, :, :, 0, :, 0].abs().max() == 0      # x offset, j=0, column 0
    assert d[0, :, :, 1, :, 2].abs().max() == 0      # x offset, j=1, last column
    assert d[1, :, 0, :, 0, :].abs().max() == 0      # y offset, i=0, row 0
