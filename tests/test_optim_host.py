"""Host side of FusedAdamW (km_unet_b200/optim.py) that needs no GPU: the chunk table kmu_adamw_step walks, argument checks, and
that the product path refuses CPU tensors (no fallback)."""
import numpy as np
import pytest
import torch


def test_chunk_codes_cover_every_element_exactly_once():
    from km_unet_b200.optim import chunk_codes
    numels = [1, 1023, 1024, 1025, 5000, 3]
    codes = chunk_codes(numels, 1024)
    ent, chunk = codes >> 32, codes & 0xFFFFFFFF
    assert codes.dtype == np.int64 and list(np.bincount(ent)) == [1, 1, 1, 2, 5, 1]
    for i, n in enumerate(numels):
        c = np.sort(chunk[ent == i])
        assert list(c) == list(range(len(c))) and (len(c) - 1) * 1024 < n <= len(c) * 1024


def test_constructor_validates_like_torch_adamw():
    from km_unet_b200 import FusedAdamW
    p = [torch.nn.Parameter(torch.zeros(3))]
    with pytest.raises(ValueError):
        FusedAdamW(p, betas=(0.9, 1.0))
    with pytest.raises(ValueError):
        FusedAdamW(p, lr=-1.0)
    with pytest.raises(ValueError):
        FusedAdamW(p, weight_decay=-0.1)
    opt = FusedAdamW(p, lr=1e-3, weight_decay=0.05)
    g = opt.param_groups[0]
    assert g["lr"] == 1e-3 and g["weight_decay"] == 0.05 and g["betas"] == (0.9, 0.999) and g["eps"] == 1e-8 and g["capturable"]


def test_step_on_cpu_parameters_raises_instead_of_falling_back():
    from km_unet_b200 import FusedAdamW
    p = torch.nn.Parameter(torch.randn(8))
    p.grad = torch.randn(8)
    before = p.detach().clone()
    with pytest.raises(RuntimeError, match="CUDA"):
        FusedAdamW([p]).step()
    assert torch.equal(p.detach(), before)


def test_graphed_train_step_rejects_a_non_capturable_optimizer_before_touching_the_gpu():
    """a default AdamW used to fail in the middle of the capture with an opaque error."""
    import pytest
    import torch
    from km_unet_b200.train import GraphedTrainStep
    net = torch.nn.Linear(3, 2)
    opt = torch.optim.AdamW(net.parameters())                       # capturable=False
    with pytest.raises(ValueError, match="capturable=True"):
        GraphedTrainStep(net, torch.nn.MSELoss(), opt, torch.zeros(1, 3), torch.zeros(1, 2))
    with pytest.raises(ValueError):
        GraphedTrainStep(net, torch.nn.MSELoss(), opt, torch.zeros(1, 3), torch.zeros(1, 2), comm="nope")
