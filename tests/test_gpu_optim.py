"""FusedAdamW (csrc/optim.cu, one launch per parameter group) against torch.optim.AdamW -- the optimizer of
train_shanghai.py:342 (`optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.05)`), its cosine schedule (:343) and a checkpoint
round trip.  fp32 arithmetic on both sides; torch evaluates the bias corrections in fp32, the kernel in fp64: 1e-6 relative."""
import copy

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

# ragged on purpose: scalars, odd lengths (scalar tail), lengths around the 1024-element chunk, a multi-chunk tensor, an unaligned view
SHAPES = [(), (1,), (3,), (17,), (16, 5, 3, 3), (1023,), (1024,), (1025,), (64, 64, 3, 3), (4, 16), (300_001,)]


def _params(seed, shapes=SHAPES):
    g = torch.Generator().manual_seed(seed)
    ps = [torch.nn.Parameter((torch.randn(s, generator=g) * 0.5).cuda()) for s in shapes]
    base = (torch.randn(64, generator=g) * 0.5).cuda()
    ps.append(torch.nn.Parameter(base[1:38]))          # storage offset of 4 bytes: not 16-byte aligned, takes the scalar path
    assert ps[-1].data_ptr() % 16 != 0
    return ps


def _grads(ps, seed):
    g = torch.Generator().manual_seed(seed)
    for p in ps:
        p.grad = (torch.randn(p.shape, generator=g) * (10.0 ** torch.randint(-3, 2, (1,), generator=g).item())).cuda()


@pytest.mark.parametrize("hyper", [dict(lr=1e-3, weight_decay=0.05), dict(lr=3e-2, betas=(0.8, 0.95), eps=1e-6, weight_decay=0.0)])
def test_fused_adamw_matches_torch_adamw_over_steps_and_a_schedule(hyper):
    from km_unet_b200 import FusedAdamW
    a, b = _params(1), _params(1)
    oa, ob = FusedAdamW(a, **hyper), torch.optim.AdamW(b, **hyper)
    sa = torch.optim.lr_scheduler.CosineAnnealingLR(oa, T_max=6)
    sb = torch.optim.lr_scheduler.CosineAnnealingLR(ob, T_max=6)
    for it in range(8):
        _grads(a, 100 + it)
        _grads(b, 100 + it)
        oa.step()
        ob.step()
        sa.step()
        sb.step()
        for x, y in zip(a, b):
            assert rel_err(x, y) <= 1e-6, (it, tuple(x.shape))
    for x, y in zip(a, b):
        assert rel_err(oa.state[x]["exp_avg"], ob.state[y]["exp_avg"]) <= 1e-6
        assert rel_err(oa.state[x]["exp_avg_sq"], ob.state[y]["exp_avg_sq"]) <= 1e-6
    assert float(oa.state[a[0]]["step"]) == 8.0


def test_fused_adamw_state_dict_round_trip_and_torch_interchange():
    from km_unet_b200 import FusedAdamW
    a, b = _params(2), _params(2)
    oa, ob = FusedAdamW(a, lr=1e-3, weight_decay=0.05), torch.optim.AdamW(b, lr=1e-3, weight_decay=0.05)
    for it in range(3):
        _grads(a, it)
        _grads(b, it)
        oa.step()
        ob.step()
    # our state into a fresh FusedAdamW and torch's state into another: both continue like torch does
    c, d = [torch.nn.Parameter(p.detach().clone()) for p in a], [torch.nn.Parameter(p.detach().clone()) for p in b]
    oc, od = FusedAdamW(c, lr=1e-3, weight_decay=0.05), FusedAdamW(d, lr=1e-3, weight_decay=0.05)
    oc.load_state_dict(copy.deepcopy(oa.state_dict()))
    od.load_state_dict(copy.deepcopy(ob.state_dict()))
    for it in range(3, 6):
        for ps in (b, c, d):
            _grads(ps, it)
        ob.step()
        oc.step()
        od.step()
    for x, y, z in zip(b, c, d):
        assert rel_err(y, x) <= 1e-6 and rel_err(z, x) <= 1e-6
    assert float(oc.state[c[0]]["step"]) == 6.0 and float(od.state[d[0]]["step"]) == 6.0


def test_fused_adamw_inside_a_cuda_graph_and_with_a_device_learning_rate():
    from km_unet_b200 import FusedAdamW
    a, b = _params(3), _params(3)
    lr = torch.tensor(1e-3, device="cuda")
    oa = FusedAdamW(a, lr=lr, weight_decay=0.05)
    ob = torch.optim.AdamW(b, lr=1e-3, weight_decay=0.05)
    static = [torch.zeros_like(p) for p in a]
    for p, s in zip(a, static):
        p.grad = s
    _grads(b, 0)
    for s, q in zip(static, b):
        s.copy_(q.grad)
    oa.step()                            # eager step 1 creates the moments
    ob.step()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(graph, stream=side):
        oa.step()
    for it in range(1, 5):
        _grads(b, it)
        for s, q in zip(static, b):
            s.copy_(q.grad)
        if it == 3:
            lr.fill_(5e-4)               # the schedule moves a device scalar: no re-capture
            ob.param_groups[0]["lr"] = 5e-4
        graph.replay()
        ob.step()
        for x, y in zip(a, b):
            assert rel_err(x, y) <= 1e-6, (it, tuple(x.shape))
    assert float(oa.state[a[0]]["step"]) == 5.0


def test_fused_adamw_rejects_what_it_does_not_cover():
    from km_unet_b200 import FusedAdamW
    p = torch.nn.Parameter(torch.randn(8))
    p.grad = torch.randn(8)
    with pytest.raises(RuntimeError, match="CUDA"):
        FusedAdamW([p]).step()
    q = torch.nn.Parameter(torch.randn(8, device="cuda", dtype=torch.float16))
    q.grad = torch.randn(8, device="cuda", dtype=torch.float16)
    with pytest.raises(RuntimeError, match="float32"):
        FusedAdamW([q]).step()
    with pytest.raises(ValueError):
        FusedAdamW([torch.nn.Parameter(torch.zeros(1, device="cuda"))], betas=(1.0, 0.9))


def test_fused_adamw_capture_without_an_eager_step_raises():
    from km_unet_b200 import FusedAdamW
    a = _params(4)
    oa = FusedAdamW(a, lr=1e-3)
    for p in a:
        p.grad = torch.zeros_like(p)
    torch.cuda.synchronize()
    g0 = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with pytest.raises(RuntimeError, match="before CUDA-graph capture"):
        with torch.cuda.graph(g0, stream=side):
            oa.step()
    del g0
    torch.cuda.synchronize()
    oa.step()                                           # and the optimizer is still usable afterwards
    assert float(oa.state[a[0]]["step"]) == 1.0


def test_fused_adamw_leaves_a_tensor_without_gradient_untouched():
    """The tensors of a group share ONE step counter (torch keeps one per tensor): a tensor that sits a step out is not updated,
    but its bias corrections advance with the group's."""
    from km_unet_b200 import FusedAdamW
    a = _params(5)
    oa = FusedAdamW(a, lr=1e-2)
    _grads(a, 0)
    oa.step()
    keep = a[2].detach().clone()
    m = oa.state[a[2]]["exp_avg"].clone()
    _grads(a, 1)
    a[2].grad = None
    before = a[3].detach().clone()
    oa.step()
    assert torch.equal(a[2].detach(), keep) and torch.equal(oa.state[a[2]]["exp_avg"], m)
    assert not torch.equal(a[3].detach(), before)
    assert float(oa.state[a[2]]["step"]) == 2.0
