"""GPU parity of DAGEM (drop-in module -> ctypes -> C ABI) against the reference's golden vectors and the oracle."""
import pytest
import torch

from conftest import Golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _close(got, want, what, tol=TOL, floor=2e-5):
    # Linear biases that feed a train-mode BatchNorm have a mathematically zero gradient (a sum of B*C*H*W fp32 terms that
    # cancel): those are compared on an absolute floor scaled by the largest gradient of the module
    err = (got.detach().double().cpu() - want.detach().double().cpu()).abs().max().item()
    assert err < tol * want.detach().abs().max().item() + floor, (what, err)


@pytest.mark.parametrize("name,train", [("dagem_8_train", True), ("dagem_8_eval", False)])
def test_dagem_golden(name, train):
    from km_unet_b200 import DAGEM
    g = Golden(name)
    m = DAGEM(input_channels=8)
    m.load_state_dict(g.sd())
    m = m.cuda().train(train)
    x = g.t("in0", "cuda").requires_grad_(True)
    y = m(x)
    assert rel_err(y, g.t("out0")) < TOL
    if g.has("gout"):
        y.backward(g.t("gout", "cuda"))
        assert rel_err(x.grad, g.t("grad_in0")) < TOL
        params = dict(m.named_parameters())
        scale = max(v.abs().max().item() for v in g.grads().values())
        for k, v in g.grads().items():
            _close(params[k].grad, v, k, floor=TOL * scale)
    if train and any(k.startswith("sd_after/") for k in g.z.files):
        after = g.sd(after=True)
        for k, v in m.state_dict().items():
            if "running" in k:
                _close(v, after[k], k)


@pytest.mark.parametrize("B,C,H,W,train", [(2, 64, 16, 16, True), (3, 16, 5, 7, True), (2, 32, 8, 8, False), (1, 8, 1, 3, True)])
def test_dagem_vs_oracle(B, C, H, W, train):
    """Forward, input gradient, every parameter gradient and the running statistics against autograd of oracle/dagem.py
    (fp64), incl. the KM-UNet bridge shape (64 channels, 16x16), a ragged pixel count and eval-mode BatchNorm."""
    from km_unet_b200 import DAGEM
    from oracle import dagem as O
    torch.manual_seed(B * C + H)
    m = DAGEM(input_channels=C)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                mod.weight.uniform_(0.5, 1.5)
                mod.bias.uniform_(-0.2, 0.2)
                mod.running_mean.uniform_(-0.1, 0.1)
                mod.running_var.uniform_(0.5, 1.5)
    m.train(train)
    x = torch.randn(B, C, H, W)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    P = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.double())
         for k, v in sd0.items()}
    xd = x.double().requires_grad_(True)
    want = O.dagem(xd, P, training=train)
    gout = torch.randn(want.shape)
    want.backward(gout.double())
    m = m.cuda()
    xc = x.cuda().requires_grad_(True)
    y = m(xc)
    assert rel_err(y, want) < TOL
    y.backward(gout.cuda())
    assert rel_err(xc.grad, xd.grad) < TOL
    scale = max(P[k].grad.abs().max().item() for k, _ in m.named_parameters())
    for k, p in m.named_parameters():
        _close(p.grad, P[k].grad, k, floor=TOL * scale)
    if train:
        # running statistics follow torch.nn.BatchNorm: momentum 0.1, unbiased variance
        agg_rows = B * C * H * W
        s = (torch.stack([torch.roll(x, 1, 2), torch.roll(x, -1, 2), torch.roll(x, 1, 3), torch.roll(x, -1, 3)], -1) * x.unsqueeze(-1))
        s = s.reshape(-1, 4) @ sd0["edge_aggregation_func.0.weight"].t() + sd0["edge_aggregation_func.0.bias"]
        want_rm = 0.9 * sd0["edge_aggregation_func.1.running_mean"] + 0.1 * s.mean(0)
        want_rv = 0.9 * sd0["edge_aggregation_func.1.running_var"] + 0.1 * s.var(0, unbiased=True)
        assert agg_rows == s.shape[0]
        _close(m.edge_aggregation_func[1].running_mean, want_rm, "running_mean")
        _close(m.edge_aggregation_func[1].running_var, want_rv, "running_var")
        assert int(m.edge_aggregation_func[1].num_batches_tracked) == int(sd0["edge_aggregation_func.1.num_batches_tracked"]) + 1


def test_dagem_is_deterministic():
    from km_unet_b200 import DAGEM
    torch.manual_seed(5)
    m = DAGEM(input_channels=64).cuda()
    x = torch.randn(4, 64, 16, 16, device="cuda")
    outs = []
    for _ in range(2):
        xx = x.clone().requires_grad_(True)
        m.zero_grad()
        y = m(xx)
        y.square().sum().backward()
        outs.append((y.detach().clone(), xx.grad.clone(), m.vertex_update_func[0].weight.grad.clone()))
    # the gating path has no atomics; only torchvision's deform-conv backward may reorder sums
    assert torch.equal(outs[0][0], outs[1][0])


def test_dagem_cpu_tensor_raises():
    from km_unet_b200 import DAGEM
    m = DAGEM(input_channels=8)
    with pytest.raises(RuntimeError):
        m(torch.randn(1, 8, 4, 4))


@pytest.mark.parametrize("B,C,Co,H,W,scale", [(2, 8, 8, 6, 6, 1.0), (2, 64, 64, 16, 16, 0.7), (1, 16, 24, 32, 32, 2.5), (1, 4, 4, 64, 64, 1.5),
                                               (3, 5, 7, 9, 4, 4.0)])
def test_deformconv3x3_vs_torchvision_semantics(B, C, Co, H, W, scale):
    """kmu_deformconv3x3_{fwd,bwd} vs the oracle restatement of torchvision.ops.deform_conv2d (pinned to torchvision's CPU op in
    tests/test_oracle_dagem.py) in fp64, all four gradients; offsets large enough to leave the image."""
    from km_unet_b200 import ops
    from oracle import dagem as OD
    g = torch.Generator().manual_seed(B * 100 + C + H)
    x = torch.randn(B, C, H, W, generator=g)
    off = torch.randn(B, 18, H, W, generator=g) * scale
    w = torch.randn(Co, C, 3, 3, generator=g) * 0.2
    bias = torch.randn(Co, generator=g)
    gout = torch.randn(B, Co, H, W, generator=g)
    ref = [t.double().requires_grad_(True) for t in (x, off, w, bias)]
    want = OD.deform_conv2d(*ref, padding=1)
    want.backward(gout.double())
    got_in = [t.cuda().requires_grad_(True) for t in (x, off, w, bias)]
    got = ops.deformconv3x3(*got_in)
    got.backward(gout.cuda())
    assert rel_err(got, want) < TOL
    for a, b, name in zip(got_in, ref, ("dx", "doffset", "dweight", "dbias")):
        assert rel_err(a.grad, b.grad) < TOL, name
    # bit-reproducible (fixed-order reductions, no atomics)
    again = [t.detach().clone().requires_grad_(True) for t in got_in]
    ops.deformconv3x3(*again).backward(gout.cuda())
    for a, b in zip(got_in, again):
        assert torch.equal(a.grad, b.grad)


def test_dagem_module_graph_replay_equals_eager():
    """torchvision's deform_conv2d is dropped from CUDA-graph captures (legacy-stream launches): the module must not depend on it."""
    from km_unet_b200 import DAGEM
    torch.manual_seed(0)
    m = DAGEM(input_channels=64).cuda().train()
    with torch.no_grad():
        m.offset_conv.weight.mul_(3.0)
    x = torch.randn(4, 64, 16, 16, device="cuda")
    state = {k: v.clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        want = m(x).clone()
        m.load_state_dict(state)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            m(x)
        torch.cuda.current_stream().wait_stream(s)
        m.load_state_dict(state)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = m(x)
        m.load_state_dict(state)
        x2 = torch.randn(4, 64, 16, 16, device="cuda")
        keep = x.clone()
        x.copy_(x2)
        g.replay()                      # different input: proves the replay recomputes the deformable columns
        x.copy_(keep)
        m.load_state_dict(state)
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, want)
