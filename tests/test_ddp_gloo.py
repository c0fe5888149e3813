"""N>1 path on CPU: world_size-2 gloo run of the bucketed gradient all-reduce (km_unet_b200/ddp.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out, views=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from km_unet_b200.ddp import BucketedGradAllReduce, broadcast_parameters
    torch.manual_seed(100 + rank)                      # different init per rank: broadcast must fix it
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    dead = torch.nn.Parameter(torch.ones(4))           # never used, like the dead KAN twin in StableHybridKANConv
    broadcast_parameters(net)
    red = BucketedGradAllReduce(list(net.parameters()), bucket_bytes=64, grad_views=views)   # tiny buckets -> several of them
    assert len(red.buckets) >= 2
    g = torch.Generator().manual_seed(7)
    data = torch.randn(world, 4, 6, generator=g)
    for step in range(2):                              # two steps: hooks must re-arm
        net.zero_grad(set_to_none=True)
        loss = net(data[rank]).square().mean()
        loss.backward()
        nbytes = red.finish()
        if views:                                      # .grad now aliases the flat buckets: no copy back
            flat_ptrs = {f.data_ptr(): f.numel() * 4 for _, f in red.buckets}
            assert all(any(b <= p.grad.data_ptr() < b + n for b, n in flat_ptrs.items()) for p in net.parameters())
    grads = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    if rank == 0:
        torch.save({"grads": grads, "state": net.state_dict(), "nbytes": nbytes}, out)
    dist.barrier()
    dist.destroy_process_group()
    assert dead.grad is None


import pytest  # noqa: E402


@pytest.mark.parametrize("views", [False, True])
def test_bucketed_allreduce_equals_mean_of_per_rank_grads(tmp_path, views):
    world = 2
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(world, _free_port(), out, views), nprocs=world, join=True)
    got = torch.load(out)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    net.load_state_dict(got["state"])
    g = torch.Generator().manual_seed(7)
    data = torch.randn(world, 4, 6, generator=g)
    per_rank = []
    for r in range(world):
        net.zero_grad(set_to_none=True)
        net(data[r]).square().mean().backward()
        per_rank.append(torch.cat([p.grad.reshape(-1) for p in net.parameters()]))
    want = torch.stack(per_rank).mean(0)
    assert torch.allclose(got["grads"], want, atol=1e-6)
    assert got["nbytes"] == 4 * want.numel()


def test_single_process_is_identity():
    from km_unet_b200.ddp import BucketedGradAllReduce
    lin = torch.nn.Linear(3, 2)
    red = BucketedGradAllReduce(list(lin.parameters()))
    lin(torch.ones(1, 3)).sum().backward()
    before = lin.weight.grad.clone()
    red.finish()
    assert torch.equal(lin.weight.grad, before)


def test_second_backward_before_finish_raises_and_no_sync_accumulates():
    """a second backward() before finish() used to re-launch buckets (gradient / world^2, racing all-reduces)."""
    import pytest
    from km_unet_b200.ddp import BucketedGradAllReduce
    lin = torch.nn.Linear(3, 2)
    red = BucketedGradAllReduce(list(lin.parameters()))
    x = torch.ones(1, 3)
    lin(x).sum().backward()
    with pytest.raises(RuntimeError, match="second gradient"):
        lin(x).sum().backward()
    red.reset()
    lin.zero_grad(set_to_none=True)
    with red.no_sync():                                 # accumulation step: hooks disarmed
        lin(x).sum().backward()
    lin(x).sum().backward()                             # the step that reduces the accumulated sum
    red.finish()
    assert torch.allclose(lin.weight.grad, torch.full((2, 3), 2.0))
    lin.zero_grad(set_to_none=True)                     # re-armed for the next step
    lin(x).sum().backward()
    red.finish()
    assert torch.allclose(lin.weight.grad, torch.ones(2, 3))


def test_bucket_with_a_missing_gradient_is_reduced_by_finish():
    """A parameter that received no gradient this step (same set on every rank, per the contract): finish() zero-fills it and reduces
    the bucket; with grad_views the gradients afterwards alias the flat bucket."""
    from km_unet_b200.ddp import BucketedGradAllReduce
    a, b = torch.nn.Linear(3, 2), torch.nn.Linear(3, 2)
    params = list(a.parameters()) + list(b.parameters())
    red = BucketedGradAllReduce(params, bucket_bytes=1 << 20, grad_views=True)       # one bucket holding all four tensors
    a(torch.ones(1, 3)).sum().backward()                                             # b gets no gradient
    assert red._launched == [False]
    nbytes = red.finish()
    assert nbytes == 4 * sum(p.numel() for p in params)
    assert torch.equal(a.weight.grad, torch.ones(2, 3)) and torch.equal(b.weight.grad, torch.zeros(2, 3))
    flat = red.buckets[0][1]
    for p in params:                                                                 # views of the flat buffer, no copy back
        assert p.grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr()
    # next step: fresh gradients accumulate into the views in place (zero_grad(set_to_none=False)) and the bucket launches from the hooks
    for p in params:
        p.grad.zero_()
    (a(torch.ones(1, 3)).sum() + b(torch.ones(1, 3)).sum()).backward()
    assert red._launched == [True]
    red.finish()
    assert torch.equal(b.weight.grad, torch.ones(2, 3))
