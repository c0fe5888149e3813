"""Data-parallel training step of KM_UNetV3 as CUDA graphs (host side of SURVEY section 8e / 7.3 item 8).

The network is tiny (7.7 GFLOP/sample) and a training step issues ~3.4 k kernels, so an eager step is bound by the host's
launch rate, not by the GPU.  Every libkmunet entry point only enqueues on the stream it is given and never synchronises
or allocates, so the whole step -- forward, HybridLoss, backward, gradient flattening, AdamW -- is captured once and
replayed:

    graph A:  zero grads -> forward -> loss -> backward -> pack the live gradients into flat buckets
    eager  :  one NCCL all-reduce per bucket (N > 1 only; 5.1 MB in total, latency-bound)
    graph B:  unpack averaged gradients -> fused AdamW update

With one GPU the two graphs are captured as one.  The reference has no counterpart (it is single-process, eager).
"""
import torch
import torch.distributed as dist


class GraphedTrainStep:
    def __init__(self, model, criterion, optimizer, x_example, t_example, world=1, group=None, bucket_bytes=4 << 20, warmup=3):
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.world, self.group = world, group
        self.x = torch.empty_like(x_example)
        self.t = torch.empty_like(t_example)
        self.x.copy_(x_example)
        self.t.copy_(t_example)
        self.params = [p for g in optimizer.param_groups for p in g["params"]]
        self.buckets = []
        if world > 1:
            cur, cur_bytes = [], 0
            for p in reversed(self.params):
                cur.append(p)
                cur_bytes += p.numel() * p.element_size()
                if cur_bytes >= bucket_bytes:
                    self.buckets.append((cur, torch.empty(sum(q.numel() for q in cur), dtype=p.dtype, device=p.device)))
                    cur, cur_bytes = [], 0
            if cur:
                self.buckets.append((cur, torch.empty(sum(q.numel() for q in cur), dtype=cur[0].dtype, device=cur[0].device)))
        # the parameters' AccumulateGrad nodes may predate this object (created on the default stream): harmless here, the
        # capture below runs every node on the capture stream
        if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._fwd_bwd()
                self._reduce()
                self._update()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        optimizer.zero_grad(set_to_none=True)
        self.graph_a = torch.cuda.CUDAGraph()
        self.graph_b = None
        with torch.cuda.graph(self.graph_a):
            self._fwd_bwd()
            if world == 1:
                self._update()
        if world > 1:
            self.graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_b, pool=self.graph_a.pool()):
                self._update()

    def _fwd_bwd(self):
        self.optimizer.zero_grad(set_to_none=True)
        self.loss = self.criterion(self.model(self.x), self.t)
        self.loss.backward()
        for ps, flat in self.buckets:
            torch._foreach_copy_(list(flat.split([p.numel() for p in ps])), [p.grad.reshape(-1) for p in ps])

    def _reduce(self):
        for _, flat in self.buckets:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)

    def _update(self):
        for ps, flat in self.buckets:
            flat.div_(self.world)
            torch._foreach_copy_([p.grad.reshape(-1) for p in ps], list(flat.split([p.numel() for p in ps])))
        self.optimizer.step()

    def __call__(self, x=None, t=None):
        """Run one step on (x, t) (copied into the graph's static inputs; None = reuse what is there) -> loss tensor."""
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if t is not None:
            self.t.copy_(t, non_blocking=True)
        self.graph_a.replay()
        if self.graph_b is not None:
            self._reduce()
            self.graph_b.replay()
        return self.loss
