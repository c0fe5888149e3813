"""Data-parallel training step of KM_UNetV3 as ONE CUDA graph (host side of SURVEY section 8e / 7.3 item 8).

The network is tiny (7.7 GFLOP/sample) and a training step issues ~3.4 k kernels, so an eager step is bound by the host's
launch rate, not by the GPU.  Every libkmunet entry point only enqueues on the stream it is given and never synchronises
or allocates, so the whole step is captured once and replayed with a single graph launch:

    zero grads -> forward -> loss -> backward  -> wait buckets -> scatter averaged gradients -> fused AdamW update
                                       |  grad-ready hooks: pack bucket k, all-reduce it on NCCL's stream
                                       |  (captured as a fork of the graph: it runs WHILE the rest of backward runs)

`comm="captured"` (default) records the NCCL all-reduces inside the graph (ddp.BucketedGradAllReduce hooks fire during the
captured backward); `comm="split"` keeps the round-1 scheme (graph A, eager all-reduce of the packed buckets, graph B) as a
fallback.  The reference has no counterpart (it is single-process, eager).
"""
import torch
import torch.distributed as dist

from .ddp import BucketedGradAllReduce


class GraphedTrainStep:
    """Captures `loss = criterion(model(x), t); loss.backward(); all-reduce; optimizer.step()` on static inputs.

    * `optimizer` must be built with `capturable=True` (its step counters live on the device).
    * The constructor runs `warmup` eager steps on the example batch (lazy allocations, cuDNN autotuning, NCCL communicator
      setup must happen outside capture) and then RESTORES parameters, buffers (BatchNorm running statistics,
      num_batches_tracked) and optimizer state in place, so the first replay is step 1 from the state the caller passed in.
    * `__call__` returns `self.loss`, a static tensor that the next replay overwrites (clone it to keep it); `self.out` is the
      network output of the last replay, `p.grad` the (averaged) gradients of the last replay.
    """

    def __init__(self, model, criterion, optimizer, x_example, t_example, world=1, group=None, bucket_bytes=1 << 20, warmup=3,
                 comm="captured"):
        if comm not in ("captured", "split"):
            raise ValueError(comm)
        for g in optimizer.param_groups:
            if not g.get("capturable", False):
                raise ValueError("GraphedTrainStep needs an optimizer built with capturable=True (its step() is captured into a CUDA graph)")
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.world, self.group, self.comm = world, group, comm
        self.x = torch.empty_like(x_example)
        self.t = torch.empty_like(t_example)
        self.x.copy_(x_example)
        self.t.copy_(t_example)
        self.params = [p for g in optimizer.param_groups for p in g["params"]]
        self.reducer = BucketedGradAllReduce(self.params, bucket_bytes=bucket_bytes, group=group,
                                             grad_views=comm == "captured") if world > 1 else None
        if self.reducer is not None and comm == "split":
            self.reducer.remove()                  # no hooks: buckets are packed after backward and reduced between two graphs
        # the parameters' AccumulateGrad nodes may predate this object (created on the default stream): harmless here, the
        # capture below runs every node on the capture stream
        if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        snap_model = [t.detach().clone() for t in list(model.parameters()) + list(model.buffers())]
        snap_opt = {id(p): {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in optimizer.state.get(p, {}).items()}
                    for p in self.params}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._fwd_bwd()
                if self.comm == "split":
                    self._reduce_eager()
                self._update()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():                      # undo the warm-up steps in place (tensor identities are what the graph captures)
            for t, s in zip(list(model.parameters()) + list(model.buffers()), snap_model):
                if not torch.equal(t, s):          # constant buffers (KAN knot tables) keep their version: no cache refresh under capture
                    t.copy_(s)
            for p in self.params:
                for k, v in optimizer.state.get(p, {}).items():
                    if torch.is_tensor(v):
                        old = snap_opt[id(p)].get(k)
                        v.copy_(old) if old is not None else v.zero_()
        optimizer.zero_grad(set_to_none=True)
        self.graph_a = torch.cuda.CUDAGraph()
        self.graph_b = None
        with torch.cuda.graph(self.graph_a):
            self._fwd_bwd()
            if self.comm == "captured" or world == 1:
                self._update()
        if world > 1 and self.comm == "split":
            self.graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_b, pool=self.graph_a.pool()):
                self._update()
        self.graph_launches_per_step = 1 if self.graph_b is None else 2

    def _fwd_bwd(self):
        self.optimizer.zero_grad(set_to_none=True)
        self.out = self.model(self.x)
        self.loss = self.criterion(self.out, self.t)
        self.loss.backward()                       # comm == "captured": the reducer's hooks launch each bucket as it fills
        if self.reducer is not None and self.comm == "split":
            for ps, flat in self.reducer.buckets:
                torch._foreach_copy_(list(flat.split([p.numel() for p in ps])), [p.grad.reshape(-1) for p in ps])

    def _reduce_eager(self):
        if self.reducer is not None:
            for _, flat in self.reducer.buckets:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)

    def _update(self):
        if self.reducer is not None:
            if self.comm == "captured":
                self.reducer.finish()
            else:
                for ps, flat in self.reducer.buckets:
                    flat.div_(self.world)
                    torch._foreach_copy_([p.grad.reshape(-1) for p in ps], list(flat.split([p.numel() for p in ps])))
        self.optimizer.step()

    def close(self):
        """Drop the captured graphs.  With comm="captured" they hold NCCL kernels: the process group must not be destroyed
        (dist.destroy_process_group hangs) while such a graph is alive."""
        self.graph_a = self.graph_b = None
        if self.reducer is not None:
            self.reducer.remove()
        import gc
        gc.collect()
        torch.cuda.synchronize()

    def __call__(self, x=None, t=None):
        """Run one step on (x, t) (copied into the graph's static inputs; None = reuse what is there) -> loss tensor (static)."""
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if t is not None:
            self.t.copy_(t, non_blocking=True)
        self.graph_a.replay()
        if self.graph_b is not None:
            self._reduce_eager()
            self.graph_b.replay()
        return self.loss
