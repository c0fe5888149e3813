"""Data-parallel gradient exchange for the hot path: one process per GPU, bucketed all-reduce (NCCL over NVLink on the
GPU box, gloo in CPU tests) launched from grad-ready hooks so it overlaps the rest of backward.

The reference is single-process: this is the only collective on the path (section 8e).  Batch shards are
independent, BatchNorm statistics stay per replica, so "parity" for N ranks = the all-reduced gradient equals the mean
of the per-rank gradients.
"""
import contextlib

import torch
import torch.distributed as dist


class BucketedGradAllReduce:
    """Average gradients of `params` across ranks.

    Parameters are packed into flat buckets (reverse registration order ~ the order backward produces them).  When the
    last gradient of a bucket has been accumulated, the bucket is copied into its flat buffer and an async all-reduce
    is issued from the hook -- while the rest of backward is still running; `finish()` waits for all buckets and scatters
    the averaged values back into `.grad`.  The hooks and `finish()` only enqueue work, so a whole step (backward, the
    all-reduces on NCCL's stream, the scatter) can be captured into ONE CUDA graph (train.GraphedTrainStep).

    Contract:
      * exactly one backward() between two finish() calls.  A second backward before finish() would re-launch a bucket whose
        all-reduce may still be in flight, so it raises; accumulate gradients under `no_sync()` instead (hooks disarmed; the
        step that follows all-reduces the accumulated sum).
      * EVERY rank must produce gradients for the same set of parameters in a step, because buckets are collectives: a
        bucket that did not fill on some rank is launched by finish() on that rank, after its peers launched it mid-backward --
        the order of collectives stays the same on all ranks only if the set of unfilled buckets is the same everywhere.
      * parameters that never receive a gradient (KM-UNet has 448k of them) must not be passed in.
    """

    def __init__(self, params, bucket_bytes=4 << 20, group=None, grad_views=False):
        self.group = group
        # grad_views: finish() does not copy the averaged values back; it re-points every `.grad` at its slice of the flat bucket
        # (no kernel at all).  The next backward produces fresh gradients as usual.
        self.grad_views = grad_views
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets = []           # list of (params, flat buffer)
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            cur.append(p)
            cur_bytes += p.numel() * p.element_size()
            if cur_bytes >= bucket_bytes:
                self._close(cur)
                cur, cur_bytes = [], 0
        if cur:
            self._close(cur)
        self._hooks = []
        self._where = {}
        self._events = {}           # CUDA: one event per parameter, recorded on the stream its gradient was accumulated on
        self._armed = True
        for bi, (ps, _) in enumerate(self.buckets):
            for p in ps:
                self._where[p] = bi
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.reset()

    def _close(self, ps):
        n = sum(p.numel() for p in ps)
        flat = torch.empty(n, dtype=ps[0].dtype, device=ps[0].device)
        self.buckets.append((list(ps), flat))

    def reset(self):
        """Re-arm every bucket for the next step (finish() does this; call it after an aborted backward)."""
        self._pending = [len(ps) for ps, _ in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._handles = [None] * len(self.buckets)
        self._order = []

    @contextlib.contextmanager
    def no_sync(self):
        """Gradient accumulation: backward passes inside this context do not count towards the buckets."""
        old, self._armed = self._armed, False
        try:
            yield
        finally:
            self._armed = old

    def _on_grad(self, p):
        if not self._armed:
            return
        bi = self._where[p]
        if self._launched[bi] or self._pending[bi] == 0:
            raise RuntimeError("BucketedGradAllReduce: a parameter received a second gradient before finish() -- call finish() "
                               "after every backward(), or accumulate under no_sync()")
        if p.is_cuda:
            # backward nodes run on the streams their forward ran on (the model forks parallel branches): the stream that packs the
            # bucket must wait for every gradient's own stream.  Under CUDA-graph capture these become edges of the graph.
            ev = self._events.get(p)
            if ev is None:
                ev = self._events[p] = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(p.device))
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi):
        ps, flat = self.buckets[bi]
        if flat.is_cuda:
            cur = torch.cuda.current_stream(flat.device)
            for p in ps:
                ev = self._events.get(p)
                if ev is not None:
                    cur.wait_event(ev)
        torch._foreach_copy_(list(flat.split([p.numel() for p in ps])), [p.grad.reshape(-1) for p in ps])
        if self.world > 1:
            self._handles[bi] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._launched[bi] = True
        self._order.append(bi)

    def finish(self):
        """Wait for every bucket, write averaged gradients back (once per bucket).  Returns the number of bytes all-reduced."""
        for bi, (ps, _) in enumerate(self.buckets):
            if not self._launched[bi]:         # a parameter got no gradient this step (same set on every rank): reduce what we have
                for p in ps:
                    if p.grad is None:
                        p.grad = torch.zeros_like(p)
                self._launch(bi)
        total = 0
        for bi in self._order:
            h = self._handles[bi]
            if h is not None:
                h.wait()
            ps, flat = self.buckets[bi]
            if self.world > 1:
                flat.div_(self.world)
            pieces = flat.split([p.numel() for p in ps])
            if self.grad_views:
                for p, piece in zip(ps, pieces):
                    p.grad = piece.view_as(p)
            else:
                torch._foreach_copy_([p.grad.reshape(-1) for p in ps], list(pieces))
            total += flat.numel() * flat.element_size()
        self.reset()
        return total

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def broadcast_parameters(module, src=0, group=None):
    """Make every replica start from rank `src`'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
