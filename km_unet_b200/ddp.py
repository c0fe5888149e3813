"""Data-parallel gradient exchange for the hot path: one process per GPU, bucketed all-reduce (NCCL over NVLink on the
GPU box, gloo in CPU tests) launched from grad-ready hooks so it overlaps the rest of backward.

The reference is single-process: this is the only collective on the path (section 8e).  Batch shards are
independent, BatchNorm statistics stay per replica, so "parity" for N ranks = the all-reduced gradient equals the mean
of the per-rank gradients.
"""
import torch
import torch.distributed as dist


class BucketedGradAllReduce:
    """Average gradients of `params` across ranks.

    Parameters are packed into flat buckets (reverse registration order ~ the order backward produces them).  When the
    last gradient of a bucket has been accumulated, the bucket is copied into its flat buffer and an async all-reduce
    is issued; `finish()` waits for all buckets and scatters the averaged values back into `.grad`.
    Parameters that never receive a gradient (KM-UNet has 448k of them) must not be passed in.
    """

    def __init__(self, params, bucket_bytes=4 << 20, group=None):
        self.group = group
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
        self._pending = {}
        self._handles = []
        self._hooks = []
        self._where = {}
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
        self._pending = {bi: len(ps) for bi, (ps, _) in enumerate(self.buckets)}
        self._handles = []

    def _on_grad(self, p):
        bi = self._where[p]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi):
        ps, flat = self.buckets[bi]
        torch._foreach_copy_(list(flat.split([p.numel() for p in ps])), [p.grad.reshape(-1) for p in ps])
        if self.world > 1:
            h = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._handles.append((bi, h))
        else:
            self._handles.append((bi, None))

    def finish(self):
        """Wait for every bucket, write averaged gradients back.  Returns the number of bytes all-reduced."""
        for bi, left in self._pending.items():
            if left != 0:                      # a parameter got no gradient this step: reduce what we have
                ps, _ = self.buckets[bi]
                for p in ps:
                    if p.grad is None:
                        p.grad = torch.zeros_like(p)
                self._launch(bi)
        total = 0
        for bi, h in self._handles:
            if h is not None:
                h.wait()
            ps, flat = self.buckets[bi]
            if self.world > 1:
                flat.div_(self.world)
            torch._foreach_copy_([p.grad.reshape(-1) for p in ps], list(flat.split([p.numel() for p in ps])))
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
