"""AdamW as the reference's training script uses it (train_shanghai.py:342 `optim.AdamW(model.parameters(), lr=1e-3,
weight_decay=0.05)`, stepped at :180), with the update of a whole parameter group as ONE kernel launch (csrc/optim.cu).

Same semantics and the same `state_dict()` layout as `torch.optim.AdamW` (per parameter: `step`, `exp_avg`, `exp_avg_sq`; no amsgrad,
no maximize), so `torch.optim.lr_scheduler.*` and checkpoints work unchanged.  The step counter lives on the device: `step()` only
enqueues work and can be captured into a CUDA graph (train.GraphedTrainStep) -- the moments must exist before capture, i.e. run at
least one eager step first (GraphedTrainStep's warm-up does).  CUDA fp32 parameters only; there is no CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import AdamWArgs, check, stream_ptr


def chunk_codes(numels, chunk):
    """The chunk list kmu_adamw_step walks: one int64 per `chunk` elements of every tensor, (tensor index << 32) | chunk number."""
    return np.concatenate([(np.int64(i) << 32) | np.arange((int(n) + chunk - 1) // chunk, dtype=np.int64) for i, n in enumerate(numels)])


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError(f"invalid betas {betas}")
        if eps < 0.0 or weight_decay < 0.0 or (not torch.is_tensor(lr) and lr < 0.0):
            raise ValueError("lr, eps and weight_decay must be non-negative")
        # `capturable` is what GraphedTrainStep checks for: this optimizer always keeps its counters on the device
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, capturable=True))
        self._tables = {}        # group index -> (key, entries tensor, chunks tensor, n_entries, n_chunks)
        self._pinned = []        # host staging buffers of tables uploaded inside a graph capture: a replay re-reads them
        self._spare = {}         # group index -> page-locked buffer set aside by the last eager step for a later capture
        self.grad_scale = 1.0    # gradients are multiplied by this on the fly (1 / loss scale when training with a GradScaler)

    def _state_of(self, group, ps, capturing):
        step = None
        for p in group["params"]:
            st = self.state.get(p)
            if st and "step" in st:
                step = st["step"]
                break
        for p in ps:
            st = self.state[p]
            if "exp_avg" not in st:
                if capturing:
                    raise RuntimeError("FusedAdamW: the moments must be created before CUDA-graph capture (run one eager step first)")
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("FusedAdamW takes contiguous float32 CUDA parameters (there is no CPU fallback)")
                if step is None:
                    step = torch.zeros((), dtype=torch.float32, device=p.device)
                st["step"] = step                  # one counter per group, shared by its tensors (state_dict stores it per tensor)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        steps = {self.state[p]["step"].data_ptr() for p in ps}
        if len(steps) != 1:                        # after load_state_dict every tensor owns a copy of the counter: share one again
            if capturing:
                raise RuntimeError("FusedAdamW: call step() once eagerly after load_state_dict() before capturing it")
            first = self.state[ps[0]]["step"]
            if not all(float(self.state[p]["step"]) == float(first) for p in ps):
                raise RuntimeError("FusedAdamW: the tensors of one parameter group must share their step count")
            first = first.detach().to(device=ps[0].device, dtype=torch.float32).reshape(()).clone()
            for p in group["params"]:
                if p in self.state and "step" in self.state[p]:
                    self.state[p]["step"] = first
        return self.state[ps[0]]["step"]

    def _table(self, gi, ps, capturing):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(), self.state[p]["exp_avg_sq"].data_ptr(),
                     p.numel()) for p in ps)
        tab = self._tables.get(gi)
        if tab is not None and tab[0] == key and not capturing:
            return tab
        chunk = int(_lib.lib().kmu_adamw_chunk_elems())
        ent = np.array(key, dtype=np.int64)                                         # (n, 5): p, g, m, v, n  == kmu_adamw_entry
        codes = chunk_codes([k[4] for k in key], chunk)
        words = torch.from_numpy(np.concatenate([ent.reshape(-1), codes]))
        if capturing:
            # page-locked memory cannot be allocated while a stream is capturing: take the buffer the last eager step set aside.  The
            # graph's memcpy node re-reads it at every replay, so it is never reused for anything else.
            host = self._spare.pop(gi, None)
            if host is None or host.numel() != words.numel():
                raise RuntimeError("FusedAdamW: run one eager step() with the same set of gradients before capturing step() into a CUDA graph")
            self._pinned.append(host)
        else:
            host = torch.empty(words.numel(), dtype=torch.int64).pin_memory()
        host.copy_(words)
        dev = torch.empty(host.numel(), dtype=torch.int64, device=ps[0].device)
        dev.copy_(host, non_blocking=True)
        tab = (key, dev, ent.size, len(key), int(codes.size), host)
        if not capturing:
            self._tables[gi] = tab
        return tab

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.lib()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                g = p.grad
                if g.is_sparse or g.dtype != torch.float32 or not g.is_contiguous() or g.device != p.device:
                    raise RuntimeError("FusedAdamW takes dense contiguous float32 gradients on the parameter's device")
            capturing = torch.cuda.is_current_stream_capturing()
            step = self._state_of(group, ps, capturing)
            _, dev, n_ent_words, n_entries, n_chunks, _ = self._table(gi, ps, capturing)
            if not capturing and (gi not in self._spare or self._spare[gi].numel() != dev.numel()):
                self._spare[gi] = torch.empty(dev.numel(), dtype=torch.int64).pin_memory()
            lr = group["lr"]
            a = AdamWArgs(dev.data_ptr(), dev.data_ptr() + 8 * n_ent_words, n_entries, n_chunks, step.data_ptr(),
                          lr.data_ptr() if torch.is_tensor(lr) else None, 0.0 if torch.is_tensor(lr) else float(lr),
                          float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]), float(group["weight_decay"]),
                          float(self.grad_scale))
            if torch.is_tensor(lr) and not (lr.is_cuda and lr.dtype == torch.float32):
                raise RuntimeError("FusedAdamW: a tensor learning rate must be a float32 CUDA scalar")
            with torch.cuda.device(ps[0].device):
                check(lib.kmu_adamw_step(C.byref(a), stream_ptr()), "kmu_adamw_step")
        return loss
