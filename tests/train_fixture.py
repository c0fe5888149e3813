"""Shared pieces of the END-TO-END train-mode parity fixtures (tests/golden/km_unetv3_{sh,laps}_train_128.npz).

Used by the generator (tests/golden/make_golden_train.py: runs the UNMODIFIED reference in fp64 on CPU) and by the GPU tests
(tests/test_gpu_model_train.py, tests/test_gpu_reference_dropin.py) so both sides build exactly the same network state,
batch and DropPath masks:

  * weights: torch.manual_seed(SEED_WEIGHTS) + the model constructor (the mirror, the reference and the reference-over-drop-in
    produce the same state_dict from the same seed: tests/test_abi.py), then `perturb_` below.  The fixture stores per-tensor
    checksums of the resulting state_dict instead of 7 MB of weights; the tests refuse to run on a mismatch.
  * `perturb_`: the default init zeroes three BatchNorm gammas per EfficientViMBlock and sets alpha = 1e-4, which would hide
    every branch behind them from an output / gradient comparison -- make every branch live (as make_golden.py does per block).
  * DropPath: the reference draws per-sample bernoulli masks from the global torch RNG (train mode).  The generator records
    them in call order; the tests replay them (same call order in the mirror: fusion sum, then FFN, per EnhancedViMBlock).
  * gradients: every live parameter.  Tensors up to FULL_LIMIT elements are stored whole, larger ones as a fixed strided
    sample + their fp64 sum and L2 norm.
"""
import zlib

import numpy as np
import torch

SEED_WEIGHTS = 1234
SEED_DATA = 20240518
SEED_DROPPATH = 4324       # ten masks over B = 2 with three zeros and no all-zero row: every block keeps a live sample
FULL_LIMIT = 16384
VARIANTS = {"sh": ("SH", 20), "laps": ("LAPS", 3)}


def make_batch(classes, batch=2, size=128, frames_in=5):
    g = torch.Generator().manual_seed(SEED_DATA)
    data = torch.rand(batch, frames_in + classes, size, size, generator=g)
    return data[:, :frames_in].contiguous(), data[:, frames_in:].contiguous()


def _u(key, shape, lo, hi):
    g = torch.Generator().manual_seed(zlib.crc32(key.encode()) & 0x7FFFFFFF)
    return torch.rand(shape, generator=g, dtype=torch.float64) * (hi - lo) + lo


def perturb_(model):
    """Deterministic (keyed by parameter name) in-place perturbation that makes every branch of the network live."""
    sd = model.state_dict()
    with torch.no_grad():
        for k, v in sd.items():
            new = None
            if k.endswith(".weight") and k[:-6] + "running_mean" in sd:                 # BatchNorm gamma (three per block init 0)
                new = _u(k, v.shape, 0.5, 1.5)
            elif k.endswith(".bias") and k[:-4] + "running_mean" in sd:
                new = _u(k, v.shape, -0.2, 0.2)
            elif k.endswith("running_mean"):
                new = _u(k, v.shape, -0.1, 0.1)
            elif k.endswith("running_var"):
                new = _u(k, v.shape, 0.5, 1.5)
            elif k.endswith(".alpha"):                                                   # sigmoid layer scales, init 1e-4
                new = _u(k, v.shape, -1.0, 1.0)
            elif k.endswith(".offset.weight"):                                           # DySample: init std 1e-3 barely moves a sample
                new = _u(k, v.shape, -0.08, 0.08)
            elif k.endswith(".offset.bias"):
                new = _u(k, v.shape, -0.3, 0.3)
            if new is not None:
                v.copy_(new.to(v.dtype))
    return model


def checksums(sd):
    """{key: (sum, sum of squares)} in fp64 -- cheap identity of a state_dict."""
    out = {}
    for k, v in sd.items():
        d = v.detach().double().cpu()
        out[k] = np.array([d.sum().item(), (d * d).sum().item()])
    return out


def assert_same_state(sd, golden, tol=1e-6):
    for k, v in checksums(sd).items():
        want = golden["cs/" + k]
        scale = max(1.0, abs(want[1]))
        assert abs(v[0] - want[0]) <= tol * max(1.0, abs(want[1]) ** 0.5 * 10) and abs(v[1] - want[1]) <= tol * scale, \
            f"state_dict differs from the fixture's at {k}: {v} vs {want}"


def sample_index(numel):
    """Indices of the stored sample of a large tensor: a fixed stride with a fixed phase."""
    if numel <= FULL_LIMIT:
        return None
    stride = -(-numel // FULL_LIMIT)
    return np.arange(stride // 2, numel, stride)


def compress(t):
    """(sample, [sum, l2]) of a tensor for the fixture."""
    d = t.detach().double().cpu().reshape(-1)
    idx = sample_index(d.numel())
    kept = d if idx is None else d[torch.from_numpy(idx)]
    return kept.numpy().astype(np.float32), np.array([d.sum().item(), d.norm().item()])


def grad_floor(maxima):
    """Denominator floor of the per-tensor relative error: a few parameters have a mathematically ZERO gradient (HSMSSD.A is
    inert, conv biases in front of a normalisation, IWP.high_freq_conv behind its one-channel softmax); their fp64 gradient is
    rounding noise (1e-21) and 'relative to its own maximum' means nothing; HSMSSD.D's is a sum with heavy cancellation that is
    1000x smaller than a typical gradient.  Errors are taken relative to
    max(max|want|, 1e-1 * median over the live tensors of max|want|): noise below a tenth of the typical gradient scale times the
    gate is not counted against a tensor whose own gradient is (near) zero."""
    live = [m for m in maxima if m > 1e-12]
    return 1e-1 * float(np.median(live))


class DropPathRecorder:
    """Wraps a DropPath class' forward: records every per-sample mask the module draws (train mode)."""

    def __init__(self, cls):
        self.cls, self.masks, self._orig = cls, [], None

    def __enter__(self):
        rec = self
        self._orig = self.cls.forward

        def forward(mod, x):
            if not mod.training or mod.drop_prob == 0.0:
                return x
            keep = 1.0 - mod.drop_prob
            mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
            rec.masks.append(mask.detach().reshape(-1).double().cpu().numpy().copy())
            if keep > 0.0 and mod.scale_by_keep:
                mask = mask / keep
            return x * mask
        self.cls.forward = forward
        return self

    def __exit__(self, *exc):
        self.cls.forward = self._orig
        return False


class DropPathReplayer:
    """Wraps a DropPath class' forward: replays recorded bernoulli masks in call order instead of drawing new ones."""

    def __init__(self, cls, masks):
        self.cls, self.masks, self.i, self._orig = cls, masks, 0, None

    def __enter__(self):
        rep = self
        self._orig = self.cls.forward
        self.i = 0

        def forward(mod, x):
            if not mod.training or mod.drop_prob == 0.0:
                return x
            keep = 1.0 - mod.drop_prob
            m = torch.as_tensor(rep.masks[rep.i], dtype=x.dtype, device=x.device).reshape((x.shape[0],) + (1,) * (x.dim() - 1))
            rep.i += 1
            if keep > 0.0 and mod.scale_by_keep:
                m = m / keep
            return x * m
        self.cls.forward = forward
        return self

    def __exit__(self, *exc):
        self.cls.forward = self._orig
        return False


def grad_errors(named_grads, golden):
    """Per-parameter relative error max|g - want| / max|want| against the fixture (full tensors or their stored samples),
    plus the relative error of the L2 norm.  Returns {key: (err, norm_err)} over the fixture's live parameters."""
    out = {}
    floor = float(golden["gfloor"])
    for k in golden.files:
        if not k.startswith("grad/"):
            continue
        name = k[5:]
        want = golden[k].astype(np.float64)
        if named_grads.get(name) is None:
            # the reference gives IWP.high_freq_conv an exactly-zero gradient (softmax over ONE channel); the mirror never runs it
            out[name] = (0.0, 0.0) if np.abs(want).max() == 0.0 else (float("inf"), float("inf"))
            continue
        g = named_grads[name].detach().double().cpu().reshape(-1)
        idx = sample_index(g.numel())
        got = g.numpy() if idx is None else g.numpy()[idx]
        err = np.abs(got - want).max() / max(np.abs(want).max(), floor)
        wsum, wnorm = golden["gstat/" + name]
        nerr = abs(g.norm().item() - wnorm) / max(wnorm, floor * np.sqrt(g.numel()))
        out[name] = (float(err), float(nerr))
    return out


def grad_global_l2(named_grads, golden):
    """||g - want|| / ||want|| over the stored samples of ALL live parameters taken as one vector."""
    num = den = 0.0
    for k in golden.files:
        if not k.startswith("grad/"):
            continue
        want = golden[k].astype(np.float64)
        if named_grads.get(k[5:]) is None:
            num += float((want ** 2).sum())
            den += float((want ** 2).sum())
            continue
        g = named_grads[k[5:]].detach().double().cpu().reshape(-1).numpy()
        idx = sample_index(g.size)
        got = g if idx is None else g[idx]
        num += float(((got - want) ** 2).sum())
        den += float((want ** 2).sum())
    return (num / den) ** 0.5
