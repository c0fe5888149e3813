"""Load the UNMODIFIED reference modules by path (test infrastructure; see oracle/__init__.py).

The reference tree is looked for at $KMU_REFERENCE_ROOT, /root/reference (the build container) and then at the
byte-for-byte mirror oracle/_ref/ that `python -m oracle.make_ref` writes (git-ignored; it travels to the GPU box with
the gpurun snapshot).  Nothing in the product imports this module.

The reference's module names (convKAN, vim_block_init, DySample_md, DAGEM_md, ...) are the same
names our drop-in package exports, so the loader imports them with a private view of
sys.modules and hands back the module objects without leaving them registered.
"""
import importlib
import os
import sys
from types import SimpleNamespace

from . import shims

MIRROR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _find_root():
    for cand in (os.environ.get("KMU_REFERENCE_ROOT"), "/root/reference", MIRROR):
        if cand and os.path.isfile(os.path.join(cand, "convKAN", "KANlayers.py")):
            return cand
    return os.environ.get("KMU_REFERENCE_ROOT", "/root/reference")


REF_ROOT = _find_root()

_NAMES = ("convKAN", "vim_block_init", "DySample_md", "DAGEM_md", "WPL",
          "KM_UNetV3_SH", "KM_UNetV3_LAPS", "metrics")
_cache = None
_cache_dropin = {}


class _NoAutocast:
    """Stand-in for torch.cuda.amp.autocast while the reference's model files are imported: the `@autocast()` decorators at
    KM_UNetV3_SH.py:71,306,327,465 bind fp16 autocast at class-definition time; an fp32 oracle on the GPU needs them inert."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, fn):
        return fn

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "convKAN", "KANlayers.py"))


def _is_ours(name):
    return any(name == n or name.startswith(n + ".") for n in _NAMES)


def load_models(dropin: bool = False, autocast: bool = False) -> SimpleNamespace:
    """Import the reference's KM_UNetV3_SH.py / KM_UNetV3_LAPS.py fresh and return their model classes.

    dropin=False: on top of the reference's own operator modules (the oracle).
    dropin=True : with km_unet_b200.enable_dropin() -- the same unmodified model files on top of the CUDA-backed operators
                  (`from convKAN.KANConv2Dlayers import *`, `from vim_block_init.efficient_vim_init import EfficientViMBlock`,
                  `from DAGEM_md import DAGEM`, `from DySample_md import DySample` resolve to km_unet_b200's modules).
    autocast=False neutralises the fp16 `@autocast()` decorators so both sides compute in fp32."""
    key = (dropin, autocast)
    if key in _cache_dropin:
        return _cache_dropin[key]
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT} (run python -m oracle.make_ref in the build container)")
    shims.install()
    import torch
    saved = {k: v for k, v in sys.modules.items() if _is_ours(k)}
    for k in saved:
        del sys.modules[k]
    saved_path = list(sys.path)
    sys.path.insert(0, REF_ROOT)
    if dropin:
        import km_unet_b200
        km_unet_b200.enable_dropin()               # drop-in directory in FRONT of the reference root
    real_autocast = torch.cuda.amp.autocast
    if not autocast:
        torch.cuda.amp.autocast = _NoAutocast
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sh = importlib.import_module("KM_UNetV3_SH")
            la = importlib.import_module("KM_UNetV3_LAPS")
        ns = SimpleNamespace(KM_UNetV3_SH=sh.KM_UNetV3, KM_UNetV3_LAPS=la.KM_UNetV3, sh_module=sh, laps_module=la,
                             KANConv2d=sh.KANConv2d, EfficientViMBlock=sh.EfficientViMBlock, DySample=sh.DySample, DAGEM=sh.DAGEM)
    finally:
        torch.cuda.amp.autocast = real_autocast
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if _is_ours(k)]:
            del sys.modules[k]
        sys.modules.update(saved)
    _cache_dropin[key] = ns
    return ns


def load(with_models: bool = True) -> SimpleNamespace:
    """Return a namespace of reference classes: KANLinear, KANConv2d, HSMSSD, EfficientViMBlock,
    LayerNorm1D, DySample, DAGEM (+ KM_UNetV3_SH / KM_UNetV3_LAPS model classes, SimplifiedEvaluator)."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    shims.install()
    saved = {k: v for k, v in sys.modules.items() if _is_ours(k)}
    for k in saved:
        del sys.modules[k]
    saved_path = list(sys.path)
    sys.path.insert(0, REF_ROOT)
    try:
        kl = importlib.import_module("convKAN.KANlayers")
        kc = importlib.import_module("convKAN.KANConv2Dlayers")
        ev = importlib.import_module("vim_block_init.efficient_vim_init")
        vu = importlib.import_module("vim_block_init.vim_utils_init")
        dy = importlib.import_module("DySample_md")
        dg = importlib.import_module("DAGEM_md")
        ns = SimpleNamespace(
            KANLinear=kl.KANLinear, KANConv2d=kc.KANConv2d,
            HSMSSD=ev.HSMSSD, EfficientViMBlock=ev.EfficientViMBlock,
            LayerNorm1D=vu.LayerNorm1D, ConvLayer1D=vu.ConvLayer1D, ConvLayer2D=vu.ConvLayer2D, FFN=vu.FFN,
            DySample=dy.DySample, DAGEM=dg.DAGEM,
        )
        if with_models:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                sh = importlib.import_module("KM_UNetV3_SH")
                la = importlib.import_module("KM_UNetV3_LAPS")
                ns.KM_UNetV3_SH = sh.KM_UNetV3
                ns.KM_UNetV3_LAPS = la.KM_UNetV3
                ns.sh_module = sh
                ns.laps_module = la
                try:
                    me = importlib.import_module("metrics")
                    ns.SimplifiedEvaluator = me.SimplifiedEvaluator
                except Exception:  # cv2 / sklearn flavour issues must not break model parity tests
                    ns.SimplifiedEvaluator = None
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if _is_ours(k)]:
            del sys.modules[k]
        sys.modules.update(saved)
    _cache = ns
    return ns
