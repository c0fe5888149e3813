"""Load the UNMODIFIED reference modules by path (test infrastructure; see oracle/__init__.py).

Only usable where the reference tree exists (this build container: /root/reference).
On the GPU box it does not exist -- `available()` is False there and nothing that runs
on the GPU box (tests -m gpu, smoke(), bench.py) calls `load()`.

The reference's module names (convKAN, vim_block_init, DySample_md, DAGEM_md, ...) are the same
names our drop-in package exports, so the loader imports them with a private view of
sys.modules and hands back the module objects without leaving them registered.
"""
import importlib
import os
import sys
from types import SimpleNamespace

from . import shims

REF_ROOT = os.environ.get("KMU_REFERENCE_ROOT", "/root/reference")

_NAMES = ("convKAN", "vim_block_init", "DySample_md", "DAGEM_md", "WPL",
          "KM_UNetV3_SH", "KM_UNetV3_LAPS", "metrics")
_cache = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "convKAN", "KANlayers.py"))


def _is_ours(name):
    return any(name == n or name.startswith(n + ".") for n in _NAMES)


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
