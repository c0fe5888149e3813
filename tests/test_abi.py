"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/kmunet.h declares (no compute calls), and the drop-in modules keep the reference's state_dict layout."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, Golden
from oracle import ref_loader


@pytest.fixture(scope="module")
def libpath():
    from km_unet_b200 import build
    return build.build()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "kmunet.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kmu_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(libpath):
    from km_unet_b200 import _lib
    handle = ctypes.CDLL(libpath)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in kmunet.h but not exported"
        assert name in _lib.SYMBOLS, f"{name} declared in kmunet.h but not bound in _lib.py"
    for name in _lib.SYMBOLS:
        assert name in declared, f"{name} bound in _lib.py but not declared in kmunet.h"
    lib = _lib.lib()
    assert lib.kmu_version() == 100
    assert isinstance(_lib.last_error(), str)


def test_bad_descriptor_is_reported_not_crashed(libpath):
    from km_unet_b200 import _lib
    lib = _lib.lib()
    d = _lib.KanDesc(1, 4, 8, 8, 8, 3, 1, 1, 7, 5, 0, 1, 0, 0.0, 0.0)       # grid_size 7 / order 5: not implemented
    assert lib.kmu_kanconv2d_fwd_workspace_bytes(ctypes.byref(d)) == 0
    assert "cubic" in _lib.last_error()
    h = _lib.HsmDesc(1, 16, 65, 8, 64)                          # L != H*H
    assert lib.kmu_hsmssd_fwd_workspace_bytes(ctypes.byref(h)) == 0
    assert "H*H" in _lib.last_error()


def test_cuda_sources_target_sm100a():
    from km_unet_b200 import build
    assert "arch=compute_100a,code=sm_100a" in build.NVCC_FLAGS
    assert "-lineinfo" in build.NVCC_FLAGS


@pytest.mark.parametrize("fixture,ctor", [
    ("kanconv2d_4_8_k3p1", lambda M: M.KANConv2d(4, 8, 3, padding=1)),
    ("kanlinear_6_5", lambda M: M.KANLinear(6, 5)),
    ("hsmssd_16_L64", lambda M: M.HSMSSD(d_model=16)),
    ("vimblock_16_train", lambda M: M.EfficientViMBlock(16)),
    ("dysample_8_g4", lambda M: M.DySample(8)),
    ("dagem_8_train", lambda M: M.DAGEM(input_channels=8)),
])
def test_state_dict_layout_matches_reference_checkpoints(fixture, ctor):
    import km_unet_b200.modules as M
    if not hasattr(M, "DAGEM") and fixture.startswith("dagem"):
        pytest.skip("DAGEM drop-in not built yet")
    m = ctor(M)
    ref_sd = Golden(fixture).sd()
    own = m.state_dict()
    assert list(own.keys()) == list(ref_sd.keys())
    for k in own:
        assert own[k].shape == ref_sd[k].shape, k
    m.load_state_dict(ref_sd)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
def test_same_seed_reproduces_reference_init():
    import km_unet_b200.modules as M
    R = ref_loader.load()
    for ours, theirs in ((lambda: M.KANConv2d(8, 16, 3, padding=1), lambda: R.KANConv2d(8, 16, 3, padding=1)),
                         (lambda: M.EfficientViMBlock(16), lambda: R.EfficientViMBlock(16)),
                         (lambda: M.DySample(64), lambda: R.DySample(64))):
        torch.manual_seed(77)
        a = ours().state_dict()
        torch.manual_seed(77)
        b = theirs().state_dict()
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert torch.allclose(a[k].float(), b[k].float(), atol=1e-6), k


def test_dropin_import_paths():
    import importlib
    import sys
    import km_unet_b200
    saved = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("convKAN", "vim_block_init", "DySample_md")}
    for k in saved_mods:
        del sys.modules[k]
    try:
        km_unet_b200.enable_dropin()
        ns = {}
        exec("from convKAN.KANConv2Dlayers import *", ns)
        assert ns["KANConv2d"].__module__.startswith("km_unet_b200")
        assert ns["KANLinear"].__module__.startswith("km_unet_b200")
        ev = importlib.import_module("vim_block_init.efficient_vim_init")
        assert ev.EfficientViMBlock.__module__.startswith("km_unet_b200")
        dy = importlib.import_module("DySample_md")
        assert dy.DySample.__module__.startswith("km_unet_b200")
    finally:
        sys.path[:] = saved
        for k in [k for k in sys.modules if k.split(".")[0] in ("convKAN", "vim_block_init", "DySample_md")]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
