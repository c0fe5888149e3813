"""CPU check of the built objects (no GPU needed): the tensor-core kernel families must compile to tcgen05 / TMEM / TMA SASS for
sm_100a -- UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (cp.async.bulk.tensor) -- and nothing may fall back to mma.sync (HMMA).
The per-object table of the same mnemonics is profiles/r02_sass_summary.md (tools/sass_summary.py)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "km_unet_b200", "_build")

# object -> mnemonics that must occur in it
EXPECT = {
    "kan_tc.o": ["UTCHMMA", "LDTM", "UBLKCP", "SYNCS"],            # KANConv2d forward: implicit GEMM, TMEM accumulators, bulk weight copies
    "kan_tc_bwd.o": ["UTCHMMA", "LDTM", "SYNCS"],                  # KANConv2d dX / dW
    "hsm_fused.o": ["UTCHMMA", "LDTM", "REDUX"],                   # HSM-SSD sweeps with P in TMEM, tile max through redux.sync
    "hsm_tc_bwd.o": ["UTCHMMA", "UTMALDG", "SYNCS"],               # HSM-SSD dgrad / wgrad: TMA-fed pipelines
    "pwconv_bwd_tc.o": ["UTCHMMA", "UTMALDG", "LDTM"],             # pointwise TMA forward + fused backward
}


def _sass(obj):
    return subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True, check=True).stdout


def _ops(sass):
    return set(m.group(1) for m in re.finditer(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", sass, flags=re.M))


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="no cuobjdump on PATH")
@pytest.mark.parametrize("obj", sorted(EXPECT))
def test_tensor_core_objects_hold_tcgen05_sass(obj):
    if not os.path.exists(os.path.join(BUILD, obj)):
        pytest.skip("objects not built yet (python -m km_unet_b200.build)")
    sass = _sass(obj)
    assert "sm_100a" in sass
    ops = _ops(sass)
    for want in EXPECT[obj]:
        assert want in ops, f"{obj}: no {want} instruction"
    assert "HMMA" not in ops, f"{obj}: mma.sync in a tcgen05 kernel family"
