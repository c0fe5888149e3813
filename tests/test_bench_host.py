"""CPU checks of bench.py's host logic: the algorithmic-work table behind `roofline.achieved` and the contract of
the reference arm under torchrun (only rank 0 works and prints).  bench.py re-points file descriptor 1 on import (its stdout carries
exactly one JSON line), so it is exercised in subprocesses."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench_eval(expr):
    code = f"import bench; bench.emit({expr})"
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_algorithmic_work_matches_survey_8d():
    got = _bench_eval("""{
        'hsm_fwd': bench.op_work('kmu_hsmssd_fwd', (32, 16, 16384)),
        'hsm_bwd': bench.op_work('kmu_hsmssd_bwd', (32, 16, 16384)),
        'kan_fwd': bench.op_work('kmu_kanconv2d_fwd', (32, 64, 128, 128, 64)),
        'kan_bwd': bench.op_work('kmu_kanconv2d_bwd', (32, 64, 128, 128, 64)),
        'dys_fwd': bench.op_work('kmu_dysample_fwd', (32, 64, 64, 64)),
        'dys_bwd': bench.op_work('kmu_dysample_bwd', (32, 64, 64, 64)),
        'unknown': bench.op_work('kmu_not_an_op', (1,)),
    }""")
    n = 32 * 16 * 16384
    assert got["hsm_fwd"] == ["hbm", 2 * 4 * n, "byte"]                   # read x, write y
    assert got["hsm_bwd"] == ["hbm", 4 * 4 * n, "byte"]                   # read x, dy; write dx; re-read x once = 134.2 MB
    f = 2.0 * 32 * 128 * 128 * 64 * 64 * 81                               # 2 M Cout Cin k^2 (G + k + 1): 10.87 GF per image
    assert got["kan_fwd"] == ["tensor", f, "flop"] and abs(f / 32 - 10.87e9) < 0.01e9
    assert got["kan_bwd"] == ["tensor", 2 * f, "flop"]
    px = 32 * 64 * 64 * 64
    assert got["dys_fwd"] == ["hbm", 20.0 * px, "byte"]                   # 4 B C H W (1 + s^2)
    assert got["dys_bwd"] == ["hbm", 24.0 * px, "byte"]                   # 4 B C H W (s^2 + 1 + 1)
    assert got["unknown"][1] == 0.0                                        # an entry point without a figure never inflates a fraction


def test_reference_arm_is_silent_on_nonzero_ranks():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], cwd=ROOT,
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0
    assert r.stdout.strip() == ""


def test_workload_table_names_the_baseline_configs():
    got = _bench_eval("{k: [bench.metric_of(k), bench.describe(k), bench.cpu_sample_batch(k)] for k in ('model', 'laps', 'infer', 'kan')}")
    assert got["model"][0] == "KM_UNetV3_SH train samples/sec" and "configs[2]" in got["model"][1] and got["model"][2] == 2
    assert "configs[3]" in got["laps"][1] and "configs[4]" in got["infer"][1] and "configs[1]" in got["kan"][1]
