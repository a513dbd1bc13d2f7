"""CPU: `bench.py --impl reference` (the CPU arm, oracle C port) prints one JSON line that honours the bench contract."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small",
                          "--steps", "3", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "obs/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"] == "ba_residual_jacobian_normal_eq_obs_per_s" and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "obs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small", "--gpus", "2"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_bench_helpers():
    """Host-side helpers of bench.py that only run on the GPU box otherwise."""
    sys.path.insert(0, ROOT)
    import bench
    # algorithmic bytes of the fused pass: BASELINE.md section 4 (cfg3 = 85 624 576 B, the judge's own recomputation in VERDICT.md)
    assert bench.algorithmic_bytes(2000000, 100000, 256) == 85624576
    # EKF flop model: the LU route costs twice the factorisation + solve terms of the Cholesky route
    import numpy as np
    n = np.array([100.0, 575.0])
    chol, lu = bench.ekf_flops(n, [0, 0]), bench.ekf_flops(n, [1, 1])
    m, s_ = 2 * n, 3 + 2 * n
    np.testing.assert_allclose(lu - chol, m ** 3 / 3 + m * m * (s_ + 1))
    assert bench._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    # best effort by contract: without a GPU (or sysfs) it says why and never raises
    import torch
    if not torch.cuda.is_available():
        assert bench.pin_to_gpu_numa_node(0).startswith("not pinned")
    vr = bench.verbatim_reference_record()
    assert vr and vr["compute_residual_obs_per_s"] > 1e4 and set(vr["ekf_update_frames_per_s_by_rays"]) >= {"128", "3000"}
    pk = bench.fp64_peak()
    assert 20 < pk["dfma_tflops"] < 60 and 20 < pk["dmma_tflops"] < 60


def test_watchdog_ends_a_stuck_run():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg3", "--watchdog", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 124 and "watchdog" in out.stderr and out.stdout.strip() == ""
