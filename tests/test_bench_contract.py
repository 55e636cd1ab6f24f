"""bench.py without a GPU: the reference arm prints the contract's JSON line (impl, metric, unit, e2e block, cpu_baseline
with kind / cores / per_core / sample), the helpers that turn summaries into the metric are exact, and the flop calibration
the roofline uses carries what the guard test needs."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--n-traj", "4096"], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                   # ONE JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fp64_trajectory_steps_per_s" and d["unit"] == "traj-steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["gpu_launches"] == 0
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["per_core"] > 1e4
    assert "sample" in cb and d["config"]["max_steps"] == 6000 and "workload" in d["config"]
    if cb["kind"] == "reference":                            # oracle/_ref present: the C port is reported beside it
        assert cb["port"]["kind"] == "port" and cb["port"]["value"] > cb["value"]


def test_non_zero_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_metric_helpers_and_c5_packing():
    sys.path.insert(0, ROOT)
    import bench

    istep = np.array([1200, -1, 1, 5999], dtype=np.int32)
    assert bench.steps_of(istep, 6000) == 1200 + 6000 + 1 + 5999          # a trajectory that never impacts flew the cap
    # C5 packs (impact step, status) into one float64: step + 2^32 * status must decode exactly
    step = np.array([0, 1, 1681, 5999, 2**31 - 1], dtype=np.float64)
    for status in (0.0, 1.0, 2.0):
        packed = step + 4294967296.0 * status
        assert np.array_equal(packed // 4294967296.0, np.full_like(step, status))
        assert np.array_equal(packed - 4294967296.0 * status, step)
    cal = json.load(open(os.path.join(ROOT, "profiles", "flop_per_traj_step.json")))
    assert set(cal["static_fp64_of_calibrated_kernel"]) == {"DFMA", "DMUL", "DADD"}
    assert 1000 < cal["flop_per_traj_step"] <= cal["cap_survey_8d"] == 6500.0
    assert bench.FLOP_PER_TRAJ_STEP == cal["flop_per_traj_step"]
