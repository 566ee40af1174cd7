"""bench.py pieces that run without a GPU: the clock-sample summary and the JSON contract of the reference arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_clock_samples_prefer_the_timed_region_and_fall_back_to_the_loaded_phases():
    import bench
    R = bench.ClockSampler.REASONS
    lines = ["max 1965", "99.0 1200 0", "100.5 1900 4", "101.0 1950 0", "101.5 1965 0", "103.0 1800 8", "garbage", "err x y"]
    a = bench.summarize_clock_samples(lines, 100.9, 101.6, 100.0, 104.0, R)
    assert a["sm_mhz"] == 1957.5 and a["samples"] == 2 and a["reasons"] == [] and a["window"] == "timed region" and a["sm_max_mhz"] == 1965.0
    # a timed region shorter than one poll: the samples of the loaded phases around it, with their throttle reasons
    b = bench.summarize_clock_samples(lines, 101.1, 101.2, 100.0, 104.0, R)
    assert b["samples"] == 4 and b["sm_mhz"] == 1925.0 and b["reasons"] == ["hw_slowdown", "sw_power_cap"] and "loaded phases" in b["window"]
    # nothing at all (NVML unavailable): explicit nulls, never a made-up clock
    c = bench.summarize_clock_samples(["err no nvml"], 1.0, 2.0, 0.0, 3.0, R)
    assert c["sm_mhz"] is None and c["samples"] == 0 and c["window"] is None


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm: the restated reference path on the host cores) on a tiny budget"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                       cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "particle_substeps_per_sec" and line["unit"] == "particle-substeps/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_reference_arm_ignores_the_single_thread_default_of_torchrun():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm must still use the host cores it may run on (round-1 finding)"""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                       cwd=ROOT, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
