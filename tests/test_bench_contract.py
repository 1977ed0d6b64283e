"""bench.py's driver-facing contract that can be checked without a GPU: the reference arm prints one JSON line with the
agreed keys, on the same metric / unit / workload description as our arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check_line(line, steps, workload):
    assert line["impl"] == "reference" and line["metric"] == "ga_iterations_per_sec" and line["unit"] == "iterations/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == steps
    assert line["value"] > 0 and abs(line["ms_per_step"] * line["value"] - 1e3) < 1e-6 * 1e3
    assert line["e2e"] == {"value": line["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, ROOT)
    import bench
    from align3r_b200.synth import CONFIGS
    assert line["config"]["workload"] == bench.workload_name(workload, CONFIGS[workload])
    assert line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]


def _reference_tree():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_harness as rh
    return rh.reference_available()


def test_reference_arm_runs_the_unmodified_reference():
    """--impl reference: the reference's own PointCloudOptimizer + global_alignment_iter on the named workload (here the
    10-frame configs[0], so that the CPU suite stays short), from /root/reference or its copy oracle/_ref."""
    import pytest
    if not _reference_tree():
        pytest.skip("no reference tree (/root/reference or oracle/_ref)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "2",
                          "--warmup", "1", "--ref-budget", "60"], check=True, capture_output=True, text=True, timeout=900, cwd=ROOT).stdout
    line = json.loads(out.strip().splitlines()[-1])
    _check_line(line, 2, "c1")
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and "unmodified reference" in cb["sample"] and line["config"]["same_config"] is True
    assert line["ms_per_step"] * line["steps"] * 1e-3 < line["seconds_total"]          # the timed steps fit inside the run


def test_reference_arm_falls_back_to_the_port_and_says_so():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-budget", "2"], check=True, capture_output=True, text=True, timeout=600, cwd=ROOT,
                         env=dict(os.environ, A3R_REFERENCE_ROOT="/nonexistent")).stdout
    line = json.loads(out.strip().splitlines()[-1])
    _check_line(line, 1, "c2")
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and "oracle port" in cb["sample"] and "UNAVAILABLE" in cb["sample"]
    assert line["config"]["same_config"] is False


def test_rank_other_than_zero_prints_nothing():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         check=True, capture_output=True, text=True, timeout=120, cwd=ROOT,
                         env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")).stdout
    assert out.strip() == ""
