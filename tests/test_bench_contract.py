"""bench.py's driver-facing contract that can be checked without a GPU: the reference arm prints one JSON line with the
agreed keys, on the same metric / unit / workload description as our arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-budget", "2"], check=True, capture_output=True, text=True, timeout=600, cwd=ROOT).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "ga_iterations_per_sec" and line["unit"] == "iterations/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0
    assert line["value"] > 0 and abs(line["ms_per_step"] * line["value"] - 1e3) < 1e-6 * 1e3
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "oracle port" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, ROOT)
    import bench
    from align3r_b200.synth import CONFIGS
    assert line["config"]["workload"] == bench.workload_name("c2", CONFIGS["c2"])


def test_rank_other_than_zero_prints_nothing():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         check=True, capture_output=True, text=True, timeout=120, cwd=ROOT,
                         env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")).stdout
    assert out.strip() == ""
