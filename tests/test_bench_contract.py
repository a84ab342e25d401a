"""bench.py contract checks that need no GPU: the reference arm runs on host cores and prints one JSON
line with the keys the driver reads; the CUDA arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "shapes/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "workload" in line["config"] and line["vs_baseline"] is None
    # like for like: the FULL configs[1] batch every step, and the very config object the CUDA arm prints
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.bench_config(1, "weak") and line["config"]["B_per_gpu"] == 4096
    assert line["config"]["score_reduce"] == "batch"
    assert abs(line["value"] - 4096 / (line["ms_per_step"] * 1e-3)) < 1e-6 * line["value"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_cuda_arm_fails_loudly_without_a_gpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_algorithmic_bytes_match_baseline_md():
    sys.path.insert(0, ROOT)
    import bench
    ab = bench.algorithmic_bytes(1, 12, 2048, 1024, 4)
    assert ab["score"] == 49248 and ab["pool_fwd"] == 106496 and ab["fwd"] == 155744
    assert ab["bwd"] == 106496 and ab["fwd_bwd"] == 262240
    ab2 = bench.algorithmic_bytes(1, 12, 2048, 1024, 2)
    assert ab2["fwd"] == 77920 and ab2["fwd_bwd"] == 131168


def test_committed_cuda_arm_line_has_the_contract_keys():
    """The CUDA arm cannot run here; the line it printed on the B200 box (profiles/r03p_bench_n1.json) is checked
    against the contract instead, so a change of bench.py's keys without a re-measurement is caught - and against the
    reference arm's line from the same box (profiles/r03p_bench_reference.json): same config object, like for like."""
    with open(os.path.join(ROOT, "profiles", "r03p_bench_n1.json")) as f:
        line = json.loads(f.read().strip().splitlines()[-1])
    with open(os.path.join(ROOT, "profiles", "r03p_bench_reference.json")) as f:
        ref = json.loads(f.read().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks",
                "shape_mode", "fwd_bwd", "api", "sweep", "config0"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["unit"] == "shapes/s" and line["dtype"] == "f32" and line["vs_baseline"] is None
    assert line["gpu_launches"] == 3 * line["steps"] and "workload" in line["config"] and "l2" in line["config"]
    assert ref["impl"] == "reference" and ref["config"] == line["config"] and ref["metric"] == line["metric"]
    assert line["config"]["score_reduce"] == "batch" and line["config"]["B_per_gpu"] == 4096
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.bench_config(1, "weak")
    r = line["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.9 < r["traffic"] / r["algorithmic_bytes_per_launch"] < 1.1      # no wasted re-reads
    c = line["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    e = line["e2e"]
    assert e["h2d_bytes_per_step"] == 4096 * 12 * (1024 + 2048) * 4 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] < line["value"]                                     # host copies are inside the e2e region
    assert 0.5 < e["frac"] <= 1.05 and e["h2d_peak_gbs"] > 0                  # against the link's measured peak
    assert e["host_path_bits_equal_device_path"] is True
    assert line["api"]["reference_sequence_bits_equal_one_call_path"] is True
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(line["clocks"]["reasons"])
