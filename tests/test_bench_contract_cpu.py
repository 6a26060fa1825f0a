"""The bench.py contract, checked on the committed output of the current engine (profiles/) and on the script's command
line: one JSON line with the driver's keys, the tier's `roofline` and `cpu_baseline` objects, an end-to-end leg that moves
host bytes, and a launch count."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def latest_bench_line():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r1_bench_resnet20_v*_fusedsum.json")))
    assert files, "profiles/ holds no bench line of the current engine"
    return json.load(open(files[-1]))


def test_committed_bench_line_has_the_contract_keys():
    d = latest_bench_line()
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "images/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["data"] == "synthetic" and "workload" in d["config"]
    assert abs(d["value"] - d["n_gpus"] * d["steps"] * d["config"]["images_per_step_per_gpu"] / (d["ms_per_step"] * d["steps"] * 1e-3)) < 1e-6
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    assert d["gpu_launches"] > 10000
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == d["unit"] and c["sample"]
    assert d["clocks"]["sm_mhz"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_bench_command_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], stdout=subprocess.PIPE, text=True, check=True).stdout
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--in-flight", "--no-hybrid"):
        assert flag in out, flag
