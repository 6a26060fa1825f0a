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
    """the newest committed single-GPU line of the default mode (profiles/r2_bench_resnet20_v<k>_*.json)"""
    best = None
    for f in glob.glob(os.path.join(ROOT, "profiles", "r2_bench_resnet20_v*.json")):
        d = json.load(open(f))
        if d.get("n_gpus") != 1 or d.get("keys", {}).get("seed_compressed"):
            continue
        v = int(os.path.basename(f).split("_v")[1].split("_")[0])
        if best is None or v > best[0]:
            best = (v, d)
    assert best, "profiles/ holds no bench line of the current engine"
    return best[1]


def test_committed_bench_line_has_the_contract_keys():
    d = latest_bench_line()
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline",
              "exact", "key_switch_us", "keys"):
        assert k in d, k
    assert d["unit"] == "images/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["data"].startswith("synthetic") and "workload" in d["config"]
    assert "model" not in d["config"]
    assert abs(d["value"] - d["n_gpus"] * d["config"]["images_per_step_per_unit"] / (d["ms_per_step"] * 1e-3)) < 1e-6
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    assert d["gpu_launches"] > 10000
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert r["traffic"] is None or r["traffic"] > 0
    ir = r["int_roofline"]            # measured in the run: per-form multiplier rates, no constants from a profile
    assert set(ir["rates_per_s"]) == {"mad_lo", "mad_wide", "mad_hi"} and all(v > 1e12 for v in ir["rates_per_s"].values())
    assert 0 < ir["frac"] <= 1 and 0 < ir["fwd_cols"]["frac"] <= 1
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == d["unit"] and c["sample"]
    assert d["clocks"]["sm_mhz"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # the exact (reference-sequence) mode and the key-switch microseconds ride in the same line
    assert 0 < d["exact"]["value"] < d["value"] and d["exact"]["unit"] == d["unit"]
    assert set(d["key_switch_us"]["engine_hybrid"]) == {"31", "17", "3"} == set(d["key_switch_us"]["reference_per_thread"])
    # evaluation keys came from a plan and hold no reference to the secret key
    assert d["keys"]["secret_key_detached_from_evaluation_keys"] is True and d["keys"]["generated_during_evaluation"] == 0


def test_bench_command_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], stdout=subprocess.PIPE, text=True, check=True).stdout
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--in-flight", "--no-hybrid", "--compress-keys", "--layers"):
        assert flag in out, flag
