"""bench.py's output contract, checked without a GPU: the reference arm (`--impl reference`, the CPU oracle port) runs here
and prints one well-formed JSON line; the candidate-arm lines committed under profiles/ carry every key the driver reads
and are internally consistent."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def test_reference_arm_runs_on_the_host_cpu():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["unit"] == "clips/s" and d["higher_is_better"] is True and d["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"]
    assert abs(cb["value"] - d["value"]) <= 1e-9 * max(1.0, d["value"])
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0 and e["unit"] == d["unit"]
    assert abs(e["value"] - d["value"]) <= 1e-9 * max(1.0, d["value"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "profiles", "r2_bench_default_final*.json"))))
def test_committed_candidate_lines_are_well_formed(path):
    d = json.loads(open(path).read().strip().splitlines()[-1])
    assert BASE_KEYS <= set(d) and d.get("impl", "candidate") != "reference"
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["dtype"] == "bf16" and d["scaling"] == "weak"
    assert "workload" in d["config"]
    # value = clips per step / step time
    B = d["config"].get("batch_per_gpu") or d["config"].get("global_batch") or 64
    assert abs(d["value"] - B * d["n_gpus"] / (d["ms_per_step"] * 1e-3)) < 1e-3 * d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6 and 0 < r["frac"] < 1
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.02
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]))
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["value"] > 0 and cb["cores"] >= 1
