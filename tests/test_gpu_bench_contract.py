"""bench.py prints ONE JSON line with the keys the driver reads, for both arms."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _run(*args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True,
                       timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    return json.loads(lines[0])


@pytest.mark.timeout(900)
def test_our_arm_prints_the_contract_line(cuda_device):
    d = _run("--steps", "4", "--warmup", "3", "--no-cpu-baseline")
    assert BASE_KEYS | {"roofline", "gpu_launches", "clocks"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 4 and d["warmup"] == 3 and d["higher_is_better"] is True
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 0 and d["vs_baseline"] is None
    assert d["gpu_launches"] == d["steps"]                          # the step is one launch: counted inside the library
    assert d["timed_steps"] >= 100 and abs(d["value"] - 128 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    ok, tot = (int(v) for v in d["parity"]["matched_anchors_identical"].split("/"))
    assert tot == 6747 and ok >= tot - 4 and d["parity"]["loss_rel_err_max"] <= (1e-5 if ok == tot else 2e-4)
    assert d["api_device_resident"]["packed_gt"]["value"] <= d["value"] * 1.05
    assert d["adverse_logits"]["ms_per_step"] > 0 and d["tal"]["ms_per_step"] > 0 and d["cfg5_bf16"]["ms_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.3 < r["frac"] < 1.0 and r["traffic"] > 0.9 * r["algorithmic_bytes_per_launch"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 600e6 and e["d2h_bytes_per_step"] == 32 and 0 < e["value"] < d["value"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])


@pytest.mark.timeout(900)
def test_reference_arm_runs_the_cpu_restatement(cuda_device):
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert d["impl"] == "reference" and BASE_KEYS | {"cpu_baseline"} <= set(d)
    # the unmodified reference (baseline/_ref) when it travelled with the snapshot, else the oracle port
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert (d["cpu_baseline"]["kind"] == "reference") == os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "src", "model", "losses.py"))
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
