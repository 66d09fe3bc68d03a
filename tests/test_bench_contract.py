"""The one-line JSON contract of bench.py, checked where it can be without a GPU: the reference arm
(`--impl reference`, the reference's own CPU path from oracle/_ref or the port) is RUN here on a tiny sample, and the
last line the B200 arm printed on a GPU box (profiles/) is checked for the same keys."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def check_common(d):
    assert BASE_KEYS <= set(d), sorted(BASE_KEYS - set(d))
    assert d["metric"].startswith("particle-updates/s") and d["unit"] == "particle-updates/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 0 and d["ms_per_step"] > 0
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)


def test_reference_arm_prints_one_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--small"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-500:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    check_common(d)
    assert d["impl"] == "reference" and d["steps"] == 1 and d["warmup"] == 1
    cb = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(cb) and cb["kind"] in ("reference", "port")
    assert cb["value"] == d["value"] and cb["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_recorded_b200_line_has_the_contract_keys():
    path = os.path.join(ROOT, "profiles", "r02l_bench.json")
    if not os.path.exists(path):
        pytest.skip("no recorded line")
    d = json.load(open(path))
    check_common(d)
    assert d["n_gpus"] == 1 and d["gpu_launches"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
