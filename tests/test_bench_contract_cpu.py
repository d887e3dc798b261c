"""CPU: bench.py's reference arm honours the driver's contract -- exactly one JSON line on stdout
with the agreed keys -- and the GPU arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--search-area", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mptc_encode_mpixel_per_s" and d["unit"] == "Mpixel/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_gpu_arm_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = run_bench("--steps", "1")
    assert r.returncode != 0
    assert r.stdout.strip() == ""
    assert "no CUDA device" in r.stderr


def test_full_fixture_comparison_reports_mismatches():
    """bench.py's parity record and tests/test_gpu_full_golden.py rest on compare_with_full_fixture:
    outputs that are not the reference's must be reported, frame by frame, with the differing block row."""
    import numpy as np
    from golden_util import compare_with_full_fixture, full_fixture_name, load
    name = full_fixture_name(1920, 1080, 16, 50, 15)
    assert name == "full_1080p60_sa16_gop15"
    assert full_fixture_name(1920, 1080, 16, 49, 15) is None
    g = load(name)
    assert g["hashes"].shape == (60, 4) and g["rows_final"].shape == (60, 270) and int(g["n_unique"].sum()) == 5837
    nb = 480 * 270
    zeros = {"blocks": np.zeros((2, nb), np.uint64), "motion": np.zeros((2, 2 * nb), np.uint8),
             "unique": np.zeros((2, nb), np.uint32), "n_unique": np.array([3, 0], np.uint32)}
    n, bad = compare_with_full_fixture(g, zeros["blocks"], zeros["motion"], zeros["unique"], zeros["n_unique"], first_frame=58)
    assert n == 2 and len(bad) == 6
    assert any("frame 58: final blocks" in b and "first differing block row 0" in b for b in bad)
    n, _ = compare_with_full_fixture(g, zeros["blocks"], zeros["motion"], zeros["unique"], zeros["n_unique"], first_frame=59)
    assert n == 1                      # frames beyond the fixture are not compared
