"""CPU tests of bench.py's command-line contract: the reference arm prints one JSON line with the keys the
driver reads; the GPU arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _have_cv2():
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


@pytest.mark.skipif(not _have_cv2(), reason="cv2 not importable")
def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-800:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "front-end frames/sec" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "frame pairs" in cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                         timeout=600, cwd=str(ROOT))
    assert out.returncode != 0
    assert "no CUDA device" in (out.stderr + out.stdout)
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")]      # no bench line from a CPU fallback
