"""bench.py's contract on a box without a GPU: the reference arm prints ONE JSON line with the agreed keys on a reduced workload,
the CUDA arm refuses to run (the product has no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--width", "160", "--height", "90", "--scale", "0.05", "--texsize", "64"]


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, cwd=ROOT, env=e, timeout=600)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "1"] + SMALL)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mpath-segments/s" and d["unit"] == "Msegments/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("C2 Atrium 160x90 batch 16 depth 9")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == os.cpu_count() and cb["value"] == d["value"] and "spp per step" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Msegments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_no_work():
    """under torchrun (N > 1) rank 0 alone runs the reference arm; the others exit 0 without output"""
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"] + SMALL, env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_cuda_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    r = _run(["--steps", "1", "--warmup", "3"] + SMALL)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
