"""bench.py's contract on a box without a GPU: the reference arm (the CPU restatement) prints one JSON
line with the keys the driver reads, and the product arm refuses to run rather than fall back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), cwd=ROOT, env=e,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)


def test_reference_arm_line():
    r = _run("--impl", "reference", "--shape", "small", "--ref-lattices", "24", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [x for x in r.stdout.splitlines() if x.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "arcs/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("lattice arcs/sec") and d["steps"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "arcs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None and d["config"]["tool"] == "frame_post"


def test_reference_arm_other_ranks_stay_silent():
    r = _run("--impl", "reference", "--shape", "small", "--ref-lattices", "8", "--steps", "1", "--warmup", "0",
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run("--shape", "tiny", "--lattices", "4", "--steps", "1", "--no-tools", "--no-cpu-baseline", "--e2e-steps", "0",
             env={"CUDA_VISIBLE_DEVICES": ""})
    assert r.returncode != 0  # klu_create fails: no CPU fallback
    assert not [x for x in r.stdout.splitlines() if x.startswith("{")]
