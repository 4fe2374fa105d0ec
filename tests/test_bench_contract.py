"""CPU tier: the parts of bench.py's contract that need no GPU -- the reference arm's JSON line (the oracle port timed
on the host cores: the one place besides tests/ and smoke() that may execute oracle/) and the product arm's refusal to
run without a CUDA device (there is no CPU path)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, env=e,
                          timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                   # ONE JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "body2hand_train_frames_per_sec" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["value"] > 0
    assert d["config"]["workload"].startswith("train step") and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_on_rank_zero_only():
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""      # the other ranks exit 0 without work


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU")
def test_product_arm_refuses_to_run_without_a_gpu():
    r = _run("--steps", "1", "--warmup", "1")
    assert r.returncode != 0
    assert "no CPU path" in (r.stderr + r.stdout)
