"""bench.py's contract, as far as a machine without a GPU can check it: the reference arm prints ONE JSON line
with the agreed keys on stdout (and nothing else), and the product arm refuses to run without CUDA instead of
falling back to anything."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--env-n", "20000", "--obj-n", "2000", "--objects", "2", "--width", "320", "--height", "240", "--views", "2"]


def run(args, **kw):
    env = dict(os.environ)
    env.pop("RANK", None); env.pop("WORLD_SIZE", None)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          cwd=ROOT, env=env, timeout=600, **kw)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    res = run(["--impl", "reference", "--steps", "1", "--warmup", "1"] + SMALL)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("frames/s") and d["unit"] == "frames/s"
    assert d["value"] > 0 and abs(d["ms_per_step"] - 1e3 / d["value"]) < 1e-6 * d["ms_per_step"]
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "passes" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_runs_on_rank_0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"] + SMALL, capture_output=True, text=True, cwd=ROOT, env=env,
                         timeout=300)
    assert res.returncode == 0 and res.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine without CUDA")
def test_product_arm_fails_loudly_without_a_gpu():
    res = run(["--steps", "1", "--warmup", "0"] + SMALL)
    assert res.returncode != 0
    assert "no CPU path" in res.stderr and res.stdout.strip() == ""
