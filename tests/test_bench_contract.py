"""bench.py's driver contract, as far as it can be checked without a GPU: the reference arm (the unmodified reference, staged
by oracle/build_ref.py, timed on the host cores) prints ONE JSON line with the agreed keys, and the product arm refuses to run without CUDA instead
of falling back to anything on the CPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=env, timeout=timeout,
                          capture_output=True, text=True)


def test_reference_arm_prints_the_contract_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["config"]["workload"].startswith("c4")          # the configuration the north star's target is quoted on
    cb = d["cpu_baseline"]
    staged = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "lnasr_reference.zip"))
    assert cb["kind"] == ("reference" if staged else "port")  # the unmodified reference whenever its archive has been staged
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a box WITHOUT a GPU")
def test_product_arm_fails_loudly_without_cuda():
    p = _run("--steps", "1", "--warmup", "0", timeout=300)
    assert p.returncode != 0
    assert not any(ln.startswith("{") for ln in p.stdout.splitlines())      # no number from a CPU fallback
