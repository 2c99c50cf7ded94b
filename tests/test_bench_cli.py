"""bench.py command line on a machine without a GPU: the reference arm (the torch-CPU port of the reference path)
prints the contract's JSON line; the B200 arm refuses to run without CUDA instead of falling back."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--batch", "2", "--seconds", "0.5")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "frames/sec" and line["unit"] == "frames/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    # "reference" when the vendored reference (baseline/_ref, written by __graft_entry__.build()) is present
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "audiozen"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and line["cpu_baseline"]["cores"] >= 1
    if have_ref:
        assert line["cpu_baseline"]["port_value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine without CUDA")
def test_b200_arm_refuses_to_run_without_cuda():
    r = _run("--steps", "1", "--warmup", "1")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
