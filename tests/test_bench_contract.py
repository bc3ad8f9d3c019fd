"""bench.py contract, the part that runs without a GPU: the reference arm (`--impl reference`) times the compiled
reference on host cores and prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

from helpers import ROOT, have_ref


@pytest.mark.skipif(not have_ref(), reason="compiled reference (oracle/_ref) not present")
def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-sample-inds", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "coancestry_chunk_pair_updates_per_sec"
    assert d["unit"] == "pair-updates/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_our_arm_refuses_to_run_without_a_gpu():
    """no CPU fallback: without a CUDA device bench.py's own arm exits with an error instead of measuring anything"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)
