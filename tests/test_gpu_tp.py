"""Tensor-parallel decode across GPUs of one box (needs >= 2 GPUs: run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_tp.py -m gpu`)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def n_gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("mode", ["small", "full"])
def test_tp_matches_single_gpu_and_golden(mode):
    n = n_gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4 if n < 8 else 8
    if mode == "small":
        world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", "29533",
           str(ROOT / "tests" / "tp_worker.py")] + (["small"] if mode == "small" else [])
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert res.returncode == 0, (res.stdout[-2000:], res.stderr[-3000:])
    out = json.loads(lines[-1])
    assert out["tokens_equal_single"] and out["all_ranks_agree"] and out["logits_max_rel_vs_single"] < 1e-2
    if mode == "full":
        assert out["tokens_equal_golden"]
