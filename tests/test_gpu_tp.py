"""Tensor-parallel decode across GPUs of one box (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_tp.py -m gpu`).
The reference is single-device (nn/llama.h:86); sharding follows BASELINE.json north_star (column/row-split blocks, one all-reduce after
wo and after w2, vocabulary-split head).  Both exchange implementations are covered: the streaming persistent kernel (all-reduce fused
into the wo / w2 epilogues: tagged fp32 words stored straight into the peers' memory) and the per-op kernels (MC_TP_NO_STREAM=1)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def n_gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def run_worker(mode, world, per_op, port, nccl=False):
    env = dict(os.environ)
    env.pop("MC_TP_NCCL", None)
    if nccl:
        env["MC_TP_NCCL"] = "1"
    if per_op:
        env["MC_TP_NO_STREAM"] = "1"
    else:
        env.pop("MC_TP_NO_STREAM", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           str(ROOT / "tests" / "tp_worker.py"), mode]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=1200, cwd=ROOT, env=env)
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    if lines:
        print(lines[-1])  # (the worker's report, also when its checks failed)
    assert res.returncode == 0 and lines, (res.stdout[-3000:], res.stderr[-4000:])
    out = json.loads(lines[-1])
    print(json.dumps(out))
    return out


@pytest.mark.parametrize("per_op", [False, True], ids=["streaming", "per_op"])
@pytest.mark.parametrize("mode", ["small", "hd128", "batch8", "batch12", "full"])
def test_tp_matches_single_gpu_and_oracle(mode, per_op):
    n = n_gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if (n < 4 or mode == "small") else 4 if n < 8 else 8  # small: 2 KV heads
    out = run_worker(mode, world, per_op, 29533 + (1 if per_op else 0))
    assert out["ok"], out
    if not per_op and mode not in ("batch8", "batch12"):
        assert out["launches_per_step"] == 1, "the streaming kernel did not take the tensor-parallel step"


def test_tp_nccl_comparator_decodes_the_same_tokens():
    """bench.py --tp-collective nccl: ncclAllReduce between the per-op kernels instead of the fused exchange (measurement comparator)."""
    n = n_gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    out = run_worker("small", 2, True, 29535, nccl=True)
    assert out["ok"], out
    assert out["launches_per_step"] > 1


@pytest.mark.parametrize("per_op", [False, True], ids=["streaming", "per_op"])
def test_tp_quantised_model_matches_single_gpu_and_oracle(per_op):
    """BASELINE.json configs[4] in small: the QLoRA layout sharded over the GPUs (column-split wq/wk/wv/w1/w3 with their scales and B rows,
    row-split wo/w2 with column-split scales and A; the all-reduce carries the main sums and the adaptor's A . x), through the streaming
    persistent kernel and through the per-op kernels."""
    n = n_gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    out = run_worker("quant", 2 if n < 4 else 4 if n < 8 else 8, per_op, 29536 + (1 if per_op else 0))
    assert out["ok"], out
    if not per_op:
        assert out["launches_per_step"] == 1, "the streaming kernel did not take the quantised tensor-parallel step"


def test_tp_quantised_batches_and_prompts_on_the_tensor_cores():
    """12 sequences of the sharded QLoRA model: prompts and decode steps run as tcgen05 GEMMs over the resident bf16 image; the row-parallel
    linears all-reduce [main sums | adaptor A . x] (tc::tp_allreduce_rows) before the adaptor epilogue."""
    n = n_gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    out = run_worker("quant12", 2 if n < 4 else 4 if n < 8 else 8, False, 29538)
    assert out["ok"], out
    assert out["launches_per_step"] > 1
