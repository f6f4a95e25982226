"""The reference's own layer code on the B200 (north_star: "a CUDA accelerator is a drop-in replacement for the Metal one").

oracle/_ref/cuda/ref_driver = the same program as tests/test_oracle_ref.py's, linked with libmc_cuda.so: nn::llama3<bf16>, the QLoRA
variant, nn::gemma3<bf16> (GELU-tanh, q/k RMSNorm, post-norms with mu = 1, sliding-window mask, sqrt(dim) embedding scale, two RoPE
bases: nn/gemma.h:43-147) and the default sampler run op by op (~30 kernels per block, like on Metal) through the façade and the
op-level sm_100a kernels.  Bars: logits within 1e-2 (max |a-b| / max |b|) of the CPU run of the same code over the oracle's kernels,
and the same greedy / sampled token wherever the CPU run's decision is not a near-tie."""
import numpy as np
import pytest

from oracle import orc
from tests import ref_driver
from tests.gpu_util import require_gpu

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (ref_driver.available("cuda") and ref_driver.available("cpu")), reason="oracle/_ref drivers were not shipped")]


def near_top(row_bits, token, ulps=2):
    lf = orc.bf16_to_f32(row_bits)
    top = float(lf.max())
    step = 2.0 ** (np.floor(np.log2(abs(top))) - 7) if top != 0 else 0.0
    return float(lf[token]) >= top - ulps * step


@pytest.mark.parametrize("kind,shape,n_prompt,n_decode,chunk", ref_driver.CASES + ref_driver.GEMMA_CASES + ref_driver.SINK_CASES)
def test_reference_layers_on_b200_match_the_cpu_run(kind, shape, n_prompt, n_decode, chunk):
    require_gpu()
    cpu = ref_driver.run("cpu", kind, shape, n_prompt, n_decode, chunk)
    gpu = ref_driver.run("cuda", kind, shape, n_prompt, n_decode, chunk)
    assert "B200" in gpu["stderr"] or "NVIDIA" in gpu["stderr"], gpu["stderr"][-500:]
    # the prompt rows see the same inputs on both sides; decode rows too as long as the greedy tokens agree
    n = len(cpu["logits"])
    assert len(gpu["logits"]) == n
    worst = 0.0
    for i in range(n):
        want, got = orc.bf16_to_f32(cpu["logits"][i]), orc.bf16_to_f32(gpu["logits"][i])
        rel = float(np.abs(got - want).max() / np.abs(want).max())
        worst = max(worst, rel)
        assert rel < 1e-2, (kind, shape, i, rel)
        assert near_top(cpu["logits"][i], int(gpu["greedy"][i])) and near_top(cpu["logits"][i], int(gpu["sampled"][i]))
        if int(gpu["greedy"][i]) != int(cpu["greedy"][i]):
            break  # a near-tie went the other way: the two runs decode different tokens from here on
    same = float(np.mean([np.mean(cpu["logits"][i] == gpu["logits"][i]) for i in range(min(n, 2))]))
    print(f"{kind} {shape}: worst max-rel {worst:.2e}, prompt rows bit-identical {same:.3f}")
