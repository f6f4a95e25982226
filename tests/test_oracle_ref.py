"""Pins the oracle's COMPOSITION with the reference's own code (VERDICT r01 "Next" #3).

oracle/_ref/cpu/ref_driver is the reference's unmodified header code -- nn::llama3<bf16> (nn/llama.h:113-134), nn::attention
(nn/attention.h:161-206), nn::sink_cache, nn::rope, nn::feed_forward, quantization::lora_linear / lora_embedding / linear swapped in
by the reference's own huggingface::llama3_qlora_safetensor_serializer::adapt, make_causal_mask, make_default_sampler -- compiled
where it lies against the façade, with the oracle's op functions as the kernels (oracle/ref/orc_mc_abi.cc).  Its logits must be
BIT-IDENTICAL to oracle/orc_model.h on the same synthetic weights and ids: prompts with the causal mask, chunked prompts (quirk Q9:
the cached prefix stays masked), decode steps, decode past max_seq_len (sink-cache roll, quirk Q14, RoPE at the absolute position),
head_dim 64 and 128 (quirk Q4), the QLoRA rounding order (Q7/Q8); the reference's
default sampler chain (top-k -> nucleus -> multinomial, quirk Q10) must pick the token the oracle's sample_default picks.

The op-level half of the pin is tests/test_ref_suite_cpu.py (the reference's own unit tests on the same backend)."""
import numpy as np
import pytest

from oracle import orc
from tests import ref_driver

pytestmark = pytest.mark.skipif(not ref_driver.available("cpu"), reason="oracle/_ref/cpu/ref_driver not built (needs /root/reference; run __graft_entry__.build())")


@pytest.mark.parametrize("kind,shape,n_prompt,n_decode,chunk", ref_driver.CASES + ref_driver.SINK_CASES)
def test_reference_composition_is_bit_identical_to_the_oracle(kind, shape, n_prompt, n_decode, chunk):
    got = ref_driver.run("cpu", kind, shape, n_prompt, n_decode, chunk)
    want = ref_driver.oracle_rows(kind, shape, n_prompt, n_decode, chunk)
    assert len(want) == len(got["logits"])
    for i, row in enumerate(want):
        assert np.array_equal(row, got["logits"][i]), f"forward {i}: the reference's composition and orc_model.h differ"
        assert int(got["greedy"][i]) == orc.argmax(orc.BF16, row)
        # the default sampler with sample_size 1 is deterministic top-1 (quirk Q10): whatever the seeds, it equals sample_default
        assert int(got["sampled"][i]) == orc.sample_default(orc.BF16, row, u=0.37)["token"]


def test_gemma3_runs_through_the_reference_code():
    # no oracle restatement of Gemma-3 exists: this run IS the reference (kind "reference"); the GPU test compares against it
    for case in ref_driver.GEMMA_CASES:
        got = ref_driver.run("cpu", *case)
        lf = orc.bf16_to_f32(got["logits"])
        assert np.isfinite(lf).all() and float(np.abs(lf).max()) > 0
        assert len(set(got["greedy"].tolist())) >= 1
