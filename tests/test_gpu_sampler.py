"""GPU parity of the on-device sampling tail (top-k -> nucleus -> multinomial, nn/sampling.h:152-316) against the
oracle's sample_default with injected uniform draws, both multinomial modes of quirk Q10."""
import numpy as np
import pytest

from oracle import orc
from oracle.orc import BF16
from tests.gpu_util import accelerator, bf, unbf

pytestmark = pytest.mark.gpu


def oracle_rows(logits, us, **kw):
    return [orc.sample_default(BF16, logits[i], u=float(us[i]), **kw) for i in range(len(logits))]


@pytest.mark.parametrize("vocab,top_k", [(5000, 50), (128256, 50), (128256, 40), (2000, 64), (64, 7), (100, 1)])
@pytest.mark.parametrize("intended", [0, 1])
def test_default_sampler_matches_oracle(rng, vocab, top_k, intended):
    from metalchat_b200 import capi

    gpu = accelerator()
    rows = 4
    logits = bf(rng.standard_normal((rows, vocab)).astype(np.float32) * 3)
    logits[0, : vocab // 3] = logits[0, 0]  # heavy ties: top-k order = lower index first, sort network fixes the rest
    us = rng.random(rows).astype(np.float32)
    got = capi.sample_default(gpu.dev, logits, us, top_k=top_k, temperature=0.6, top_p=0.9, intended=intended)
    want = oracle_rows(logits, us, topk=top_k, temperature=0.6, top_p=0.9, intended=intended)
    for i in range(rows):
        w = want[i]
        assert np.array_equal(got["topk_idx"][i], w["topk_idx"]), i                      # integer work: bit-exact
        gp, wp = unbf(got["probs_sorted"][i]), w["probs_sorted"]
        # probabilities: expf (CUDA) vs libm differs by <= 2 ulp in fp32 -> at most one bf16 step, usually none
        assert np.all(np.abs(got["probs_sorted"][i].astype(np.int32) - bf(wp).astype(np.int32)) <= 1), (i, gp, wp)
        if np.array_equal(got["probs_sorted"][i], bf(wp)):
            assert np.array_equal(got["probs_idx"][i], w["probs_idx"]), i               # same sort network, same tie order
            assert got["choice"][i] == w["choice"] and got["token"][i] == w["token"], i
        if intended == 0:
            assert got["choice"][i] == 0  # quirk Q10: the default sampler is deterministic top-1
            assert got["token"][i] == orc.argmax(BF16, logits[i]) or top_k > 1


def test_sampler_mask_and_order_properties(rng):
    from metalchat_b200 import capi

    gpu = accelerator()
    logits = bf(rng.standard_normal((8, 128256)).astype(np.float32) * 4)
    us = rng.random(8).astype(np.float32)
    got = capi.sample_default(gpu.dev, logits, us, top_k=50, temperature=0.6, top_p=0.9, intended=1)
    lf = unbf(logits)
    for i in range(8):
        order = np.lexsort((np.arange(128256), -lf[i]))[:50]
        assert np.array_equal(got["topk_idx"][i], order)
        p = unbf(got["probs_sorted"][i])
        nz = p[p > 0]
        assert np.all(nz[:-1] >= nz[1:]) and np.all(p[len(nz):] == 0)  # descending, masked tail
        assert set(got["probs_idx"][i].tolist()) == set(order.tolist())
        assert 0 <= got["choice"][i] < 50 and got["token"][i] == got["probs_idx"][i][got["choice"][i]]


def test_sampler_argument_validation(rng):
    from metalchat_b200 import capi

    gpu = accelerator()
    logits = bf(rng.standard_normal((1, 100)).astype(np.float32))
    with pytest.raises(capi.McInvalidArgument, match="top_k"):
        capi.sample_default(gpu.dev, logits, [0.5], top_k=65)
    with pytest.raises(capi.McInvalidArgument, match="temperature"):
        capi.sample_default(gpu.dev, logits, [0.5], top_k=5, temperature=0.0)


def test_engine_sampled_decode_matches_oracle_chain():
    # the engine's sampled step == the oracle's sampler chain applied to the engine's own logits with the same uniform
    # (the logits themselves are compared in test_gpu_engine.py); then the device-side feedback loop == per-call decode
    from metalchat_b200 import capi
    from tests.test_gpu_engine import SMALL, make_engine

    m = make_engine(SMALL)
    ids = [3, 77, 512, 999, 0, 41]
    m.prefill(ids)
    steps = 12
    us = np.random.default_rng(7).random(steps).astype(np.float32)
    sc = capi.SamplerConfig(1, 50, 0.6, 0.9, 1)
    first = int(np.argmax(unbf(m.logits())))
    tok, pos, calls = first, len(ids), []
    for s in range(steps):
        out = int(m.decode([tok], [pos], uniforms=us[s : s + 1], sampler=sc)[0])
        want = orc.sample_default(BF16, m.logits(), topk=50, temperature=0.6, top_p=0.9, u=float(us[s]), intended=1)
        assert out == want["token"], (s, out, want["token"], want["choice"])
        calls.append(out)
        tok, pos = out, pos + 1
    m2 = make_engine(SMALL)
    m2.prefill(ids)
    toks, _ = m2.decode_loop([first], [len(ids)], steps, uniforms=us, sampler=sc)
    assert toks[:, 0].tolist() == calls
    # the reference-exact multinomial (quirk Q10) makes the default sampler greedy
    m3 = make_engine(SMALL)
    m3.prefill(ids)
    g, _ = m3.decode_loop([first], [len(ids)], steps)
    m4 = make_engine(SMALL)
    m4.prefill(ids)
    q10, _ = m4.decode_loop([first], [len(ids)], steps, uniforms=us, sampler=capi.SamplerConfig(1, 50, 0.6, 0.9, 0))
    assert q10[:, 0].tolist() == g[:, 0].tolist()
