"""Full-size parity (Llama-3.2-1B shape, 16 layers, vocabulary 128 256) against the committed oracle fixtures.

BASELINE.json configs[1] (bf16) and configs[2] (QLoRA int4), plus an untied-head bf16 variant whose free-running greedy
continuation visits 64 distinct tokens.  See tests/golden/make_golden.py for what a fixture holds and why the old
"64 identical greedy tokens" check was not discriminating (a tied random-init head decodes into a fixed point).

Bars, at the prompt and at decode steps 0 / 15 / 31 / 63 (every 16th logit of the row + the oracle's eight best):
  * against the FP32 oracle (no bf16 rounding anywhere): the engine is no further from it than the bf16 ORACLE is -- both are bf16
    chains of 16 layers, a rounding-noise distance away from the fp32 truth: mean distance <= 1.25 x the bf16 oracle's + 1e-3
    (measured: 0.93-1.02 x), max distance (a single logit) <= 2 x + 2e-3;
  * against the bf16 oracle: max |engine - oracle| / max |oracle| <= 3e-2 and mean |engine - oracle| / mean |oracle| <= 2.5e-2;
  * argmax identical at EVERY step whose oracle decision is not a near-tie (top-2 gap >= 4 bf16 ulps -- the two chains are ~1 ulp
    apart on average, 2-3 ulps at worst; below that the engine must pick a token the oracle scores within 4 ulps of its best);
    free-running greedy tokens identical to the oracle's up to the first near-tie step.

Why not 1e-2 against the bf16 oracle: north_star's figure is PER LAYER against the fp32 oracle (tests/test_gpu_engine.py config 1
keeps it).  Two bf16 chains of 16 layers round every intermediate (1 ulp = 0.4-0.8 % of a value); fp32 re-association flips a few
of those roundings and a flip propagates through the residual stream, so after 16 blocks the logits of the two chains are ~1 ulp
apart on average (measured: mean 0.4-1.4e-2, max 0.5-1.4e-2 of the row maximum) -- the same distance the bf16 oracle itself keeps
from the fp32 oracle.  A wrong kernel moves the engine AWAY from the fp32 rows; rounding noise does not.
"""
import base64
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import orc
from tests.gpu_util import accelerator, unbf

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"
SHAPE_1B = dict(dim=2048, n_layers=16, n_heads=32, n_kv_heads=8, head_dim=64, ffn_dim=8192, vocab=128256, max_seq_len=1024)
NEAR_TIE_ULPS = 4.0
LOGITS_TOL = 3e-2        # max |engine - oracle| / max |oracle| over the compared logits of a 16-layer bf16 chain (see above)
LOGITS_MEAN_TOL = 2.5e-2  # mean |engine - oracle| / mean |oracle|
F32_SLACK = 1.25          # engine-to-fp32 distance <= F32_SLACK x (bf16 oracle-to-fp32 distance) + 1e-3


def hash_ids(n, vocab, tid):
    return [int(orc.lib().orc_hash_int(0x5EED, tid, i, 0, vocab)) for i in range(n)]


def load(variant):
    path = GOLDEN / f"llama1b_L16_{variant}_p512_s64.json"
    assert path.exists(), f"{path.name} is missing: run tests/golden/make_golden.py --variant {variant}"
    return json.loads(path.read_text())


def make_engine(variant, embed_mult=1.0):
    from metalchat_b200 import capi

    gpu = accelerator()
    quant = 1 if variant == "w4" else 0
    m = capi.Llama(gpu.dev, capi.llama_config(**SHAPE_1B, quant=quant, flags=capi.LLAMA_W4_PACKED if quant else 0))
    m.init_random(0x5EED)
    if variant == "bf16-untied":
        # the fixture's model: own output matrix (generator id G_OUT = 2, U/sqrt(dim)) and the embedding table times config.embed_mult, both
        # produced by the oracle's generator and uploaded by name (set_tensor("output.weight") unties the head)
        o = orc.Llama(orc.make_cfg(**{**SHAPE_1B, "n_layers": 0}, flags=orc.UNTIED_HEAD), orc.BF16)
        o.init_random(0x5EED)
        tok = o.tensor("tok_embeddings.weight", np.uint16)
        m.set_tensor("tok_embeddings.weight", orc.f32_to_bf16(orc.bf16_to_f32(tok) * float(embed_mult)))
        m.set_tensor("output.weight", o.tensor("output.weight", np.uint16))
        o.close()
    m.finalize()
    return m


def check_row(logits_bits, cp, what, f32_b64=None):
    """Engine logits row vs one oracle checkpoint: strided subsample + the oracle's eight best (+ the fp32 oracle's subsample)."""
    want_sub = unbf(np.frombuffer(base64.b64decode(cp["every16_b64"]), dtype=np.uint16))
    got_sub = unbf(logits_bits[::16])
    if f32_b64 is not None:
        ref = np.frombuffer(base64.b64decode(f32_b64), dtype=np.float32)
        rs, rm = float(np.abs(ref).max()), float(np.abs(ref).mean())
        e_max, e_mean = float(np.abs(got_sub - ref).max() / rs), float(np.abs(got_sub - ref).mean() / rm)
        o_max, o_mean = float(np.abs(want_sub - ref).max() / rs), float(np.abs(want_sub - ref).mean() / rm)
        print(f"  {what}: distance to the fp32 oracle: engine max {e_max:.2e} mean {e_mean:.2e} | bf16 oracle max {o_max:.2e} mean {o_mean:.2e}")
        assert e_mean <= F32_SLACK * o_mean + 1e-3 and e_max <= 2.0 * o_max + 2e-3, (what, e_max, o_max, e_mean, o_mean)
    scale = float(np.abs(want_sub).max())
    rel = float(np.abs(got_sub - want_sub).max() / scale)
    mean = float(np.abs(got_sub - want_sub).mean() / np.abs(want_sub).mean())
    top_want = unbf(np.array(cp["top8_bits"], np.uint16))
    top_got = unbf(logits_bits[np.array(cp["top8_ids"])])
    rel_top = float(np.abs(top_got - top_want).max() / scale)
    same = float((logits_bits[::16] == np.frombuffer(base64.b64decode(cp["every16_b64"]), dtype=np.uint16)).mean())
    print(f"  {what}: max-rel {rel:.2e} (every 16th) / {rel_top:.2e} (oracle top-8), mean-rel {mean:.2e}, bit-identical {same:.3f}")
    assert rel < LOGITS_TOL and rel_top < LOGITS_TOL and mean < LOGITS_MEAN_TOL, f"{what}: logits max-rel {rel:.3e} (every 16th), {rel_top:.3e} (oracle top-8), mean-rel {mean:.3e}"
    return max(rel, rel_top)


def check_choice(got, seq, s, what):
    """Step s of a continuation: the engine's sampled id against the oracle's decision."""
    want, gap = seq["argmax_after"][s], seq["top2_gap_ulps"][s]
    if gap >= NEAR_TIE_ULPS:
        assert got == want, f"{what}: argmax {got} != oracle {want} (oracle top-2 gap {gap} ulps)"
        return 1
    ids, vals = seq["top4_ids"][s], unbf(np.array(seq["top4_bits"][s], np.uint16))
    ulp = 2.0 ** (np.floor(np.log2(abs(float(vals[0])))) - 7)
    near = [i for i, v in zip(ids, vals) if float(vals[0]) - float(v) < NEAR_TIE_ULPS * ulp]
    assert got in near, f"{what}: {got} is not among the tokens the oracle scores within {NEAR_TIE_ULPS} ulps of its best {near} (gap {gap} ulps)"
    return int(got == want)


def argmax_low(logits_bits):
    lf = unbf(logits_bits)
    return int(np.lexsort((np.arange(len(lf)), -lf))[0])


def teacher_forced(m, g, seq, P, what, f32=None):
    """Feeds seq['inputs'] one token per step through the per-token call; every step's sampled id and the checkpoint logits
    are compared with the oracle."""
    exact, worst = 0, 0.0
    for s, tok in enumerate(seq["inputs"]):
        got = int(m.decode([tok], [P + s])[0])
        exact += check_choice(got, seq, s, f"{what} step {s}")
        cp = seq["checkpoints"].get(str(s))
        if cp is not None:
            logits = m.logits()
            assert argmax_low(logits) == got, "the sampled id is not the argmax of the stored logits row"
            worst = max(worst, check_row(logits, cp, f"{what} step {s}", f32.get(str(s)) if f32 else None))
    return exact, worst


@pytest.mark.parametrize("variant", ["bf16", "w4", "bf16-untied"])
def test_full_1b_logits_and_tokens_match_fixture(variant):
    g = load(variant)
    P, steps = g["prompt_len"], g["steps"]
    m = make_engine(variant, g["config"]["embed_mult"])
    m.prefill(hash_ids(P, SHAPE_1B["vocab"], 0xFFFF))
    first_logits = m.logits()
    worst = check_row(first_logits, g["prompt_checkpoint"], "prompt", g["f32_every16_b64"]["prompt"])
    first = argmax_low(first_logits)
    if g["prompt_gap_ulps"] >= NEAR_TIE_ULPS:
        assert first == g["prompt_argmax"]

    # (1) teacher-forced hash continuation: 64 distinct inputs, every step judged on its own
    tf = g["teacher"]
    assert tf["inputs"] == hash_ids(steps, SHAPE_1B["vocab"], 0xFFFE) and tf["distinct_inputs"] >= 60
    exact, w = teacher_forced(m, g, tf, P, f"{variant} teacher-forced", g["f32_every16_b64"])
    worst = max(worst, w)
    n_clear = sum(gp >= NEAR_TIE_ULPS for gp in tf["top2_gap_ulps"])
    assert exact >= n_clear

    # (2) free-running greedy through the device-side loop (several steps per launch of the streaming kernel): identical to
    # the oracle's continuation up to its first near-tie decision
    gr = g["greedy"]
    want = gr["inputs"][1:] + [gr["argmax_after"][-1]]  # token produced by step s
    toks, _ = m.decode_loop([gr["inputs"][0]], [P], steps)
    got = toks[:, 0].tolist()
    fragile = [s for s in range(steps) if gr["top2_gap_ulps"][s] < NEAR_TIE_ULPS]
    horizon = fragile[0] if fragile else steps
    assert got[:horizon] == want[:horizon], (variant, horizon, got[:horizon], want[:horizon])
    if fragile and horizon < steps:
        check_choice(got[horizon], gr, horizon, f"{variant} free-running step {horizon}")
    # (3) the same continuation teacher-forced along the ORACLE's path, so that steps after a near-tie are still checked
    exact_g, w = teacher_forced(m, g, gr, P, f"{variant} greedy path")
    worst = max(worst, w)
    print(f"{variant}: worst logits max-rel {worst:.3e}; teacher-forced argmax equal {exact}/{steps} ({n_clear} clear); greedy path: "
          f"{gr['distinct_inputs']} distinct tokens, free-run identical for {horizon} steps, teacher-forced equal {exact_g}/{steps}")
    if variant == "bf16-untied":
        assert gr["distinct_inputs"] >= 32
