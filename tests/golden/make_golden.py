"""Generates the committed full-size golden fixtures with the CPU oracle (test infrastructure).

    python tests/golden/make_golden.py --variant bf16|w4|bf16-untied [--prompt 512] [--steps 64]

Full-size (Llama-3.2-1B-shaped, 16 layers, vocabulary 128 256) runs take minutes on host cores, so the GPU
tests compare against these fixtures instead of re-running the oracle.  Weights: init_random(0x5EED); the
prompt is hash-generated (seed 0x5EED, tensor id 0xFFFF).

A random-init model with a tied head decodes into a fixed point (the current token's own embedding wins the
argmax), so "64 identical greedy tokens" alone carries one token of information.  Every fixture therefore holds
TWO continuations of the same 512-token prompt, each with per-step evidence:

  * ``teacher``: 64 hash-generated input tokens (seed 0x5EED, tensor id 0xFFFE) fed one per step whatever the
    model predicts -- 64 distinct inputs, 64 independent logits rows;
  * ``greedy``: the oracle's own free-running greedy continuation (token t+1 = argmax of step t).

Per step: the oracle's four best tokens (ids + bf16 logit bits) and the gap between its two largest logits in bf16 ulps of the
winner.  After 16 blocks two bf16 chains are ~1 ulp apart on average (see tests/test_gpu_golden.py), so a decision with a gap
below 4 ulps is a near-tie: the test then requires the engine's choice to be a token the oracle scores within 4 ulps of its best.  At steps 0, 15, 31 and 63: the eight largest logits
(ids + bf16 bits), every 16th logit of the row (bf16 bits, base64) and the CRC-32 of the whole row; for the prompt and the teacher
path also every 16th logit of the FP32 oracle's row (``f32_every16_b64``): the yardstick of what bf16 rounding alone costs.

Variant ``bf16-untied``: the same model with its own output matrix (orc flag UNTIED_HEAD, generator id G_OUT)
and the embedding table scaled by 64 -- the residual stream is then dominated by the current token, the head
carries no self-token bias, and the free-running greedy continuation visits 64 distinct tokens.
"""
import argparse
import base64
import json
import sys
import time
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import orc  # noqa: E402

CHECKPOINTS = (0, 15, 31, 63)
EMBED_MULT = 64.0  # bf16-untied variant (a power of two: the scaling is exact in bf16)


def hash_ids(n, vocab, tid, seed=0x5EED):
    return [int(orc.lib().orc_hash_int(seed, tid, i, 0, vocab)) for i in range(n)]


def top2_gap_ulps(lf):
    top2 = np.partition(lf, -2)[-2:]
    ulp = 2.0 ** (np.floor(np.log2(abs(float(top2[1])))) - 7)
    return float((top2[1] - top2[0]) / ulp)


def checkpoint(logits_bits):
    lf = orc.bf16_to_f32(logits_bits)
    order = np.lexsort((np.arange(len(lf)), -lf))[:8]
    return dict(top8_ids=[int(i) for i in order], top8_bits=[int(logits_bits[i]) for i in order],
                every16_b64=base64.b64encode(np.ascontiguousarray(logits_bits[::16]).tobytes()).decode(),
                crc32=int(zlib.crc32(np.ascontiguousarray(logits_bits).tobytes())))


def build_model(variant, layers, dtype=orc.BF16):
    quant = 1 if variant == "w4" else 0
    flags = orc.UNTIED_HEAD if variant == "bf16-untied" else 0
    cfg = orc.make_cfg(n_layers=layers, quant=quant, max_seq_len=1024, flags=flags)
    m = orc.Llama(cfg, dtype)
    m.init_random(0x5EED)  # the fp32 model holds the same (bf16-representable) weight values
    if variant == "bf16-untied":
        if dtype == orc.BF16:
            w = m.tensor("tok_embeddings.weight", np.uint16)
            w[:] = orc.f32_to_bf16(orc.bf16_to_f32(w) * EMBED_MULT)  # exact: a power of two
        else:
            m.tensor("tok_embeddings.weight", np.float32)[:] *= EMBED_MULT
    return cfg, m


def f32_reference(variant, layers, prompt_ids, teacher_ids):
    """The fp32 oracle (no bf16 rounding anywhere) on the prompt and along the teacher-forced path: every 16th logit at the
    checkpoints.  Both the engine and the bf16 oracle are a rounding-noise distance away from these rows; the GPU test
    requires the engine to be no further from them than the bf16 oracle is."""
    cfg, m = build_model(variant, layers, orc.F32)
    enc = lambda row: base64.b64encode(np.ascontiguousarray(row[::16], dtype=np.float32).tobytes()).decode()
    out = {"prompt": enc(m.forward(prompt_ids, 0))}
    pos = len(prompt_ids)
    for s, tok in enumerate(teacher_ids):
        row = m.forward([tok], pos)
        pos += 1
        if s in CHECKPOINTS:
            out[str(s)] = enc(row)
    m.close()
    return out


def continuation(m, first_logits, start_pos, steps, teacher):
    """teacher = list of input tokens, or None for free-running greedy."""
    logits = first_logits
    pos = start_pos
    inputs, argmaxes, seconds, gaps, top4_ids, top4_bits, cps = [], [], [], [], [], [], {}
    for s in range(steps):
        am = orc.argmax(orc.BF16, logits)  # prediction made BEFORE step s consumes its input
        tok = teacher[s] if teacher is not None else am
        inputs.append(int(tok))
        logits = m.forward([tok], pos)
        pos += 1
        lf = orc.bf16_to_f32(logits)
        best = np.lexsort((np.arange(len(lf)), -lf))[:4]
        argmaxes.append(int(best[0]))
        seconds.append(int(best[1]))
        top4_ids.append([int(i) for i in best])
        top4_bits.append([int(logits[i]) for i in best])
        gaps.append(top2_gap_ulps(lf))
        if s in CHECKPOINTS:
            cps[str(s)] = checkpoint(logits)
    return dict(inputs=inputs, argmax_after=argmaxes, second_after=seconds, top2_gap_ulps=gaps, top4_ids=top4_ids, top4_bits=top4_bits, checkpoints=cps,
                distinct_inputs=len(set(inputs)), distinct_argmax=len(set(argmaxes)), min_gap_ulps=min(gaps))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="bf16", choices=["bf16", "w4", "bf16-untied"])
    ap.add_argument("--prompt", type=int, default=512)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--layers", type=int, default=16)
    ap.add_argument("--out", type=str, default=None)
    a = ap.parse_args()
    cfg, m = build_model(a.variant, a.layers)
    ids = hash_ids(a.prompt, cfg.vocab, 0xFFFF)
    t0 = time.time()
    first = m.forward(ids, 0)
    t_prefill = time.time() - t0
    t0 = time.time()
    teacher = continuation(m, first, a.prompt, a.steps, hash_ids(a.steps, cfg.vocab, 0xFFFE))
    greedy = continuation(m, first, a.prompt, a.steps, None)  # overwrites cache rows >= prompt as it goes
    t_dec = time.time() - t0
    m.close()
    t0 = time.time()
    f32 = f32_reference(a.variant, a.layers, ids, teacher["inputs"])
    t_f32 = time.time() - t0
    out = dict(config=dict(shape="llama-3.2-1b", n_layers=a.layers, variant=a.variant, quant=1 if a.variant == "w4" else 0,
                           untied_head=a.variant == "bf16-untied", embed_mult=EMBED_MULT if a.variant == "bf16-untied" else 1.0,
                           max_seq_len=1024, seed=0x5EED),
               prompt_len=a.prompt, steps=a.steps, prompt_checkpoint=checkpoint(first), prompt_argmax=int(orc.argmax(orc.BF16, first)),
               prompt_gap_ulps=top2_gap_ulps(orc.bf16_to_f32(first)), teacher=teacher, greedy=greedy, f32_every16_b64=f32,
               oracle_threads=orc.num_threads(), prefill_s=t_prefill, decode_s=t_dec, f32_s=t_f32)
    name = a.out or f"llama1b_L{a.layers}_{a.variant}_p{a.prompt}_s{a.steps}.json"
    (Path(__file__).parent / name).write_text(json.dumps(out, indent=0, separators=(",", ":")))
    print(name, "prefill", round(t_prefill, 1), "s decode", round(t_dec, 1), "s; teacher: distinct argmax", teacher["distinct_argmax"],
          "min gap", teacher["min_gap_ulps"], "| greedy: distinct", greedy["distinct_inputs"], "min gap", greedy["min_gap_ulps"])


if __name__ == "__main__":
    main()
