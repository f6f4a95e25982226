"""Generates the committed golden fixtures of the full-size configs with the CPU oracle.

    python tests/golden/make_golden.py [--quant 0|1] [--prompt 512] [--steps 64]

Full-size (Llama-3.2-1B-shaped) runs take minutes on host cores, so the GPU tests compare against
these fixtures instead of re-running the oracle.  The prompt is hash-generated (seed 0x5EED,
tensor id 0xFFFF); weights come from init_random(0x5EED).  Output: greedy token ids, plus the
top-2 logit gap (in bf16 ulps of the winner) at every step so that a near-tie can be recognised.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import orc  # noqa: E402


def prompt_ids(n, vocab, seed=0x5EED):
    return [int(orc.lib().orc_hash_int(seed, 0xFFFF, i, 0, vocab)) for i in range(n)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quant", type=int, default=0)
    ap.add_argument("--prompt", type=int, default=512)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--layers", type=int, default=16)
    ap.add_argument("--out", type=str, default=None)
    a = ap.parse_args()
    cfg = orc.make_cfg(n_layers=a.layers, quant=a.quant, max_seq_len=1024)
    m = orc.Llama(cfg, orc.BF16)
    m.init_random(0x5EED)
    ids = prompt_ids(a.prompt, cfg.vocab)
    t0 = time.time()
    logits = m.forward(ids, 0)
    t_prefill = time.time() - t0
    toks, gaps = [], []
    pos = a.prompt
    t0 = time.time()
    for _ in range(a.steps):
        lf = orc.bf16_to_f32(logits)
        tok = orc.argmax(orc.BF16, logits)
        top2 = np.partition(lf, -2)[-2:]
        ulp = 2.0 ** (np.floor(np.log2(abs(float(top2[1])))) - 7)
        gaps.append(float((top2[1] - top2[0]) / ulp))
        toks.append(int(tok))
        logits = m.forward([tok], pos)
        pos += 1
    t_dec = time.time() - t0
    out = dict(config=dict(shape="llama-3.2-1b", n_layers=a.layers, quant=a.quant, max_seq_len=1024, seed=0x5EED),
               prompt_len=a.prompt, steps=a.steps, tokens=toks, top2_gap_ulps=gaps,
               first_logits_head=[int(x) for x in m.forward([toks[-1]], pos)[:16]] if False else None,
               oracle_threads=orc.num_threads(), prefill_s=t_prefill, decode_s=t_dec)
    name = a.out or f"llama1b_L{a.layers}_q{a.quant}_p{a.prompt}_s{a.steps}.json"
    (Path(__file__).parent / name).write_text(json.dumps(out, indent=1))
    print(name, "prefill", round(t_prefill, 1), "s decode", round(t_dec, 1), "s; min gap", min(gaps))


if __name__ == "__main__":
    main()
