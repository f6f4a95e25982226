"""Shared helper of the reference-driver tests: runs oracle/_ref/{cpu,cuda}/ref_driver (oracle/ref/ref_driver.cc: the reference's own
nn::llama3 / QLoRA / nn::gemma3 / default sampler on the synthetic hash weights, through the façade) and parses its output."""
import subprocess
import tempfile
from pathlib import Path

import numpy as np

from oracle import orc

ROOT = Path(__file__).resolve().parent.parent
REF_BIN = ROOT / "oracle" / "_ref"
SHAPES = {"small": dict(dim=512, n_layers=3, n_heads=8, n_kv_heads=2, head_dim=64, ffn_dim=1024, vocab=2000, max_seq_len=96),
          "hd128": dict(dim=512, n_layers=2, n_heads=6, n_kv_heads=2, head_dim=128, ffn_dim=768, vocab=1500, max_seq_len=160)}
# (kind, shape, n_prompt, n_decode, chunk): chunk > 0 feeds the prompt in two calls (quirk Q9: the second call's mask hides the prefix)
CASES = [("llama", "small", 10, 6, 0), ("llama", "hd128", 20, 4, 0), ("llama", "small", 24, 3, 9), ("qlora", "small", 10, 5, 0),
         ("qlora", "hd128", 12, 3, 5)]
GEMMA_CASES = [("gemma", "small", 20, 5, 0), ("gemma", "hd128", 40, 4, 0), ("gemma", "small", 30, 3, 12)]
# decode past max_seq_len (96): nn::sink_cache keeps log2(max_seq_len) sink rows, rolls the rest left and writes at the end
# (nn/cache.h:183-204, kernel/roll.metal); RoPE keeps the absolute position (nn/embedding.h:193-198)
SINK_CASES = [("llama", "small", 90, 14, 0), ("qlora", "small", 94, 6, 0)]


def available(backend: str) -> bool:
    return (REF_BIN / backend / "ref_driver").exists()


def run(backend: str, kind: str, shape: str, n_prompt: int, n_decode: int, chunk: int = 0):
    with tempfile.TemporaryDirectory() as td:
        out = Path(td) / "out.bin"
        cmd = [str(REF_BIN / backend / "ref_driver"), kind, shape, str(out), str(n_prompt), str(n_decode)] + ([str(chunk)] if chunk else [])
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
        assert res.returncode == 0, (cmd, res.stdout[-1000:], res.stderr[-3000:])
        b = out.read_bytes()
    assert b[:4] == b"MCRF"
    vocab, n = (int(x) for x in np.frombuffer(b[4:12], np.uint32))
    off = 12
    logits = np.frombuffer(b[off:off + 2 * vocab * n], np.uint16).reshape(n, vocab)
    off += 2 * vocab * n
    sampled = np.frombuffer(b[off:off + 4 * n], np.int32)
    greedy = np.frombuffer(b[off + 4 * n:off + 8 * n], np.int32)
    return dict(logits=logits, sampled=sampled, greedy=greedy, stderr=res.stderr)


def oracle_rows(kind: str, shape: str, n_prompt: int, n_decode: int, chunk: int = 0, dtype=orc.BF16):
    """The same schedule through oracle/orc_model.h: prompt (one or two calls), then greedy decode steps."""
    cfg = SHAPES[shape]
    o = orc.Llama(orc.make_cfg(**cfg, quant=1 if kind == "qlora" else 0), dtype)
    o.init_random(0x5EED)
    ids = [int(orc.lib().orc_hash_int(0x5EED, 0xFFFF, i, 0, cfg["vocab"])) for i in range(n_prompt)]
    rows = [o.forward(ids[:chunk], 0), o.forward(ids[chunk:], chunk)] if chunk else [o.forward(ids, 0)]
    tok = orc.argmax(dtype, rows[-1])
    for s in range(n_decode):
        rows.append(o.forward([tok], n_prompt + s))
        tok = orc.argmax(dtype, rows[-1])
    o.close()
    return rows
