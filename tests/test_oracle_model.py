"""Cross-checks the composition part of the oracle (oracle/orc_model.h), which the
reference's tests do not pin, against an independent fp32 PyTorch Llama forward
(HF "rotate-half" RoPE, GQA, SwiGLU, tied head) on the oracle's own random weights.
"""
import numpy as np
import pytest
import torch

from oracle import orc
from oracle.orc import BF16, F32

CFG = dict(dim=256, n_layers=2, n_heads=8, n_kv_heads=2, head_dim=32, ffn_dim=512, vocab=1000, max_seq_len=64)


def torch_llama(model: orc.Llama, ids, quant=False):
    c = model.cfg
    D, H, KV, hd = c.dim, c.n_heads, c.n_kv_heads, c.head_dim

    def W(name, shape):
        return torch.from_numpy(model.tensor(name, np.float32).reshape(shape).copy()).double()

    def lin(prefix, N, K):
        if not quant:
            return W(prefix + ".weight", (N, K))
        q = torch.from_numpy(model.tensor(prefix + ".weight", np.int8).reshape(N, K).copy()).double()
        s = torch.from_numpy(model.tensor(prefix + ".scales", np.float32).reshape(N, K // 32).copy()).double()
        return q * s.repeat_interleave(32, dim=1)

    def apply(prefix, N, K, x):
        y = x @ lin(prefix, N, K).T
        if quant:
            A = W(prefix + ".adaptor.A.weight", (c.lora_rank, K))
            B = W(prefix + ".adaptor.B.weight", (N, c.lora_rank))
            y = y + c.lora_scale * ((x @ A.T) @ B.T)
        return y

    def rms(x, w):
        return w * x * torch.rsqrt((x * x).mean(-1, keepdim=True) + c.norm_eps)

    ids = torch.tensor(ids, dtype=torch.long)
    T = len(ids)
    if quant:
        tq = torch.from_numpy(model.tensor("tok_embeddings.weight", np.int8).reshape(c.vocab, D).copy()).double()
        ts = torch.from_numpy(model.tensor("tok_embeddings.scales", np.float32).copy()).double()
        emb = tq * ts[:, None]
        oq = torch.from_numpy(model.tensor("output.weight", np.int8).reshape(c.vocab, D).copy()).double()
        osc = torch.from_numpy(model.tensor("output.scales", np.float32).copy()).double()
        head = oq * osc[:, None]
    else:
        emb = W("tok_embeddings.weight", (c.vocab, D))
        head = emb
    x = emb[ids]
    pos = torch.arange(T).double()
    inv = 1.0 / (c.rope_theta ** (torch.arange(0, hd, 2).double() / hd))
    ang = pos[:, None] * inv[None, :]
    cos, sin = torch.cos(ang), torch.sin(ang)

    def rope(t):  # [T, heads, hd]
        a, b = t[..., : hd // 2], t[..., hd // 2 :]
        return torch.cat([a * cos[:, None, :] - b * sin[:, None, :], a * sin[:, None, :] + b * cos[:, None, :]], -1)

    for i in range(c.n_layers):
        p = f"layers.{i}."
        n = rms(x, W(p + "attention_norm.weight", (D,)))
        q = rope(apply(p + "attention.wq", H * hd, D, n).view(T, H, hd))
        k = rope(apply(p + "attention.wk", KV * hd, D, n).view(T, KV, hd))
        v = apply(p + "attention.wv", KV * hd, D, n).view(T, KV, hd)
        k = k.repeat_interleave(H // KV, dim=1)
        v = v.repeat_interleave(H // KV, dim=1)
        s = torch.einsum("thd,shd->hts", q, k) / np.sqrt(hd)
        s = s + torch.triu(torch.full((T, T), float("-inf")), 1)
        o = torch.einsum("hts,shd->thd", torch.softmax(s, -1), v).reshape(T, H * hd)
        h = x + apply(p + "attention.wo", D, H * hd, o)
        m = rms(h, W(p + "ffn_norm.weight", (D,)))
        g = apply(p + "feed_forward.w1", c.ffn_dim, D, m)
        u = apply(p + "feed_forward.w3", c.ffn_dim, D, m)
        x = h + apply(p + "feed_forward.w2", D, c.ffn_dim, torch.nn.functional.silu(g) * u)
    out = rms(x, W("norm.weight", (D,)))
    return (out[-1] @ head.T).numpy(), x.numpy()


@pytest.mark.parametrize("quant", [0, 1])
def test_fp32_oracle_matches_torch(quant):
    m = orc.Llama(orc.make_cfg(**CFG, quant=quant), F32)
    m.init_random(0x5EED)
    ids = [3, 77, 512, 999, 0, 41, 41, 7]
    logits, hidden = m.forward(ids, 0, want_hidden=True)
    ref_logits, ref_hidden = torch_llama(m, ids, quant=bool(quant))
    assert np.allclose(hidden, ref_hidden, rtol=2e-4, atol=2e-5)
    assert np.allclose(logits, ref_logits, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("quant", [0, 1])
def test_prefill_equals_incremental_decode(quant):
    # KV cache semantics (nn/cache.h:207-214): prefill(0..T) then decode must see the same
    # history as token-by-token decode; bf16 rounding points are identical, so the
    # results agree bit for bit except for the softmax reduction partition (same here).
    cfg = orc.make_cfg(**CFG, quant=quant)
    a, b = orc.Llama(cfg, BF16), orc.Llama(cfg, BF16)
    a.init_random(1)
    b.init_random(1)
    ids = [5, 9, 200, 31, 8, 640]
    la = a.forward(ids, 0)
    for t, tok in enumerate(ids):
        lb = b.forward([tok], t)
    assert np.array_equal(a.cache(0, 1, 0), b.cache(0, 1, 0))
    assert np.array_equal(la, lb)


def test_bf16_oracle_close_to_fp32_oracle():
    # north_star tolerance: bf16 path within 1e-2 max relative error per layer of the fp32 oracle
    cfg = orc.make_cfg(**{**CFG, "n_layers": 1})
    a, b = orc.Llama(cfg, BF16), orc.Llama(cfg, F32)
    a.init_random(2)
    b.init_random(2)
    ids = [1, 2, 3, 4]
    _, ha = a.forward(ids, 0, want_hidden=True)
    _, hb = b.forward(ids, 0, want_hidden=True)
    ha = orc.bf16_to_f32(ha)
    rel = np.abs(ha - hb).max() / np.abs(hb).max()
    assert rel < 1e-2, rel


def test_generator_is_counter_based():
    # weights are a pure function of (seed, tensor id, index): two models agree, a different
    # seed differs, and the documented formula reproduces an element.
    cfg = orc.make_cfg(**CFG)
    a, b, c = orc.Llama(cfg, BF16), orc.Llama(cfg, BF16), orc.Llama(cfg, BF16)
    a.init_random(7)
    b.init_random(7)
    c.init_random(8)
    name = "layers.1.attention.wk.weight"
    wa, wb, wc = (m.tensor(name, np.uint16) for m in (a, b, c))
    assert np.array_equal(wa, wb) and not np.array_equal(wa, wc)
    tid = (1 + 1) * 256 + 3  # tid_layer(layer=1, K_WK)
    u = orc.lib().orc_hash_uniform(7, tid, 12345)
    expect = orc.f32_to_bf16(np.float32(u) * (np.float32(1.0) / np.sqrt(np.float32(cfg.dim))))
    assert wa[12345] == expect


def test_default_sampler_chain():
    # nn/sampling.h:183-200,244-264,289-297: top-k(50) -> nucleus(0.6, 0.9) -> multinomial(1)
    rng = np.random.default_rng(3)
    logits = orc.f32_to_bf16(rng.standard_normal(5000).astype(np.float32) * 3)
    r = orc.sample_default(BF16, logits, u=0.7, intended=0)
    lf = orc.bf16_to_f32(logits)
    order = np.lexsort((np.arange(5000), -lf))[:50]
    assert np.array_equal(r["topk_idx"], order)
    assert r["choice"] == 0 and r["token"] == r["probs_idx"][0]  # quirk Q10
    assert r["token"] == orc.argmax(BF16, logits)
    p = r["probs_sorted"]
    nz = p[p > 0]
    assert np.all(nz[:-1] >= nz[1:]) and np.all(p[len(nz):] == 0)
    r2 = orc.sample_default(BF16, logits, u=0.95, intended=1)
    assert 0 <= r2["choice"] < 50 and r2["token"] == r2["probs_idx"][r2["choice"]]
