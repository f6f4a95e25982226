"""Isolated parity of the attention kernels and of the int8 embedding gather against the oracle (through the C ABI).

Reference chain (nn/attention.h:195-203): scores = T(q . K^T) -> T(scores * T(1/sqrt(hd))) -> [T(scores + mask)] -> softmax without
max shift (kernel/softmax.metal:40-80) -> T(P . V); oracle: orc_model.h sdpa() / causal_mask().  The engine's kernels fuse the
chain and re-associate the fp32 sums (tensor-core tiles, split key ranges), so outputs agree to bf16 rounding noise, not bit for
bit: max |got - want| / max |want| <= 1e-2 (the north_star tolerance) and a mean distance far below it; the share of bit-identical
outputs is printed.  Integer work (which keys are visible, which cache rows are read, the embedding row gather and its dequant)
is exact.
"""
import numpy as np
import pytest

from oracle import orc
from oracle.orc import BF16
from tests.gpu_util import accelerator, bf, unbf

pytestmark = pytest.mark.gpu
LENGTHS = [1, 63, 64, 65, 513, 700]


def rand_bf(rng, shape, scale=1.0):
    return bf(rng.uniform(-1.0, 1.0, size=shape).astype(np.float32) * scale)


def compare(got, want, what):
    g, w = unbf(got).ravel(), unbf(want).ravel()
    scale = float(np.abs(w).max())
    rel = float(np.abs(g - w).max() / scale)
    mean = float(np.abs(g - w).mean() / np.abs(w).mean())
    same = float((np.asarray(got).ravel() == np.asarray(want).ravel()).mean())
    assert rel <= 1e-2 and mean <= 2e-3, f"{what}: max-rel {rel:.3e}, mean-rel {mean:.3e}, bit-identical {same:.3f}"
    return rel, mean, same


@pytest.mark.parametrize("kernel", [0, 1], ids=["cluster_split", "grouped_query"])
@pytest.mark.parametrize("hd,H,KV", [(64, 8, 2), (128, 8, 2), (64, 32, 8)])
def test_decode_attention_matches_oracle(rng, kernel, hd, H, KV):
    from metalchat_b200 import capi

    gpu = accelerator()
    max_seq = 704
    n_seqs = len(LENGTHS)
    K, V = rand_bf(rng, (n_seqs, max_seq, KV, hd)), rand_bf(rng, (n_seqs, max_seq, KV, hd))
    q = rand_bf(rng, (n_seqs, H, hd))
    row_seq = list(range(n_seqs))[::-1]  # rows and sequences deliberately not aligned
    row_pos = [LENGTHS[s] - 1 for s in row_seq]
    got = capi.attn_decode(gpu.dev, q, K, V, row_seq, row_pos, kernel=kernel)
    stats = []
    for r, s in enumerate(row_seq):
        P = LENGTHS[s]
        want = orc.sdpa(BF16, q[r:r + 1], K[s, :P], V[s, :P], causal=False)
        stats.append(compare(got[r], want[0], f"decode attention kernel {kernel} hd {hd} H {H} KV {KV} P {P}"))
    print(f"kernel {kernel} hd {hd} H/KV {H}/{KV}: worst max-rel {max(s[0] for s in stats):.2e}, bit-identical {min(s[2] for s in stats):.3f}..{max(s[2] for s in stats):.3f}")


@pytest.mark.parametrize("hd,H,KV", [(64, 8, 2), (128, 8, 2), (128, 6, 2)])
@pytest.mark.parametrize("start_pos,rows", [(0, 1), (0, 63), (0, 65), (0, 513), (187, 513), (640, 60)])
def test_prefill_attention_matches_oracle_both_chunk_masks(rng, hd, H, KV, start_pos, rows):
    from metalchat_b200 import capi

    gpu = accelerator()
    max_seq = 704
    S = start_pos + rows
    K, V = rand_bf(rng, (1, max_seq, KV, hd)), rand_bf(rng, (1, max_seq, KV, hd))
    q = rand_bf(rng, (rows, H, hd))
    # (a) the engine's default: the cached prefix is visible (the intended reading of make_causal_mask)
    got = capi.attn_prefill(gpu.dev, q, K, V, start_pos, key_begin=0)
    want = orc.sdpa(BF16, q, K[0, :S], V[0, :S], causal=True, prefix_visible=True)
    a = compare(got, want, f"prefill attention hd {hd} start {start_pos} rows {rows} (prefix visible)")
    # (b) the reference's literal mask: the prefix columns stay at -inf when len > 1 (quirk Q9, nn/attention.h:283-299)
    got_ref = capi.attn_prefill(gpu.dev, q, K, V, start_pos, key_begin=start_pos if rows > 1 else 0)
    want_ref = orc.sdpa(BF16, q, K[0, :S], V[0, :S], causal=True, prefix_visible=False)
    b = compare(got_ref, want_ref, f"prefill attention hd {hd} start {start_pos} rows {rows} (reference chunk mask)")
    if start_pos > 0 and rows > 1:
        assert not np.array_equal(want, want_ref), "the two masks must differ when a prefix exists"
    print(f"hd {hd} H/KV {H}/{KV} start {start_pos} rows {rows}: max-rel {a[0]:.2e} / {b[0]:.2e}, bit-identical {a[2]:.3f} / {b[2]:.3f}")


def test_int8_embedding_rows_are_bit_exact(rng):
    """lora_embedding (quantization/lora.h:160-170): out = r(r(q) * r(s)) with one fp32 scale per row (kernel/mul.metal:76-77);
    integer gather + two roundings -> bit-exact against the oracle's hadamard_broadcast."""
    from metalchat_b200 import capi

    gpu = accelerator()
    vocab, D = 4096, 2048
    table = rng.integers(-127, 128, size=(vocab, D), dtype=np.int8)
    scales = (rng.uniform(0.5, 1.5, size=vocab) * 0.0625 / 127.0).astype(np.float32)
    ids = np.array([0, vocab - 1, 17, 17, 2048, 1, 4000, 333], np.int32)
    got = capi.embed_rows(gpu.dev, table, ids, row_scales=scales)
    want = np.zeros((len(ids), D), np.uint16)
    orc.hadamard_broadcast(BF16, orc.F32, want, np.ascontiguousarray(table[ids]), np.ascontiguousarray(scales[ids]))
    assert np.array_equal(got, want)
    # bf16 table: pure gather
    tb = rand_bf(rng, (vocab, D))
    assert np.array_equal(capi.embed_rows(gpu.dev, tb, ids), tb[ids])


def test_chunked_prompt_matches_oracle_in_both_mask_modes(rng):
    """mc_llama_prefill in two calls (start_pos > 0, len > 1): the default attends the cached prefix = the oracle with
    ORC_PREFIX_VISIBLE; MC_LLAMA_REF_CHUNK_MASK = the oracle's literal reference mask (quirk Q9).  Both prompt paths."""
    from metalchat_b200 import capi

    gpu = accelerator()
    cfgd = dict(dim=512, n_layers=3, n_heads=8, n_kv_heads=2, head_dim=64, ffn_dim=1024, vocab=2000, max_seq_len=128)
    ids = [int(x) for x in rng.integers(0, cfgd["vocab"], size=50)]
    for ref_mask in (False, True):
        o = orc.Llama(orc.make_cfg(**cfgd, flags=0 if ref_mask else orc.PREFIX_VISIBLE), BF16)
        o.init_random(0x5EED)
        o.forward(ids[:30], 0)
        want = o.forward(ids[30:], 30)
        for path_flags in (0, capi.LLAMA_NO_TC_PREFILL):
            m = capi.Llama(gpu.dev, capi.llama_config(**cfgd, flags=path_flags | (capi.LLAMA_REF_CHUNK_MASK if ref_mask else 0)))
            m.init_random(0x5EED)
            m.finalize()
            m.prefill(ids[:30], 0)
            m.prefill(ids[30:], 30)
            rel = float(np.abs(unbf(m.logits()) - unbf(want)).max() / np.abs(unbf(want)).max())
            assert rel < 1e-2, (ref_mask, path_flags, rel)
            for layer in range(cfgd["n_layers"]):
                kc = m.cache(0, layer, 0, 50)
                ko = o.cache(0, layer, 0).reshape(cfgd["max_seq_len"], cfgd["n_kv_heads"], cfgd["head_dim"])[:50]
                assert np.abs(unbf(kc) - unbf(ko)).max() / np.abs(unbf(ko)).max() < 1e-2
            m.close()
        # the two modes really differ on this input
        if ref_mask:
            o2 = orc.Llama(orc.make_cfg(**cfgd, flags=orc.PREFIX_VISIBLE), BF16)
            o2.init_random(0x5EED)
            o2.forward(ids[:30], 0)
            assert not np.array_equal(o2.forward(ids[30:], 30), want)
