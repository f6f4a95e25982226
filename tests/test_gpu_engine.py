"""GPU parity of the fused decode engine (mc_llama_* through the C ABI) against the CPU oracle.

Bars (BASELINE.json north_star): integer/index work bit-exact (embedding gather, KV-cache append of
identical values, argmax ties); bf16 activations within max-rel 1e-2 per layer of the fp32 oracle;
identical greedy tokens vs the bf16 oracle.  Full-size runs compare with tests/golden fixtures."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import orc
from oracle.orc import BF16, F32
from tests.gpu_util import accelerator, unbf

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"

SMALL = dict(dim=512, n_layers=3, n_heads=8, n_kv_heads=2, head_dim=64, ffn_dim=1024, vocab=2000, max_seq_len=96)
NAMES = ["attention_norm.weight", "ffn_norm.weight", "attention.wq.weight", "attention.wk.weight", "attention.wv.weight",
         "attention.wo.weight", "feed_forward.w1.weight", "feed_forward.w2.weight", "feed_forward.w3.weight"]


def make_engine(cfgd, seed=0x5EED, flags=0, n_seqs=1, from_oracle=None):
    from metalchat_b200 import capi

    gpu = accelerator()
    m = capi.Llama(gpu.dev, capi.llama_config(**cfgd, n_seqs=n_seqs, flags=flags))
    if from_oracle is None:
        m.init_random(seed)
    else:
        for i in range(cfgd["n_layers"]):
            for n in NAMES:
                name = f"layers.{i}.{n}"
                m.set_tensor(name, from_oracle.tensor(name, np.uint16))
        m.set_tensor("norm.weight", from_oracle.tensor("norm.weight", np.uint16))
        m.set_tensor("tok_embeddings.weight", from_oracle.tensor("tok_embeddings.weight", np.uint16))
    m.finalize()
    return m


def near_top(oracle_logits_bf16, token, ulps=2):
    """True when the oracle scores `token` within `ulps` bf16 steps of its best logit (a near-tie)."""
    lf = unbf(oracle_logits_bf16)
    top = float(lf.max())
    step = 2.0 ** (np.floor(np.log2(abs(top))) - 7) if top != 0 else 0.0
    return float(lf[token]) >= top - ulps * step


def max_rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def test_device_generator_matches_oracle_bits():
    # weights generated on the device == the oracle's generator (same counter hash), checked through
    # the two load paths giving bit-identical logits and caches
    o = orc.Llama(orc.make_cfg(**SMALL), BF16)
    o.init_random(0x5EED)
    a = make_engine(SMALL)
    b = make_engine(SMALL, from_oracle=o)
    ids = [5, 9, 200, 31, 8, 640, 77]
    a.prefill(ids)
    b.prefill(ids)
    assert np.array_equal(a.logits(), b.logits())
    assert np.array_equal(a.cache(0, 2, 0, len(ids)), b.cache(0, 2, 0, len(ids)))


# streaming persistent kernel (default), then the per-op path: graph+PDL, no graph, no PDL, and the GEMV prompt path (no tensor-core prefill)
@pytest.mark.parametrize("flags", [0, 16, 18, 20, 32, 48])
def test_prefill_and_greedy_decode_match_oracle(flags):
    o = orc.Llama(orc.make_cfg(**SMALL), BF16)
    o.init_random(0x5EED)
    of = orc.Llama(orc.make_cfg(**SMALL), F32)
    of.init_random(0x5EED)
    m = make_engine(SMALL, flags=flags)
    ids = [3, 77, 512, 999, 0, 41, 41, 7, 1500, 2]
    m.prefill(ids)
    want_logits, want_hidden = o.forward(ids, 0, want_hidden=True)
    f32_logits, f32_hidden = of.forward(ids, 0, want_hidden=True)
    # KV cache rows: layer 0 keys depend only on embedding -> norm -> wk -> rope: near bit-exact
    for layer in range(SMALL["n_layers"]):
        for which in (0, 1):
            got = m.cache(0, layer, which, len(ids)).reshape(-1)
            want = o.cache(0, layer, which)[: got.size]
            assert max_rel(unbf(got), unbf(want)) < 1e-2
    k0 = m.cache(0, 0, 0, len(ids)).reshape(-1)
    assert np.mean(k0 == o.cache(0, 0, 0)[: k0.size]) > 0.99
    # hidden state after the last block and logits: bf16 engine vs fp32 oracle within 1e-2 max-rel
    assert max_rel(unbf(m.hidden()), f32_hidden[-1]) < 1e-2
    assert max_rel(unbf(m.logits()), f32_logits) < 1e-2
    assert max_rel(unbf(m.logits()), unbf(want_logits)) < 1e-2
    # greedy tokens: engine loop (token fed back on the device) vs oracle loop
    first = orc.argmax(BF16, want_logits)
    assert int(np.argmax(unbf(m.logits()))) == first
    steps = 24
    toks, _ = m.decode_loop([first], [len(ids)], steps)
    want = []
    tok, pos = first, len(ids)
    for _ in range(steps):
        lg = o.forward([tok], pos)
        tok = orc.argmax(BF16, lg)
        want.append(tok)
        pos += 1
    assert toks[:, 0].tolist() == want


@pytest.mark.parametrize("flags", [0, 16])
def test_decode_call_matches_decode_loop_and_prefill(flags):
    m = make_engine(SMALL, flags=flags)
    ids = [11, 22, 33, 44, 55]
    m.prefill(ids)
    l_prefill = m.logits().copy()
    m2 = make_engine(SMALL, flags=flags)
    for t, tok in enumerate(ids):
        out = m2.decode([tok], [t])
    if flags & 16:
        # per-op path: prefill (4 rows per pass) and token-by-token decode run the same kernels: same bits
        assert np.array_equal(m2.logits(), l_prefill)
        assert out[0] == int(np.lexsort((np.arange(SMALL["vocab"]), -unbf(l_prefill)))[0])
    else:
        # streaming kernel (tensor-core k-blocks) vs the per-op prefill: same rounding points, fp32 sums re-associated
        assert max_rel(unbf(m2.logits()), unbf(l_prefill)) < 1e-2
        assert np.abs(unbf(m2.logits()) - unbf(l_prefill)).mean() / np.abs(unbf(l_prefill)).mean() < 1e-2
        assert near_top(l_prefill, int(out[0]))
        for layer in range(SMALL["n_layers"]):
            for which in (0, 1):
                assert max_rel(unbf(m2.cache(0, layer, which, len(ids))), unbf(m.cache(0, layer, which, len(ids)))) < 1e-2
        assert np.array_equal(m2.cache(0, 0, 1, len(ids)), m.cache(0, 0, 1, len(ids))) or np.mean(m2.cache(0, 0, 1, len(ids)) == m.cache(0, 0, 1, len(ids))) > 0.99
    # the device-side loop (several steps per launch) and the per-token call give the same tokens from the same state
    m3 = make_engine(SMALL, flags=flags)
    m3.prefill(ids)
    a, _ = m.decode_loop([out[0]], [len(ids)], 8)
    b = []
    tok = out[0]
    for s in range(8):
        tok = int(m3.decode([tok], [len(ids) + s])[0])
        b.append(tok)
    assert a[:, 0].tolist() == b


@pytest.mark.parametrize("flags", [0, 16])
def test_multi_sequence_decode_is_independent(flags):
    # "batch" = independent bs=1 sequences (quirk Q2/Q15): row r of a 6-sequence step equals a single-sequence run
    n = 6
    m = make_engine(SMALL, n_seqs=n, flags=flags)
    single = make_engine(SMALL, flags=flags)
    prompts = [[(7 * s + 3 * t) % SMALL["vocab"] for t in range(3 + s)] for s in range(n)]
    for s, p in enumerate(prompts):
        m.prefill(p, seq=s)
    firsts = [int(np.argmax(unbf(m.logits(s)))) for s in range(n)]
    toks, _ = m.decode_loop(firsts, [len(p) for p in prompts], 6)
    single.prefill(prompts[4])
    want, _ = single.decode_loop([firsts[4]], [len(prompts[4])], 6)
    assert toks[:, 4].tolist() == want[:, 0].tolist()


def test_argument_validation():
    from metalchat_b200 import capi

    gpu = accelerator()
    with pytest.raises(capi.McInvalidArgument):
        capi.Llama(gpu.dev, capi.llama_config(**{**SMALL, "head_dim": 48}))
    m = make_engine(SMALL)
    with pytest.raises(capi.McInvalidArgument, match="token id out of range"):
        m.prefill([SMALL["vocab"]])
    with pytest.raises(capi.McInvalidArgument, match="max_seq_len"):
        m.prefill(list(range(SMALL["max_seq_len"] + 1)))
    with pytest.raises(capi.McNotFound):
        m.set_tensor("layers.0.nope.weight", np.zeros(4, np.uint16))
    with pytest.raises(capi.McInvalidArgument, match="expected"):
        m.set_tensor("norm.weight", np.zeros(4, np.uint16))


def test_config1_single_1b_layer_seq128():
    # BASELINE.json configs[0]: one Llama-3.2-1B-shaped decoder layer, random-init bf16, 127 cached + 1 new position
    cfgd = dict(dim=2048, n_layers=1, n_heads=32, n_kv_heads=8, head_dim=64, ffn_dim=8192, vocab=128256, max_seq_len=256)
    o = orc.Llama(orc.make_cfg(**cfgd), BF16)
    o.init_random(0x5EED)
    of = orc.Llama(orc.make_cfg(**cfgd), F32)
    of.init_random(0x5EED)
    m = make_engine(cfgd)
    ids = [int(orc.lib().orc_hash_int(0x5EED, 0xFFFF, i, 0, cfgd["vocab"])) for i in range(128)]
    m.prefill(ids[:127])
    o.forward(ids[:127], 0)
    of.forward(ids[:127], 0)
    m.prefill(ids[127:], start_pos=127)  # the decode step of the config, logits kept
    lb, hb = o.forward(ids[127:], 127, want_hidden=True)
    lf, hf = of.forward(ids[127:], 127, want_hidden=True)
    assert max_rel(unbf(m.hidden()), hf[-1]) < 1e-2
    assert max_rel(unbf(m.logits()), lf) < 1e-2
    # vs the bf16-rounding oracle: a single 1-ulp flip (fp32 re-association) cascades through the bf16 chain, so
    # elements agree to within a couple of bf16 ulps rather than bit for bit at this width
    ulps = np.abs(m.hidden().astype(np.int32) - hb[-1].astype(np.int32))
    assert np.median(ulps) <= 1, np.median(ulps)
    # mean distance to the bf16 oracle: ~3e-3 for either prompt path (tools/diag_config1.py: 3.2e-3 tensor-core, 3.0e-3 GEMV), half of
    # the bf16 oracle's own distance to the fp32 oracle (6.6e-3)
    assert np.abs(unbf(m.hidden()) - unbf(hb[-1])).mean() / np.abs(unbf(hb[-1])).mean() < 4e-3
    assert max_rel(unbf(m.hidden()), unbf(hb[-1])) < 1e-2
    assert int(np.argmax(unbf(m.logits()))) == orc.argmax(BF16, lb)


# (the full-size 16-layer checks of configs[1] / configs[2] live in tests/test_gpu_golden.py)


# ---- quantised (QLoRA layout) path: BASELINE.json configs[2] -----------------------------------------------------------
QNAMES = [f"{l}.{s}" for l in ("attention.wq", "attention.wk", "attention.wv", "attention.wo", "feed_forward.w1", "feed_forward.w2", "feed_forward.w3")
          for s in ("weight", "scales", "adaptor.A.weight", "adaptor.B.weight")]
QDT = {"weight": np.int8, "scales": np.float32, "adaptor.A.weight": np.uint16, "adaptor.B.weight": np.uint16}


def make_qengine(cfgd, seed=0x5EED, n_seqs=1, from_oracle=None, flags=0):
    from metalchat_b200 import capi

    gpu = accelerator()
    m = capi.Llama(gpu.dev, capi.llama_config(**cfgd, quant=1, n_seqs=n_seqs, flags=flags))
    if from_oracle is None:
        m.init_random(seed)
    else:
        o = from_oracle
        for i in range(cfgd["n_layers"]):
            for n in QNAMES:
                name = f"layers.{i}.{n}"
                m.set_tensor(name, o.tensor(name, QDT[n.split(".", 2)[2]]))
            for n in ("attention_norm.weight", "ffn_norm.weight"):
                m.set_tensor(f"layers.{i}.{n}", o.tensor(f"layers.{i}.{n}", np.uint16))
        m.set_tensor("norm.weight", o.tensor("norm.weight", np.uint16))
        for t in ("tok_embeddings", "output"):
            m.set_tensor(t + ".weight", o.tensor(t + ".weight", np.int8))
            m.set_tensor(t + ".scales", o.tensor(t + ".scales", np.float32))
    m.finalize()
    return m


def test_w4_pack_roundtrip_and_linear(rng):
    # int4 unpack must be bit-exact; the base linear r(x . r(r(q) r(s))^T) against the oracle's dequant + bmm
    from metalchat_b200 import capi

    gpu = accelerator()
    for (N, K, M) in [(64, 256, 1), (512, 2048, 3), (2048, 8192, 8), (34, 512, 2)]:
        q = rng.integers(-8, 8, size=(N, K), dtype=np.int8)
        s = (rng.random((N, K // 32), dtype=np.float32) + 0.5) * 0.01
        w4, sp = capi.pack_w4(gpu.dev, q, s)
        assert np.array_equal(capi.unpack_w4(gpu.dev, w4, N, K), q)
        x = orc.f32_to_bf16(rng.standard_normal((M, K)).astype(np.float32))
        dx, dy = gpu.dev.upload(x), gpu.dev.alloc(M * N * 2)
        capi.linear_w4(gpu.dev, dy, dx, w4, sp, M, N, K)
        got = dy.read(np.uint16).reshape(M, N)
        # oracle: hadamard_broadcast (group rows of 32) then bmm through the transposed view
        wd = np.zeros((N * (K // 32), 32), np.uint16)
        orc.hadamard_broadcast(BF16, F32, wd, q.reshape(-1, 32), s.reshape(-1))
        want = np.zeros((1, M, N), np.uint16)
        orc.bmm(BF16, want, x.reshape(1, M, K), wd.reshape(1, N, K).transpose(0, 2, 1))
        # fp32 accumulation order differs (tensor-core k-blocks vs ascending k): almost every output is bit-identical,
        # the rest within one bf16 ulp of the largest output (near-zero sums can be many of their own ulps apart)
        exact = np.mean(got == want[0])
        err = np.abs(unbf(got) - unbf(want[0])).max() / np.abs(unbf(want[0])).max()
        assert exact > 0.98 and err < 2.0 ** -8, (N, K, M, exact, err)
    with pytest.raises(capi.McInvalidArgument, match="int4 range"):
        capi.pack_w4(gpu.dev, np.full((16, 256), 9, np.int8), np.ones((16, 8), np.float32))


def test_quant_engine_matches_oracle():
    cfgd = SMALL
    o = orc.Llama(orc.make_cfg(**cfgd, quant=1), BF16)
    o.init_random(0x5EED)
    of = orc.Llama(orc.make_cfg(**cfgd, quant=1), F32)
    of.init_random(0x5EED)
    a = make_qengine(cfgd)
    b = make_qengine(cfgd, from_oracle=o)
    ids = [3, 77, 512, 999, 0, 41, 41, 7, 1500, 2]
    a.prefill(ids)
    b.prefill(ids)
    assert np.array_equal(a.logits(), b.logits())  # device generator == oracle generator, both load paths
    want_logits, want_hidden = o.forward(ids, 0, want_hidden=True)
    f32_logits, f32_hidden = of.forward(ids, 0, want_hidden=True)
    # north_star tolerance is 1e-2 max-rel PER LAYER vs the fp32 oracle (whose weights are not rounded to bf16)
    assert max_rel(unbf(a.hidden()), f32_hidden[-1]) < 1e-2 * cfgd["n_layers"]
    assert max_rel(unbf(a.logits()), f32_logits) < 1e-2 * cfgd["n_layers"]
    assert max_rel(unbf(a.hidden()), unbf(want_hidden[-1])) < 1e-2
    assert max_rel(unbf(a.logits()), unbf(want_logits)) < 1e-2
    # greedy tokens, teacher-forced with the oracle's tokens: a random-init model this small has near-ties between the two
    # best logits, so a token may differ only where the oracle itself scores it within 2 bf16 ulps of its maximum
    tok, pos, exact = orc.argmax(BF16, want_logits), len(ids), 0
    steps = 24
    assert near_top(want_logits, int(np.argmax(unbf(a.logits()))))
    for _ in range(steps):
        got = int(a.decode([tok], [pos])[0])
        lg = o.forward([tok], pos)
        tok = orc.argmax(BF16, lg)
        exact += got == tok
        assert near_top(lg, got), (got, tok)
        pos += 1
    assert exact >= steps - 3, exact


# ---- streaming persistent kernel: shapes and lengths beyond the default small config -------------------------------------
def _stream_vs_perop(cfgd, quant, prompt_len, steps, n_seqs=1):
    """Greedy decode through the streaming kernel and through the per-op path from the same prefill state."""
    from metalchat_b200 import capi

    gpu = accelerator()
    outs = []
    for flags in (0, capi.LLAMA_NO_STREAM):
        m = capi.Llama(gpu.dev, capi.llama_config(**cfgd, n_seqs=n_seqs, flags=flags, quant=quant))
        m.init_random(0x5EED)
        m.finalize()
        firsts = []
        for s in range(n_seqs):
            ids = [int(orc.lib().orc_hash_int(0x5EED, 0xFFFF + s, i, 0, cfgd["vocab"])) for i in range(prompt_len + s)]
            m.prefill(ids, seq=s)
            firsts.append(int(np.argmax(unbf(m.logits(s)))))
        toks, _ = m.decode_loop(firsts, [prompt_len + s for s in range(n_seqs)], steps)
        assert m.launches_per_step() == (1 if flags == 0 else m.launches_per_step())
        outs.append((toks.copy(), [m.logits(s).copy() for s in range(n_seqs)], m.launches_per_step()))
    (ta, la, na), (tb, lb, nb) = outs
    assert na == 1 and nb > 1, (na, nb)  # the first engine really took the streaming kernel
    return ta, la, tb, lb


@pytest.mark.parametrize("quant", [0, 1])
def test_stream_head_dim_128_and_eight_sequences(quant):
    # head_dim 128 instantiation, 8 sequences at different positions in one launch, several steps per launch
    cfgd = dict(dim=768, n_layers=2, n_heads=6, n_kv_heads=2, head_dim=128, ffn_dim=1536, vocab=4000, max_seq_len=128)
    ta, la, tb, lb = _stream_vs_perop(cfgd, quant, 9, 12, n_seqs=8)
    agree = np.mean(ta == tb)
    assert agree > 0.9, agree  # near-ties of a random-init model may flip a token; the logits must stay within tolerance
    first_div = [int(np.argmax(ta[:, s] != tb[:, s])) if (ta[:, s] != tb[:, s]).any() else None for s in range(8)]
    for s in range(8):
        if first_div[s] is None:  # same token history: same state, logits comparable
            # two decode paths, 12 steps, each with its own fp32 association; the QLoRA chain has three more bf16 rounding points
            # per linear, so its bar is 1.5e-2 (measured 0.9e-2 .. 1.2e-2 depending on the prompt path that filled the cache)
            assert max_rel(unbf(la[s]), unbf(lb[s])) < (1.5e-2 if quant else 1e-2)


@pytest.mark.parametrize("quant", [0, 1])
def test_stream_long_history(quant):
    # more cached positions than the attention phase keeps in registers (576): exercises its streaming tail loops
    cfgd = dict(dim=512, n_layers=2, n_heads=8, n_kv_heads=2, head_dim=64, ffn_dim=1024, vocab=2000, max_seq_len=1024)
    ta, la, tb, lb = _stream_vs_perop(cfgd, quant, 700, 6)
    if (ta == tb).all():
        assert max_rel(unbf(la[0]), unbf(lb[0])) < 1e-2
    assert np.mean(ta == tb) >= 0.5


def test_stream_many_launches_keep_tags_unique():
    # per-token calls: one launch each, the tag sequence advances; results equal the multi-step launch
    m = make_engine(SMALL)
    m.prefill([5, 6, 7])
    first = int(np.argmax(unbf(m.logits())))
    a, _ = m.decode_loop([first], [3], 40)
    m2 = make_engine(SMALL)
    m2.prefill([5, 6, 7])
    tok, b = first, []
    for s in range(40):
        tok = int(m2.decode([tok], [3 + s])[0])
        b.append(tok)
    assert a[:, 0].tolist() == b


# ---- decode beyond max_seq_len: the sink-cache roll (nn/cache.h:183-204, kernel/roll.metal) ----------------------------------------------
@pytest.mark.parametrize("quant,flags_name", [(0, "per_op"), (0, "batched_tc"), (1, "per_op")])
def test_decode_past_the_cache_rolls_like_sink_cache(quant, flags_name):
    """max_seq_len 96: positions 96.. keep the log2(96) = 6 sink rows, shift the rest left and write the last row; RoPE stays at the
    absolute position (the tables are regrown past 2 * max_seq_len).  Oracle: orc_model.h forward(), itself bit-identical to the
    reference's nn::sink_cache code (tests/test_oracle_ref.py SINK_CASES)."""
    from metalchat_b200 import capi

    n_seqs = 6 if flags_name == "batched_tc" else 1
    o = orc.Llama(orc.make_cfg(**SMALL, quant=quant, n_seqs=n_seqs), BF16)
    o.init_random(0x5EED)
    m = make_qengine(SMALL, n_seqs=n_seqs) if quant else make_engine(SMALL, n_seqs=n_seqs)
    rng = np.random.default_rng(7)
    prompts = [[int(x) for x in rng.integers(0, SMALL["vocab"], size=90 - 3 * s)] for s in range(n_seqs)]
    toks, pos = [], []
    for s, p in enumerate(prompts):
        m.prefill(p, 0, s)
        toks.append(orc.argmax(BF16, o.forward(p, 0, seq=s)))
        pos.append(len(p))
    steps = 118  # sequence 0 runs from position 90 to 207: past the cache (96) and past the initial RoPE tables (192)
    for step in range(steps):
        got = m.decode(toks, pos)
        for s in range(n_seqs):
            lg = o.forward([toks[s]], pos[s], seq=s)
            want = orc.argmax(BF16, lg)
            assert near_top(lg, int(got[s])), (flags_name, step, s, pos[s], int(got[s]), want)
            if s == 0 and step in (5, 6, 7, 60, 117):
                rel = max_rel(unbf(m.logits(0)), unbf(lg))
                assert rel < 2e-2, (step, pos[0], rel)
                kc = m.cache(0, 1, 0, SMALL["max_seq_len"]).reshape(-1)
                ko = o.cache(0, 1, 0)[: kc.size]
                assert max_rel(unbf(kc), unbf(ko)) < 2e-2, (step, pos[0])
            toks[s] = want  # teacher-forced along the oracle's path
            pos[s] += 1
    # the device-side loop crosses the boundary too: same tokens as the per-token call from the same state (both on the per-op kernels:
    # a call that reaches past the cache takes them for all of its steps, and the int4 streaming kernel re-associates its sums)
    a = make_qengine(SMALL, flags=capi.LLAMA_NO_STREAM) if quant else make_engine(SMALL, flags=capi.LLAMA_NO_STREAM)
    b = make_qengine(SMALL, flags=capi.LLAMA_NO_STREAM) if quant else make_engine(SMALL, flags=capi.LLAMA_NO_STREAM)
    for e in (a, b):
        e.prefill(prompts[0], 0, 0)
    t_loop, _ = a.decode_loop([5], [90], 12)
    tok, out = 5, []
    for i in range(12):
        tok = int(b.decode([tok], [90 + i])[0])
        out.append(tok)
    assert t_loop[:, 0].tolist() == out
