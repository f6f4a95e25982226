"""safetensors -> device loader (mc_llama_load_safetensors) on a B200: a checkpoint written from the oracle's
tensors and loaded from disk must give the same bits as the same tensors handed over one by one (mc_llama_set_tensor), for
HuggingFace names (huggingface/llama.h:88-103), Meta-format head order (reference.h:73-94), fp32 files, QLoRA layouts
(huggingface/llama.h:152-171) and sharded checkpoints; and the model must then follow the oracle."""
import numpy as np
import pytest

from oracle import orc
from oracle.orc import BF16
from tests import st_files
from tests.gpu_util import accelerator, unbf
from tests.test_gpu_engine import SMALL, make_engine, make_qengine, max_rel

pytestmark = pytest.mark.gpu
PROMPT = [5, 17, 1999, 3, 250, 77, 1024, 9, 12, 640, 31]


def shapes(c, quant):
    D, H, KV, hd, F, V, r, g = c["dim"], c["n_heads"], c["n_kv_heads"], c["head_dim"], c["ffn_dim"], c["vocab"], 16, 32
    lin = {"attention.wq": (H * hd, D), "attention.wk": (KV * hd, D), "attention.wv": (KV * hd, D), "attention.wo": (D, H * hd),
           "feed_forward.w1": (F, D), "feed_forward.w2": (D, F), "feed_forward.w3": (F, D)}
    out = {}
    for i in range(c["n_layers"]):
        p = f"layers.{i}."
        out[p + "attention_norm.weight"] = out[p + "ffn_norm.weight"] = (D,)
        for n, (N, K) in lin.items():
            out[p + n + ".weight"] = (N, K)
            if quant:
                out[p + n + ".scales"] = (N, K // g)
                out[p + n + ".adaptor.A.weight"] = (r, K)
                out[p + n + ".adaptor.B.weight"] = (N, r)
    out["norm.weight"] = (D,)
    out["tok_embeddings.weight"] = (V, D)
    if quant:
        out["tok_embeddings.scales"] = out["output.scales"] = (V, 1)
        out["output.weight"] = (V, D)
    return out


def oracle_tensors(o, c, quant):
    """{registered name: ndarray (bf16 as uint16)} of every parameter of the oracle model"""
    out = {}
    for name, shp in shapes(c, quant).items():
        dt = np.uint16
        if quant and name.endswith(".scales"):
            dt = np.float32
        elif quant and name.endswith(".weight") and "norm" not in name and "adaptor" not in name:
            dt = np.int8
        out[name] = o.tensor(name, dt).reshape(shp).copy()
    return out


def tagged(tensors):
    return {n: ((t, "BF16") if t.dtype == np.uint16 else t) for n, t in tensors.items()}


def load_from(path, quant=0, flags=None, cfgd=SMALL):
    from metalchat_b200 import capi

    gpu = accelerator()
    m = capi.Llama(gpu.dev, capi.llama_config(**cfgd, quant=quant))
    st = capi.Safetensors(path)
    n = m.load_safetensors(st, capi.LOAD_STRICT if flags is None else flags)
    st.close()  # the device holds its own copy
    m.finalize()
    return m, n


def run(m):
    m.prefill(PROMPT)
    lg = m.logits().copy()
    toks = [int(np.argmax(unbf(lg)))]
    for s in range(6):
        toks.append(int(m.decode([toks[-1]], [len(PROMPT) + s])[0]))
    return lg, toks


def test_hf_named_checkpoint_equals_set_tensor_and_follows_the_oracle(tmp_path):
    from metalchat_b200 import capi

    o = orc.Llama(orc.make_cfg(**SMALL), BF16)
    o.init_random(0x5EED)
    t = oracle_tensors(o, SMALL, 0)
    st_files.write(tmp_path / "model.safetensors", {st_files.to_hf_name(n): v for n, v in tagged(t).items()}, metadata={"format": "pt"})
    m, n = load_from(tmp_path / "model.safetensors", flags=capi.LOAD_STRICT | capi.LOAD_HF_NAMES)
    assert n == len(t)
    ref = make_engine(SMALL, from_oracle=o)
    lg, toks = run(m)
    lg_ref, toks_ref = run(ref)
    np.testing.assert_array_equal(lg, lg_ref)
    assert toks == toks_ref
    want = o.forward(PROMPT, 0)
    assert max_rel(unbf(lg), unbf(want)) < 1e-2
    # without the rename flag nothing matches: strict loading names the first missing parameter
    with pytest.raises(capi.McInvalidArgument, match="attention_norm.weight is missing"):
        load_from(tmp_path / "model.safetensors", flags=capi.LOAD_STRICT)


def test_untied_lm_head_and_sharded_directory(tmp_path):
    from metalchat_b200 import capi

    o = orc.Llama(orc.make_cfg(**SMALL, flags=orc.UNTIED_HEAD), BF16)
    o.init_random(0x5EED)
    t = oracle_tensors(o, SMALL, 0)
    t["output.weight"] = o.tensor("output.weight", np.uint16).reshape(SMALL["vocab"], SMALL["dim"]).copy()
    names = sorted(t)
    half = len(names) // 2
    for k, part in enumerate((names[:half], names[half:])):
        st_files.write(tmp_path / f"model-0000{k + 1}-of-00002.safetensors", {st_files.to_hf_name(n): tagged(t)[n] for n in part})
    m, n = load_from(tmp_path, flags=capi.LOAD_STRICT | capi.LOAD_HF_NAMES)
    assert n == len(t)
    lg, toks = run(m)
    want = o.forward(PROMPT, 0)
    assert max_rel(unbf(lg), unbf(want)) < 1e-2
    assert toks[0] == orc.argmax(BF16, want)
    # the same files without lm_head: the head stays tied to the embedding and the logits change
    (tmp_path / "tied").mkdir()
    st_files.write(tmp_path / "tied" / "model.safetensors", {st_files.to_hf_name(n): v for n, v in tagged(t).items() if n != "output.weight"})
    tied, n2 = load_from(tmp_path / "tied", flags=capi.LOAD_STRICT | capi.LOAD_HF_NAMES)
    assert n2 == len(t) - 1
    assert not np.array_equal(run(tied)[0], lg)


def test_meta_format_head_order_and_fp32_files(tmp_path):
    from metalchat_b200 import capi

    o = orc.Llama(orc.make_cfg(**SMALL), BF16)
    o.init_random(0x5EED)
    t = oracle_tensors(o, SMALL, 0)
    ref_lg, ref_toks = run(make_engine(SMALL, from_oracle=o))
    hd = SMALL["head_dim"]
    meta = dict(t)
    for name, w in t.items():
        if name.endswith("attention.wq.weight") or name.endswith("attention.wk.weight"):
            # the inverse of nn/attention.h:232-247: meta[head, j, k] = hf[head, k, j]
            heads = w.shape[0] // hd
            meta[name] = w.reshape(heads, 2, hd // 2, w.shape[1]).transpose(0, 2, 1, 3).reshape(w.shape).copy()
    st_files.write(tmp_path / "consolidated.safetensors", tagged(meta))
    m, _ = load_from(tmp_path / "consolidated.safetensors", flags=capi.LOAD_STRICT | capi.LOAD_META_PERMUTE)
    lg, toks = run(m)
    np.testing.assert_array_equal(lg, ref_lg)
    assert toks == ref_toks
    plain, _ = load_from(tmp_path / "consolidated.safetensors")
    assert not np.array_equal(run(plain)[0], ref_lg)
    # an fp32 file of the same (bf16-representable) values rounds to the same bits; fp16 goes through fp32
    as_f32 = {n: (v.astype(np.uint32) << 16).view(np.float32) for n, v in t.items()}
    st_files.write(tmp_path / "f32.safetensors", as_f32)
    m32, _ = load_from(tmp_path / "f32.safetensors")
    np.testing.assert_array_equal(run(m32)[0], ref_lg)
    as_f32["norm.weight"] = as_f32["norm.weight"].astype(np.float16)
    st_files.write(tmp_path / "f16.safetensors", as_f32)
    m16, _ = load_from(tmp_path / "f16.safetensors")
    want = make_engine(SMALL, from_oracle=o)
    want.set_tensor("norm.weight", orc.f32_to_bf16(as_f32["norm.weight"].astype(np.float32)))
    want.finalize()
    np.testing.assert_array_equal(run(m16)[0], run(want)[0])


def test_qlora_checkpoint_and_errors(tmp_path):
    from metalchat_b200 import capi

    o = orc.Llama(orc.make_cfg(**SMALL, quant=1), BF16)
    o.init_random(0x5EED)
    t = oracle_tensors(o, SMALL, 1)
    st_files.write(tmp_path / "qlora.safetensors", tagged(t))
    m, n = load_from(tmp_path / "qlora.safetensors", quant=1)
    assert n == len(t) == len(st_files.param_names(SMALL["n_layers"], True))
    ref = make_qengine(SMALL, from_oracle=o)
    lg, toks = run(m)
    lg_ref, toks_ref = run(ref)
    np.testing.assert_array_equal(lg, lg_ref)
    assert toks == toks_ref
    # a bf16 model cannot take the int8 weights, a quantised model cannot take bf16 ones
    with pytest.raises(capi.McInvalidArgument, match="is I8, the model stores BF16"):
        load_from(tmp_path / "qlora.safetensors", quant=0)
    bad = dict(tagged(t))
    del bad["layers.1.feed_forward.w2.scales"]
    st_files.write(tmp_path / "missing.safetensors", bad)
    with pytest.raises(capi.McInvalidArgument, match="layers.1.feed_forward.w2.scales is missing"):
        load_from(tmp_path / "missing.safetensors", quant=1)
    m2, n2 = load_from(tmp_path / "missing.safetensors", quant=1, flags=0)  # not strict: the parameter keeps its zero initialisation
    assert n2 == len(t) - 1
    bad = dict(tagged(t))
    bad["layers.0.attention.wq.weight"] = t["layers.0.attention.wq.weight"][:-1]
    st_files.write(tmp_path / "shape.safetensors", bad)
    with pytest.raises(capi.McInvalidArgument, match="expected .* bytes"):
        load_from(tmp_path / "shape.safetensors", quant=1)
