"""Writes safetensors files for the loader tests (format: src/safetensor.cc:83-133 of the reference: u64 little-endian header
length, JSON header {name: {dtype, shape, data_offsets}}, raw data) and lists the parameters of a model in the reference's
registered-path names."""
import json
import struct

import numpy as np

_TAG = {np.dtype(np.int8): "I8", np.dtype(np.uint8): "U8", np.dtype(np.float32): "F32", np.dtype(np.float16): "F16", np.dtype(np.int32): "I32",
        np.dtype(np.int64): "I64", np.dtype(np.float64): "F64", np.dtype(np.uint32): "U32", np.dtype(np.bool_): "BOOL", np.dtype(np.int16): "I16"}


def write(path, tensors, metadata=None, align=8):
    """tensors: {name: ndarray | (ndarray of uint16, "BF16")}; returns the header dict."""
    header, blobs, off = {}, [], 0
    if metadata:
        header["__metadata__"] = metadata
    for name, t in tensors.items():
        tag = None
        if isinstance(t, tuple):
            t, tag = t
        t = np.asarray(t)
        tag = tag or ("U16" if t.dtype == np.uint16 else _TAG[t.dtype])
        raw = t.tobytes()
        header[name] = {"dtype": tag, "shape": list(t.shape), "data_offsets": [off, off + len(raw)]}
        blobs.append(raw)
        off += len(raw)
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * ((-len(hj)) % align)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for b in blobs:
            f.write(b)
    return header


def param_names(n_layers, quant):
    names = []
    for i in range(n_layers):
        p = f"layers.{i}."
        names += [p + "attention_norm.weight", p + "ffn_norm.weight"]
        for w in ("attention.wq", "attention.wk", "attention.wv", "attention.wo", "feed_forward.w1", "feed_forward.w2", "feed_forward.w3"):
            names += [p + w + s for s in ((".weight", ".scales", ".adaptor.A.weight", ".adaptor.B.weight") if quant else (".weight",))]
    names += ["norm.weight", "tok_embeddings.weight"]
    if quant:
        names += ["tok_embeddings.scales", "output.weight", "output.scales"]
    return names


HF = [("attention_norm", "input_layernorm"), ("ffn_norm", "post_attention_layernorm"), ("feed_forward.w1", "mlp.gate_proj"), ("feed_forward.w2", "mlp.down_proj"),
      ("feed_forward.w3", "mlp.up_proj"), ("attention.wq", "self_attn.q_proj"), ("attention.wk", "self_attn.k_proj"), ("attention.wv", "self_attn.v_proj"),
      ("attention.wo", "self_attn.o_proj")]


def to_hf_name(name):
    """registered path -> HuggingFace name (the inverse of huggingface/llama.h:88-103)"""
    if name.startswith("layers."):
        for meta, hf in HF:
            if f".{meta}." in name:
                return "model." + name.replace(meta, hf)
    return {"norm.weight": "model.norm.weight", "tok_embeddings.weight": "model.embed_tokens.weight", "output.weight": "lm_head.weight"}.get(name, name)
