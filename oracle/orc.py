"""ctypes binding of the CPU oracle (oracle/_build/liborc.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, bench.py's cpu_baseline /
``--impl reference`` legs and ``__graft_entry__.smoke()``.  The product package
``metalchat_b200`` never imports this module.

Arrays are numpy; bf16 tensors are carried as ``uint16`` bit patterns.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liborc.so"

BF16, F32, I32 = 0, 1, 2


def build(force: bool = False) -> Path:
    srcs = [_HERE / n for n in ("orc_api.cc", "orc_model.h", "orc_ops.h", "orc_common.h", "Makefile")]
    stale = not _LIB_PATH.exists() or any(s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.orc_last_error.restype = C.c_char_p
        _lib.orc_f32_to_bf16.restype = C.c_uint16
        _lib.orc_f32_to_bf16.argtypes = [C.c_float]
        _lib.orc_f32_to_bf16_host.restype = C.c_uint16
        _lib.orc_f32_to_bf16_host.argtypes = [C.c_float]
        _lib.orc_bf16_to_f32.restype = C.c_float
        _lib.orc_bf16_to_f32.argtypes = [C.c_uint16]
        _lib.orc_hash3.restype = C.c_uint64
        _lib.orc_hash3.argtypes = [C.c_uint64] * 3
        _lib.orc_hash_uniform.restype = C.c_float
        _lib.orc_hash_uniform.argtypes = [C.c_uint64] * 3
        _lib.orc_hash_int.restype = C.c_int32
        _lib.orc_hash_int.argtypes = [C.c_uint64] * 3 + [C.c_int32, C.c_uint32]
        _lib.orc_pcg32_uniform.restype = C.c_float
        _lib.orc_pcg32_uniform.argtypes = [C.c_uint64, C.c_uint64]
        _lib.orc_llama_create.restype = C.c_void_p
        _lib.orc_llama_tensor.restype = C.c_void_p
        _lib.orc_llama_cache.restype = C.c_void_p
        _lib.orc_llama_decode_timed.restype = C.c_double
        _lib.orc_sample_default.restype = C.c_uint32
        _lib.orc_argmax.restype = C.c_int32
    return _lib


# ---- bf16 helpers (numpy, vectorised; same RNE rule as orc_common.h) ----------------
def f32_to_bf16(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32)
    nan = (u & 0x7FFFFFFF) > 0x7F800000
    bias = np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))
    r = ((u + bias) >> np.uint32(16)).astype(np.uint16)
    r[nan] = ((u[nan] >> np.uint32(16)) | np.uint32(0x40)).astype(np.uint16)
    return r


def bf16_to_f32(b) -> np.ndarray:
    b = np.ascontiguousarray(b, dtype=np.uint16)
    return (b.astype(np.uint32) << np.uint32(16)).view(np.float32)


def np_dtype(dt: int):
    return {BF16: np.uint16, F32: np.float32, I32: np.int32}[dt]


# ---- layouts -------------------------------------------------------------------------
def layout_of(a: np.ndarray, ndim: int | None = None) -> np.ndarray:
    """tensor_layout<N> POD {sizes, strides, offsets} (element units) of a numpy view."""
    n = a.ndim if ndim is None else ndim
    assert a.ndim == n
    item = a.itemsize
    strides = [s // item for s in a.strides]
    assert all(s >= 0 for s in strides)
    return np.array(list(a.shape) + strides + [0] * n, dtype=np.uint32)


def _p(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def _lp(l: np.ndarray):
    return l.ctypes.data_as(C.POINTER(C.c_uint32))


def _args(*arrs):
    out = []
    for a in arrs:
        out += [_p(a), _lp(layout_of(a))]
    return out


class _Keep(list):
    pass


# Each op takes numpy arrays (possibly strided views) and writes into `out`.
def bmm(dt, out, a, b):
    la, lb, lo = layout_of(a), layout_of(b), layout_of(out)
    lib().orc_bmm(dt, _p(out), _lp(lo), _p(a), _lp(la), _p(b), _lp(lb))
    return out


def rmsnorm(dt, out, a, w, eps=1e-5, mu=0.0, block=None):
    D = a.shape[1]
    block = block or -(-D // 1024)
    lo, la, lw = layout_of(out), layout_of(a), layout_of(w)
    lib().orc_rmsnorm(dt, _p(out), _lp(lo), _p(a), _lp(la), _p(w), _lp(lw), C.c_float(eps), C.c_float(mu), C.c_uint32(block))
    return out


def softmax(dt, out, a, block=None):
    D = a.shape[1]
    block = block or -(-D // 1024)
    lo, la = layout_of(out), layout_of(a)
    lib().orc_softmax(dt, _p(out), _lp(lo), _p(a), _lp(la), C.c_uint32(block))
    return out


def rowsum(dt, out, a, block=None):
    D = a.shape[1]
    block = block or -(-D // 1024)
    lo, la = layout_of(out), layout_of(a)
    lib().orc_sum(dt, _p(out), _lp(lo), _p(a), _lp(la), C.c_uint32(block))
    return out


def rope(dt, out, a, fcos, fsin, bs, n_head, start_pos):
    lo, la, lc, ls = layout_of(out), layout_of(a), layout_of(fcos), layout_of(fsin)
    lib().orc_rope(dt, _p(out), _lp(lo), _p(a), _lp(la), _p(fcos), _lp(lc), _p(fsin), _lp(ls), C.c_uint32(bs), C.c_uint32(n_head), C.c_uint32(start_pos))
    return out


def rope_freqs(fcos, fsin, dim, start_pos, theta):
    lc, ls = layout_of(fcos), layout_of(fsin)
    lib().orc_rope_freqs(_p(fcos), _lp(lc), _p(fsin), _lp(ls), C.c_uint32(dim), C.c_uint32(start_pos), C.c_float(theta))
    return fcos, fsin


def embedding(dt, out, ids, w):
    lo, li, lw = layout_of(out), layout_of(ids), layout_of(w)
    lib().orc_embedding(dt, _p(out), _lp(lo), _p(ids), _lp(li), _p(w), _lp(lw))
    return out


def sort(dt, values, indices, a):
    lv, lx, la = layout_of(values), layout_of(indices), layout_of(a)
    lib().orc_sort(dt, _p(values), _lp(lv), _p(indices), _lp(lx), _p(a), _lp(la))
    return values, indices


def cumsum(dt, out, a, block=None):
    D = a.shape[1]
    if block is None:
        b = -(-D // 1024)
        p = 1
        while p < b:
            p *= 2
        block = max(2, p)
    lo, la = layout_of(out), layout_of(a)
    lib().orc_cumsum(dt, _p(out), _lp(lo), _p(a), _lp(la), C.c_uint32(block))
    return out


def multinomial(dt, out, a, init_state=0, init_seq=0, uniforms=None, intended=0):
    lo, la = layout_of(out), layout_of(a)
    up = None
    if uniforms is not None:
        uniforms = np.ascontiguousarray(uniforms, dtype=np.float32)
        up = _p(uniforms)
    lib().orc_multinomial(dt, _p(out), _lp(lo), _p(a), _lp(la), C.c_uint64(init_state), C.c_uint64(init_seq), up, C.c_int(intended))
    return out


def binary(dt, op, out, a, b):
    code = {"add": 0, "sub": 1, "div": 2, "hadamard": 3}[op]
    lo, la, lb = layout_of(out), layout_of(a), layout_of(b)
    lib().orc_binary(dt, code, _p(out), _lp(lo), _p(a), _lp(la), _p(b), _lp(lb))
    return out


def add_broadcast(dt, out, a, b):
    lo, la, lb = layout_of(out), layout_of(a), layout_of(b)
    lib().orc_add_broadcast(dt, _p(out), _lp(lo), _p(a), _lp(la), _p(b), _lp(lb))
    return out


def hadamard_broadcast(odt, sdt, out, a, b):
    lo, la, lb = layout_of(out), layout_of(a), layout_of(b)
    lib().orc_hadamard_broadcast(odt, sdt, _p(out), _lp(lo), _p(a), _lp(la), _p(b), _lp(lb))
    return out


def scalar_mul(dt, out, a, c):
    lo, la = layout_of(out), layout_of(a)
    lib().orc_scalar_mul(dt, _p(out), _lp(lo), _p(a), _lp(la), C.c_float(c))
    return out


def activation(dt, op, out, a):
    lo, la = layout_of(out), layout_of(a)
    lib().orc_activation(dt, {"silu": 0, "gelu": 1}[op], _p(out), _lp(lo), _p(a), _lp(la))
    return out


def copy(dt, out, a):
    lo, la = layout_of(out), layout_of(a)
    lib().orc_copy(0 if dt == BF16 else 1, _p(out), _lp(lo), _p(a), _lp(la))
    return out


def gather(dt, out, a, idx):
    lo, la, li = layout_of(out), layout_of(a), layout_of(idx)
    lib().orc_gather(0 if dt == BF16 else 1, _p(out), _lp(lo), _p(a), _lp(la), _p(idx), _lp(li))
    return out


def scatter(dt, out, mask, value):
    lo, lm = layout_of(out), layout_of(mask)
    lib().orc_scatter(dt, _p(out), _lp(lo), _p(mask), _lp(lm), C.c_float(value))
    return out


def compare(dt, op, out, a, value):
    lo, la = layout_of(out), layout_of(a)
    lib().orc_compare(dt, {"gt": 0, "le": 1}[op], _p(out), _lp(lo), _p(a), _lp(la), C.c_float(value))
    return out


def roll(dt, out, a, shift, size, stride):
    lo, la = layout_of(out), layout_of(a)
    lib().orc_roll(0 if dt == BF16 else 1, _p(out), _lp(lo), _p(a), _lp(la), C.c_uint32(shift), C.c_uint32(size), C.c_uint32(stride))
    return out


# ---- model ----------------------------------------------------------------------------
class LlamaCfg(C.Structure):
    _fields_ = [
        ("dim", C.c_uint32), ("n_layers", C.c_uint32), ("n_heads", C.c_uint32), ("n_kv_heads", C.c_uint32),
        ("head_dim", C.c_uint32), ("ffn_dim", C.c_uint32), ("vocab", C.c_uint32), ("max_seq_len", C.c_uint32),
        ("rope_theta", C.c_float), ("norm_eps", C.c_float),
        ("quant", C.c_uint32), ("lora_rank", C.c_uint32), ("lora_scale", C.c_float), ("group_size", C.c_uint32),
        ("n_seqs", C.c_uint32), ("flags", C.c_uint32),
    ]


UNTIED_HEAD, PREFIX_VISIBLE = 1, 2  # LlamaCfg.flags (orc_model.h ORC_*)


def make_cfg(dim=2048, n_layers=16, n_heads=32, n_kv_heads=8, head_dim=64, ffn_dim=8192, vocab=128256,
             max_seq_len=1024, rope_theta=500000.0, norm_eps=1e-5, quant=0, lora_rank=16, lora_scale=2.0,
             group_size=32, n_seqs=1, flags=0) -> LlamaCfg:
    return LlamaCfg(dim, n_layers, n_heads, n_kv_heads, head_dim, ffn_dim, vocab, max_seq_len, rope_theta,
                    norm_eps, quant, lora_rank, lora_scale, group_size, n_seqs, flags)


class Llama:
    """Scalar CPU Llama-3 forward following nn/llama.h:113-134 (dtype BF16 or F32)."""

    def __init__(self, cfg: LlamaCfg, dtype: int = BF16):
        self.cfg, self.dtype = cfg, dtype
        self.h = lib().orc_llama_create(C.byref(cfg), dtype)
        if not self.h:
            raise RuntimeError(lib().orc_last_error().decode())

    def close(self):
        if self.h:
            lib().orc_llama_destroy(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        self.close()

    def init_random(self, seed: int = 0x5EED):
        lib().orc_llama_init_random(C.c_void_p(self.h), C.c_uint64(seed))

    def tensor(self, name: str, dtype) -> np.ndarray:
        """Writable numpy view on a named parameter (reference layer path names)."""
        n = C.c_uint64(0)
        p = lib().orc_llama_tensor(C.c_void_p(self.h), name.encode(), C.byref(n))
        if not p:
            raise KeyError(name)
        buf = (C.c_uint8 * n.value).from_address(p)
        return np.frombuffer(buf, dtype=dtype)

    def has(self, name: str) -> bool:
        return bool(lib().orc_llama_tensor(C.c_void_p(self.h), name.encode(), None))

    def cache(self, seq: int, layer: int, which: int) -> np.ndarray:
        n = C.c_uint64(0)
        p = lib().orc_llama_cache(C.c_void_p(self.h), seq, layer, which, C.byref(n))
        buf = (C.c_uint8 * n.value).from_address(p)
        return np.frombuffer(buf, dtype=np_dtype(self.dtype))

    def forward(self, ids, start_pos: int, seq: int = 0, want_hidden: bool = False):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        logits = np.zeros(self.cfg.vocab, dtype=np_dtype(self.dtype))
        hidden = np.zeros((len(ids), self.cfg.dim), dtype=np_dtype(self.dtype)) if want_hidden else None
        rc = lib().orc_llama_forward(C.c_void_p(self.h), C.c_uint32(seq), _p(ids), C.c_uint32(len(ids)),
                                     C.c_uint32(start_pos), _p(logits), _p(hidden) if want_hidden else None)
        if rc:
            raise RuntimeError(lib().orc_last_error().decode())
        return (logits, hidden) if want_hidden else logits

    def decode_timed(self, first_id: int, start_pos: int, steps: int):
        toks = np.zeros(steps, dtype=np.int32)
        sec = lib().orc_llama_decode_timed(C.c_void_p(self.h), C.c_int32(first_id), C.c_uint32(start_pos),
                                           C.c_uint32(steps), _p(toks))
        if sec < 0:
            raise RuntimeError(lib().orc_last_error().decode())
        return sec, toks


def sdpa(dt, q, K, V, causal=True, prefix_visible=False):
    """nn::attention from the scores on: q [len, H, hd], K / V [S, KV, hd] -> o [len, H, hd]."""
    q, K, V = (np.ascontiguousarray(a) for a in (q, K, V))
    length, H, hd = q.shape
    S, KV, _ = K.shape
    o = np.zeros_like(q)
    lib().orc_sdpa(dt, _p(o), _p(q), _p(K), _p(V), C.c_uint32(length), C.c_uint32(S), C.c_uint32(H), C.c_uint32(KV), C.c_uint32(hd),
                   C.c_int(int(causal)), C.c_int(int(prefix_visible)))
    return o


def argmax(dt, logits) -> int:
    return int(lib().orc_argmax(dt, _p(logits), C.c_uint32(len(logits))))


def sample_default(dt, logits, topk=50, temperature=0.6, top_p=0.9, u=0.5, intended=0):
    vocab = len(logits)
    k = min(vocab, topk)
    topk_idx = np.zeros(k, np.int32)
    probs = np.zeros(k, np.float32)
    pidx = np.zeros(k, np.int32)
    choice, token = C.c_int32(0), C.c_int32(0)
    lib().orc_sample_default(dt, _p(logits), C.c_uint32(vocab), C.c_uint32(topk), C.c_float(temperature),
                             C.c_float(top_p), C.c_float(u), C.c_int(intended), _p(topk_idx), _p(probs), _p(pidx),
                             C.byref(choice), C.byref(token))
    return dict(topk_idx=topk_idx, probs_sorted=probs, probs_idx=pidx, choice=choice.value, token=token.value)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int):
    lib().orc_set_num_threads(C.c_int(int(n)))
