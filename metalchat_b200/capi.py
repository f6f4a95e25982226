"""ctypes binding of libmc_cuda.so (include/mc_cuda.h).

Thin, allocation-explicit mirror of the C ABI used by the tests, bench.py and the Python host
layer.  There is NO fallback: if the shared library is missing or no B200 is present the
constructors raise.  Nothing in this package imports the test oracle.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libmc_cuda.so"

MC_OK, MC_ERR_INVALID, MC_ERR_RUNTIME, MC_ERR_ALLOC, MC_ERR_NOT_FOUND, MC_ERR_FULL = range(6)
MEM_DEVICE, MEM_SHARED, MEM_PINNED = 0, 1, 2
LLAMA_W4_PACKED, LLAMA_NO_GRAPH, LLAMA_NO_PDL, LLAMA_NO_STREAM, LLAMA_NO_TC_PREFILL, LLAMA_NO_SHADOW, LLAMA_REF_CHUNK_MASK = 1, 2, 4, 16, 32, 64, 128


class McError(RuntimeError):
    """std::runtime_error of the reference (command-buffer / library failures)."""

    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


class McInvalidArgument(McError, ValueError):
    """std::invalid_argument of the reference (shape / launch validation, kernel.h:126-140)."""


class McAllocError(McError, MemoryError):
    """alloc_error of the reference (allocator.h:20-34)."""


class McNotFound(McError, KeyError):
    pass


class LlamaConfig(C.Structure):
    _fields_ = [
        ("dim", C.c_uint32), ("n_layers", C.c_uint32), ("n_heads", C.c_uint32), ("n_kv_heads", C.c_uint32),
        ("head_dim", C.c_uint32), ("ffn_dim", C.c_uint32), ("vocab", C.c_uint32), ("max_seq_len", C.c_uint32),
        ("rope_theta", C.c_float), ("norm_eps", C.c_float),
        ("quant", C.c_uint32), ("lora_rank", C.c_uint32), ("lora_scale", C.c_float), ("group_size", C.c_uint32),
        ("n_seqs", C.c_uint32), ("tp_rank", C.c_uint32), ("tp_world", C.c_uint32), ("flags", C.c_uint32),
    ]


class SamplerConfig(C.Structure):
    _fields_ = [("mode", C.c_uint32), ("top_k", C.c_uint32), ("temperature", C.c_float), ("top_p", C.c_float),
                ("intended", C.c_uint32)]


def llama_config(dim=2048, n_layers=16, n_heads=32, n_kv_heads=8, head_dim=64, ffn_dim=8192, vocab=128256,
                 max_seq_len=1024, rope_theta=500000.0, norm_eps=1e-5, quant=0, lora_rank=16, lora_scale=2.0,
                 group_size=32, n_seqs=1, tp_rank=0, tp_world=1, flags=0) -> LlamaConfig:
    return LlamaConfig(dim, n_layers, n_heads, n_kv_heads, head_dim, ffn_dim, vocab, max_seq_len, rope_theta, norm_eps,
                       quant, lora_rank, lora_scale, group_size, n_seqs, tp_rank, tp_world, flags)


class SafetensorsEntry(C.Structure):
    """mc_safetensors_entry of include/mc_cuda.h"""
    _fields_ = [("name", C.c_char_p), ("dtype", C.c_char_p), ("rank", C.c_uint32), ("shape", C.c_uint64 * 8), ("data", C.c_void_p), ("nbytes", C.c_uint64)]


LOAD_HF_NAMES, LOAD_META_PERMUTE, LOAD_STRICT = 1, 2, 4

# every exported symbol of include/mc_cuda.h (checked by tests/test_abi.py against the header)
_SIGNATURES = {
    "mc_last_error": (C.c_char_p, []),
    "mc_version": (C.c_char_p, []),
    "mc_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "mc_device_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "mc_device_destroy": (C.c_int, [C.c_void_p]),
    "mc_device_name": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "mc_device_max_buffer": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t)]),
    "mc_device_sm_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "mc_device_synchronize": (C.c_int, [C.c_void_p]),
    "mc_device_stream": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "mc_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "mc_alloc_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "mc_wrap_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "mc_buffer_retain": (C.c_int, [C.c_void_p]),
    "mc_buffer_release": (C.c_int, [C.c_void_p]),
    "mc_buffer_host_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "mc_buffer_dev_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "mc_buffer_size": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t)]),
    "mc_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "mc_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "mc_memset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_size_t]),
    "mc_heap_create": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "mc_heap_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "mc_heap_reset": (C.c_int, [C.c_void_p]),
    "mc_heap_destroy": (C.c_int, [C.c_void_p]),
    "mc_kernel_lookup": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "mc_kernel_release": (C.c_int, [C.c_void_p]),
    "mc_kernel_name": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p)]),
    "mc_kernel_max_threads": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t)]),
    "mc_kernel_count": (C.c_int, [C.POINTER(C.c_int)]),
    "mc_kernel_name_at": (C.c_int, [C.c_int, C.POINTER(C.c_char_p)]),
    "mc_stream_begin": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "mc_set_bytes": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]),
    "mc_set_buffer": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]),
    "mc_barrier": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mc_dispatch": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "mc_on_completed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mc_commit": (C.c_int, [C.c_void_p]),
    "mc_wait": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "mc_cmdbuf_size": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t)]),
    "mc_cmdbuf_release": (C.c_int, [C.c_void_p]),
    "mc_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "mc_llama_create": (C.c_int, [C.c_void_p, C.POINTER(LlamaConfig), C.POINTER(C.c_void_p)]),
    "mc_llama_destroy": (C.c_int, [C.c_void_p]),
    "mc_llama_set_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]),
    "mc_llama_get_config": (C.c_int, [C.c_void_p, C.POINTER(LlamaConfig)]),
    "mc_safetensors_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "mc_safetensors_close": (C.c_int, [C.c_void_p]),
    "mc_safetensors_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32)]),
    "mc_safetensors_entry_at": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(SafetensorsEntry)]),
    "mc_safetensors_find": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(SafetensorsEntry)]),
    "mc_safetensors_metadata": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p)]),
    "mc_llama_load_safetensors": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]),
    "mc_llama_init_random": (C.c_int, [C.c_void_p, C.c_uint64]),
    "mc_llama_finalize": (C.c_int, [C.c_void_p]),
    "mc_llama_weight_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "mc_llama_prefill": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]),
    "mc_llama_decode": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(SamplerConfig), C.c_void_p]),
    "mc_llama_decode_loop": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                       C.POINTER(SamplerConfig), C.c_void_p, C.POINTER(C.c_float)]),
    "mc_llama_logits": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]),
    "mc_llama_hidden": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]),
    "mc_llama_cache": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_void_p, C.c_size_t]),
    "mc_llama_launches_per_step": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32)]),
    "mc_llama_tp_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mc_llama_tp_connect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mc_nccl_unique_id": (C.c_int, [C.c_void_p, C.c_size_t]),
    "mc_llama_tp_use_nccl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mc_llama_profile_step": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]),
    "mc_sample_default": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(SamplerConfig), C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "mc_linear_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]),
    "mc_gemm_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32,
                               C.POINTER(C.c_float)]),
    "mc_attn_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                 C.c_uint32, C.c_uint32, C.c_int]),
    "mc_attn_prefill": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                  C.c_uint32, C.c_uint32, C.c_uint32]),
    "mc_embed_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]),
    "mc_linear_w4": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]),
    "mc_pack_w4": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]),
    "mc_w4_sizes": (C.c_int, [C.c_uint32, C.c_uint32, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "mc_unpack_w4": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]),
}

_lib = None


def lib():
    """Loads libmc_cuda.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise McError(MC_ERR_RUNTIME, f"{LIB_PATH} is missing: run `python -m metalchat_b200.build` "
                                          "(the CUDA backend has no CPU fallback)")
        l = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc: int):
    if rc == MC_OK:
        return
    msg = lib().mc_last_error().decode(errors="replace")
    cls = {MC_ERR_INVALID: McInvalidArgument, MC_ERR_ALLOC: McAllocError, MC_ERR_NOT_FOUND: McNotFound}.get(rc, McError)
    raise cls(rc, msg)


def _vp(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    return a


class Device:
    """mc_device: one GPU + one in-order stream (hardware_accelerator's device + queue)."""

    def __init__(self, ordinal: int = 0):
        h = C.c_void_p()
        check(lib().mc_device_create(ordinal, C.byref(h)))
        self.h = h
        self.ordinal = ordinal

    @staticmethod
    def count() -> int:
        n = C.c_int()
        check(lib().mc_device_count(C.byref(n)))
        return n.value

    def name(self) -> str:
        buf = C.create_string_buffer(256)
        check(lib().mc_device_name(self.h, buf, 256))
        return buf.value.decode()

    def sm_count(self) -> int:
        n = C.c_int()
        check(lib().mc_device_sm_count(self.h, C.byref(n)))
        return n.value

    def max_buffer(self) -> int:
        n = C.c_size_t()
        check(lib().mc_device_max_buffer(self.h, C.byref(n)))
        return n.value

    def synchronize(self):
        check(lib().mc_device_synchronize(self.h))

    def stream(self) -> int:
        s = C.c_void_p()
        check(lib().mc_device_stream(self.h, C.byref(s)))
        return s.value or 0

    def launches(self) -> int:
        n = C.c_uint64()
        check(lib().mc_launch_count(self.h, C.byref(n)))
        return n.value

    def alloc(self, size: int, flags: int = MEM_DEVICE) -> "Buffer":
        h = C.c_void_p()
        check(lib().mc_alloc(self.h, size, flags, C.byref(h)))
        return Buffer(self, h)

    def upload(self, a: np.ndarray, flags: int = MEM_DEVICE) -> "Buffer":
        a = np.ascontiguousarray(a)
        h = C.c_void_p()
        check(lib().mc_alloc_copy(self.h, _vp(a), a.nbytes, flags, C.byref(h)))
        return Buffer(self, h)

    def kernel(self, name: str) -> "Kernel":
        h = C.c_void_p()
        check(lib().mc_kernel_lookup(self.h, name.encode(), C.byref(h)))
        return Kernel(h, name)

    def command_buffer(self, capacity: int = 64) -> "CommandBuffer":
        h = C.c_void_p()
        check(lib().mc_stream_begin(self.h, capacity, C.byref(h)))
        return CommandBuffer(self, h)

    def close(self):
        if self.h:
            lib().mc_device_destroy(self.h)
            self.h = None


class Buffer:
    def __init__(self, dev: Device, h):
        self.dev, self.h = dev, h

    @property
    def size(self) -> int:
        n = C.c_size_t()
        check(lib().mc_buffer_size(self.h, C.byref(n)))
        return n.value

    def dev_ptr(self) -> int:
        p = C.c_void_p()
        check(lib().mc_buffer_dev_ptr(self.h, C.byref(p)))
        return p.value or 0

    def host_ptr(self) -> int:
        p = C.c_void_p()
        check(lib().mc_buffer_host_ptr(self.h, C.byref(p)))
        return p.value or 0

    def write(self, a: np.ndarray, offset: int = 0):
        a = np.ascontiguousarray(a)
        check(lib().mc_memcpy_h2d(self.dev.h, self.h, offset, _vp(a), a.nbytes))

    def read(self, dtype, count: int | None = None, offset: int = 0) -> np.ndarray:
        dt = np.dtype(dtype)
        n = (self.size - offset) // dt.itemsize if count is None else count
        out = np.empty(n, dtype=dt)
        check(lib().mc_memcpy_d2h(self.dev.h, _vp(out), self.h, offset, out.nbytes))
        return out

    def release(self):
        if self.h:
            lib().mc_buffer_release(self.h)
            self.h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Kernel:
    def __init__(self, h, name):
        self.h, self.name = h, name

    def max_threads(self) -> int:
        n = C.c_size_t()
        check(lib().mc_kernel_max_threads(self.h, C.byref(n)))
        return n.value

    def release(self):
        if self.h:
            lib().mc_kernel_release(self.h)
            self.h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class CommandBuffer:
    """mc_cmdbuf: the kernel_thread of the reference (kernel_thread.h:57-294)."""

    def __init__(self, dev: Device, h):
        self.dev, self.h = dev, h
        self._keep = []

    def set_bytes(self, index: int, data: bytes):
        check(lib().mc_set_bytes(self.h, index, data, len(data)))

    def set_buffer(self, index: int, buf: Buffer, offset: int = 0):
        check(lib().mc_set_buffer(self.h, index, buf.h, offset))

    def dispatch(self, kernel: Kernel, grid, group):
        g = (C.c_uint32 * 3)(*grid)
        t = (C.c_uint32 * 3)(*group)
        check(lib().mc_dispatch(self.h, kernel.h, g, t))

    def commit(self):
        check(lib().mc_commit(self.h))

    def wait(self):
        err = C.create_string_buffer(512)
        rc = lib().mc_wait(self.h, err, 512)
        if rc != MC_OK:
            raise McError(rc, err.value.decode(errors="replace"))

    def size(self) -> int:
        n = C.c_size_t()
        check(lib().mc_cmdbuf_size(self.h, C.byref(n)))
        return n.value

    def release(self):
        if self.h:
            lib().mc_cmdbuf_release(self.h)
            self.h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


_ST_NUMPY = {"BOOL": np.bool_, "I8": np.int8, "U8": np.uint8, "I16": np.int16, "U16": np.uint16, "F16": np.float16, "BF16": np.uint16, "I32": np.int32,
             "U32": np.uint32, "F32": np.float32, "F64": np.float64, "I64": np.int64, "U64": np.uint64}


class Safetensors:
    """A safetensors file (or a directory of shards) mapped read-only by the library (mc_safetensors_*)."""

    def __init__(self, path):
        h = C.c_void_p()
        check(lib().mc_safetensors_open(str(path).encode(), C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            check(lib().mc_safetensors_close(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        n = C.c_uint32()
        check(lib().mc_safetensors_count(self.h, C.byref(n)))
        return n.value

    @staticmethod
    def _unpack(e):
        return {"name": e.name.decode(), "dtype": e.dtype.decode(), "shape": tuple(int(e.shape[i]) for i in range(e.rank)), "nbytes": int(e.nbytes), "data": e.data}

    def entries(self):
        out = []
        for i in range(len(self)):
            e = SafetensorsEntry()
            check(lib().mc_safetensors_entry_at(self.h, i, C.byref(e)))
            out.append(self._unpack(e))
        return out

    def tensor(self, name: str) -> np.ndarray:
        """A copy of one tensor (BF16 comes back as uint16 bit patterns)."""
        e = SafetensorsEntry()
        check(lib().mc_safetensors_find(self.h, name.encode(), C.byref(e)))
        d = self._unpack(e)
        raw = C.string_at(d["data"], d["nbytes"]) if d["nbytes"] else b""
        return np.frombuffer(raw, dtype=_ST_NUMPY[d["dtype"]]).reshape(d["shape"]).copy()

    def metadata(self, key: str):
        v = C.c_char_p()
        check(lib().mc_safetensors_metadata(self.h, key.encode(), C.byref(v)))
        return v.value.decode() if v.value is not None else None


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(lib().mc_nccl_unique_id(buf, 128))
    return buf.raw


class Llama:
    """mc_llama: the fused decode engine (transformer<llama3>::transform, transformer.h:357-364)."""

    def __init__(self, dev: Device, cfg: LlamaConfig):
        self.dev, self.cfg = dev, cfg
        h = C.c_void_p()
        check(lib().mc_llama_create(dev.h, C.byref(cfg), C.byref(h)))
        self.h = h

    def set_tensor(self, name: str, a: np.ndarray):
        a = np.ascontiguousarray(a)
        check(lib().mc_llama_set_tensor(self.h, name.encode(), _vp(a), a.nbytes))

    def init_random(self, seed: int = 0x5EED):
        check(lib().mc_llama_init_random(self.h, seed))

    def config(self) -> LlamaConfig:
        cfg = LlamaConfig()
        check(lib().mc_llama_get_config(self.h, C.byref(cfg)))
        return cfg

    def load_safetensors(self, st: "Safetensors", flags: int = LOAD_STRICT) -> int:
        """safetensors -> device (the loader behind safetensor_document::open / load, src/safetensor.cc:145-153); returns the
        number of parameters loaded.  Call finalize() afterwards."""
        n = C.c_uint32()
        check(lib().mc_llama_load_safetensors(self.h, st.h, flags, C.byref(n)))
        return n.value

    def finalize(self):
        check(lib().mc_llama_finalize(self.h))

    def weight_bytes(self):
        s, r = C.c_uint64(), C.c_uint64()
        check(lib().mc_llama_weight_bytes(self.h, C.byref(s), C.byref(r)))
        return s.value, r.value

    def prefill(self, ids, start_pos: int = 0, seq: int = 0):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        check(lib().mc_llama_prefill(self.h, seq, _vp(ids), len(ids), start_pos))

    def decode(self, ids, pos, uniforms=None, sampler: SamplerConfig | None = None) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        pos = np.ascontiguousarray(pos, dtype=np.int32)
        out = np.zeros(len(ids), dtype=np.int32)
        u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float32)
        check(lib().mc_llama_decode(self.h, len(ids), _vp(ids), _vp(pos), _vp(u), C.byref(sampler) if sampler else None, _vp(out)))
        return out

    def decode_loop(self, first_ids, first_pos, steps: int, uniforms=None, sampler: SamplerConfig | None = None):
        ids = np.ascontiguousarray(first_ids, dtype=np.int32)
        pos = np.ascontiguousarray(first_pos, dtype=np.int32)
        out = np.zeros((steps, len(ids)), dtype=np.int32)
        ms = C.c_float()
        u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float32)
        check(lib().mc_llama_decode_loop(self.h, len(ids), _vp(ids), _vp(pos), steps, _vp(u), C.byref(sampler) if sampler else None,
                                         _vp(out), C.byref(ms)))
        return out, ms.value

    def logits(self, seq: int = 0) -> np.ndarray:
        vl = self.cfg.vocab // max(1, self.cfg.tp_world)
        out = np.zeros(vl, dtype=np.uint16)
        check(lib().mc_llama_logits(self.h, seq, _vp(out), out.nbytes))
        return out

    def hidden(self, seq: int = 0) -> np.ndarray:
        out = np.zeros(self.cfg.dim, dtype=np.uint16)
        check(lib().mc_llama_hidden(self.h, seq, _vp(out), out.nbytes))
        return out

    def cache(self, seq: int, layer: int, which: int, n_pos: int) -> np.ndarray:
        kvl = self.cfg.n_kv_heads // max(1, self.cfg.tp_world)
        out = np.zeros((n_pos, kvl, self.cfg.head_dim), dtype=np.uint16)
        check(lib().mc_llama_cache(self.h, seq, layer, which, n_pos, _vp(out), out.nbytes))
        return out

    def profile_step(self, n: int = 1) -> np.ndarray:
        us = np.zeros(1 << 18, dtype=np.float32)
        cnt = C.c_uint32()
        check(lib().mc_llama_profile_step(self.h, n, _vp(us), len(us), C.byref(cnt)))
        return us[: cnt.value]

    def tp_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        check(lib().mc_llama_tp_export(self.h, buf, 64))
        return buf.raw

    def tp_connect(self, handles: list[bytes]):
        blob = b"".join(handles)
        check(lib().mc_llama_tp_connect(self.h, blob, len(blob)))

    def tp_use_nccl(self, unique_id: bytes):
        """Comparator: the block all-reduces as ncclAllReduce calls between the per-op kernels (collective call, every rank)."""
        check(lib().mc_llama_tp_use_nccl(self.h, unique_id, len(unique_id)))

    def launches_per_step(self) -> int:
        n = C.c_uint32()
        check(lib().mc_llama_launches_per_step(self.h, C.byref(n)))
        return n.value

    def close(self):
        if self.h:
            lib().mc_llama_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def linear_bf16(dev: Device, y: Buffer, x: Buffer, w: Buffer, M: int, N: int, K: int):
    check(lib().mc_linear_bf16(dev.h, y.h, x.h, w.h, M, N, K))


GEMM_STORE, GEMM_RESIDUAL, GEMM_SWIGLU = 0, 2, 3


def gemm_bf16(dev: Device, y: Buffer, x: Buffer, w: Buffer, M: int, N: int, K: int, mode: int = GEMM_STORE, res: Buffer | None = None, iters: int = 1) -> float:
    """The tcgen05 prefill GEMM (mc_gemm_bf16): returns the CUDA-event time of `iters` launches in ms."""
    ms = C.c_float()
    check(lib().mc_gemm_bf16(dev.h, y.h, x.h, w.h, res.h if res is not None else None, M, N, K, mode, iters, C.byref(ms)))
    return ms.value


def w4_sizes(N: int, K: int):
    a, b = C.c_size_t(), C.c_size_t()
    check(lib().mc_w4_sizes(N, K, C.byref(a), C.byref(b)))
    return a.value, b.value


def pack_w4(dev: Device, q8: np.ndarray, scales: np.ndarray):
    """int8 [N,K] in [-8,7] + fp32 scales [N,K/32] -> (packed weights, packed bf16 scales) device buffers."""
    N, K = q8.shape
    wb, sb = w4_sizes(N, K)
    dq, ds = dev.upload(np.ascontiguousarray(q8, np.int8)), dev.upload(np.ascontiguousarray(scales, np.float32))
    w4, sp = dev.alloc(wb), dev.alloc(sb)
    check(lib().mc_pack_w4(dev.h, w4.h, sp.h, dq.h, ds.h, N, K))
    return w4, sp


def unpack_w4(dev: Device, w4: Buffer, N: int, K: int) -> np.ndarray:
    out = dev.alloc(N * K)
    check(lib().mc_unpack_w4(dev.h, out.h, w4.h, N, K))
    return out.read(np.int8).reshape(N, K)


def linear_w4(dev: Device, y: Buffer, x: Buffer, w4: Buffer, sp: Buffer, M: int, N: int, K: int):
    check(lib().mc_linear_w4(dev.h, y.h, x.h, w4.h, sp.h, M, N, K))


def sample_default(dev: Device, logits_bf16: np.ndarray, uniforms, top_k=50, temperature=0.6, top_p=0.9, intended=0):
    """make_default_sampler on the device; logits_bf16 is [rows, vocab] uint16."""
    logits_bf16 = np.ascontiguousarray(logits_bf16, np.uint16)
    rows, vocab = logits_bf16.shape
    buf = dev.upload(logits_bf16)
    cfg = SamplerConfig(1, top_k, temperature, top_p, intended)
    u = np.ascontiguousarray(uniforms, np.float32)
    k = min(top_k, vocab)
    topk = np.zeros((rows, k), np.int32)
    ps = np.zeros((rows, k), np.uint16)
    pi = np.zeros((rows, k), np.int32)
    ch, tok = np.zeros(rows, np.int32), np.zeros(rows, np.int32)
    check(lib().mc_sample_default(dev.h, buf.h, rows, vocab, C.byref(cfg), _vp(u), _vp(topk), _vp(ps), _vp(pi), _vp(ch), _vp(tok)))
    return dict(topk_idx=topk, probs_sorted=ps, probs_idx=pi, choice=ch, token=tok)


def _cache_to_device_layout(c: np.ndarray) -> np.ndarray:
    """[n_seqs, pos, KV, hd] (the reference's cache layout, nn/cache.h:150-160) -> the engine's [n_seqs, KV, pos, hd]."""
    return np.ascontiguousarray(np.transpose(c, (0, 2, 1, 3)))


def attn_decode(dev: Device, q: np.ndarray, kcache: np.ndarray, vcache: np.ndarray, row_seq, row_pos, kernel: int = 0) -> np.ndarray:
    """q [rows, H, hd] bf16 bits (rotated); caches [n_seqs, max_seq, KV, hd] bf16 bits -> out [rows, H, hd]."""
    rows, H, hd = q.shape
    n_seqs, max_seq, KV, _ = kcache.shape
    dq, dk, dv = dev.upload(np.ascontiguousarray(q, np.uint16)), dev.upload(_cache_to_device_layout(kcache)), dev.upload(_cache_to_device_layout(vcache))
    out = dev.alloc(rows * H * hd * 2)
    rs, rp = np.ascontiguousarray(row_seq, np.int32), np.ascontiguousarray(row_pos, np.int32)
    check(lib().mc_attn_decode(dev.h, out.h, dq.h, dk.h, dv.h, rows, _vp(rs), _vp(rp), n_seqs, H, KV, hd, max_seq, kernel))
    return out.read(np.uint16).reshape(rows, H, hd)


def attn_prefill(dev: Device, q: np.ndarray, kcache: np.ndarray, vcache: np.ndarray, start_pos: int, key_begin: int = 0, seq: int = 0) -> np.ndarray:
    rows, H, hd = q.shape
    n_seqs, max_seq, KV, _ = kcache.shape
    dq, dk, dv = dev.upload(np.ascontiguousarray(q, np.uint16)), dev.upload(_cache_to_device_layout(kcache)), dev.upload(_cache_to_device_layout(vcache))
    out = dev.alloc(rows * H * hd * 2)
    check(lib().mc_attn_prefill(dev.h, out.h, dq.h, dk.h, dv.h, rows, seq, start_pos, key_begin, n_seqs, H, KV, hd, max_seq))
    return out.read(np.uint16).reshape(rows, H, hd)


def embed_rows(dev: Device, table: np.ndarray, ids, row_scales: np.ndarray | None = None) -> np.ndarray:
    """table [vocab, D]: uint16 bf16 bits, or int8 with fp32 row_scales [vocab]."""
    vocab, D = table.shape
    ids = np.ascontiguousarray(ids, np.int32)
    dt = dev.upload(np.ascontiguousarray(table))
    ds = dev.upload(np.ascontiguousarray(row_scales, np.float32)) if row_scales is not None else None
    out = dev.alloc(len(ids) * D * 2)
    check(lib().mc_embed_rows(dev.h, out.h, dt.h, ds.h if ds is not None else None, _vp(ids), len(ids), D, vocab))
    return out.read(np.uint16).reshape(len(ids), D)
