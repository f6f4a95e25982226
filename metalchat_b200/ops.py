"""Host-side mirror of the reference's kernel wrappers (include/metalchat/kernel/*.h).

One function per wrapper class, same names and argument meaning: each computes the launch
shape the reference computes, allocates the output, binds the arguments in the reference's bind
order (tensor = tensor_layout<N> bytes + buffer, scalar = raw bytes) and dispatches the kernel
named ``<op>[_<block>]_<dtype>`` through the C ABI.  Tensors are strided device views
(sizes / strides / per-dimension offsets, tensor/basic.h:180-1187).  Everything runs on the
GPU; there is no CPU path.
"""
from __future__ import annotations

import math
import struct

import numpy as np

from . import capi

BF16, F32, I32, I8, BOOL = "bfloat", "float", "int32_t", "int8_t", "bool"
_ITEM = {BF16: 2, F32: 4, I32: 4, I8: 1, BOOL: 1}
_NP = {BF16: np.uint16, F32: np.float32, I32: np.int32, I8: np.int8, BOOL: np.uint8}
MAX_THREADS = 1024  # basic_kernel::max_threads_per_threadgroup (src/kernel.cc:75-79)


def ceil_div(a, b):
    return -(-a // b)


def ceil_pow2(x):
    p = 1
    while p < x:
        p *= 2
    return p


class Tensor:
    """A strided view over a device buffer: the tensor<T,N,hardware_memory_container> of the reference."""

    def __init__(self, buf: capi.Buffer, dtype: str, sizes, strides=None, offsets=None):
        self.buf, self.dtype = buf, dtype
        self.sizes = [int(s) for s in sizes]
        if strides is None:
            strides, acc = [], 1
            for s in reversed(self.sizes):
                strides.insert(0, acc)
                acc *= s
        self.strides = [int(s) for s in strides]
        self.offsets = [int(o) for o in (offsets or [0] * len(self.sizes))]

    @property
    def dim(self):
        return len(self.sizes)

    def numel(self):
        return int(np.prod(self.sizes)) if self.sizes else 1

    def layout_bytes(self) -> bytes:
        return struct.pack(f"{3 * self.dim}I", *self.sizes, *self.strides, *self.offsets)

    def is_contiguous(self):
        exp, acc = [], 1
        for s in reversed(self.sizes):
            exp.insert(0, acc)
            acc *= s
        return self.strides == exp and not any(self.offsets)

    def view(self, sizes) -> "Tensor":
        sizes = list(sizes)
        if -1 in sizes:
            known = int(np.prod([s for s in sizes if s != -1])) or 1
            sizes[sizes.index(-1)] = self.numel() // known
        if not self.is_contiguous():
            raise capi.McInvalidArgument(capi.MC_ERR_INVALID, "view: tensor is not contiguous")
        return Tensor(self.buf, self.dtype, sizes)

    def flatten2(self) -> "Tensor":
        return self if self.dim == 2 else self.view([-1, self.sizes[-1]])

    def narrow(self, dim, start, length) -> "Tensor":
        t = Tensor(self.buf, self.dtype, self.sizes, self.strides, self.offsets)
        t.offsets[dim] += start * t.strides[dim]
        t.sizes[dim] = length
        return t

    def transpose(self, d0, d1) -> "Tensor":
        t = Tensor(self.buf, self.dtype, self.sizes, self.strides, self.offsets)
        for arr in (t.sizes, t.strides, t.offsets):
            arr[d0], arr[d1] = arr[d1], arr[d0]
        return t

    def numpy(self) -> np.ndarray:
        """Device -> host copy of the viewed elements (row-major)."""
        raw = self.buf.read(_NP[self.dtype])
        idx = np.zeros(self.sizes, dtype=np.int64)
        for d, (n, st, off) in enumerate(zip(self.sizes, self.strides, self.offsets)):
            shape = [1] * self.dim
            shape[d] = n
            idx = idx + (np.arange(n, dtype=np.int64) * st + off).reshape(shape)
        return raw[idx]


class Accelerator:
    """hardware_accelerator (accelerator.h:55-219): device + kernel cache + current command buffer."""

    def __init__(self, ordinal: int = 0, thread_capacity: int = 64):
        self.dev = capi.Device(ordinal)
        self.capacity = thread_capacity
        self._kernels = {}
        self._cb = None

    def name(self):
        return self.dev.name()

    def load(self, name: str) -> capi.Kernel:
        if name not in self._kernels:
            self._kernels[name] = self.dev.kernel(name)
        return self._kernels[name]

    def thread(self) -> capi.CommandBuffer:
        # recursive_kernel_thread: roll to a new buffer when full or committed (src/kernel_thread.cc:213-224)
        if self._cb is None or self._cb.size() >= self.capacity:
            if self._cb is not None:
                self._cb.commit()
                self._cb.release()
            self._cb = self.dev.command_buffer(self.capacity)
        return self._cb

    def wait(self):
        if self._cb is not None:
            self._cb.wait()
            self._cb.release()
            self._cb = None

    def empty(self, dtype: str, sizes) -> Tensor:
        n = int(np.prod(sizes)) if len(sizes) else 1
        return Tensor(self.dev.alloc(max(1, n) * _ITEM[dtype]), dtype, sizes)

    def tensor(self, a: np.ndarray, dtype: str) -> Tensor:
        a = np.ascontiguousarray(a, dtype=_NP[dtype])
        return Tensor(self.dev.upload(a), dtype, a.shape)

    def dispatch(self, name: str, grid, group, *args):
        """kernel_task::encode (kernel.h:282-297): bind in order, then dispatchThreads."""
        cb = self.thread()
        k = self.load(name)
        slot = 0
        for a in args:
            if isinstance(a, Tensor):
                cb.set_bytes(slot, a.layout_bytes())
                cb.set_buffer(slot + 1, a.buf, 0)
                slot += 2
            else:
                cb.set_bytes(slot, a)
                slot += 1
        cb.dispatch(k, grid, group)


def _scalar(dtype: str, v) -> bytes:
    if dtype == BF16:
        u = np.array([v], dtype=np.float32).view(np.uint32)[0]
        # host bf16 conversion: RNE with subnormal/zero flush (dtype.h:36-57)
        f = float(np.float32(v))
        if f == 0.0 or (abs(f) < 1.1754943508222875e-38):
            return struct.pack("H", (int(u) >> 16) & 0x8000)
        if (u & 0x7FFFFFFF) > 0x7F800000:
            return struct.pack("H", ((int(u) >> 16) | 0x40) & 0xFFFF)
        return struct.pack("H", ((int(u) + 0x7FFF + ((int(u) >> 16) & 1)) >> 16) & 0xFFFF)
    if dtype == F32:
        return struct.pack("f", v)
    raise ValueError(dtype)


def make_kernel_grid_2d(t: Tensor, max_threads=MAX_THREADS):
    """src/kernel.cc:14-38 — launch shape for element-wise kernels over a [rows, dim] view."""
    rows, dim = max(t.sizes[0], 1), max(t.sizes[1], 1)
    if rows * dim <= max_threads:
        return (dim, rows, 1), (dim, rows, 1)
    if dim <= max_threads:
        return (dim, rows, 1), (dim, 1, 1)
    return (max_threads * ceil_div(dim, max_threads), rows, 1), (max_threads, 1, 1)


# ---- wrappers -----------------------------------------------------------------------------------------
def bmm(gpu: Accelerator, a: Tensor, b: Tensor) -> Tensor:
    """kernel::bmm (kernel/bmm.h:27-89): [.., M, K] x [.., K, N]."""
    a3 = a if a.dim == 3 else Tensor(a.buf, a.dtype, [1] + a.sizes, [0] + a.strides, [0] + a.offsets)
    b3 = b if b.dim == 3 else Tensor(b.buf, b.dtype, [a3.sizes[0]] + b.sizes, [0] + b.strides, [0] + b.offsets)
    if a3.sizes[2] != b3.sizes[1]:
        raise capi.McInvalidArgument(capi.MC_ERR_INVALID, "bmm: inner dimensions differ")
    B, M, N = a3.sizes[0], a3.sizes[1], b3.sizes[2]
    out = gpu.empty(a.dtype, [B, M, N])
    grid = (ceil_div(max(M, 1), 8) * 8, ceil_div(max(N, 1), 8) * 8, max(B, 1))
    gpu.dispatch(f"bmm_8_{a.dtype}", grid, (8, 8, 1), out, a3, b3)
    return out if a.dim == 3 else out.view([M, N])


def _row_partition(dim):
    block = ceil_div(max(dim, 1), MAX_THREADS)
    threads = ceil_div(max(dim, 1), block)
    return block, threads


def rmsnorm(gpu, x: Tensor, weight: Tensor, eps=1e-5, mu=0.0) -> Tensor:
    """kernel::rmsnorm (kernel/rmsnorm.h:28-55)."""
    x2 = x.flatten2()
    if weight.sizes[-1] != x2.sizes[1]:
        raise capi.McInvalidArgument(capi.MC_ERR_INVALID, "rmsnorm: weight size differs from the last dimension")
    out = gpu.empty(x.dtype, x2.sizes)
    block, threads = _row_partition(x2.sizes[1])
    rows = max(x2.sizes[0], 1)
    gpu.dispatch(f"rmsnorm_{x.dtype}", (threads * rows, 1, 1), (threads, 1, 1), out, x2, weight,
                 struct.pack("f", eps), struct.pack("f", mu), struct.pack("I", block))
    return out.view(x.sizes)


def softmax(gpu, x: Tensor) -> Tensor:
    """kernel::softmax (kernel/softmax.h:27-51)."""
    x2 = x.flatten2()
    out = gpu.empty(x.dtype, x2.sizes)
    block, threads = _row_partition(x2.sizes[1])
    rows = max(x2.sizes[0], 1)
    gpu.dispatch(f"softmax_{x.dtype}", (threads * rows, 1, 1), (threads, 1, 1), out, x2, struct.pack("I", block))
    return out.view(x.sizes)


def sum_(gpu, x: Tensor) -> Tensor:
    """kernel::sum (kernel/sum.h:75-115)."""
    x2 = x.flatten2()
    out = gpu.empty(x.dtype, [x2.sizes[0]])
    block, threads = _row_partition(x2.sizes[1])
    rows = max(x2.sizes[0], 1)
    gpu.dispatch(f"sum_{x.dtype}", (threads * rows, 1, 1), (threads, 1, 1), out, x2, struct.pack("I", block))
    return out


def cumsum(gpu, x: Tensor) -> Tensor:
    """kernel::cumsum (kernel/sum.h:27-61): block = max(2, pow2(ceil(dim / max_threads)))."""
    x2 = x.flatten2()
    dim = x2.sizes[1]
    block = max(2, ceil_pow2(ceil_div(max(dim, 1), MAX_THREADS)))
    threads = ceil_div(max(dim, 1), block)
    out = gpu.empty(x.dtype, x2.sizes)
    rows = max(x2.sizes[0], 1)
    gpu.dispatch(f"cumsum_{block}_{x.dtype}", (threads * rows, 1, 1), (threads, 1, 1), out, x2)
    return out.view(x.sizes)


def sort(gpu, x: Tensor):
    """kernel::sort (kernel/sort.h:25-62): descending, returns (values, indices) sliced to the input size."""
    x2 = x.flatten2()
    rows, dim = x2.sizes
    P = ceil_pow2(max(dim, 1))
    values = gpu.empty(x.dtype, [rows, P])
    indices = gpu.empty(I32, [rows, P])
    block = ceil_div(P, MAX_THREADS)
    threads = ceil_div(P, block)
    gpu.dispatch(f"sort_{x.dtype}", (threads * max(rows, 1), 1, 1), (threads, 1, 1), values, indices, x2, struct.pack("I", block))
    return values.narrow(1, 0, dim), indices.narrow(1, 0, dim)


def embedding(gpu, ids: Tensor, weight: Tensor) -> Tensor:
    """kernel::embedding (kernel/embedding.h:31-71)."""
    out = gpu.empty(weight.dtype, [ids.sizes[0], ids.sizes[1], weight.sizes[1]])
    dim_size, emb = ids.sizes[1], weight.sizes[1]
    q = dim_size / float(emb + dim_size) if (emb + dim_size) else 0.0
    mtx = max(1, int(math.sqrt(MAX_THREADS) * q))
    mty = max(1, MAX_THREADS // mtx)
    bx, by = ceil_div(max(dim_size, 1), mtx), ceil_div(max(emb, 1), mty)
    tx, ty = ceil_div(max(dim_size, 1), bx), ceil_div(max(emb, 1), by)
    grid = (tx * ceil_div(max(dim_size, 1), tx), ty * ceil_div(max(emb, 1), ty), max(ids.sizes[0], 1))
    gpu.dispatch(f"embedding_{weight.dtype}", grid, (tx, ty, 1), out, ids, weight, struct.pack("I", bx))
    return out


def rope_freqs(gpu, seq_len: int, dim: int, start_pos: int, theta: float):
    """kernel::rope_freqs (kernel/embedding.h:131-166): fp32 cos/sin tables [seq_len, dim/2]."""
    fcos, fsin = gpu.empty(F32, [seq_len, dim // 2]), gpu.empty(F32, [seq_len, dim // 2])
    grid, group = make_kernel_grid_2d(fcos)
    gpu.dispatch("rope_freqs_float", grid, group, fcos, fsin, struct.pack("I", dim), struct.pack("I", start_pos), struct.pack("f", theta))
    return fcos, fsin


def rope(gpu, x: Tensor, fcos: Tensor, fsin: Tensor, start_pos: int) -> Tensor:
    """kernel::rope (kernel/embedding.h:87-125): x is [bs, len, n_head, head_dim]."""
    bs, n_head = x.sizes[0], x.sizes[2]
    x2 = x.flatten2()
    out = gpu.empty(x.dtype, x2.sizes)
    grid, group = make_kernel_grid_2d(x2)
    gpu.dispatch(f"rope_{x.dtype}", grid, group, out, x2, fcos, fsin, struct.pack("I", bs), struct.pack("I", n_head), struct.pack("i", start_pos))
    return out.view(x.sizes)


def multinomial(gpu, x: Tensor, sample_size: int, init_state: int, init_seq: int) -> Tensor:
    """kernel::multinomial (kernel/multinomial.h:24-60); the reference draws the two seeds from mt19937."""
    x2 = x.flatten2()
    out = gpu.empty(I32, [x2.sizes[0], sample_size])
    grid, group = make_kernel_grid_2d(out)
    gpu.dispatch(f"multinomial_{x.dtype}", grid, group, out, x2, struct.pack("Q", init_state), struct.pack("Q", init_seq))
    return out


def _binary(name):
    def op(gpu, a: Tensor, b: Tensor) -> Tensor:
        a2, b2 = a.flatten2(), b.flatten2()
        if a2.sizes != b2.sizes:
            raise capi.McInvalidArgument(capi.MC_ERR_INVALID, f"{name}: operand shapes differ")
        out = gpu.empty(a.dtype, a2.sizes)
        grid, group = make_kernel_grid_2d(a2)
        gpu.dispatch(f"{name}_{a.dtype}", grid, group, out, a2, b2)
        return out.view(a.sizes)

    op.__doc__ = f"kernel::{name} via binary_kernel_wrapper (kernel.h:301-400)."
    return op


add, sub, div, hadamard = _binary("add"), _binary("sub"), _binary("div"), _binary("hadamard")


def add_broadcast(gpu, a: Tensor, b: Tensor) -> Tensor:
    """kernel::add_broadcast (kernel/arithmetic.h:41-78): b is flattened and indexed j mod numel."""
    a2 = a.view([-1, b.numel()]) if a.sizes[-1] != b.numel() else a.flatten2()
    b1 = b.view([-1])
    out = gpu.empty(a.dtype, a2.sizes)
    grid, group = make_kernel_grid_2d(a2)
    gpu.dispatch(f"add_broadcast_{a.dtype}", grid, group, out, a2, b1)
    return out.view(a.sizes)


def hadamard_broadcast(gpu, out_dtype: str, a: Tensor, b: Tensor) -> Tensor:
    """kernel::hadamard_broadcast<T, int8, S> (kernel/mul.h:37-75): the dequantisation kernel."""
    a2, b1 = a.flatten2(), b.view([-1])
    if a2.sizes[0] != b1.sizes[0]:
        raise capi.McInvalidArgument(capi.MC_ERR_INVALID, "hadamard_broadcast: first dimensions differ")  # same_first_dim, mul.h:50
    out = gpu.empty(out_dtype, a2.sizes)
    grid, group = make_kernel_grid_2d(a2)
    gpu.dispatch(f"hadamard_broadcast_{out_dtype}_int8_t_{b.dtype}", grid, group, out, a2, b1)
    return out.view(a.sizes)


def scalar_mul(gpu, a: Tensor, c: float) -> Tensor:
    a2 = a.flatten2()
    out = gpu.empty(a.dtype, a2.sizes)
    grid, group = make_kernel_grid_2d(a2)
    gpu.dispatch(f"scalar_mul_{a.dtype}", grid, group, out, a2, _scalar(a.dtype, c))
    return out.view(a.sizes)


def _unary(name):
    def op(gpu, a: Tensor) -> Tensor:
        a2 = a.flatten2()
        out = gpu.empty(a.dtype, a2.sizes)
        grid, group = make_kernel_grid_2d(a2)
        gpu.dispatch(f"{name}_{a.dtype}", grid, group, out, a2)
        return out.view(a.sizes)

    return op


silu, gelu = _unary("silu"), _unary("gelu")


def clone(gpu, a: Tensor, out: Tensor | None = None) -> Tensor:
    """kernel::clone (kernel/copy.h:30-81), optionally into a (strided) target view."""
    a2 = a if a.dim == 2 else Tensor(a.buf, a.dtype, [int(np.prod(a.sizes[:-1]))] + [a.sizes[-1]]) if a.is_contiguous() else a
    if out is None:
        out = gpu.empty(a.dtype, a.sizes)
    o2 = out if out.dim == 2 else out.flatten2()
    grid, group = make_kernel_grid_2d(a2)
    gpu.dispatch(f"copy_{a.dtype}", grid, group, o2, a2)
    return out


def scatter(gpu, out: Tensor, mask: Tensor, value: float) -> Tensor:
    o2, m2 = out.flatten2(), mask.flatten2()
    grid, group = make_kernel_grid_2d(o2)
    gpu.dispatch(f"scatter_{out.dtype}", grid, group, o2, m2, _scalar(out.dtype, value))
    return out


def gather(gpu, a: Tensor, index: Tensor) -> Tensor:
    a2, i2 = a.flatten2(), index.flatten2()
    out = gpu.empty(a.dtype, i2.sizes)
    grid, group = make_kernel_grid_2d(i2)
    gpu.dispatch(f"gather_{a.dtype}", grid, group, out, a2, i2)
    return out.view(index.sizes)


def _compare(name):
    def op(gpu, a: Tensor, value: float) -> Tensor:
        a2 = a.flatten2()
        out = gpu.empty(BOOL, a2.sizes)
        grid, group = make_kernel_grid_2d(a2)
        gpu.dispatch(f"{name}_{a.dtype}", grid, group, out, a2, _scalar(a.dtype, value))
        return out.view(a.sizes)

    return op


gt, le = _compare("gt"), _compare("le")


def roll(gpu, a: Tensor, shift: int, dim: int) -> Tensor:
    """kernel::roll (kernel/roll.h:30-88) on a contiguous tensor."""
    size = a.sizes[dim]
    stride = a.strides[dim]
    flat = a.view([-1])
    out = gpu.empty(a.dtype, flat.sizes)
    n = max(flat.sizes[0], 1)
    threads = min(n, MAX_THREADS)
    gpu.dispatch(f"roll_{a.dtype}", (ceil_div(n, threads) * threads, 1, 1), (threads, 1, 1), out, flat,
                 struct.pack("i", shift % size if size else 0), struct.pack("i", size), struct.pack("i", stride))
    return out.view(a.sizes)
