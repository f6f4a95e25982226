"""GPU parity of the op-level kernels (called through the C ABI by their reference names) against
the CPU oracle on seeded inputs, plus the reference's own known-answer tests.
Integer / index / byte work must be bit-exact; floating point within the stated tolerance."""
import numpy as np
import pytest

from oracle import orc
from oracle.orc import BF16, F32
from tests.gpu_util import accelerator, bf, unbf

pytestmark = pytest.mark.gpu

DT = {"bfloat": BF16, "float": F32}


def host(rng, shape, dtype, scale=1.0):
    a = (rng.standard_normal(shape) * scale).astype(np.float32)
    return bf(a) if dtype == "bfloat" else a


def as_f32(a, dtype):
    return unbf(a) if dtype == "bfloat" else a


def ulp_close(got, want, dtype, ulps=1):
    """bf16: at most `ulps` bf16 steps apart; float: rtol 1e-5."""
    if dtype == "bfloat":
        g, w = got.astype(np.int32), want.astype(np.int32)
        return np.all(np.abs(g - w) <= ulps)
    return np.allclose(got, want, rtol=2e-5, atol=1e-6)


@pytest.fixture()
def gpu():
    g = accelerator()
    yield g
    g.wait()


# ---- reference KATs ---------------------------------------------------------------------------------
def test_softmax_golden_vector(gpu):
    # test/test_kernel_softmax.cc:19-39
    from metalchat_b200 import ops

    x = gpu.tensor(bf(np.arange(5, dtype=np.float32)).reshape(1, 5), "bfloat")
    out = ops.softmax(gpu, x)
    gpu.wait()
    expect = np.array([0.0116577, 0.0317383, 0.0859375, 0.234375, 0.636719], dtype=np.float32)
    assert np.allclose(unbf(out.numpy())[0], expect, atol=1e-5)


def test_rmsnorm_ones(gpu):
    # test/test_kernel_rmsnorm.cc:18-37
    from metalchat_b200 import ops

    x = gpu.tensor(bf(np.ones((4 * 3 * 5, 7), dtype=np.float32)), "bfloat")
    w = gpu.tensor(bf(np.full(7, 3.0, dtype=np.float32)), "bfloat")
    out = ops.rmsnorm(gpu, x, w)
    gpu.wait()
    assert np.all(unbf(out.numpy()) == 3.0)


def test_dequant_kat(gpu, rng):
    # test/test_kernel_mul.cc:41-65: <float, int8, float> [512,64,32] * [512,64,1]
    from metalchat_b200 import ops

    q = rng.integers(-128, 128, size=(512 * 64, 32), dtype=np.int8)
    s = rng.random(512 * 64, dtype=np.float32)
    out = ops.hadamard_broadcast(gpu, "float", gpu.tensor(q, "int8_t"), gpu.tensor(s, "float"))
    gpu.wait()
    assert np.allclose(out.numpy(), q.astype(np.float32) * s[:, None], atol=1e-5)


def test_kernel_thread_chain(gpu):
    # test/test_kernel_thread.cc:16-40: three chained adds give 8.0 (stream ordering)
    from metalchat_b200 import ops

    a = gpu.tensor(np.ones((3, 4), np.float32), "float")
    b = ops.add(gpu, a, a)
    c = ops.add(gpu, b, b)
    d = ops.add(gpu, c, c)
    gpu.wait()
    assert np.all(d.numpy() == 8.0)


def test_kernel_not_found_and_bad_launch(gpu):
    from metalchat_b200 import capi

    with pytest.raises(capi.McNotFound):
        gpu.dev.kernel("no_such_kernel_bfloat")
    cb = gpu.dev.command_buffer(4)
    k = gpu.dev.kernel("add_float")
    with pytest.raises(capi.McInvalidArgument, match="exceeds maximum number of threads"):
        cb.dispatch(k, (2048, 1, 1), (2048, 1, 1))  # kernel.h:126-133
    with pytest.raises(capi.McInvalidArgument, match="less threads in grid"):
        cb.dispatch(k, (16, 1, 1), (32, 1, 1))  # kernel.h:134-140
    with pytest.raises(capi.McInvalidArgument):
        cb.dispatch(k, (32, 1, 1), (32, 1, 1))  # nothing bound
    cb.release()


def test_command_buffer_capacity(gpu):
    from metalchat_b200 import capi, ops

    a = gpu.tensor(np.ones((2, 2), np.float32), "float")
    out = gpu.empty("float", [2, 2])
    cb = gpu.dev.command_buffer(2)
    k = gpu.dev.kernel("add_float")
    for _ in range(2):
        for i, t in enumerate((out, a, a)):
            cb.set_bytes(2 * i, t.layout_bytes())
            cb.set_buffer(2 * i + 1, t.buf)
        cb.dispatch(k, (2, 2, 1), (2, 2, 1))
    with pytest.raises(capi.McError) as e:
        cb.dispatch(k, (2, 2, 1), (2, 2, 1))
    assert e.value.code == capi.MC_ERR_FULL
    cb.commit()
    cb.wait()
    assert np.all(out.numpy() == 2.0)


def test_memory_kinds(gpu):
    from metalchat_b200 import capi

    d = gpu.dev.alloc(1024, capi.MEM_DEVICE)
    with pytest.raises(capi.McInvalidArgument):
        d.host_ptr()
    s = gpu.dev.alloc(1024, capi.MEM_SHARED)
    assert s.host_ptr() != 0 and s.size == 1024
    with pytest.raises(capi.McAllocError):
        gpu.dev.alloc(1 << 50, capi.MEM_DEVICE)


# ---- oracle parity per kernel -------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["bfloat", "float"])
@pytest.mark.parametrize("shape", [(1, 5, 64, 48), (2, 3, 17, 9), (1, 1, 2048, 40)])
def test_bmm(gpu, rng, dtype, shape):
    from metalchat_b200 import ops

    B, M, K, N = shape
    a, b = host(rng, (B, M, K), dtype), host(rng, (B, N, K), dtype)  # b is used through a transposed view like nn::linear
    ta, tb = gpu.tensor(a, dtype), gpu.tensor(b, dtype).transpose(1, 2)
    out = ops.bmm(gpu, ta, tb)
    gpu.wait()
    want = np.zeros((B, M, N), dtype=a.dtype)
    orc.bmm(DT[dtype], want, a, b.transpose(0, 2, 1))
    assert np.array_equal(out.numpy(), want)  # same ascending-k fp32 chain: bit-exact


@pytest.mark.parametrize("dtype", ["bfloat", "float"])
@pytest.mark.parametrize("shape", [(3, 7), (15, 2048), (2, 5000)])
def test_rmsnorm_softmax_sum(gpu, rng, dtype, shape):
    from metalchat_b200 import ops

    x = host(rng, shape, dtype)
    w = host(rng, (shape[1],), dtype)
    tx, tw = gpu.tensor(x, dtype), gpu.tensor(w, dtype)
    got_n, got_s, got_sum = ops.rmsnorm(gpu, tx, tw, 1e-5), ops.softmax(gpu, tx), ops.sum_(gpu, tx)
    gpu.wait()
    want = np.zeros_like(x)
    orc.rmsnorm(DT[dtype], want, x, w, eps=1e-5)
    assert np.array_equal(got_n.numpy(), want)  # same partition, IEEE sqrt/div: bit-exact
    orc.softmax(DT[dtype], want, x)
    assert ulp_close(got_s.numpy(), want, dtype)  # expf differs from libm by <= 2 ulp fp32
    ws = np.zeros(shape[0], dtype=x.dtype)
    orc.rowsum(DT[dtype], ws, x)
    assert np.array_equal(got_sum.numpy(), ws)


@pytest.mark.parametrize("dtype", ["bfloat", "float"])
def test_rope_and_freqs(gpu, rng, dtype):
    # rope_freqs KAT: test/test_kernel_embedding.cc:72-137 (theta 5e5, dim 64, start 100, abs 1e-4)
    from metalchat_b200 import ops

    fc, fs = ops.rope_freqs(gpu, 32, 64, 100, 500000.0)
    gpu.wait()
    wc, wsn = np.zeros((32, 32), np.float32), np.zeros((32, 32), np.float32)
    orc.rope_freqs(wc, wsn, 64, 100, 500000.0)
    assert np.allclose(fc.numpy(), wc, atol=1e-4) and np.allclose(fs.numpy(), wsn, atol=1e-4)
    # rotation against the oracle using the oracle's tables
    bs, ln, nh, hd = 1, 6, 4, 64
    x = host(rng, (bs, ln, nh, hd), dtype)
    tc, ts = gpu.tensor(wc, "float"), gpu.tensor(wsn, "float")
    out = ops.rope(gpu, gpu.tensor(x, dtype), tc, ts, 3)
    gpu.wait()
    want = np.zeros((ln * nh, hd), dtype=x.dtype)
    orc.rope(DT[dtype], want, x.reshape(ln * nh, hd), wc, wsn, bs, nh, 3)
    assert np.array_equal(out.numpy().reshape(ln * nh, hd), want)


@pytest.mark.parametrize("dtype", ["bfloat", "float"])
def test_embedding_exact(gpu, rng, dtype):
    # test/test_kernel_embedding.cc:19-53
    from metalchat_b200 import ops

    w = host(rng, (1000, 256), dtype)
    ids = rng.integers(0, 1000, size=(3, 11), dtype=np.int32)
    out = ops.embedding(gpu, gpu.tensor(ids, "int32_t"), gpu.tensor(w, dtype))
    gpu.wait()
    assert np.array_equal(out.numpy(), w[ids])


@pytest.mark.parametrize("dtype", ["bfloat", "float"])
@pytest.mark.parametrize("n", [1, 5, 50, 64, 1000, 5000])
def test_sort_bit_exact(gpu, rng, dtype, n):
    from metalchat_b200 import ops

    x = host(rng, (3, n), dtype)
    x[0, : n // 2] = x[0, 0]  # ties: the network fixes their order (quirk Q12)
    v, i = ops.sort(gpu, gpu.tensor(x, dtype))
    gpu.wait()
    P = ops.ceil_pow2(n)
    wv, wi = np.zeros((3, P), dtype=x.dtype), np.zeros((3, P), dtype=np.int32)
    orc.sort(DT[dtype], wv, wi, x)
    assert np.array_equal(v.numpy(), wv[:, :n]) and np.array_equal(i.numpy(), wi[:, :n])


def test_sort_100k(gpu, rng):
    # test/test_kernel_sort.cc:17-52
    from metalchat_b200 import ops

    x = rng.standard_normal((1, 100000)).astype(np.float32)
    v, i = ops.sort(gpu, gpu.tensor(x, "float"))
    gpu.wait()
    vv, ii = v.numpy()[0], i.numpy()[0]
    assert np.all(vv[:-1] >= vv[1:]) and np.array_equal(x[0][ii], vv) and len(set(ii.tolist())) == 100000


@pytest.mark.parametrize("dtype", ["bfloat", "float"])
@pytest.mark.parametrize("n", [1, 7, 50, 1024, 3000])
def test_cumsum_bit_exact(gpu, rng, dtype, n):
    from metalchat_b200 import ops

    x = host(rng, (2, n), dtype)
    x = np.abs(as_f32(x, dtype)).astype(np.float32)
    x = bf(x) if dtype == "bfloat" else x
    out = ops.cumsum(gpu, gpu.tensor(x, dtype))
    gpu.wait()
    want = np.zeros_like(x)
    orc.cumsum(DT[dtype], want, x)
    assert np.array_equal(out.numpy(), want)


@pytest.mark.parametrize("dtype", ["bfloat", "float"])
def test_multinomial_bit_exact(gpu, rng, dtype):
    from metalchat_b200 import ops

    p = np.sort(rng.random((6, 40)).astype(np.float32), axis=1)[:, ::-1].copy()
    x = bf(p) if dtype == "bfloat" else p
    out = ops.multinomial(gpu, gpu.tensor(x, dtype), 8, 12345, 678)
    gpu.wait()
    want = np.zeros((6, 8), np.int32)
    orc.multinomial(DT[dtype], want, x, 12345, 678)
    assert np.array_equal(out.numpy(), want)


def test_multinomial_frequencies(gpu):
    # test/test_kernel_multinomial.cc:16-54 reads a [4,5] input with 8192 samples; here the row is widened
    # to the sample count so that the kernel stays in bounds (quirk Q10)
    from metalchat_b200 import ops

    row = np.array([1.0, 0.8, 0.4, 0.3, 0.1], np.float32)
    x = np.zeros((2, 2048), np.float32)
    x[:, :5] = row
    out = ops.multinomial(gpu, gpu.tensor(x, "float"), 2048, 42, 7)
    gpu.wait()
    assert out.numpy().min() >= 0 and out.numpy().max() < 2048


@pytest.mark.parametrize("dtype", ["bfloat", "float"])
def test_elementwise(gpu, rng, dtype):
    # test/test_kernel_arithmetic.cc, test_kernel_mul.cc, test_kernel_activation.cc, test_kernel_logical.cc
    from metalchat_b200 import ops

    a, b = host(rng, (5, 160), dtype), host(rng, (5, 160), dtype)
    ta, tb = gpu.tensor(a, dtype), gpu.tensor(b, dtype)
    got = {n: getattr(ops, n)(gpu, ta, tb) for n in ("add", "sub", "div", "hadamard")}
    got_sm = ops.scalar_mul(gpu, ta, 0.125)
    got_silu, got_gelu = ops.silu(gpu, ta), ops.gelu(gpu, ta)
    got_gt, got_le = ops.gt(gpu, ta, 0.25), ops.le(gpu, ta, 0.25)
    bvec = host(rng, (160,), dtype)
    got_ab = ops.add_broadcast(gpu, ta, gpu.tensor(bvec, dtype))
    gpu.wait()
    want = np.zeros_like(a)
    for n in ("add", "sub", "div", "hadamard"):
        orc.binary(DT[dtype], n, want, a, b)
        assert np.array_equal(got[n].numpy(), want), n
    orc.scalar_mul(DT[dtype], want, a, 0.125)
    assert np.array_equal(got_sm.numpy(), want)
    orc.activation(DT[dtype], "silu", want, a)
    assert ulp_close(got_silu.numpy(), want, dtype)
    orc.activation(DT[dtype], "gelu", want, a)
    assert ulp_close(got_gelu.numpy(), want, dtype)
    m = np.zeros(a.shape, np.uint8)
    orc.compare(DT[dtype], "gt", m, a, 0.25)
    assert np.array_equal(got_gt.numpy(), m)
    orc.compare(DT[dtype], "le", m, a, 0.25)
    assert np.array_equal(got_le.numpy(), m)
    orc.add_broadcast(DT[dtype], want, a, bvec)
    assert np.array_equal(got_ab.numpy(), want)


def test_gelu_large_bf16_no_nan(gpu):
    # test/test_kernel_activation.cc: bf16 gelu(12) == 12
    from metalchat_b200 import ops

    out = ops.gelu(gpu, gpu.tensor(bf(np.full((1, 8), 12.0, np.float32)), "bfloat"))
    gpu.wait()
    assert np.all(unbf(out.numpy()) == 12.0)


@pytest.mark.parametrize("odt", ["bfloat", "float"])
@pytest.mark.parametrize("sdt", ["bfloat", "float"])
def test_hadamard_broadcast_bit_exact(gpu, rng, odt, sdt):
    from metalchat_b200 import ops

    q = rng.integers(-8, 8, size=(64, 32), dtype=np.int8)
    s = host(rng, (64,), sdt, 0.01)
    out = ops.hadamard_broadcast(gpu, odt, gpu.tensor(q, "int8_t"), gpu.tensor(s, sdt))
    gpu.wait()
    want = np.zeros((64, 32), dtype=np.uint16 if odt == "bfloat" else np.float32)
    orc.hadamard_broadcast(DT[odt], DT[sdt], want, q, s)
    assert np.array_equal(out.numpy(), want)


@pytest.mark.parametrize("dtype", ["bfloat", "float"])
def test_copy_scatter_gather_roll(gpu, rng, dtype):
    # test/test_kernel_copy.cc:14-130, test/test_kernel_roll.cc:16-71
    from metalchat_b200 import ops

    a = host(rng, (6, 32), dtype)
    ta = gpu.tensor(a, dtype)
    c = ops.clone(gpu, ta)
    big = gpu.tensor(np.zeros((6, 64), dtype=a.dtype), dtype)
    ops.clone(gpu, ta, big.narrow(1, 16, 32))  # copy into a strided slice (test_kernel_copy.cc:50-71)
    mask = (rng.random((6, 32)) > 0.5).astype(np.uint8)
    sc = ops.scatter(gpu, ops.clone(gpu, ta), gpu.tensor(mask, "bool"), 0.0)
    idx = rng.integers(0, 32, size=(6, 10), dtype=np.int32)
    g = ops.gather(gpu, ta, gpu.tensor(idx, "int32_t"))
    r = host(rng, (2, 128, 8, 64), dtype)
    rolled = ops.roll(gpu, gpu.tensor(r, dtype), 5, 1)
    gpu.wait()
    assert np.array_equal(c.numpy(), a)
    exp_big = np.zeros((6, 64), dtype=a.dtype)
    exp_big[:, 16:48] = a
    assert np.array_equal(big.numpy(), exp_big)
    exp_sc = a.copy()
    exp_sc[mask.astype(bool)] = 0
    assert np.array_equal(sc.numpy(), exp_sc)
    assert np.array_equal(g.numpy(), np.take_along_axis(a, idx.astype(np.int64), axis=1))
    want = np.zeros(r.size, dtype=r.dtype)
    orc.roll(DT[dtype], want, r.reshape(-1), 5, 128, 8 * 64)
    assert np.array_equal(rolled.numpy().reshape(-1), want)


def test_empty_inputs(gpu):
    from metalchat_b200 import ops

    e = gpu.tensor(np.zeros((0, 16), np.float32), "float")
    for fn in (ops.softmax, ops.silu, ops.clone):
        out = fn(gpu, e)
        assert out.numel() == 0
    gpu.wait()
