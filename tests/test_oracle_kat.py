"""Pins the CPU oracle against the reference's own known-answer tests.

Each test cites the reference test it mirrors (paths relative to ybubnov/metalchat).
"""
import numpy as np

from oracle import orc
from oracle.orc import BF16, F32


def bf(a):
    return orc.f32_to_bf16(np.asarray(a, dtype=np.float32))


def test_bf16_rne_and_host_flush():
    # include/metalchat/dtype.h:26-57
    assert orc.lib().orc_f32_to_bf16(1.0) == 0x3F80
    assert orc.lib().orc_f32_to_bf16(1.00390625) == 0x3F80  # tie -> even
    assert orc.lib().orc_f32_to_bf16(1.01171875) == 0x3F82  # tie -> even (up)
    assert orc.lib().orc_f32_to_bf16_host(1e-40) == 0x0000  # subnormal flushed (Q17)
    assert orc.lib().orc_f32_to_bf16_host(-1e-40) == 0x8000
    x = np.random.default_rng(1).standard_normal(10000).astype(np.float32)
    mine = orc.f32_to_bf16(x)
    ref = np.array([orc.lib().orc_f32_to_bf16(float(v)) for v in x[:2000]], dtype=np.uint16)
    assert np.array_equal(mine[:2000], ref)


def test_softmax_golden_vector():
    # test/test_kernel_softmax.cc:19-39
    x = bf(np.arange(5, dtype=np.float32)).reshape(1, 5)
    out = np.zeros_like(x)
    orc.softmax(BF16, out, x)
    expect = np.array([0.0116577, 0.0317383, 0.0859375, 0.234375, 0.636719], dtype=np.float32)
    assert np.allclose(orc.bf16_to_f32(out)[0], orc.bf16_to_f32(bf(expect)), atol=1e-5)


def test_softmax_sums_to_one(rng):
    # test/test_kernel_softmax.cc:42-72
    x = rng.random((32 * 4, 4), dtype=np.float32)
    out = np.zeros_like(x)
    orc.softmax(F32, out, x)
    assert np.allclose(out.sum(axis=1), 1.0, atol=1e-5)


def test_rmsnorm_ones():
    # test/test_kernel_rmsnorm.cc:18-37: bf16 ones with weight 3 -> exactly 3.0
    x = bf(np.ones((4 * 3 * 5, 7), dtype=np.float32))
    w = bf(np.full(7, 3.0, dtype=np.float32))
    out = np.zeros_like(x)
    orc.rmsnorm(BF16, out, x, w)
    assert np.all(orc.bf16_to_f32(out) == 3.0)


def test_rmsnorm_random(rng):
    # test/test_kernel_rmsnorm.cc:40-70
    x = rng.random((15, 2048), dtype=np.float32)
    w = rng.random(2048, dtype=np.float32)
    out = np.zeros_like(x)
    orc.rmsnorm(F32, out, x, w, eps=1e-5)
    inv = 1.0 / np.sqrt((x.astype(np.float64) ** 2).mean(axis=1, keepdims=True) + 1e-5)
    assert np.allclose(out, w * x * inv, atol=1e-5)


def test_bmm_transposed_view(rng):
    # test/test_kernel_bmm.cc:33-61: [1,5,2048] x [2048,8192] given as a transposed view
    a = rng.random((1, 5, 2048), dtype=np.float32)
    wt = rng.random((1, 1024, 2048), dtype=np.float32)  # N reduced 8192 -> 1024 for CPU time
    b = wt.transpose(0, 2, 1)
    out = np.zeros((1, 5, 1024), dtype=np.float32)
    orc.bmm(F32, out, a, b)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    assert np.allclose(out, ref, atol=2e-3, rtol=1e-5)


def test_bmm_bf16_ascending_k():
    # kernel/bmm.metal:55-76: fp32 accumulate in ascending k, one rounding at the end:
    # 1 + 2^-9 + 2^-9 = 1 + 2^-8 in fp32, a bf16 tie that rounds to even (1.0); rounding
    # after every add would also give 1.0, so add a case that distinguishes them.
    a = bf(np.array([[[1.0, 2.0**-9, 2.0**-9, 2.0**-9]]], dtype=np.float32))
    b = bf(np.ones((1, 4, 1), dtype=np.float32))
    out = np.zeros((1, 1, 1), dtype=np.uint16)
    orc.bmm(BF16, out, a, b)
    assert orc.bf16_to_f32(out)[0, 0, 0] == 1.0 + 2.0**-7  # 1 + 3*2^-9 rounds up


def test_embedding_exact(rng):
    # test/test_kernel_embedding.cc:19-53
    w = rng.random((1024, 64), dtype=np.float32)
    ids = rng.integers(0, 1024, size=(3, 7), dtype=np.int32)
    out = np.zeros((3, 7, 64), dtype=np.float32)
    orc.embedding(F32, out, ids, w)
    assert np.array_equal(out, w[ids])


def test_rope_freqs_vs_libm():
    # test/test_kernel_embedding.cc:72-137: theta 5e5, dim 64, start 100, no scaling
    fc = np.zeros((1024, 32), dtype=np.float32)
    fs = np.zeros_like(fc)
    orc.rope_freqs(fc, fs, 64, 100, 500000.0)
    # the reference test computes freqs with std::powf (test_kernel_embedding.cc:86)
    import ctypes

    libm = ctypes.CDLL("libm.so.6")
    libm.powf.restype = ctypes.c_float
    libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
    freqs = np.array([np.float32(1.0) / np.float32(libm.powf(500000.0, 2.0 * j / 64)) for j in range(32)], dtype=np.float32)
    ang = (np.arange(100, 1124)[:, None].astype(np.float32) * freqs[None, :]).astype(np.float32)
    assert np.allclose(fc, np.cos(ang.astype(np.float64)), atol=1e-4)
    assert np.allclose(fs, np.sin(ang.astype(np.float64)), atol=1e-4)


def test_sort_descending(rng):
    # test/test_kernel_sort.cc:17-52: 100000 floats
    x = rng.random((1, 100000), dtype=np.float32)
    P = 131072
    v = np.zeros((1, P), dtype=np.float32)
    ix = np.zeros((1, P), dtype=np.int32)
    orc.sort(F32, v, ix, x)
    vals, idx = v[0, :100000], ix[0, :100000]
    assert np.all(vals[:-1] >= vals[1:])
    assert np.array_equal(x[0, idx], vals)
    assert np.all(np.isneginf(v[0, 100000:]))


def test_cumsum_and_sum(rng):
    # test/test_kernel_sum.cc:17-68
    x = rng.random((4, 300), dtype=np.float32)
    out = np.zeros_like(x)
    orc.cumsum(F32, out, x)
    assert np.allclose(out, np.cumsum(x.astype(np.float64), axis=1), atol=1e-4)
    s = np.zeros(4, dtype=np.float32)
    orc.rowsum(F32, s, x)
    assert np.allclose(s, x.sum(axis=1), atol=1e-2)


def test_cumsum_bf16_order():
    # kernel/cumsum.metal:45-67: accumulation happens in T (bf16), block totals are added
    # nearest block first.
    x = bf(np.ones((1, 8), dtype=np.float32))
    out = np.zeros_like(x)
    orc.cumsum(BF16, out, x, block=2)
    assert list(orc.bf16_to_f32(out)[0]) == [1, 2, 3, 4, 5, 6, 7, 8]
    # 300 ones: bf16 has 8 significand bits, so counting by adding block totals of 2
    # saturates differently from an fp32 scan — the order is observable.
    x = bf(np.ones((1, 600), dtype=np.float32))
    out = np.zeros_like(x)
    orc.cumsum(BF16, out, x, block=2)
    got = orc.bf16_to_f32(out)[0]
    assert got[255] == 256 and got[511] == 512  # +2 steps stay exact up to 512
    assert got[599] != 600  # beyond 512 the bf16 spacing is 4


def test_pcg32_known_answer():
    # kernel/multinomial.metal:17-57 is PCG32 XSH-RR (pcg-random.org); the published demo
    # stream for srandom(42, 54) starts 0xa15c02b7
    u = orc.lib().orc_pcg32_uniform(42, 54)
    bits = (0xA15C02B7 >> 9) | 0x3F800000
    expect = np.array([bits], dtype=np.uint32).view(np.float32)[0] - np.float32(1.0)
    assert u == expect


def test_multinomial_frequencies():
    # test/test_kernel_multinomial.cc:16-54: 8192 draws from reverse CDF {1,.8,.4,.3,.1}
    # within +-0.02 of {.2,.4,.1,.2,.1}.  The reference test reads a = input[row, 8191]
    # past the 5-wide row (quirk Q10), i.e. a == 0; a trailing 0 column with the
    # "intended" reading reproduces that.
    cdf = np.tile(np.array([1.0, 0.8, 0.4, 0.3, 0.1, 0.0], dtype=np.float32), (4, 1))
    out = np.zeros((4, 8192), dtype=np.int32)
    orc.multinomial(F32, out, cdf, init_state=1234, init_seq=99, intended=1)
    expect = np.array([0.2, 0.4, 0.1, 0.2, 0.1])
    for r in range(4):
        freq = np.bincount(out[r], minlength=6)[:5] / 8192.0
        assert np.all(np.abs(freq - expect) < 0.02), freq


def test_multinomial_reference_default_is_top1():
    # quirk Q10: sample_size == 1 -> a == b == input[row, 0] -> always index 0
    probs = np.array([[0.5, 0.3, 0.2, 0.0]], dtype=np.float32)
    out = np.zeros((1, 1), dtype=np.int32)
    for u in (0.0, 0.3, 0.999):
        orc.multinomial(F32, out, probs, uniforms=np.array([u]), intended=0)
        assert out[0, 0] == 0


def test_hadamard_broadcast_dequant(rng):
    # test/test_kernel_mul.cc:41-65: <float, int8, float> [512,64,32] * [512,64,1]
    w = rng.integers(1, 10, size=(512 * 64, 32), dtype=np.int8)
    s = rng.random(512 * 64, dtype=np.float32)
    out = np.zeros(w.shape, dtype=np.float32)
    orc.hadamard_broadcast(F32, F32, out, w, s)
    assert np.allclose(out, w.astype(np.float32) * s[:, None], atol=1e-5)
    # bf16 output: double rounding r(r(q) * r(s))  (kernel/mul.metal:76-77)
    outb = np.zeros(w.shape, dtype=np.uint16)
    orc.hadamard_broadcast(BF16, F32, outb, w, s)
    sb = orc.bf16_to_f32(bf(s))
    assert np.array_equal(outb, bf(w.astype(np.float32) * sb[:, None]))


def test_elementwise(rng):
    # test/test_kernel_mul.cc:16-38,68-93; test/test_kernel_arithmetic.cc:18-149
    a = rng.random((5 * 32, 160), dtype=np.float32)
    b = rng.random((5 * 32, 160), dtype=np.float32) + 0.5
    out = np.zeros_like(a)
    for op, f in (("add", np.add), ("sub", np.subtract), ("div", np.divide), ("hadamard", np.multiply)):
        orc.binary(F32, op, out, a, b)
        assert np.allclose(out, f(a, b), atol=1e-5)
    orc.scalar_mul(F32, out, a, 2.5)
    assert np.allclose(out, a * 2.5, atol=1e-5)
    m = rng.random(160, dtype=np.float32)
    orc.add_broadcast(F32, out, a, m)
    assert np.allclose(out, a + m[None, :], atol=1e-5)


def test_activation(rng):
    # test/test_kernel_activation.cc:19-83
    x = (rng.random((3, 64), dtype=np.float32) - 0.5) * 8
    out = np.zeros_like(x)
    orc.activation(F32, "silu", out, x)
    assert np.allclose(out, x / (1 + np.exp(-x)), atol=1e-5)
    orc.activation(F32, "gelu", out, x)
    ref = 0.5 * x * (1 + np.tanh(np.sqrt(2 / np.pi) * (x + 0.044715 * x**3)))
    assert np.allclose(out, ref, atol=1e-5)
    big = bf(np.array([[12.0]], dtype=np.float32))
    ob = np.zeros_like(big)
    orc.activation(BF16, "gelu", ob, big)
    assert orc.bf16_to_f32(ob)[0, 0] == 12.0


def test_copy_scatter_gather(rng):
    # test/test_kernel_copy.cc:14-130 (incl. copy into a strided slice)
    a = rng.random((16, 64), dtype=np.float32)
    big = np.zeros((16, 128), dtype=np.float32)
    orc.copy(F32, big[:, 32:96], a)
    assert np.array_equal(big[:, 32:96], a) and not big[:, :32].any() and not big[:, 96:].any()
    idx = rng.integers(0, 64, size=(16, 10), dtype=np.int32)
    g = np.zeros((16, 10), dtype=np.float32)
    orc.gather(F32, g, a, idx)
    assert np.array_equal(g, np.take_along_axis(a, idx, axis=1))
    mask = (a > 0.5).astype(np.uint8)
    s = a.copy()
    orc.scatter(F32, s, mask, -1.0)
    assert np.array_equal(s, np.where(a > 0.5, np.float32(-1.0), a))


def test_compare(rng):
    # test/test_kernel_logical.cc:14-34
    a = rng.random((4, 33), dtype=np.float32)
    o = np.zeros(a.shape, dtype=np.uint8)
    orc.compare(F32, "gt", o, a, 0.5)
    assert np.array_equal(o.astype(bool), a > 0.5)
    orc.compare(F32, "le", o, a, 0.5)
    assert np.array_equal(o.astype(bool), a <= 0.5)


def test_roll(rng):
    # test/test_kernel_roll.cc:16-71: [2,128,8,64] rolled along dim 1
    a = rng.random((2, 128, 8, 64), dtype=np.float32)
    out = np.zeros(a.size, dtype=np.float32)
    orc.roll(F32, out, a.reshape(-1), 5, 128, 8 * 64)
    assert np.array_equal(out.reshape(a.shape), np.roll(a, -5, axis=1))
