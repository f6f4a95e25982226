// oracle/orc_ops.h — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Scalar CPU restatement of every Metal kernel of ybubnov/metalchat v1.2.1
// (kernel/*.metal), one function per kernel, same argument order, same
// tensor_layout addressing, same rounding points.  Where the GPU's reduction
// partition is observable (rmsnorm/softmax/sum/cumsum) the threadgroup partition
// chosen by the reference wrapper (include/metalchat/kernel/*.h) is emulated with
// max_threads_per_threadgroup fixed at 1024 (Apple GPUs; src/kernel.cc:75-79).
//
// Floating-point transcendental functions (`metal::precise::exp/rsqrt/pow/cos/
// sin/tanh`) are replaced by <cmath> fp32 functions; the file is compiled with
// -ffp-contract=off so that every fp32 multiply/add is individually rounded.
#pragma once
#include "orc_common.h"

#include <algorithm>
#include <limits>
#include <vector>

namespace orc {

constexpr uint32_t kMaxThreads = 1024; // src/kernel.cc:75-79 on Apple GPUs
constexpr uint32_t kSimd = 32;

static inline uint32_t ceil_div(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
static inline uint32_t ceil_pow2(uint32_t x)
{
    uint32_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

// metal::simd_sum over 32 lanes, modelled as an xor-butterfly (all lanes end up
// with the same value because fp32 addition is commutative).
static inline float simd_sum32(const float* lanes)
{
    float v[kSimd];
    for (uint32_t i = 0; i < kSimd; i++) v[i] = lanes[i];
    for (uint32_t off = kSimd / 2; off > 0; off >>= 1) {
        float n[kSimd];
        for (uint32_t i = 0; i < kSimd; i++) n[i] = v[i] + v[i ^ off];
        for (uint32_t i = 0; i < kSimd; i++) v[i] = n[i];
    }
    return v[0];
}

// Threadgroup-wide sum of one fp32 partial per thread: simd_sum per simdgroup, the
// per-simdgroup results through a zero-initialised 32-slot threadgroup array, then
// simd_sum again (kernel/rmsnorm.metal:58-82, kernel/softmax.metal:49-73,
// kernel/sum.metal:49-70).
static inline float threadgroup_sum(const std::vector<float>& partial)
{
    float tg[kSimd];
    for (uint32_t i = 0; i < kSimd; i++) tg[i] = 0.0f;
    const uint32_t n = uint32_t(partial.size());
    for (uint32_t g = 0; g * kSimd < n; g++) {
        float lanes[kSimd];
        for (uint32_t l = 0; l < kSimd; l++) {
            const uint32_t t = g * kSimd + l;
            lanes[l] = t < n ? partial[t] : 0.0f;
        }
        tg[g % kSimd] = simd_sum32(lanes);
    }
    return simd_sum32(tg);
}

// ---- bmm  (kernel/bmm.metal:24-82; wrapper kernel/bmm.h:27-89) -----------------
// C[b,M,N] = A[b,M,K] * B[b,K,N]; operands promoted to fp32, fp32 accumulation in
// ascending k (the 8x8 tiling only zero-pads), single rounding to T at the end.
template <typename T>
void bmm(T* out, const layout<3>& lo, const T* m1, const layout<3>& l1, const T* m2, const layout<3>& l2)
{
    view3<T> o{out, &lo};
    view3<const T> a{m1, &l1};
    view3<const T> b{m2, &l2};
    const uint32_t B = a.size(0), M = a.size(1), K = a.size(2), N = b.size(2);
#pragma omp parallel for collapse(2) schedule(static)
    for (uint32_t bi = 0; bi < B; bi++) {
        for (uint32_t n = 0; n < N; n++) {
            for (uint32_t m = 0; m < M; m++) {
                float acc = 0.0f;
                for (uint32_t k = 0; k < K; k++) {
                    acc += float(a.at(bi, m, k)) * float(b.at(bi, k, n));
                }
                o.at(bi, m, n) = T(acc);
            }
        }
    }
}

// ---- rmsnorm (kernel/rmsnorm.metal:28-91; wrapper kernel/rmsnorm.h:28-55) ------
template <typename T>
void rmsnorm(
    T* out, const layout<2>& lo, const T* in, const layout<2>& li, const T* w, const layout<1>& lw,
    float eps, float mu, uint32_t block_size
)
{
    view2<T> o{out, &lo};
    view2<const T> x{in, &li};
    view1<const T> wt{w, &lw};
    const uint32_t rows = x.size(0), D = x.size(1);
    const uint32_t threads = ceil_div(D, block_size);
    for (uint32_t i = 0; i < rows; i++) {
        std::vector<float> partial(threads, 0.0f);
        for (uint32_t t = 0; t < threads; t++) {
            float s = 0.0f;
            for (uint32_t j = t * block_size; j < (t + 1) * block_size && j < D; j++) {
                const float xj = float(x.at(i, j));
                s += xj * xj;
            }
            partial[t] = s;
        }
        const float acc = threadgroup_sum(partial);
        const float mean_sq = acc / float(D);
        const float inv = 1.0f / std::sqrt(mean_sq + eps); // precise::rsqrt, :80
        for (uint32_t j = 0; j < D; j++) {
            const float xv = float(x.at(i, j));
            const float weight = mu + float(wt.at(j));
            o.at(i, j) = T(weight * xv * inv); // (weight*x)*inv, :88-89
        }
    }
}

// ---- softmax (kernel/softmax.metal:24-88): NO max subtraction (quirk Q1) -------
template <typename T>
void softmax(T* out, const layout<2>& lo, const T* in, const layout<2>& li, uint32_t block_size)
{
    view2<T> o{out, &lo};
    view2<const T> x{in, &li};
    const uint32_t rows = x.size(0), D = x.size(1);
    const uint32_t threads = ceil_div(D, block_size);
    for (uint32_t i = 0; i < rows; i++) {
        std::vector<float> partial(threads, 0.0f);
        for (uint32_t t = 0; t < threads; t++) {
            float s = 0.0f;
            for (uint32_t j = t * block_size; j < (t + 1) * block_size && j < D; j++) {
                s += std::exp(float(x.at(i, j)));
            }
            partial[t] = s;
        }
        const float inv = 1.0f / threadgroup_sum(partial);
        for (uint32_t j = 0; j < D; j++) {
            o.at(i, j) = T(std::exp(float(x.at(i, j))) * inv);
        }
    }
}

// ---- sum (kernel/sum.metal:27-75) ----------------------------------------------
template <typename T>
void sum(T* out, const layout<1>& lo, const T* in, const layout<2>& li, uint32_t block_size)
{
    view1<T> o{out, &lo};
    view2<const T> x{in, &li};
    const uint32_t rows = x.size(0), D = x.size(1);
    const uint32_t threads = ceil_div(D, block_size);
    for (uint32_t i = 0; i < rows; i++) {
        std::vector<float> partial(threads, 0.0f);
        for (uint32_t t = 0; t < threads; t++) {
            float s = 0.0f;
            for (uint32_t j = t * block_size; j < (t + 1) * block_size && j < D; j++) {
                s += float(x.at(i, j));
            }
            partial[t] = s;
        }
        o.at(i) = T(threadgroup_sum(partial));
    }
}

// ---- rope (kernel/rope.metal:28-63): half-split pairs (k, k + D/2), quirk Q2 ----
template <typename T>
void rope(
    T* out, const layout<2>& lo, const T* in, const layout<2>& li, const float* fcos,
    const layout<2>& lc, const float* fsin, const layout<2>& ls, uint32_t batch_size,
    uint32_t n_head, uint32_t start_pos
)
{
    view2<T> o{out, &lo};
    view2<const T> x{in, &li};
    view2<const float> c{fcos, &lc};
    view2<const float> s{fsin, &ls};
    const uint32_t rows = x.size(0), half = c.size(1);
    for (uint32_t i = 0; i < rows; i++) {
        const uint32_t pos = i / (batch_size * n_head);
        for (uint32_t k = 0; k < half; k++) {
            const float x1 = float(x.at(i, k));
            const float x2 = float(x.at(i, half + k));
            const float fc = c.at(start_pos + pos, k);
            const float fs = s.at(start_pos + pos, k);
            o.at(i, k) = T(fc * x1 - fs * x2);
            o.at(i, half + k) = T(fs * x1 + fc * x2);
        }
    }
}

// ---- rope_freqs (kernel/rope.metal:76-102) --------------------------------------
static inline void rope_freqs(
    float* fcos, const layout<2>& lc, float* fsin, const layout<2>& ls, uint32_t dim,
    uint32_t start_pos, float theta
)
{
    view2<float> c{fcos, &lc};
    view2<float> s{fsin, &ls};
    for (uint32_t i = 0; i < c.size(0); i++) {
        for (uint32_t j = 0; j < dim / 2; j++) {
            const float freq = 1.0f / std::pow(theta, 2.0f * float(j) / float(dim));
            const float angle = float(start_pos + i) * freq;
            c.at(i, j) = std::cos(angle);
            s.at(i, j) = std::sin(angle);
        }
    }
}

// ---- embedding (kernel/embedding.metal:38-66): out[i,j,:] = W[ids[i,j],:] --------
template <typename T>
void embedding(
    T* out, const layout<3>& lo, const int32_t* ids, const layout<2>& li, const T* w,
    const layout<2>& lw
)
{
    view3<T> o{out, &lo};
    view2<const int32_t> id{ids, &li};
    view2<const T> wt{w, &lw};
    for (uint32_t i = 0; i < id.size(0); i++)
        for (uint32_t j = 0; j < id.size(1); j++)
            for (uint32_t k = 0; k < wt.size(1); k++) o.at(i, j, k) = wt.at(uint32_t(id.at(i, j)), k);
}

// ---- sort (kernel/sort.metal:31-86; wrapper kernel/sort.h:33-62) ----------------
// Descending bitonic network over the row padded to a power of two with -inf.
// Each (k, j) stage touches disjoint pairs, so a sequential sweep reproduces the
// parallel network exactly, including the tie order (quirk Q12).
template <typename T> static inline bool lt(const T& a, const T& b) { return float(a) < float(b); }
template <typename T> static inline bool gt_(const T& a, const T& b) { return float(a) > float(b); }

template <typename T>
void sort(
    T* values, const layout<2>& lv, int32_t* indices, const layout<2>& lx, const T* in,
    const layout<2>& li
)
{
    view2<T> v{values, &lv};
    view2<int32_t> ix{indices, &lx};
    view2<const T> x{in, &li};
    const uint32_t rows = x.size(0), D = x.size(1), P = v.size(1);
    for (uint32_t b = 0; b < rows; b++) {
        for (uint32_t k = 0; k < P; k++) {
            v.at(b, k) = k < D ? x.at(b, k) : T(-std::numeric_limits<float>::infinity());
            ix.at(b, k) = int32_t(k);
        }
        for (uint32_t k = 2; k <= P; k *= 2) {
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t i = 0; i < P; i++) {
                    const uint32_t ij = i ^ j;
                    if (i < ij) {
                        T& a = v.at(b, i);
                        T& c = v.at(b, ij);
                        const bool up = (i & k) == 0;
                        if ((up && lt(a, c)) || (!up && gt_(a, c))) {
                            std::swap(a, c);
                            std::swap(ix.at(b, i), ix.at(b, ij));
                        }
                    }
                }
            }
        }
    }
}

// ---- cumsum (kernel/cumsum.metal:24-98; wrapper kernel/sum.h:33-61) --------------
// Accumulates in T (bf16!): thread-serial prefix per block of `block` elements,
// then every thread adds the totals of all preceding blocks one at a time,
// nearest block first.  The reference's `group_sums[256]` overflows for more than
// 256 threads (quirk Q11); the oracle models an unbounded array (intended
// semantics) — rows that need more than 256 threads are UB in the reference.
template <typename T> void cumsum(T* out, const layout<2>& lo, const T* in, const layout<2>& li, uint32_t block)
{
    view2<T> o{out, &lo};
    view2<const T> x{in, &li};
    const uint32_t rows = x.size(0), D = x.size(1);
    const uint32_t threads = ceil_div(D, block);
    std::vector<T> group(threads);
    std::vector<T> local(size_t(threads) * block);
    for (uint32_t i = 0; i < rows; i++) {
        for (uint32_t t = 0; t < threads; t++) {
            const uint32_t begin = t * block, end = begin + block;
            const uint32_t bs = end > D ? D % block : block;
            T* ls = &local[size_t(t) * block];
            for (uint32_t k = begin, j = 0; k < end && k < D; k++, j++) {
                ls[j] = j > 0 ? T(float(x.at(i, k)) + float(ls[j - 1])) : x.at(i, k);
            }
            group[t] = ls[bs - 1];
        }
        for (uint32_t t = 0; t < threads; t++) {
            const uint32_t begin = t * block, end = begin + block;
            const uint32_t bs = end > D ? D % block : block;
            T* ls = &local[size_t(t) * block];
            for (uint32_t a = 1; a <= t; a++) {
                const T acc = group[t - a];
                for (uint32_t j = 0; j < bs; j++) ls[j] = T(float(ls[j]) + float(acc));
            }
            for (uint32_t k = begin; k < end && k < D; k++) o.at(i, k) = ls[k - begin];
        }
    }
}

// ---- multinomial (kernel/multinomial.metal:15-123) -------------------------------
struct pcg32 {
    uint64_t state, inc;
    uint32_t next()
    {
        const uint64_t pre = state;
        state = pre * 6364136223846793005ull + inc;
        const uint32_t xs = uint32_t(((pre >> 18u) ^ pre) >> 27u);
        const uint32_t rot = uint32_t(pre >> 59u);
        return (xs >> rot) | (xs << ((~rot + 1u) & 31));
    }
    pcg32(uint64_t init_state, uint64_t init_seq) : state(0), inc((init_seq << 1u) | 1u)
    {
        next();
        state += init_state;
        next();
    }
    float uniform()
    {
        const uint32_t u = (next() >> 9) | 0x3f800000u;
        float f;
        std::memcpy(&f, &u, 4);
        return f - 1.0f;
    }
};

template <typename T> uint32_t reverse_cdf_search(const view2<const T>& d, uint32_t row, T value)
{
    int low = 0, high = int(d.size(1));
    while (low < high) {
        const uint32_t mid = uint32_t(low + high) / 2;
        if (float(d.at(row, mid)) > float(value)) {
            low = int(mid) + 1;
        } else {
            high = int(mid);
        }
    }
    return uint32_t(std::max(low, 1) - 1);
}

// `uniforms` (rows x samples, may be null) injects the draws; otherwise PCG32 with
// (init_state + row, init_seq + sample).  `intended` selects a = input[row, N-1]
// instead of the reference's a = input[row, samples-1] (quirk Q10,
// kernel/multinomial.metal:107).  With samples > N that read leaves the row: it lands in a later row, or -- beyond the
// `avail` elements of the buffer -- outside the allocation, where a Metal device read yields 0; the reference's own test
// depends on it (test/test_kernel_multinomial.cc:16-54 draws 8192 samples from rows of 5: a = 0, r = u * p_max).
template <typename T>
void multinomial(
    int32_t* out, const layout<2>& lo, const T* in, const layout<2>& li, uint64_t init_state,
    uint64_t init_seq, const float* uniforms, int intended, uint64_t avail = ~uint64_t(0)
)
{
    view2<int32_t> o{out, &lo};
    view2<const T> x{in, &li};
    const uint32_t rows = o.size(0), S = o.size(1), N = x.size(1);
    if (avail == ~uint64_t(0)) avail = uint64_t(x.size(0) - 1) * li.strides[0] + uint64_t(N - 1) * li.strides[1] + li.offsets[0] + li.offsets[1] + 1;
    for (uint32_t i = 0; i < rows; i++) {
        for (uint32_t k = 0; k < S; k++) {
            const uint32_t ca = intended ? N - 1 : S - 1;
            const uint64_t flat = uint64_t(i) * li.strides[0] + uint64_t(ca) * li.strides[1] + li.offsets[0] + li.offsets[1];
            const float a = flat < avail ? float(in[flat]) : 0.0f;
            const float b = float(x.at(i, 0));
            float u;
            if (uniforms) {
                u = uniforms[size_t(i) * S + k];
            } else {
                pcg32 g(init_state + i, init_seq + k);
                u = g.uniform();
            }
            const T r = T(u * (b - a) + a);
            o.at(i, k) = int32_t(reverse_cdf_search(x, i, r));
        }
    }
}

// ---- elementwise (kernel/mul.metal, arithmetic.metal, activation.metal) ---------
// Binary ops are evaluated in T: operands are T, the exact result is rounded once.
template <typename T, typename F>
void binary2(T* out, const layout<2>& lo, const T* a, const layout<2>& la, const T* b, const layout<2>& lb, F f)
{
    view2<T> o{out, &lo};
    view2<const T> x{a, &la};
    view2<const T> y{b, &lb};
    for (uint32_t i = 0; i < x.size(0); i++)
        for (uint32_t k = 0; k < x.size(1); k++) o.at(i, k) = T(f(float(x.at(i, k)), float(y.at(i, k))));
}

// add_broadcast (kernel/arithmetic.metal:59-80): out[i,j] = a[i,j] + b[j mod n]
template <typename T>
void add_broadcast(T* out, const layout<2>& lo, const T* a, const layout<2>& la, const T* b, const layout<1>& lb)
{
    view2<T> o{out, &lo};
    view2<const T> x{a, &la};
    view1<const T> y{b, &lb};
    for (uint32_t i = 0; i < x.size(0); i++)
        for (uint32_t j = 0; j < x.size(1); j++)
            o.at(i, j) = T(float(x.at(i, j)) + float(y.at(j % y.size(0))));
}

// hadamard_broadcast (kernel/mul.metal:59-85): the dequant kernel.
// out[i,j] = O(in1[i,j]) * O(in2[i mod n]) evaluated in O (quirk Q7: the scale is
// rounded to O first, the product is rounded to O).
template <typename O, typename S>
void hadamard_broadcast(O* out, const layout<2>& lo, const int8_t* a, const layout<2>& la, const S* b, const layout<1>& lb)
{
    view2<O> o{out, &lo};
    view2<const int8_t> x{a, &la};
    view1<const S> y{b, &lb};
    for (uint32_t i = 0; i < x.size(0); i++) {
        const O s = O(float(y.at(i % y.size(0))));
        for (uint32_t j = 0; j < x.size(1); j++) {
            const O q = O(float(x.at(i, j)));
            o.at(i, j) = O(float(q) * float(s));
        }
    }
}

// scalar_mul (kernel/mul.metal:97-117): out = in * c in T.
template <typename T> void scalar_mul(T* out, const layout<2>& lo, const T* a, const layout<2>& la, T c)
{
    view2<T> o{out, &lo};
    view2<const T> x{a, &la};
    for (uint32_t i = 0; i < x.size(0); i++)
        for (uint32_t k = 0; k < x.size(1); k++) o.at(i, k) = T(float(x.at(i, k)) * float(c));
}

// silu (kernel/activation.metal:18-37): x / (T(1) + T(exp(-x))) evaluated in T
// (quirk Q5): e = T(exp(-x)); d = T(1 + e); out = T(x / d).
template <typename T> static inline T silu1(T x)
{
    const T e = T(std::exp(-float(x)));
    const T d = T(1.0f + float(e));
    return T(float(x) / float(d));
}
template <typename T> void silu(T* out, const layout<2>& lo, const T* a, const layout<2>& la)
{
    view2<T> o{out, &lo};
    view2<const T> x{a, &la};
    for (uint32_t i = 0; i < x.size(0); i++)
        for (uint32_t k = 0; k < x.size(1); k++) o.at(i, k) = silu1(x.at(i, k));
}

// gelu (kernel/activation.metal:50-75): tanh approximation in fp32, one rounding.
template <typename T> static inline T gelu1(T xv)
{
    const float beta = 1.41421356237309504880f * 1.12837916709551257390f * 0.5f;
    const float kappa = 0.044715f;
    const float x = float(xv);
    const float x3 = x * x * x;
    const float inner = beta * (x + kappa * x3);
    return T(0.5f * x * (1.0f + std::tanh(inner)));
}
template <typename T> void gelu(T* out, const layout<2>& lo, const T* a, const layout<2>& la)
{
    view2<T> o{out, &lo};
    view2<const T> x{a, &la};
    for (uint32_t i = 0; i < x.size(0); i++)
        for (uint32_t k = 0; k < x.size(1); k++) o.at(i, k) = gelu1(x.at(i, k));
}

// ---- copy / scatter / gather (kernel/copy.metal) ---------------------------------
template <typename T> void copy(T* out, const layout<2>& lo, const T* a, const layout<2>& la)
{
    view2<T> o{out, &lo};
    view2<const T> x{a, &la};
    for (uint32_t i = 0; i < x.size(0); i++)
        for (uint32_t k = 0; k < x.size(1); k++) o.at(i, k) = x.at(i, k);
}
template <typename T> void scatter(T* out, const layout<2>& lo, const uint8_t* mask, const layout<2>& lm, T value)
{
    view2<T> o{out, &lo};
    view2<const uint8_t> m{mask, &lm};
    for (uint32_t i = 0; i < o.size(0); i++)
        for (uint32_t k = 0; k < o.size(1); k++)
            if (m.at(i, k)) o.at(i, k) = value;
}
template <typename T>
void gather(T* out, const layout<2>& lo, const T* a, const layout<2>& la, const int32_t* idx, const layout<2>& li)
{
    view2<T> o{out, &lo};
    view2<const T> x{a, &la};
    view2<const int32_t> ix{idx, &li};
    for (uint32_t i = 0; i < ix.size(0); i++)
        for (uint32_t k = 0; k < ix.size(1); k++) o.at(i, k) = x.at(i, uint32_t(ix.at(i, k)));
}

// ---- gt / le (kernel/logical.metal) ----------------------------------------------
template <typename T> void gt(uint8_t* out, const layout<2>& lo, const T* a, const layout<2>& la, T value)
{
    view2<uint8_t> o{out, &lo};
    view2<const T> x{a, &la};
    for (uint32_t i = 0; i < x.size(0); i++)
        for (uint32_t k = 0; k < x.size(1); k++) o.at(i, k) = float(x.at(i, k)) > float(value);
}
template <typename T> void le(uint8_t* out, const layout<2>& lo, const T* a, const layout<2>& la, T value)
{
    view2<uint8_t> o{out, &lo};
    view2<const T> x{a, &la};
    for (uint32_t i = 0; i < x.size(0); i++)
        for (uint32_t k = 0; k < x.size(1); k++) o.at(i, k) = float(x.at(i, k)) <= float(value);
}

// ---- roll (kernel/roll.metal:22-45) ----------------------------------------------
template <typename T>
void roll(T* out, const layout<1>& lo, const T* a, const layout<1>& la, uint32_t shift, uint32_t size, uint32_t stride)
{
    view1<T> o{out, &lo};
    view1<const T> x{a, &la};
    const uint32_t stride_size = size * stride;
    for (uint32_t k = 0; k < x.size(0); k++) {
        const uint32_t base = (k / stride_size) * stride_size;
        const uint32_t i = (k / stride + shift) % size;
        const uint32_t j = k % stride;
        o.at(k) = x.at(base + i * stride + j);
    }
}

} // namespace orc
