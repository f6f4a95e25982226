// metalchat_b200/csrc/mc_sample_kernels.cuh — the sampling tail on the device.
//
// Reference (nn/sampling.h:152-316, make_default_sampler): top-k on the HOST (clone x2, .get() sync, std::partial_sort
// over 128 256 entries, nn/sampling.h:244-264) -> nucleus: scalar_mul(1/T), softmax, sort, cumsum, sub, gt, scatter,
// gather (:183-200) -> multinomial + 2 gathers (:289-297): ~15 launches and 2 host round trips per token.  Here: two
// kernels, no host involvement, the token lands in the next step's input buffer.
//   K7a sample_select_kernel  per (block, row): bitonic sort of a 2048-logit slice by (value desc, index asc), keeps 64
//   K7b sample_finish_kernel  per row: sorts the 64 x blocks candidates, takes top-k, then replays the reference chain
//                             with its rounding points and reduction partitions: scale in T, softmax (no max shift,
//                             simd/threadgroup sum order), the SAME bitonic network for the descending sort (tie order,
//                             quirk Q12), cumsum_2 in T (kernel/cumsum.metal:45-67 order), sub, gt, scatter, multinomial
//                             with the injected uniform and the reference's `a = input[row, sample_size-1]` read (Q10).
// Integer results (top-k ids, sort order, mask, choice, token) are bit-exact against the oracle whenever the
// probabilities round to the same bf16 (expf vs libm differs by <= 2 ulp in fp32).
#pragma once
#include "mc_decode_kernels.cuh"

namespace mc {

constexpr int kSampleSlice = 2048;  // logits per select CTA
constexpr int kSampleKeep = 64;     // candidates kept per slice (>= top_k)
constexpr int kSampleMaxK = 64;

// total order "larger logit first, lower index first on ties" as a descending u64 key
__device__ __forceinline__ unsigned long long sample_key(uint16_t bf, uint32_t index)
{
    const uint32_t o = (bf & 0x8000u) ? (~uint32_t(bf) & 0xffffu) : (uint32_t(bf) | 0x8000u);
    return (static_cast<unsigned long long>(o) << 32) | (0xffffffffu - index);
}
// in-place descending bitonic sort of n (power of two) keys in shared memory by the whole CTA
__device__ __forceinline__ void bitonic_desc_u64(unsigned long long* s, uint32_t n)
{
    for (uint32_t k = 2; k <= n; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (uint32_t t = threadIdx.x; t < n / 2; t += blockDim.x) {
                const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), ij = i | j;
                const unsigned long long a = s[i], c = s[ij];
                const bool up = (i & k) == 0;
                if (up ? (a < c) : (a > c)) s[i] = c, s[ij] = a;
            }
        }
    }
    __syncthreads();
}

// grid (blocks, rows): candidates[row][block][64] = the 64 best (key-descending) of logits[row][block*2048 ...]
MC_KERNEL void __launch_bounds__(256) sample_select_kernel(const uint16_t* logits, uint32_t ld, uint32_t vocab, uint32_t index_base,
                                                            unsigned long long* cand)
{
    pdl_launch_dependents();
    pdl_wait();
    __shared__ unsigned long long s[kSampleSlice];
    const uint32_t row = blockIdx.y, beg = blockIdx.x * kSampleSlice;
    const uint16_t* l = logits + size_t(row) * ld;
    for (uint32_t i = threadIdx.x; i < uint32_t(kSampleSlice); i += blockDim.x) {
        const uint32_t v = beg + i;
        s[i] = v < vocab ? sample_key(l[v], index_base + v) : 0ull; // key 0 sorts last
    }
    bitonic_desc_u64(s, kSampleSlice);
    unsigned long long* out = cand + (size_t(row) * gridDim.x + blockIdx.x) * kSampleKeep;
    for (uint32_t i = threadIdx.x; i < uint32_t(kSampleKeep); i += blockDim.x) out[i] = s[i];
}

struct sample_params {
    const unsigned long long* cand; // [rows][n_cand]
    uint32_t n_cand;                // candidates per row (multiple of 64)
    uint32_t sort_n;                // n_cand rounded up to a power of two
    const uint16_t* logits;         // to fetch the exact bf16 of a candidate
    uint32_t ld;
    uint32_t index_base;            // first vocabulary id of this logits shard
    uint32_t top_k;
    float inv_t;                    // r(1 / r(T)) as fp32 (nn/sampling.h:183-186)
    float top_p;                    // r(p)
    uint32_t intended;              // 0: a = input[row, sample_size-1] (reference), 1: a = input[row, N-1]
    const float* uniforms;          // [steps][rows] injected draws (or [rows] when step_counter is null)
    // outputs (any may be null)
    int32_t* topk_idx;              // [rows][k]
    uint16_t* probs_sorted;         // [rows][k] bf16 after the top-p mask
    int32_t* probs_idx;             // [rows][k]
    int32_t* choice;                // [rows]
    int32_t* token;                 // [rows]
    // engine feedback (null for the stand-alone call)
    int32_t* ids;
    int32_t* pos;
    int32_t* out_log;
    int32_t* step_counter;
    uint32_t rows;
    int32_t advance;
};

// one CTA (1024 threads) per row
MC_KERNEL void __launch_bounds__(1024) sample_finish_kernel(const sample_params p)
{
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* s = reinterpret_cast<unsigned long long*>(smem_raw); // [sort_n]
    __shared__ float lg[kSampleMaxK], pr[kSampleMaxK], sv[kSampleMaxK], cs[kSampleMaxK], grp[kSampleMaxK], tg[32];
    __shared__ int32_t idx[kSampleMaxK], si[kSampleMaxK];
    const uint32_t row = blockIdx.x, k = p.top_k;
    for (uint32_t i = threadIdx.x; i < p.sort_n; i += blockDim.x) s[i] = i < p.n_cand ? p.cand[size_t(row) * p.n_cand + i] : 0ull;
    bitonic_desc_u64(s, p.sort_n);
    // top-k ids and their logits (topk_sampler, nn/sampling.h:244-264; ties: lower index first)
    if (threadIdx.x < k) {
        const uint32_t id = 0xffffffffu - uint32_t(s[threadIdx.x] & 0xffffffffu);
        idx[threadIdx.x] = int32_t(id);
        lg[threadIdx.x] = bf16_bits_to_f32(p.logits[size_t(row) * p.ld + (id - p.index_base)]);
        if (p.topk_idx) p.topk_idx[size_t(row) * k + threadIdx.x] = int32_t(id);
    }
    __syncthreads();
    // nucleus_sampler (nn/sampling.h:183-200): scalar_mul in T, softmax with one element per thread
    const uint32_t t = threadIdx.x;
    float e = 0.0f;
    if (t < k) e = expf(rbf(__fmul_rn(lg[t], p.inv_t)));
    if (t < 32) tg[t] = 0.0f;
    __syncthreads();
    if (t < 64) {
        float v = e;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
        if ((t & 31) == 0) tg[t >> 5] = v;
    }
    __syncthreads();
    if (t < 32) {
        float v = tg[t];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
        if (t == 0) tg[0] = v;
    }
    __syncthreads();
    const float inv = 1.0f / tg[0];
    uint32_t P = 1;
    while (P < k) P <<= 1;
    if (t < P) {
        pr[t] = t < k ? rbf(__fmul_rn(e, inv)) : -INFINITY; // sort pads with -inf (kernel/sort.metal:50-55)
        si[t] = int32_t(t);
    }
    // the reference's bitonic network (kernel/sort.metal:57-79), strict comparisons: ties keep their places
    for (uint32_t kk = 2; kk <= P; kk <<= 1) {
        for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
            __syncthreads();
            if (t < P / 2) {
                const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), ij = i | j;
                const float a = pr[i], c = pr[ij];
                const bool up = (i & kk) == 0;
                if (up ? (a < c) : (a > c)) {
                    pr[i] = c, pr[ij] = a;
                    const int32_t x = si[i];
                    si[i] = si[ij], si[ij] = x;
                }
            }
        }
    }
    __syncthreads();
    if (t < k) sv[t] = pr[t];
    __syncthreads();
    // cumsum_2 in T (kernel/cumsum.metal:45-67; wrapper block = max(2, pow2(ceil(k/1024))) = 2)
    const uint32_t threads = (k + 1) / 2;
    if (t < threads) {
        const uint32_t b = 2 * t;
        cs[b] = sv[b];
        float last = sv[b];
        if (b + 1 < k) last = rbf(__fadd_rn(sv[b + 1], sv[b])), cs[b + 1] = last;
        grp[t] = last;
    }
    __syncthreads();
    if (t < threads) {
        const uint32_t b = 2 * t;
        for (uint32_t a = 1; a <= t; a++) {
            const float acc = grp[t - a];
            cs[b] = rbf(__fadd_rn(cs[b], acc));
            if (b + 1 < k) cs[b + 1] = rbf(__fadd_rn(cs[b + 1], acc));
        }
    }
    __syncthreads();
    // mask = (cum - probs) > p ; probs[mask] = 0 ; ids gathered (nn/sampling.h:192-199)
    if (t < k) {
        const float diff = rbf(__fsub_rn(cs[t], sv[t]));
        const float ps = diff > p.top_p ? 0.0f : sv[t];
        pr[t] = ps;
        const int32_t id = idx[si[t]];
        si[t] = id;
        if (p.probs_sorted) p.probs_sorted[size_t(row) * k + t] = f32_to_bf16_bits(ps);
        if (p.probs_idx) p.probs_idx[size_t(row) * k + t] = id;
    }
    __syncthreads();
    // multinomial_sampler, sample_size 1 (kernel/multinomial.metal:94-123)
    if (t == 0) {
        const int32_t step = p.step_counter ? *p.step_counter : 0;
        const float u = p.uniforms[size_t(step) * p.rows + row];
        const float a = pr[p.intended ? k - 1 : 0], b = pr[0];
        const float r = rbf(__fadd_rn(__fmul_rn(u, __fsub_rn(b, a)), a));
        int low = 0, high = int(k);
        while (low < high) {
            const int mid = (low + high) / 2;
            if (pr[mid] > r) low = mid + 1;
            else high = mid;
        }
        const int32_t ch = (low > 1 ? low : 1) - 1;
        const int32_t tok = si[ch];
        if (p.choice) p.choice[row] = ch;
        if (p.token) p.token[row] = tok;
        if (p.out_log) p.out_log[size_t(step) * p.rows + row] = tok;
        if (p.advance) {
            p.ids[row] = tok;
            p.pos[row] += 1;
        }
    }
}

// last launch of a sampled decode step: bump the step counter once all rows are done
MC_KERNEL void sample_step_kernel(int32_t* step_counter)
{
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x == 0) *step_counter += 1;
}

} // namespace mc
