// metalchat_b200/csrc/mc_ops.cu — op-level sm_100a kernels behind the reference's kernel names.
//
// One CUDA kernel per kernel of metalchat.metallib (kernel/*.metal, 71 host names), taking the
// same arguments in the same bind order: a tensor = tensor_layout<N>
// bytes + buffer, a scalar = raw bytes.  These serve the operator API (include/metalchat/
// kernel/*.h) one-for-one; the decode engine (mc_engine.cu) uses fused kernels instead.
// The Metal launch shape is ignored except where it defines the reduction partition
// (rmsnorm/softmax/sum block_size, cumsum_B, sort), which is reproduced so that results match
// the reference's partial-sum order.  Written from the kernels' observable semantics; no code
// is shared with the Metal sources.
#include "mc_common.cuh"

namespace mc {
namespace {

constexpr uint32_t kMaxThreads = 1024;

inline uint32_t ceil_div(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// grid for flat element-wise work: enough 256-thread CTAs to cover n, capped so that large
// tensors are grid-strided by a few resident waves of the 148 SMs
inline dim3 flat_grid(uint64_t n)
{
    const uint64_t blocks = (n + 255) / 256;
    const uint64_t cap = 148ull * 16;
    return dim3(unsigned(blocks < 1 ? 1 : (blocks > cap ? cap : blocks)));
}

// ---- threadgroup-wide fp32 sum with the reference's partition ----------------------------------
// simd_sum per 32 lanes (xor butterfly), per-simdgroup results through a zeroed 32-slot array,
// simd_sum again (kernel/rmsnorm.metal:58-82, kernel/softmax.metal:49-73, kernel/sum.metal:49-70).
__device__ __forceinline__ float warp_butterfly(float v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ float threadgroup_sum(float partial, float* tg /* [32] shared */)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 32) tg[threadIdx.x] = 0.0f;
    __syncthreads();
    const float s = warp_butterfly(partial);
    if (lane == 0) tg[warp] = s;
    __syncthreads();
    const float total = warp_butterfly(tg[lane]);
    __syncthreads();
    return total;
}

// ---- bmm (kernel/bmm.metal:24-82) ------------------------------------------------------------------
// C[b,m,n] = sum_k A[b,m,k] * B[b,k,n]: fp32 accumulation in ascending k from 0, one rounding.
// 16x16 output tile per CTA, operands staged through shared memory as fp32.
template <typename T>
__global__ void __launch_bounds__(256) bmm_kernel(tview<T, 3> out, tview<const T, 3> a, tview<const T, 3> b)
{
    __shared__ float sa[16][17];
    __shared__ float sb[16][17];
    const uint32_t M = a.size(1), K = a.size(2), N = b.size(2);
    const uint32_t bi = blockIdx.z;
    const uint32_t tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const uint32_t m = blockIdx.y * 16 + ty, n = blockIdx.x * 16 + tx;
    float acc = 0.0f;
    for (uint32_t k0 = 0; k0 < K; k0 += 16) {
        // A tile: row m, col k0+tx ; B tile: row k0+ty, col n
        sa[ty][tx] = (m < M && k0 + tx < K) ? to_f32(a.at(bi, m, k0 + tx)) : 0.0f;
        sb[ty][tx] = (k0 + ty < K && n < N) ? to_f32(b.at(bi, k0 + ty, n)) : 0.0f;
        __syncthreads();
#pragma unroll
        for (uint32_t kk = 0; kk < 16; kk++) acc = __fadd_rn(acc, __fmul_rn(sa[ty][kk], sb[kk][tx]));
        __syncthreads();
    }
    if (m < M && n < N) out.at(bi, m, n) = from_f32<T>(acc);
}
template <typename T> void launch_bmm(const arg_pack& p, cudaStream_t s)
{
    auto out = p.tensor<T, 3>(0);
    auto a = p.tensor<const T, 3>(2);
    auto b = p.tensor<const T, 3>(4);
    MC_REQUIRE(a.size(2) == b.size(1), "bmm: inner dimensions differ");
    // no broadcasting over the batch (kernel/bmm.h:36-39 same_dim(weight, 0)); a stride-0 batch view of a 2-D weight is fine
    MC_REQUIRE(a.size(0) == b.size(0), "bmm: batch mismatch");
    const uint32_t B = a.size(0), M = a.size(1), N = b.size(2);
    if (B == 0 || M == 0 || N == 0) return;
    MC_REQUIRE(B <= 65535, "bmm: batch too large");
    dim3 grid(ceil_div(N, 16), ceil_div(M, 16), B);
    bmm_kernel<T><<<grid, 256, 0, s>>>(out, a, b);
}

// ---- rmsnorm (kernel/rmsnorm.metal:28-91) -------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024) rmsnorm_kernel(
    tview<T, 2> out, tview<const T, 2> in, tview<const T, 1> w, float eps, float mu, uint32_t block, uint32_t threads
)
{
    __shared__ float tg[32];
    const uint32_t row = blockIdx.x, D = in.size(1), t = threadIdx.x;
    float part = 0.0f;
    if (t < threads) {
        const uint32_t begin = t * block;
        for (uint32_t j = begin; j < begin + block && j < D; j++) {
            const float x = to_f32(in.at(row, j));
            part = __fadd_rn(part, __fmul_rn(x, x));
        }
    }
    const float acc = threadgroup_sum(part, tg);
    const float mean_sq = acc / float(D);
    const float inv = 1.0f / sqrtf(__fadd_rn(mean_sq, eps));
    if (t < threads) {
        const uint32_t begin = t * block;
        for (uint32_t j = begin; j < begin + block && j < D; j++) {
            const float x = to_f32(in.at(row, j));
            const float weight = __fadd_rn(mu, to_f32(w.at(j)));
            out.at(row, j) = from_f32<T>(__fmul_rn(__fmul_rn(weight, x), inv));
        }
    }
}
inline uint32_t partition_threads(uint32_t D, uint32_t block, const char* what)
{
    MC_REQUIRE(block > 0, std::string(what) + ": block_size must be positive");
    const uint32_t threads = ceil_div(D, block);
    MC_REQUIRE(threads <= kMaxThreads, std::string(what) + ": row needs more than 1024 threads for this block_size");
    return threads;
}
template <typename T> void launch_rmsnorm(const arg_pack& p, cudaStream_t s)
{
    auto out = p.tensor<T, 2>(0);
    auto in = p.tensor<const T, 2>(2);
    auto w = p.tensor<const T, 1>(4);
    const float eps = p.scalar<float>(6), mu = p.scalar<float>(7);
    const uint32_t block = p.scalar<uint32_t>(8);
    MC_REQUIRE(w.size(0) == in.size(1), "rmsnorm: weight size differs from the row size");
    if (in.size(0) == 0 || in.size(1) == 0) return;
    const uint32_t threads = partition_threads(in.size(1), block, "rmsnorm");
    rmsnorm_kernel<T><<<in.size(0), ceil_div(threads, 32) * 32, 0, s>>>(out, in, w, eps, mu, block, threads);
}

// ---- softmax (kernel/softmax.metal:24-88): no max subtraction ---------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024) softmax_kernel(tview<T, 2> out, tview<const T, 2> in, uint32_t block, uint32_t threads)
{
    __shared__ float tg[32];
    const uint32_t row = blockIdx.x, D = in.size(1), t = threadIdx.x;
    float part = 0.0f;
    if (t < threads) {
        const uint32_t begin = t * block;
        for (uint32_t j = begin; j < begin + block && j < D; j++) part = __fadd_rn(part, expf(to_f32(in.at(row, j))));
    }
    const float inv = 1.0f / threadgroup_sum(part, tg);
    if (t < threads) {
        const uint32_t begin = t * block;
        for (uint32_t j = begin; j < begin + block && j < D; j++)
            out.at(row, j) = from_f32<T>(__fmul_rn(expf(to_f32(in.at(row, j))), inv));
    }
}
template <typename T> void launch_softmax(const arg_pack& p, cudaStream_t s)
{
    auto out = p.tensor<T, 2>(0);
    auto in = p.tensor<const T, 2>(2);
    const uint32_t block = p.scalar<uint32_t>(4);
    if (in.size(0) == 0 || in.size(1) == 0) return;
    const uint32_t threads = partition_threads(in.size(1), block, "softmax");
    softmax_kernel<T><<<in.size(0), ceil_div(threads, 32) * 32, 0, s>>>(out, in, block, threads);
}

// ---- sum (kernel/sum.metal:27-75) ----------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024) sum_kernel(tview<T, 1> out, tview<const T, 2> in, uint32_t block, uint32_t threads)
{
    __shared__ float tg[32];
    const uint32_t row = blockIdx.x, D = in.size(1), t = threadIdx.x;
    float part = 0.0f;
    if (t < threads) {
        const uint32_t begin = t * block;
        for (uint32_t j = begin; j < begin + block && j < D; j++) part = __fadd_rn(part, to_f32(in.at(row, j)));
    }
    const float total = threadgroup_sum(part, tg);
    if (t == 0) out.at(row) = from_f32<T>(total);
}
template <typename T> void launch_sum(const arg_pack& p, cudaStream_t s)
{
    auto out = p.tensor<T, 1>(0);
    auto in = p.tensor<const T, 2>(2);
    const uint32_t block = p.scalar<uint32_t>(4);
    if (in.size(0) == 0) return;
    const uint32_t threads = in.size(1) ? partition_threads(in.size(1), block, "sum") : 1;
    sum_kernel<T><<<in.size(0), ceil_div(threads, 32) * 32, 0, s>>>(out, in, block, threads);
}

// ---- rope (kernel/rope.metal:28-63): pairs (k, k + D/2), fp32 tables ---------------------------------------------
template <typename T>
__global__ void rope_kernel(
    tview<T, 2> out, tview<const T, 2> in, tview<const float, 2> fcos, tview<const float, 2> fsin, uint32_t bs,
    uint32_t n_head, uint32_t start_pos
)
{
    const uint32_t rows = in.size(0), half = fcos.size(1);
    const uint64_t n = uint64_t(rows) * half;
    for (uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; idx < n; idx += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t i = uint32_t(idx / half), k = uint32_t(idx % half);
        const uint32_t pos = i / (bs * n_head);
        const float x1 = to_f32(in.at(i, k)), x2 = to_f32(in.at(i, half + k));
        const float c = fcos.at(start_pos + pos, k), sn = fsin.at(start_pos + pos, k);
        out.at(i, k) = from_f32<T>(__fsub_rn(__fmul_rn(c, x1), __fmul_rn(sn, x2)));
        out.at(i, half + k) = from_f32<T>(__fadd_rn(__fmul_rn(sn, x1), __fmul_rn(c, x2)));
    }
}
template <typename T> void launch_rope(const arg_pack& p, cudaStream_t s)
{
    auto out = p.tensor<T, 2>(0);
    auto in = p.tensor<const T, 2>(2);
    auto fcos = p.tensor<const float, 2>(4);
    auto fsin = p.tensor<const float, 2>(6);
    const uint32_t bs = p.scalar<uint32_t>(8), n_head = p.scalar<uint32_t>(9), start_pos = p.scalar<uint32_t>(10);
    MC_REQUIRE(bs > 0 && n_head > 0, "rope: batch_size and n_head must be positive");
    MC_REQUIRE(in.size(1) == 2 * fcos.size(1), "rope: head dimension differs from 2 x table width");
    const uint64_t n = uint64_t(in.size(0)) * fcos.size(1);
    if (n == 0) return;
    const uint32_t last_pos = start_pos + (in.size(0) - 1) / (bs * n_head);
    MC_REQUIRE(last_pos < fcos.size(0) && last_pos < fsin.size(0), "rope: position outside the frequency tables");
    rope_kernel<T><<<flat_grid(n), 256, 0, s>>>(out, in, fcos, fsin, bs, n_head, start_pos);
}

// ---- rope_freqs (kernel/rope.metal:76-102) ---------------------------------------------------------------------------
__global__ void rope_freqs_kernel(tview<float, 2> fcos, tview<float, 2> fsin, uint32_t dim, uint32_t start_pos, float theta)
{
    const uint32_t rows = fcos.size(0), half = dim / 2;
    const uint64_t n = uint64_t(rows) * half;
    for (uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; idx < n; idx += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t i = uint32_t(idx / half), j = uint32_t(idx % half);
        const float freq = 1.0f / powf(theta, __fmul_rn(2.0f, float(j)) / float(dim));
        const float angle = __fmul_rn(float(start_pos + i), freq);
        fcos.at(i, j) = cosf(angle);
        fsin.at(i, j) = sinf(angle);
    }
}
void launch_rope_freqs(const arg_pack& p, cudaStream_t s)
{
    auto fcos = p.tensor<float, 2>(0);
    auto fsin = p.tensor<float, 2>(2);
    const uint32_t dim = p.scalar<uint32_t>(4), start_pos = p.scalar<uint32_t>(5);
    const float theta = p.scalar<float>(6);
    MC_REQUIRE(dim >= 2 && fcos.size(1) >= dim / 2 && fsin.size(1) >= dim / 2, "rope_freqs: tables narrower than dim/2");
    const uint64_t n = uint64_t(fcos.size(0)) * (dim / 2);
    if (n == 0) return;
    rope_freqs_kernel<<<flat_grid(n), 256, 0, s>>>(fcos, fsin, dim, start_pos, theta);
}

// ---- embedding (kernel/embedding.metal:38-66) -----------------------------------------------------------------------------
template <typename T>
__global__ void embedding_kernel(tview<T, 3> out, tview<const int32_t, 2> ids, tview<const T, 2> w, int* bad)
{
    const uint32_t J = ids.size(1), E = w.size(1);
    const uint32_t i = blockIdx.z;
    for (uint32_t j = blockIdx.y; j < J; j += gridDim.y) {
        const int32_t id = ids.at(i, j);
        if (id < 0 || uint32_t(id) >= w.size(0)) {
            if (threadIdx.x == 0 && blockIdx.x == 0) *reinterpret_cast<volatile int*>(bad) = 1; // mapped host word: a plain store
            continue;
        }
        for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < E; k += gridDim.x * blockDim.x)
            out.at(i, j, k) = w.at(uint32_t(id), k);
    }
}
template <typename T> void launch_embedding(const arg_pack& p, cudaStream_t s)
{
    auto out = p.tensor<T, 3>(0);
    auto ids = p.tensor<const int32_t, 2>(2);
    auto w = p.tensor<const T, 2>(4);
    if (ids.size(0) == 0 || ids.size(1) == 0 || w.size(1) == 0) return;
    MC_REQUIRE(ids.size(0) <= 65535, "embedding: batch too large");
    // out-of-range ids are skipped (the Metal kernel would read out of bounds) and reported by mc_wait; the flag is a mapped pinned
    // word owned by the device the kernel runs on
    MC_REQUIRE(p.dev != nullptr, "embedding: no device bound to the dispatch");
    if (!p.dev->bad_ids) {
        MC_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&p.dev->bad_ids_host), sizeof(int), cudaHostAllocMapped));
        *p.dev->bad_ids_host = 0;
        MC_CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&p.dev->bad_ids), p.dev->bad_ids_host, 0));
    }
    int* bad = p.dev->bad_ids;
    dim3 grid(ceil_div(w.size(1), 256) > 64 ? 64 : ceil_div(w.size(1), 256), ids.size(1) > 65535 ? 65535 : ids.size(1), ids.size(0));
    embedding_kernel<T><<<grid, 256, 0, s>>>(out, ids, w, bad);
}

// ---- sort (kernel/sort.metal:31-86): descending bitonic network, pad with -inf ----------------------------------------------
template <typename T> __device__ __forceinline__ bool less_than(T a, T b) { return to_f32(a) < to_f32(b); }
template <typename T> __device__ __forceinline__ T neg_inf();
template <> __device__ __forceinline__ float neg_inf<float>() { return __uint_as_float(0xff800000u); }
template <> __device__ __forceinline__ bf16 neg_inf<bf16>() { return bf16{0xff80}; }

// whole row in shared memory (P <= 4096)
template <typename T>
__global__ void __launch_bounds__(1024) sort_smem_kernel(tview<T, 2> values, tview<int32_t, 2> indices, tview<const T, 2> in)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t P = values.size(1), D = in.size(1), row = blockIdx.x;
    int32_t* si = reinterpret_cast<int32_t*>(smem_raw);
    T* sv = reinterpret_cast<T*>(si + P);
    for (uint32_t k = threadIdx.x; k < P; k += blockDim.x) {
        sv[k] = k < D ? in.at(row, k) : neg_inf<T>();
        si[k] = int32_t(k);
    }
    for (uint32_t k = 2; k <= P; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (uint32_t t = threadIdx.x; t < P / 2; t += blockDim.x) {
                // t-th pair (i, i^j) with i < i^j
                const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const uint32_t ij = i | j;
                const T a = sv[i], c = sv[ij];
                const bool up = (i & k) == 0;
                if (up ? less_than(a, c) : less_than(c, a)) {
                    sv[i] = c, sv[ij] = a;
                    const int32_t x = si[i];
                    si[i] = si[ij], si[ij] = x;
                }
            }
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < P; k += blockDim.x) {
        values.at(row, k) = sv[k];
        indices.at(row, k) = si[k];
    }
}
// large rows: the network runs in the output buffers (L2-resident), one CTA per row
template <typename T>
__global__ void __launch_bounds__(1024) sort_gmem_kernel(tview<T, 2> values, tview<int32_t, 2> indices, tview<const T, 2> in)
{
    const uint32_t P = values.size(1), D = in.size(1), row = blockIdx.x;
    for (uint32_t k = threadIdx.x; k < P; k += blockDim.x) {
        values.at(row, k) = k < D ? in.at(row, k) : neg_inf<T>();
        indices.at(row, k) = int32_t(k);
    }
    for (uint32_t k = 2; k <= P; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (uint32_t t = threadIdx.x; t < P / 2; t += blockDim.x) {
                const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const uint32_t ij = i | j;
                T& ra = values.at(row, i);
                T& rc = values.at(row, ij);
                const T a = ra, c = rc;
                const bool up = (i & k) == 0;
                if (up ? less_than(a, c) : less_than(c, a)) {
                    ra = c, rc = a;
                    int32_t& xa = indices.at(row, i);
                    int32_t& xc = indices.at(row, ij);
                    const int32_t x = xa;
                    xa = xc, xc = x;
                }
            }
        }
    }
}
template <typename T> void launch_sort(const arg_pack& p, cudaStream_t s)
{
    auto values = p.tensor<T, 2>(0);
    auto indices = p.tensor<int32_t, 2>(2);
    auto in = p.tensor<const T, 2>(4);
    const uint32_t P = values.size(1);
    MC_REQUIRE(P > 0 && (P & (P - 1)) == 0, "sort: output row size must be a power of two");
    MC_REQUIRE(in.size(1) <= P && indices.size(1) == P, "sort: output narrower than the input row");
    if (in.size(0) == 0) return;
    if (P <= 4096) {
        const uint32_t threads = P / 2 < 32 ? 32 : (P / 2 > 1024 ? 1024 : P / 2);
        sort_smem_kernel<T><<<in.size(0), threads, size_t(P) * (4 + sizeof(T)), s>>>(values, indices, in);
    } else {
        sort_gmem_kernel<T><<<in.size(0), 1024, 0, s>>>(values, indices, in);
    }
}

// ---- cumsum_B (kernel/cumsum.metal:24-98): accumulates in T ----------------------------------------------------------------------
// Thread t owns elements [t*B, (t+1)*B): a serial prefix in T, then the totals of all preceding
// blocks are added one at a time, nearest block first (kernel/cumsum.metal:45-67).  The output
// row is used as the per-thread scratch.  The reference's group_sums[256] limit (quirk Q11) is
// lifted to 1024 threads.
template <typename T>
__global__ void __launch_bounds__(1024) cumsum_kernel(tview<T, 2> out, tview<const T, 2> in, uint32_t B, uint32_t threads)
{
    __shared__ float group[1024];
    const uint32_t row = blockIdx.x, D = in.size(1), t = threadIdx.x;
    const uint32_t begin = t * B, end = (begin + B < D) ? begin + B : D;
    if (t < threads) {
        float run = 0.0f;
        for (uint32_t k = begin; k < end; k++) {
            const float x = to_f32(in.at(row, k));
            run = (k == begin) ? x : round_to<T>(__fadd_rn(x, run));
            out.at(row, k) = from_f32<T>(run);
        }
        group[t] = run;
    }
    __syncthreads();
    if (t < threads) {
        for (uint32_t k = begin; k < end; k++) {
            float v = to_f32(out.at(row, k));
            for (uint32_t a = 1; a <= t; a++) v = round_to<T>(__fadd_rn(v, group[t - a]));
            out.at(row, k) = from_f32<T>(v);
        }
    }
}
template <typename T, uint32_t B> void launch_cumsum(const arg_pack& p, cudaStream_t s)
{
    auto out = p.tensor<T, 2>(0);
    auto in = p.tensor<const T, 2>(2);
    if (in.size(0) == 0 || in.size(1) == 0) return;
    const uint32_t threads = partition_threads(in.size(1), B, "cumsum");
    cumsum_kernel<T><<<in.size(0), ceil_div(threads, 32) * 32, 0, s>>>(out, in, B, threads);
}

// ---- multinomial (kernel/multinomial.metal:15-123) -----------------------------------------------------------------------------------
struct pcg32 {
    uint64_t state, inc;
    __device__ uint32_t next()
    {
        const uint64_t pre = state;
        state = pre * 6364136223846793005ull + inc;
        const uint32_t xs = uint32_t(((pre >> 18u) ^ pre) >> 27u);
        const uint32_t rot = uint32_t(pre >> 59u);
        return (xs >> rot) | (xs << ((~rot + 1u) & 31u));
    }
    __device__ pcg32(uint64_t init_state, uint64_t init_seq) : state(0), inc((init_seq << 1u) | 1u)
    {
        next();
        state += init_state;
        next();
    }
    __device__ float uniform() { return __uint_as_float((next() >> 9) | 0x3f800000u) - 1.0f; }
};
template <typename T>
__global__ void multinomial_kernel(tview<int32_t, 2> out, tview<const T, 2> in, uint64_t init_state, uint64_t init_seq, uint64_t avail)
{
    const uint32_t rows = out.size(0), S = out.size(1), N = in.size(1);
    const uint64_t n = uint64_t(rows) * S;
    for (uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; idx < n; idx += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t i = uint32_t(idx / S), k = uint32_t(idx % S);
        // quirk Q10: `a` is read at column sample_size-1, not at the last column (kernel/multinomial.metal:107).  With more samples than
        // columns that address lies in a later row or outside the buffer, where a Metal device read yields 0 -- the reference's own
        // test draws 8192 samples from rows of 5 and relies on it (test/test_kernel_multinomial.cc:16-54)
        const uint64_t flat = uint64_t(i) * in.l.strides[0] + uint64_t(S - 1) * in.l.strides[1] + in.l.offsets[0] + in.l.offsets[1];
        const float a = flat < avail ? to_f32(in.data[flat]) : 0.0f;
        const float b = to_f32(in.at(i, 0));
        pcg32 g(init_state + i, init_seq + k);
        const float r = round_to<T>(__fadd_rn(__fmul_rn(g.uniform(), __fsub_rn(b, a)), a));
        int low = 0, high = int(N);
        while (low < high) {
            const int mid = (low + high) / 2;
            if (to_f32(in.at(i, uint32_t(mid))) > r) low = mid + 1;
            else high = mid;
        }
        out.at(i, k) = (low > 1 ? low : 1) - 1;
    }
}
template <typename T> void launch_multinomial(const arg_pack& p, cudaStream_t s)
{
    auto out = p.tensor<int32_t, 2>(0);
    auto in = p.tensor<const T, 2>(2);
    const uint64_t st = p.scalar<uint64_t>(4), sq = p.scalar<uint64_t>(5);
    const uint64_t n = uint64_t(out.size(0)) * out.size(1);
    if (n == 0) return;
    MC_REQUIRE(in.size(1) > 0, "multinomial: empty distribution");
    const arg_slot& buf = p.slot[3];
    const uint64_t avail = (buf.buf->size - buf.offset) / sizeof(T); // elements of the bound buffer readable from the tensor's base
    multinomial_kernel<T><<<flat_grid(n), 256, 0, s>>>(out, in, st, sq, avail);
}

// ---- element-wise (kernel/mul.metal, arithmetic.metal, activation.metal, logical.metal, copy.metal) ----------------------------------------
enum { OP_ADD, OP_SUB, OP_DIV, OP_MUL };
template <typename T, int OP>
__global__ void binary_kernel(tview<T, 2> out, tview<const T, 2> a, tview<const T, 2> b)
{
    const uint32_t cols = a.size(1);
    const uint64_t n = uint64_t(a.size(0)) * cols;
    for (uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; idx < n; idx += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t i = uint32_t(idx / cols), k = uint32_t(idx % cols);
        const float x = to_f32(a.at(i, k)), y = to_f32(b.at(i, k));
        float r;
        if (OP == OP_ADD) r = __fadd_rn(x, y);
        else if (OP == OP_SUB) r = __fsub_rn(x, y);
        else if (OP == OP_DIV) r = __fdiv_rn(x, y);
        else r = __fmul_rn(x, y);
        out.at(i, k) = from_f32<T>(r);
    }
}
template <typename T, int OP> void launch_binary(const arg_pack& p, cudaStream_t s)
{
    auto out = p.tensor<T, 2>(0);
    auto a = p.tensor<const T, 2>(2);
    auto b = p.tensor<const T, 2>(4);
    MC_REQUIRE(a.size(0) == b.size(0) && a.size(1) == b.size(1), "binary kernel: operand shapes differ");
    const uint64_t n = uint64_t(a.size(0)) * a.size(1);
    if (n == 0) return;
    binary_kernel<T, OP><<<flat_grid(n), 256, 0, s>>>(out, a, b);
}

template <typename T>
__global__ void add_broadcast_kernel(tview<T, 2> out, tview<const T, 2> a, tview<const T, 1> b)
{
    const uint32_t cols = a.size(1), nb = b.size(0);
    const uint64_t n = uint64_t(a.size(0)) * cols;
    for (uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; idx < n; idx += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t i = uint32_t(idx / cols), j = uint32_t(idx % cols);
        out.at(i, j) = from_f32<T>(__fadd_rn(to_f32(a.at(i, j)), to_f32(b.at(j % nb))));
    }
}
template <typename T> void launch_add_broadcast(const arg_pack& p, cudaStream_t s)
{
    auto out = p.tensor<T, 2>(0);
    auto a = p.tensor<const T, 2>(2);
    auto b = p.tensor<const T, 1>(4);
    const uint64_t n = uint64_t(a.size(0)) * a.size(1);
    if (n == 0) return;
    MC_REQUIRE(b.size(0) > 0, "add_broadcast: empty broadcast operand");
    add_broadcast_kernel<T><<<flat_grid(n), 256, 0, s>>>(out, a, b);
}

// hadamard_broadcast (kernel/mul.metal:59-85) — the dequantisation kernel: the scale is rounded to
// O first, the product is rounded to O (quirk Q7).
template <typename O, typename S>
__global__ void hadamard_broadcast_kernel(tview<O, 2> out, tview<const int8_t, 2> a, tview<const S, 1> b)
{
    const uint32_t cols = a.size(1), nb = b.size(0);
    const uint64_t n = uint64_t(a.size(0)) * cols;
    for (uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; idx < n; idx += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t i = uint32_t(idx / cols), j = uint32_t(idx % cols);
        const float sc = round_to<O>(to_f32(b.at(i % nb)));
        const float q = round_to<O>(float(a.at(i, j)));
        out.at(i, j) = from_f32<O>(__fmul_rn(q, sc));
    }
}
template <typename O, typename S> void launch_hadamard_broadcast(const arg_pack& p, cudaStream_t s)
{
    auto out = p.tensor<O, 2>(0);
    auto a = p.tensor<const int8_t, 2>(2);
    auto b = p.tensor<const S, 1>(4);
    const uint64_t n = uint64_t(a.size(0)) * a.size(1);
    if (n == 0) return;
    MC_REQUIRE(b.size(0) > 0, "hadamard_broadcast: empty scale operand");
    hadamard_broadcast_kernel<O, S><<<flat_grid(n), 256, 0, s>>>(out, a, b);
}

enum { U_SCALAR_MUL, U_SILU, U_GELU, U_COPY };
template <typename T, int OP>
__global__ void unary_kernel(tview<T, 2> out, tview<const T, 2> a, T c)
{
    const uint32_t cols = a.size(1);
    const uint64_t n = uint64_t(a.size(0)) * cols;
    for (uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; idx < n; idx += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t i = uint32_t(idx / cols), k = uint32_t(idx % cols);
        const T xv = a.at(i, k);
        if (OP == U_COPY) {
            out.at(i, k) = xv;
            continue;
        }
        const float x = to_f32(xv);
        float r;
        if (OP == U_SCALAR_MUL) {
            r = __fmul_rn(x, to_f32(c));
        } else if (OP == U_SILU) {
            // x / (T(1) + T(exp(-x))) evaluated in T (kernel/activation.metal:34-35, quirk Q5)
            const float e = round_to<T>(expf(-x));
            const float d = round_to<T>(__fadd_rn(1.0f, e));
            r = __fdiv_rn(x, d);
        } else {
            // tanh approximation in fp32 (kernel/activation.metal:59-72)
            const float beta = 1.41421356237309504880f * 1.12837916709551257390f * 0.5f;
            const float x3 = __fmul_rn(__fmul_rn(x, x), x);
            const float inner = __fmul_rn(beta, __fadd_rn(x, __fmul_rn(0.044715f, x3)));
            r = __fmul_rn(__fmul_rn(0.5f, x), __fadd_rn(1.0f, tanhf(inner)));
        }
        out.at(i, k) = from_f32<T>(r);
    }
}
template <typename T, int OP> void launch_unary(const arg_pack& p, cudaStream_t s)
{
    auto out = p.tensor<T, 2>(0);
    auto a = p.tensor<const T, 2>(2);
    T c{};
    if (OP == U_SCALAR_MUL) c = p.scalar<T>(4);
    const uint64_t n = uint64_t(a.size(0)) * a.size(1);
    if (n == 0) return;
    if (OP == U_COPY) MC_REQUIRE(uint64_t(out.size(0)) * out.size(1) == n && out.size(1) == a.size(1), "copy: output shape differs from the input");
    unary_kernel<T, OP><<<flat_grid(n), 256, 0, s>>>(out, a, c);
}

template <typename T> __global__ void scatter_kernel(tview<T, 2> out, tview<const uint8_t, 2> mask, T value)
{
    const uint32_t cols = out.size(1);
    const uint64_t n = uint64_t(out.size(0)) * cols;
    for (uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; idx < n; idx += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t i = uint32_t(idx / cols), k = uint32_t(idx % cols);
        if (mask.at(i, k)) out.at(i, k) = value;
    }
}
template <typename T> void launch_scatter(const arg_pack& p, cudaStream_t s)
{
    auto out = p.tensor<T, 2>(0);
    auto mask = p.tensor<const uint8_t, 2>(2);
    const T value = p.scalar<T>(4);
    const uint64_t n = uint64_t(out.size(0)) * out.size(1);
    if (n == 0) return;
    MC_REQUIRE(mask.size(0) == out.size(0) && mask.size(1) == out.size(1), "scatter: mask shape differs from the output");
    scatter_kernel<T><<<flat_grid(n), 256, 0, s>>>(out, mask, value);
}

template <typename T>
__global__ void gather_kernel(tview<T, 2> out, tview<const T, 2> in, tview<const int32_t, 2> index)
{
    const uint32_t cols = index.size(1), width = in.size(1);
    const uint64_t n = uint64_t(index.size(0)) * cols;
    for (uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; idx < n; idx += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t i = uint32_t(idx / cols), k = uint32_t(idx % cols);
        const int32_t src = index.at(i, k);
        if (src >= 0 && uint32_t(src) < width) out.at(i, k) = in.at(i, uint32_t(src));
    }
}
template <typename T> void launch_gather(const arg_pack& p, cudaStream_t s)
{
    auto out = p.tensor<T, 2>(0);
    auto in = p.tensor<const T, 2>(2);
    auto index = p.tensor<const int32_t, 2>(4);
    const uint64_t n = uint64_t(index.size(0)) * index.size(1);
    if (n == 0) return;
    gather_kernel<T><<<flat_grid(n), 256, 0, s>>>(out, in, index);
}

template <typename T, bool GT> __global__ void compare_kernel(tview<uint8_t, 2> out, tview<const T, 2> a, T value)
{
    const uint32_t cols = a.size(1);
    const uint64_t n = uint64_t(a.size(0)) * cols;
    const float v = to_f32(value);
    for (uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; idx < n; idx += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t i = uint32_t(idx / cols), k = uint32_t(idx % cols);
        const float x = to_f32(a.at(i, k));
        out.at(i, k) = GT ? (x > v) : (x <= v);
    }
}
template <typename T, bool GT> void launch_compare(const arg_pack& p, cudaStream_t s)
{
    auto out = p.tensor<uint8_t, 2>(0);
    auto a = p.tensor<const T, 2>(2);
    const T value = p.scalar<T>(4);
    const uint64_t n = uint64_t(a.size(0)) * a.size(1);
    if (n == 0) return;
    compare_kernel<T, GT><<<flat_grid(n), 256, 0, s>>>(out, a, value);
}

// ---- roll (kernel/roll.metal:22-45) ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void roll_kernel(tview<T, 1> out, tview<const T, 1> in, uint32_t shift, uint32_t size, uint32_t stride)
{
    const uint32_t n = in.size(0);
    const uint32_t stride_size = size * stride;
    for (uint64_t k64 = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; k64 < n; k64 += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t k = uint32_t(k64);
        const uint32_t base = (k / stride_size) * stride_size;
        const uint32_t i = (k / stride + shift) % size;
        const uint32_t j = k % stride;
        out.at(k) = in.at(base + i * stride + j);
    }
}
template <typename T> void launch_roll(const arg_pack& p, cudaStream_t s)
{
    auto out = p.tensor<T, 1>(0);
    auto in = p.tensor<const T, 1>(2);
    const uint32_t shift = p.scalar<uint32_t>(4), size = p.scalar<uint32_t>(5), stride = p.scalar<uint32_t>(6);
    if (in.size(0) == 0) return;
    MC_REQUIRE(size > 0 && stride > 0, "roll: size and stride must be positive");
    roll_kernel<T><<<flat_grid(in.size(0)), 256, 0, s>>>(out, in, shift, size, stride);
}

std::vector<kernel_entry> build_registry()
{
    std::vector<kernel_entry> r;
#define BOTH(name, fn) \
    r.push_back({name "_bfloat", fn<bf16>}); \
    r.push_back({name "_float", fn<float>});
    BOTH("bmm_8", launch_bmm)
    BOTH("rmsnorm", launch_rmsnorm)
    BOTH("softmax", launch_softmax)
    BOTH("sum", launch_sum)
    BOTH("rope", launch_rope)
    r.push_back({"rope_freqs_float", launch_rope_freqs});
    BOTH("embedding", launch_embedding)
    BOTH("sort", launch_sort)
    BOTH("multinomial", launch_multinomial)
    BOTH("add_broadcast", launch_add_broadcast)
    BOTH("scatter", launch_scatter)
    BOTH("roll", launch_roll)
#undef BOTH
#define BOTH2(name, fn, arg) \
    r.push_back({name "_bfloat", fn<bf16, arg>}); \
    r.push_back({name "_float", fn<float, arg>});
    BOTH2("add", launch_binary, OP_ADD)
    BOTH2("sub", launch_binary, OP_SUB)
    BOTH2("div", launch_binary, OP_DIV)
    BOTH2("hadamard", launch_binary, OP_MUL)
    BOTH2("scalar_mul", launch_unary, U_SCALAR_MUL)
    BOTH2("silu", launch_unary, U_SILU)
    BOTH2("gelu", launch_unary, U_GELU)
    BOTH2("copy", launch_unary, U_COPY)
    BOTH2("gt", launch_compare, true)
    BOTH2("le", launch_compare, false)
    BOTH2("cumsum_2", launch_cumsum, 2)
    BOTH2("cumsum_4", launch_cumsum, 4)
    BOTH2("cumsum_8", launch_cumsum, 8)
    BOTH2("cumsum_16", launch_cumsum, 16)
    BOTH2("cumsum_32", launch_cumsum, 32)
    BOTH2("cumsum_64", launch_cumsum, 64)
    BOTH2("cumsum_128", launch_cumsum, 128)
    BOTH2("cumsum_256", launch_cumsum, 256)
    BOTH2("cumsum_512", launch_cumsum, 512)
    BOTH2("cumsum_1024", launch_cumsum, 1024)
#undef BOTH2
    // int32 storage is copied/gathered as 4-byte words
    r.push_back({"copy_int32_t", launch_unary<float, U_COPY>});
    r.push_back({"gather_bfloat", launch_gather<bf16>});
    r.push_back({"gather_float", launch_gather<float>});
    r.push_back({"gather_int32_t", launch_gather<float>});
    r.push_back({"hadamard_broadcast_bfloat_int8_t_bfloat", launch_hadamard_broadcast<bf16, bf16>});
    r.push_back({"hadamard_broadcast_bfloat_int8_t_float", launch_hadamard_broadcast<bf16, float>});
    r.push_back({"hadamard_broadcast_float_int8_t_bfloat", launch_hadamard_broadcast<float, bf16>});
    r.push_back({"hadamard_broadcast_float_int8_t_float", launch_hadamard_broadcast<float, float>});
    return r;
}

} // namespace

const std::vector<kernel_entry>& kernel_registry()
{
    static const std::vector<kernel_entry> r = build_registry();
    return r;
}

} // namespace mc
