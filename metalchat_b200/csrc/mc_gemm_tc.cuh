// metalchat_b200/csrc/mc_gemm_tc.cuh — prefill on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Y[M, N] = epilogue( X[M, K] . W[N, K]^T )   bf16 x bf16 -> fp32 accumulators in tensor memory -> ONE bf16 rounding,
// i.e. the rounding point of kernel/bmm.metal:76 (the fp32 sum is re-associated by the tensor core: stated tolerance).
//
// Persistent, warp-specialised kernel, one CTA per SM, 192 threads:
//   warp 0      TMA producer: cp.async.bulk.tensor.2d (SASS UTMALDG) moves a 128 x 64 tile of X and a BN x 64 tile of W
//               (both K-major, 128-byte swizzle) into a ring of stages guarded by full/empty mbarriers;
//   warp 1      MMA issuer: one thread issues tcgen05.mma (cta_group::1, kind::f16, M=128, N=BN, K=16), four per stage;
//               tcgen05.commit frees the stage and, after the last k block, publishes the accumulator;
//   warps 2..5  epilogue: tcgen05.ld (32 lanes x 32 columns) -> r(.) -> fused tail (store | residual add | SiLU*mul on
//               the row-interleaved w1|w3) -> global.  The accumulator is double-buffered in TMEM (2 x BN columns), so the
//               epilogue of tile i overlaps the MMAs of tile i+1.
// Tiles are dealt m-fastest: the CTAs running at the same time share a handful of W tiles, so every weight byte comes from
// HBM once per prefill chunk and the activations (a few MB) stay in L2.
// The row-wise kernels around the GEMMs (RMSNorm, RoPE + KV append) and the causal prefill attention (mma.sync, the
// reference's rounding points: nn/attention.h:195-200) are at the end of the file.
#pragma once
#include <cuda.h>
#include "mc_common.cuh"
#include "mc_prefill.h"

namespace mc {
namespace tc {

// epilogues of the prefill GEMM (same meaning as the EPI_* of the decode GEMV kernels)
enum { EPI_NONE = 0, EPI_RESIDUAL = 2, EPI_SWIGLU = 3, EPI_PARTIAL = 4 }; // EPI_PARTIAL: unrounded fp32 sums (tensor parallel, row-parallel linear)

// ---- helpers (this translation unit is independent of the decode kernels) -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t a, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ float rbf(float f) { return bf16_bits_to_f32(f32_to_bf16_bits(f)); } // r(.)
__device__ __forceinline__ uint32_t pack2(float lo, float hi) { return uint32_t(f32_to_bf16_bits(lo)) | (uint32_t(f32_to_bf16_bits(hi)) << 16); }
// silu evaluated in bf16 steps: x / (T(1) + T(exp(-x)))  (kernel/activation.metal:34-35, quirk Q5)
__device__ __forceinline__ float silu_bf16(float g)
{
    const float e = rbf(expf(-g));
    const float d = rbf(__fadd_rn(1.0f, e));
    return rbf(__fdiv_rn(g, d));
}
// block-wide sum of 256 threads with a fixed partition (warp butterfly, then the 8 warp totals in order)
__device__ __forceinline__ float block_sum_256(float v, float* scratch /* [8] */)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) t += scratch[i];
    __syncthreads();
    return t;
}

// programmatic dependent launch: let the next kernel of the stream start its prologue / wait for the previous kernel's results
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

constexpr int kTcBM = 128;      // prompt rows per tile (the M of the MMA = the 128 TMEM lanes)
constexpr int kTcBK = 64;       // k per stage: 128 bytes per row = one swizzle atom, four MMAs of K = 16
constexpr int kTcThreads = 192; // producer warp | MMA warp | 4 epilogue warps
constexpr int kTcHdr = 1024;    // barriers + TMEM base slot

// AROWS = prompt rows actually staged per A tile.  A prompt chunk stages all 128; a decode batch of <= 32 rows stages 32 and lets
// the MMA read the remaining (stale) rows of its 128-row operand window -- they only reach accumulator rows nobody reads -- so
// that a stage costs 4 KiB + the weight tile and the ring holds several times more weight bytes in flight (HBM-bound regime).
constexpr int kTcRing = 192 * 1024;
__host__ __device__ constexpr int tc_stage_bytes(int BN, int AROWS) { return (AROWS + BN) * kTcBK * 2; }
__host__ __device__ constexpr int tc_a_slack(int AROWS) { return (kTcBM - AROWS) * kTcBK * 2; }
__host__ __device__ constexpr int tc_stages(int BN, int AROWS)
{
    return (kTcRing - tc_a_slack(AROWS)) / tc_stage_bytes(BN, AROWS) > 24 ? 24 : (kTcRing - tc_a_slack(AROWS)) / tc_stage_bytes(BN, AROWS);
}
__host__ __device__ constexpr int tc_smem_bytes(int BN, int AROWS) { return kTcHdr + tc_stages(BN, AROWS) * tc_stage_bytes(BN, AROWS) + tc_a_slack(AROWS) + 1024; }

struct gemm_tc_params {
    uint16_t* Y;         // EPI_NONE / EPI_RESIDUAL: [M, ldy]; EPI_SWIGLU: [M, ldy] holding N/2 columns
    const uint16_t* res; // EPI_RESIDUAL: [M, ldy]
    uint32_t M, N, K, ldy;
    int* err;            // set when a bounded wait times out (never hang the device)
};

// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, 8-row groups 1024 bytes apart (UMMA descriptor v1)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= uint64_t((smem_addr >> 4) & 0x3fffu);  // start address
    d |= uint64_t(1) << 16;                     // leading byte offset: unused for swizzled K-major operands
    d |= uint64_t(1024 >> 4) << 32;             // stride byte offset between 8-row core-matrix groups
    d |= uint64_t(1) << 46;                     // descriptor version (Blackwell)
    d |= uint64_t(2) << 61;                     // layout: SWIZZLE_128B
    return d;
}
// instruction descriptor: D = fp32, A = B = bf16, both K-major, M x N
__host__ __device__ constexpr uint32_t tc_idesc(uint32_t M, uint32_t N)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b),
                 "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t mbar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void tc_tma_load(uint32_t dst, const CUtensorMap* map, uint32_t c0, uint32_t c1, uint32_t mbar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(mbar), "r"(c0), "r"(c1)
                 : "memory");
}
// bounded mbarrier wait: a lost arrival raises the error flag and lets every later wait fall through
__device__ __forceinline__ void tc_wait(uint32_t bar, uint32_t parity, volatile int* dead, int* err)
{
    if (mbar_try_wait(bar, parity)) return;
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 1023u) == 0) {
            if (*dead) return;
            if (spins > (1u << 22)) {
                *dead = 1;
                atomicExch(err, 5);
                return;
            }
        }
    }
}

template <int BN, int EPI, int AROWS>
__global__ void __launch_bounds__(kTcThreads, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const gemm_tc_params p)
{
    constexpr int STAGES = tc_stages(BN, AROWS);
    constexpr uint32_t A_BYTES = AROWS * kTcBK * 2, B_BYTES = BN * kTcBK * 2;
    extern __shared__ unsigned char smem_raw[];
    // 128-byte swizzle: tiles on 1024-byte boundaries
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t base = smem_u32(smem);
    const uint32_t bar_full = base, bar_empty = base + 192, bar_tfull = base + 384, bar_tempty = base + 400, tmem_slot = base + 416;
    volatile int* dead = reinterpret_cast<volatile int*>(smem + 432);
    const uint32_t a0 = base + kTcHdr, b0 = a0 + STAGES * A_BYTES + tc_a_slack(AROWS);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(bar_full + s * 8, 1), mbar_init(bar_empty + s * 8, 1);
        for (int s = 0; s < 2; s++) mbar_init(bar_tfull + s * 8, 1), mbar_init(bar_tempty + s * 8, 4);
        *dead = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
    }
    if (warp == 1) {
        // two accumulator buffers of BN fp32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(uint32_t(2 * BN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 416);

    const uint32_t m_blocks = (p.M + kTcBM - 1) / kTcBM, n_blocks = (p.N + BN - 1) / BN;
    const uint32_t n_tiles = m_blocks * n_blocks, k_blocks = p.K / kTcBK;

    pdl_trigger();
    if (warp == 0) {
        if (lane == 0) {
            // every tile walks k from its own starting block: CTAs that share an X (or W) tile would otherwise request the same L2
            // lines in lockstep.  The start depends only on the 256-column group and the 128-row block of the outputs, so a row's
            // bits do not depend on the tile width or on how many rows share the call.
            const uint32_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
            const uint32_t total = my_tiles * k_blocks;
            auto coords = [&](uint32_t it, uint32_t& kx, uint32_t& mrow, uint32_t& nrow) {
                const uint32_t tile = blockIdx.x + (it / k_blocks) * gridDim.x, kb = it % k_blocks;
                const uint32_t mb = tile % m_blocks, nb = tile / m_blocks;
                const uint32_t rot = (((nb * BN) >> 8) * 11u + mb * 5u) % k_blocks;
                kx = (kb + rot < k_blocks ? kb + rot : kb + rot - k_blocks) * kTcBK, mrow = mb * kTcBM, nrow = nb * BN;
            };
            // the weights do not depend on the previous kernel of the stream: the first ring-full of W tiles is requested before
            // griddepcontrol.wait, the X tiles (its output) after
            const uint32_t pre = total < uint32_t(STAGES) ? total : uint32_t(STAGES);
            uint32_t kx, mrow, nrow;
            for (uint32_t it = 0; it < pre; it++) {
                coords(it, kx, mrow, nrow);
                mbar_expect_tx(bar_full + it * 8, A_BYTES + B_BYTES);
                tc_tma_load(b0 + it * B_BYTES, &tmW, kx, nrow, bar_full + it * 8);
            }
            pdl_sync();
            for (uint32_t it = 0; it < pre; it++) {
                coords(it, kx, mrow, nrow);
                tc_tma_load(a0 + it * A_BYTES, &tmX, kx, mrow, bar_full + it * 8);
            }
            for (uint32_t it = pre; it < total; it++) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                coords(it, kx, mrow, nrow);
                tc_wait(bar_empty + s * 8, ph ^ 1u, dead, p.err);
                mbar_expect_tx(bar_full + s * 8, A_BYTES + B_BYTES);
                tc_tma_load(a0 + s * A_BYTES, &tmX, kx, mrow, bar_full + s * 8);
                tc_tma_load(b0 + s * B_BYTES, &tmW, kx, nrow, bar_full + s * 8);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc_idesc(kTcBM, BN);
            uint32_t it = 0, lt = 0;
            for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, lt++) {
                const uint32_t as = lt & 1u, aph = (lt >> 1) & 1u;
                tc_wait(bar_tempty + as * 8, aph ^ 1u, dead, p.err); // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem + as * BN;
                for (uint32_t kb = 0; kb < k_blocks; kb++, it++) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                    tc_wait(bar_full + s * 8, ph, dead, p.err);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t da = tc_smem_desc(a0 + s * A_BYTES), db = tc_smem_desc(b0 + s * B_BYTES);
#pragma unroll
                    for (uint32_t k = 0; k < kTcBK / 16; k++) // 16 elements = 32 bytes along k inside the swizzle atom: +2 in the address field
                        tc_mma(d_tmem, da + k * 2, db + k * 2, idesc, (kb | k) != 0);
                    tc_commit(bar_empty + s * 8); // arrives when these MMAs have read the stage
                }
                tc_commit(bar_tfull + as * 8); // ... and when the accumulator is complete
            }
        }
    } else {
        const uint32_t quarter = warp & 3u; // a warp reaches TMEM lanes 32 (warp % 4) .. +31
        uint32_t lt = 0;
        pdl_sync(); // the residual rows and the output buffer belong to earlier kernels of the stream until they have finished
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, lt++) {
            const uint32_t mb = tile % m_blocks, nb = tile / m_blocks;
            const uint32_t as = lt & 1u, aph = (lt >> 1) & 1u;
            tc_wait(bar_tfull + as * 8, aph, dead, p.err);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t row = mb * kTcBM + quarter * 32 + lane;
            const uint32_t n0 = nb * BN;
#pragma unroll 1
            for (uint32_t cb = 0; cb < uint32_t(BN); cb += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem + ((quarter * 32) << 16) + as * BN + cb;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                      "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                      "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                      "=r"(v[31])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (EPI == EPI_PARTIAL) {
                    // this rank's k range of a row-parallel linear: the fp32 sums go out unrounded (Y is a float matrix, ldy in floats);
                    // tp_allreduce_rows_kernel sums the ranks in rank order and applies the one rounding of kernel/bmm.metal:76
                    if (row < p.M && n0 + cb < p.N) {
                        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<float*>(p.Y) + size_t(row) * p.ldy + n0 + cb);
#pragma unroll
                        for (int q = 0; q < 8; q++) dst[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    }
                } else if (row < p.M && n0 + cb < p.N) {
                    float y[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) y[j] = rbf(__uint_as_float(v[j])); // the bmm output buffer is T (kernel/bmm.metal:76)
                    if (EPI == EPI_SWIGLU) {
                        // z = r(silu_T(g) * u), columns (2i, 2i+1) = (w1 row i, w3 row i)  (nn/transformer.h:57-59)
                        uint32_t o[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) o[j] = pack2(__fmul_rn(silu_bf16(y[4 * j]), y[4 * j + 1]), __fmul_rn(silu_bf16(y[4 * j + 2]), y[4 * j + 3]));
                        uint4* dst = reinterpret_cast<uint4*>(p.Y + size_t(row) * p.ldy + ((n0 + cb) >> 1));
                        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
                    } else {
                        uint32_t o[16];
                        if (EPI == EPI_RESIDUAL) {
                            // h = r(x + a)  (nn/transformer.h:133,139)
                            const uint4* rp = reinterpret_cast<const uint4*>(p.res + size_t(row) * p.ldy + n0 + cb);
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const uint4 r = rp[q];
                                const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                                for (int j = 0; j < 4; j++)
                                    o[q * 4 + j] = pack2(__fadd_rn(bf_lo(rw[j]), y[(q * 4 + j) * 2]), __fadd_rn(bf_hi(rw[j]), y[(q * 4 + j) * 2 + 1]));
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; j++) o[j] = pack2(y[2 * j], y[2 * j + 1]);
                        }
                        uint4* dst = reinterpret_cast<uint4*>(p.Y + size_t(row) * p.ldy + n0 + cb);
#pragma unroll
                        for (int q = 0; q < 4; q++) dst[q] = make_uint4(o[q * 4], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + as * 8);
        }
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(uint32_t(2 * BN)) : "memory");
    }
}

// ---- decode-batch GEMM: split-K over a thread-block cluster --------------------------------------------------------------------
// M <= 32 rows (a decode batch) against a matrix too small to give every SM a 256-row tile: an M = 128 MMA occupies the tensor
// core for >= 128 cycles whatever its N (the A operand is read at 32 bytes per cycle), so the weight stream of an SM is
// N x 32 bytes per 128 cycles -- wide weight tiles are mandatory, and parallelism has to come from K.  A cluster of `ks` CTAs
// shares one 128-row weight tile: CTA r multiplies k blocks [r, r+1) * K/ks, parks its fp32 partial tile in shared memory, and
// after one cluster barrier every CTA sums a 128/ks-column slice over the ranks IN RANK ORDER through distributed shared memory
// (ld.shared::cluster), rounds once and applies the fused tail.  No global scratch, no atomics, deterministic.
constexpr int kSkBN = 128, kSkRows = 32, kSkStages = 4, kSkRedPitch = kSkBN + 4;
constexpr int kSkABytes = kSkRows * kTcBK * 2, kSkBBytes = kSkBN * kTcBK * 2;
constexpr int kSkSmem = kTcHdr + kSkStages * kSkABytes + kSkStages * kSkBBytes + kSkRows * kSkRedPitch * 4 + 1024;
static_assert(kSkStages * kSkABytes == kTcBM * kTcBK * 2, "the MMA reads a 128-row window from every A stage: it must stay inside the ring");

__device__ __forceinline__ uint32_t cluster_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_size()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_barrier()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ld_cluster_f4(uint32_t local_addr, uint32_t rank)
{
    uint32_t ra;
    float4 v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra) : "memory");
    return v;
}

template <int EPI>
__global__ void __launch_bounds__(kTcThreads, 2) gemm_tc_splitk_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const gemm_tc_params p)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t base = smem_u32(smem);
    const uint32_t bar_full = base, bar_empty = base + 64, bar_tfull = base + 128, tmem_slot = base + 136;
    volatile int* dead = reinterpret_cast<volatile int*>(smem + 144);
    const uint32_t a0 = base + kTcHdr, b0 = a0 + kSkStages * kSkABytes;
    float* red = reinterpret_cast<float*>(smem + kTcHdr + kSkStages * (kSkABytes + kSkBBytes)); // [32][kSkRedPitch] fp32 partial tile
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank(), ks = cluster_size();
    const uint32_t n0 = (blockIdx.x / ks) * kSkBN;
    const uint32_t kb_per = (p.K / kTcBK) / ks, kb0 = rank * kb_per;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kSkStages; s++) mbar_init(bar_full + s * 8, 1), mbar_init(bar_empty + s * 8, 1);
        mbar_init(bar_tfull, 1);
        *dead = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(uint32_t(kSkBN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 136);

    pdl_trigger();
    if (warp == 0) {
        if (lane == 0) {
            // weight tiles first (independent of the previous kernel), then griddepcontrol.wait, then the X tiles
            const uint32_t pre = kb_per < uint32_t(kSkStages) ? kb_per : uint32_t(kSkStages);
            for (uint32_t kb = 0; kb < pre; kb++) {
                mbar_expect_tx(bar_full + kb * 8, kSkABytes + kSkBBytes);
                tc_tma_load(b0 + kb * kSkBBytes, &tmW, (kb0 + kb) * kTcBK, n0, bar_full + kb * 8);
            }
            pdl_sync();
            for (uint32_t kb = 0; kb < pre; kb++) tc_tma_load(a0 + kb * kSkABytes, &tmX, (kb0 + kb) * kTcBK, 0, bar_full + kb * 8);
            for (uint32_t kb = pre; kb < kb_per; kb++) {
                const uint32_t s = kb % kSkStages, ph = (kb / kSkStages) & 1u;
                tc_wait(bar_empty + s * 8, ph ^ 1u, dead, p.err);
                mbar_expect_tx(bar_full + s * 8, kSkABytes + kSkBBytes);
                tc_tma_load(a0 + s * kSkABytes, &tmX, (kb0 + kb) * kTcBK, 0, bar_full + s * 8);
                tc_tma_load(b0 + s * kSkBBytes, &tmW, (kb0 + kb) * kTcBK, n0, bar_full + s * 8);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc_idesc(kTcBM, kSkBN);
            for (uint32_t kb = 0; kb < kb_per; kb++) {
                const uint32_t s = kb % kSkStages, ph = (kb / kSkStages) & 1u;
                tc_wait(bar_full + s * 8, ph, dead, p.err);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // rows 32..127 of the A window are whatever the ring holds: they only reach accumulator lanes nobody reads
                const uint64_t da = tc_smem_desc(a0 + s * kSkABytes), db = tc_smem_desc(b0 + s * kSkBBytes);
#pragma unroll
                for (uint32_t k = 0; k < kTcBK / 16; k++) tc_mma(tmem, da + k * 2, db + k * 2, idesc, (kb | k) != 0);
                tc_commit(bar_empty + s * 8);
            }
            tc_commit(bar_tfull);
        }
    } else if (warp == 4) {
        // TMEM lanes 0..31 (the batch rows) belong to the warp with warp % 4 == 0: park the fp32 partial tile in shared memory
        tc_wait(bar_tfull, 0, dead, p.err);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (uint32_t cb = 0; cb < uint32_t(kSkBN); cb += 32) {
            uint32_t v[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                  "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                  "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                  "=r"(v[31])
                : "r"(tmem + cb)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float4* dst = reinterpret_cast<float4*>(red + lane * kSkRedPitch + cb);
#pragma unroll
            for (int q = 0; q < 8; q++)
                dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncwarp();
    __syncthreads();
    cluster_barrier(); // every rank's partial tile is in its shared memory
    pdl_sync();        // residual rows / output buffer: earlier kernels of the stream have finished

    // this CTA joins columns [rank, rank + 1) * 128 / ks of the tile: 8 columns per thread and step, ranks in order
    const uint32_t w8 = (kSkBN / ks) >> 3, col0 = rank * (kSkBN / ks);
    const uint32_t red_addr = smem_u32(red);
    for (uint32_t u = threadIdx.x; u < p.M * w8; u += kTcThreads) {
        const uint32_t row = u / w8, c = col0 + (u % w8) * 8;
        if (n0 + c >= p.N) continue;
        float y[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        for (uint32_t r = 0; r < ks; r++) {
            const uint32_t a = red_addr + (row * kSkRedPitch + c) * 4;
            const float4 lo = ld_cluster_f4(a, r), hi = ld_cluster_f4(a + 16, r);
            y[0] += lo.x, y[1] += lo.y, y[2] += lo.z, y[3] += lo.w, y[4] += hi.x, y[5] += hi.y, y[6] += hi.z, y[7] += hi.w;
        }
#pragma unroll
        for (int j = 0; j < 8; j++) y[j] = rbf(y[j]); // the bmm output buffer is T (kernel/bmm.metal:76)
        if (EPI == EPI_SWIGLU) {
            const uint2 o = make_uint2(pack2(__fmul_rn(silu_bf16(y[0]), y[1]), __fmul_rn(silu_bf16(y[2]), y[3])),
                                       pack2(__fmul_rn(silu_bf16(y[4]), y[5]), __fmul_rn(silu_bf16(y[6]), y[7])));
            *reinterpret_cast<uint2*>(p.Y + size_t(row) * p.ldy + ((n0 + c) >> 1)) = o;
        } else {
            uint32_t o[4];
            if (EPI == EPI_RESIDUAL) {
                const uint4 rv = *reinterpret_cast<const uint4*>(p.res + size_t(row) * p.ldy + n0 + c);
                const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int j = 0; j < 4; j++) o[j] = pack2(__fadd_rn(bf_lo(rw[j]), y[2 * j]), __fadd_rn(bf_hi(rw[j]), y[2 * j + 1]));
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) o[j] = pack2(y[2 * j], y[2 * j + 1]);
            }
            *reinterpret_cast<uint4*>(p.Y + size_t(row) * p.ldy + n0 + c) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
    cluster_barrier(); // nobody leaves while a peer may still read its partial tile
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(uint32_t(kSkBN)) : "memory");
    }
}

// ---- the row-wise kernels around the prefill GEMMs --------------------------------------------------------------
// gather the embedding rows of a prompt chunk (kernel/embedding.metal:25-70), bf16 table
__global__ void __launch_bounds__(256) embed_rows_kernel(uint16_t* out, const uint16_t* table, const int32_t* ids, uint32_t D)
{
    pdl_trigger();
    pdl_sync();
    const uint4* src = reinterpret_cast<const uint4*>(table + size_t(ids[blockIdx.x]) * D);
    uint4* dst = reinterpret_cast<uint4*>(out + size_t(blockIdx.x) * D);
    for (uint32_t k = threadIdx.x; k < D / 8; k += 256) dst[k] = src[k];
}
// n = r((0 + w) * x * rsqrt(mean(x^2) + eps))  (kernel/rmsnorm.metal:53-89), one CTA per row
__global__ void __launch_bounds__(256) rmsnorm_rows_kernel(uint16_t* out, const uint16_t* x, const uint16_t* w, uint32_t D, float eps)
{
    __shared__ float scr[8];
    pdl_trigger();
    pdl_sync();
    const uint16_t* xr = x + size_t(blockIdx.x) * D;
    float part = 0.0f;
    for (uint32_t k = threadIdx.x * 8; k < D; k += 256 * 8) {
        const uint4 v = *reinterpret_cast<const uint4*>(xr + k);
        float f;
        f = bf_lo(v.x), part = fmaf(f, f, part), f = bf_hi(v.x), part = fmaf(f, f, part);
        f = bf_lo(v.y), part = fmaf(f, f, part), f = bf_hi(v.y), part = fmaf(f, f, part);
        f = bf_lo(v.z), part = fmaf(f, f, part), f = bf_hi(v.z), part = fmaf(f, f, part);
        f = bf_lo(v.w), part = fmaf(f, f, part), f = bf_hi(v.w), part = fmaf(f, f, part);
    }
    const float total = block_sum_256(part, scr);
    const float inv = 1.0f / sqrtf(__fadd_rn(total / float(D), eps));
    for (uint32_t k = threadIdx.x * 8; k < D; k += 256 * 8) {
        const uint4 v = *reinterpret_cast<const uint4*>(xr + k);
        const uint4 g = *reinterpret_cast<const uint4*>(w + k);
        uint4 o;
#define MC_NORM2(d, vv, gg)                                                                          \
    d = uint32_t(f32_to_bf16_bits(__fmul_rn(__fmul_rn(bf_lo(gg), bf_lo(vv)), inv))) |                \
        (uint32_t(f32_to_bf16_bits(__fmul_rn(__fmul_rn(bf_hi(gg), bf_hi(vv)), inv))) << 16)
        MC_NORM2(o.x, v.x, g.x);
        MC_NORM2(o.y, v.y, g.y);
        MC_NORM2(o.z, v.z, g.z);
        MC_NORM2(o.w, v.w, g.w);
#undef MC_NORM2
        *reinterpret_cast<uint4*>(out + size_t(blockIdx.x) * D + k) = o;
    }
}
// rotate q and k (kernel/rope.metal:47-58), store q, append k', v to the cache (nn/cache.h:207-214); one CTA per prompt row
__global__ void __launch_bounds__(256) rope_append_kernel(const uint16_t* qkv, uint32_t ld, uint16_t* q, uint16_t* kcache, uint16_t* vcache, const float* fcos,
                                                          const float* fsin, const int32_t* row_seq, const int32_t* row_pos, uint32_t seq, uint32_t start_pos, uint32_t H,
                                                          uint32_t KV, uint32_t hd, uint32_t max_seq)
{
    pdl_trigger();
    pdl_sync();
    // prompt: rows are consecutive positions of one sequence; batched decode: every row carries its own (sequence, position)
    const uint32_t row = blockIdx.x, half = hd >> 1, pos = row_pos ? uint32_t(row_pos[row]) : start_pos + row;
    if (row_seq) seq = uint32_t(row_seq[row]);
    const uint16_t* src_row = qkv + size_t(row) * ld;
    // rotated heads (q heads then k heads): thread = one pair of adjacent j -> (j, j+1) and (j + half, j + half + 1)
    const uint32_t pairs = (H + KV) * (half >> 1);
    for (uint32_t i = threadIdx.x; i < pairs; i += 256) {
        const uint32_t head = i / (half >> 1), j = (i % (half >> 1)) * 2;
        const uint16_t* src = src_row + size_t(head) * hd;
        const uint32_t lo = *reinterpret_cast<const uint32_t*>(src + j), hi = *reinterpret_cast<const uint32_t*>(src + j + half);
        const float2 cs = *reinterpret_cast<const float2*>(fcos + size_t(pos) * half + j), sn = *reinterpret_cast<const float2*>(fsin + size_t(pos) * half + j);
        const float a0 = bf_lo(lo), a1 = bf_hi(lo), b0 = bf_lo(hi), b1 = bf_hi(hi);
        const uint32_t o_lo = pack2(__fsub_rn(__fmul_rn(cs.x, a0), __fmul_rn(sn.x, b0)), __fsub_rn(__fmul_rn(cs.y, a1), __fmul_rn(sn.y, b1)));
        const uint32_t o_hi = pack2(__fadd_rn(__fmul_rn(sn.x, a0), __fmul_rn(cs.x, b0)), __fadd_rn(__fmul_rn(sn.y, a1), __fmul_rn(cs.y, b1)));
        uint16_t* dst = head < H ? q + size_t(row) * H * hd + size_t(head) * hd : kcache + ((size_t(seq) * KV + (head - H)) * max_seq + size_t(min(pos, max_seq - 1))) * hd;
        *reinterpret_cast<uint32_t*>(dst + j) = o_lo;
        *reinterpret_cast<uint32_t*>(dst + j + half) = o_hi;
    }
    // values: bit copy, 16 bytes per thread
    const uint32_t chunks = KV * (hd >> 3);
    for (uint32_t i = threadIdx.x; i < chunks; i += 256) {
        const uint32_t kvh = i / (hd >> 3), c = (i % (hd >> 3)) * 8;
        *reinterpret_cast<uint4*>(vcache + ((size_t(seq) * KV + kvh) * max_seq + size_t(min(pos, max_seq - 1))) * hd + c) =
            *reinterpret_cast<const uint4*>(src_row + size_t(H + KV + kvh) * hd + c);
    }
}

// ---- causal prefill attention -------------------------------------------------------------------------------------------
// One CTA = 8 warps x 16 prompt rows: `rep` query heads of one KV head (grouped-query attention: they read the same K/V tiles)
// times 128 / rep consecutive rows; keys in tiles of 64 from the KV cache, double-buffered with cp.async.  The heaviest
// row tiles (most visible keys) are scheduled first.
// The reference's chain (nn/attention.h:195-200, kernel/softmax.metal:40-80) has no running maximum and rounds to bf16 after
// each of bmm / scale / softmax / bmm, so the kernel makes two passes over the keys:
//   pass 1   s = r(r(q.K) * scale) for every visible key, total = sum exp(s)               (fp32, fixed order per thread;
//            exp = ex2.approx(s * log2 e), 2 ulp of fp32 -- far inside the bf16 rounding that follows)
//   pass 2   p = r(exp(s) * (1 / total)), o += p . V with p as the bf16 A operand of the second mma; o = r(o) at the end
// Masked keys (t > pos) contribute exactly 0, as exp(r(s + -inf)) does in the reference.
struct pattn_params {
    const uint16_t* q;  // [rows, H*hd] rotated
    const uint16_t* kc; // this layer, this sequence: [KV][max_seq][hd]
    const uint16_t* vc;
    uint16_t* out;      // [rows, H*hd]
    uint32_t rows, start_pos, H, KV, max_seq;
    uint32_t key_begin; // first visible key (0; start_pos under the reference's chunk mask, quirk Q9: nn/attention.h:283-299)
    uint32_t rep;       // query heads per CTA: H / KV when that is 1, 2, 4 or 8, else 1
    float scale;        // r(1/sqrt(hd)) stored as T (quirk Q4)
};
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// two fp32 -> packed bf16 pair with one cvt.rn.bf16x2.f32 (round to nearest even; same bits as pack2 for every non-NaN input)
__device__ __forceinline__ uint32_t pack2_rn(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float exp2f_approx(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void pa_cp16(uint32_t dst, const void* src, bool valid)
{
    const uint32_t n = valid ? 16u : 0u; // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
// RNE to bf16 on the integer pipe (the conversion unit is the busiest pipe of this kernel); same bits as rbf() for every non-NaN input
__device__ __forceinline__ float rbf_alu(float x)
{
    const uint32_t u = __float_as_uint(x);
    return __uint_as_float((u + 0x7fffu + ((u >> 16) & 1u)) & 0xffff0000u);
}
// POW2: the stored scale r(1/sqrt(hd)) is a power of two (head_dim 64: 0.125), so r(r(q.K) * scale) == r(q.K) * scale exactly
// and the second rounding is skipped.
template <int HD, bool POW2> __global__ void __launch_bounds__(256, HD == 64 ? 2 : 1) prefill_attn_kernel(const pattn_params p)
{
    constexpr int PITCH = HD + 8;             // bf16 elements per smem row: 16-byte aligned, conflict-free fragment loads
    constexpr int TILE = 64 * PITCH * 2;      // bytes of one 64-key K (or V) tile
    extern __shared__ __align__(16) unsigned char psm[]; // [2 buffers][K tile | V tile]
    const uint32_t sbase = smem_u32(psm);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const uint32_t head = blockIdx.x * p.rep + warp % p.rep, kvh = head / (p.H / p.KV);
    const uint32_t qt = 128u / p.rep;                       // prompt rows of this CTA
    const uint32_t q0 = (gridDim.y - 1 - blockIdx.y) * qt;  // row tile = the slow grid index, heaviest (most visible keys) first
    const uint32_t wr0 = q0 + (warp / p.rep) * 16;          // first row of this warp
    const uint32_t r_lo = wr0 + g, r_hi = r_lo + 8;         // the two prompt rows of this thread
    const uint32_t rows_end = min(q0 + qt, p.rows);
    const uint32_t n_keys = p.start_pos + rows_end; // keys visible to the last row of the CTA
    const uint32_t kt_first = p.key_begin / 64;     // tiles below the first visible key are never touched
    const uint32_t n_tiles = (n_keys + 63) / 64 - kt_first;
    const uint16_t* kbase = p.kc + size_t(kvh) * p.max_seq * HD;
    const uint16_t* vbase = p.vc + size_t(kvh) * p.max_seq * HD;

    pdl_trigger();
    pdl_sync();
    // Q fragments (A operand, row-major 16 x HD)
    uint32_t qa[HD / 16][4];
    {
        const uint16_t* ql = p.q + size_t(min(r_lo, p.rows - 1)) * p.H * HD + size_t(head) * HD;
        const uint16_t* qh = p.q + size_t(min(r_hi, p.rows - 1)) * p.H * HD + size_t(head) * HD;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ks++) {
            qa[ks][0] = *reinterpret_cast<const uint32_t*>(ql + ks * 16 + 2 * t);
            qa[ks][1] = *reinterpret_cast<const uint32_t*>(qh + ks * 16 + 2 * t);
            qa[ks][2] = *reinterpret_cast<const uint32_t*>(ql + ks * 16 + 8 + 2 * t);
            qa[ks][3] = *reinterpret_cast<const uint32_t*>(qh + ks * 16 + 8 + 2 * t);
        }
    }
    const uint32_t pos_lo = p.start_pos + r_lo, pos_hi = p.start_pos + r_hi; // last visible key of each row

    // job j < n_tiles: pass 1 on key tile j (K only); job j >= n_tiles: pass 2 on key tile j - n_tiles (K and V)
    auto issue = [&](uint32_t job) {
        const bool second = job >= n_tiles;
        const uint32_t kt = kt_first + (second ? job - n_tiles : job);
        const uint32_t buf = sbase + (job & 1u) * 2 * TILE;
        constexpr int CH = HD / 8; // 16-byte chunks per row
        for (uint32_t c = tid; c < 64 * CH; c += 256) {
            const uint32_t r = c / CH, cc = c % CH, key = kt * 64 + r;
            const bool ok = key < n_keys;
            const size_t off = size_t(ok ? key : 0) * HD + cc * 8;
            pa_cp16(buf + (r * PITCH + cc * 8) * 2, kbase + off, ok);
            if (second) pa_cp16(buf + TILE + (r * PITCH + cc * 8) * 2, vbase + off, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    float sum_lo = 0.0f, sum_hi = 0.0f, inv_lo = 0.0f, inv_hi = 0.0f;
    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; i++) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.0f;

    const uint32_t n_jobs = 2 * n_tiles;
    issue(0);
    for (uint32_t job = 0; job < n_jobs; job++) {
        if (job + 1 < n_jobs) {
            issue(job + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const bool second = job >= n_tiles;
        const uint32_t kt = kt_first + (second ? job - n_tiles : job);
        const uint32_t kbuf = sbase + (job & 1u) * 2 * TILE, vbuf = kbuf + TILE;
        if (job == n_tiles) {
            // between the passes: the four lanes of a quad hold the partial sums of one row
            sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1), sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
            sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1), sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);
            inv_lo = 1.0f / sum_lo, inv_hi = 1.0f / sum_hi;
        }
        // this warp's rows see keys <= pos_hi: skip tiles entirely above the diagonal
        if (kt * 64 <= p.start_pos + wr0 + 15) {
            // S = Q . K^T for 64 keys: 8 n-tiles of 8 keys
            float s[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; nt++) {
                s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f;
#pragma unroll
                for (int k2 = 0; k2 < HD / 32; k2++) {
                    // B[k = d][n = key] = K[key][d]: ldmatrix.x4 = the 8 keys of the n-tile x d chunks (0-7 | 8-15 | 16-23 | 24-31) of this k pair
                    const uint32_t addr = kbuf + ((nt * 8 + (lane & 7)) * PITCH + k2 * 32 + (lane >> 3) * 8) * 2;
                    uint32_t b0, b1, b2, b3;
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(addr));
                    mma_bf16_16816(s[nt], qa[2 * k2], b0, b1);
                    mma_bf16_16816(s[nt], qa[2 * k2 + 1], b2, b3);
                }
            }
            // s = r(r(q.K) * scale), e = exp(s) = ex2(s * log2 e); masked keys drop out  (kernel/bmm.metal:76, scalar_mul,
            // kernel/softmax.metal:46)
            const bool diagonal = kt * 64 + 63 > p.start_pos + wr0 || kt * 64 < p.key_begin; // some key of the tile is hidden from some row of the warp
            const float sl2e = p.scale * 1.4426950408889634f;
#pragma unroll
            for (int nt = 0; nt < 8; nt++) {
                float e0, e1, e2, e3;
                if (POW2) {
                    e0 = exp2f_approx(rbf_alu(s[nt][0]) * sl2e), e1 = exp2f_approx(rbf_alu(s[nt][1]) * sl2e);
                    e2 = exp2f_approx(rbf_alu(s[nt][2]) * sl2e), e3 = exp2f_approx(rbf_alu(s[nt][3]) * sl2e);
                } else {
                    const uint32_t lo2 = pack2_rn(__fmul_rn(rbf_alu(s[nt][0]), p.scale), __fmul_rn(rbf_alu(s[nt][1]), p.scale));
                    const uint32_t hi2 = pack2_rn(__fmul_rn(rbf_alu(s[nt][2]), p.scale), __fmul_rn(rbf_alu(s[nt][3]), p.scale));
                    e0 = __expf(bf_lo(lo2)), e1 = __expf(bf_hi(lo2)), e2 = __expf(bf_lo(hi2)), e3 = __expf(bf_hi(hi2));
                }
                if (diagonal) {
                    const uint32_t key = kt * 64 + nt * 8 + 2 * t;
                    const bool v0 = key >= p.key_begin, v1 = key + 1 >= p.key_begin;
                    e0 = v0 && key <= pos_lo ? e0 : 0.0f, e1 = v1 && key + 1 <= pos_lo ? e1 : 0.0f;
                    e2 = v0 && key <= pos_hi ? e2 : 0.0f, e3 = v1 && key + 1 <= pos_hi ? e3 : 0.0f;
                }
                if (!second) {
                    sum_lo += e0 + e1, sum_hi += e2 + e3;
                } else {
                    // p = r(exp(s) * (1/total))  (kernel/softmax.metal:79), kept packed: the A operand of the second product
                    s[nt][0] = __uint_as_float(pack2_rn(__fmul_rn(e0, inv_lo), __fmul_rn(e1, inv_lo)));
                    s[nt][2] = __uint_as_float(pack2_rn(__fmul_rn(e2, inv_hi), __fmul_rn(e3, inv_hi)));
                }
            }
            if (second) {
                // O += P . V: k = 16 keys per step, P from the S registers (C layout of two n-tiles = A layout of one k step)
#pragma unroll
                for (int kk = 0; kk < 4; kk++) {
                    uint32_t pa[4];
                    pa[0] = __float_as_uint(s[2 * kk][0]);
                    pa[1] = __float_as_uint(s[2 * kk][2]);
                    pa[2] = __float_as_uint(s[2 * kk + 1][0]);
                    pa[3] = __float_as_uint(s[2 * kk + 1][2]);
#pragma unroll
                    for (int dp = 0; dp < HD / 16; dp++) {
                        // ldmatrix.x4.trans: matrices (keys 0-7 | 8-15) x (d 0-7 | 8-15) of the 16 x 16 block of V
                        const uint32_t mi = lane >> 3, mr = lane & 7;
                        const uint32_t addr = vbuf + ((kk * 16 + (mi & 1) * 8 + mr) * PITCH + dp * 16 + (mi >> 1) * 8) * 2;
                        uint32_t v0, v1, v2, v3;
                        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(addr));
                        mma_bf16_16816(o[2 * dp], pa, v0, v1);
                        mma_bf16_16816(o[2 * dp + 1], pa, v2, v3);
                    }
                }
            }
        }
        __syncthreads(); // the buffer is refilled by the job after next
    }
    // o = r(sum p.V)  (kernel/bmm.metal:76); the reference's transposed copy back to [rows, H*hd] (nn/attention.h:90-102)
#pragma unroll
    for (int nt = 0; nt < HD / 8; nt++) {
        if (r_lo < p.rows) *reinterpret_cast<uint32_t*>(p.out + size_t(r_lo) * p.H * HD + size_t(head) * HD + nt * 8 + 2 * t) = pack2(rbf(o[nt][0]), rbf(o[nt][1]));
        if (r_hi < p.rows) *reinterpret_cast<uint32_t*>(p.out + size_t(r_hi) * p.H * HD + size_t(head) * HD + nt * 8 + 2 * t) = pack2(rbf(o[nt][2]), rbf(o[nt][3]));
    }
}


// ---- batched decode attention (grouped-query) -------------------------------------------------------------------------------
// One CTA = one (sequence, KV head): the H / KV query heads that share the KV head are the rows of one m16 mma tile, so the
// cached keys / values of the head are read ONCE for all of them (the per-(row, head) decode kernel reads them H / KV times).
// Keys go in tiles of 64; warp w owns keys [16w, 16w + 16) of every tile (two score n-tiles, one k step of the value product).
// Same two passes and rounding points as the prompt kernel above; the four per-warp partial sums / outputs are joined in
// warp order through shared memory.
struct dattn_params {
    const uint16_t* q;  // [rows, H*hd] rotated (when qkv == nullptr)
    const uint16_t* qkv; // or: [rows, (H + 2 KV) * hd] un-rotated q|k|v rows of this step -- the kernel rotates q and k (kernel/rope.metal:47-58)
                         // and appends k', v to the cache itself (nn/cache.h:207-214), one launch less per block
    const float* fcos;
    const float* fsin;
    uint16_t* kc;       // this layer: [n_seqs][KV][max_seq][hd]
    uint16_t* vc;
    uint16_t* out;      // [rows, H*hd]
    const int32_t* row_seq;
    const int32_t* row_pos;
    uint32_t H, KV, max_seq;
    float scale;
};
template <int HD, bool POW2> __global__ void __launch_bounds__(128) decode_attn_gqa_kernel(const dattn_params p)
{
    constexpr int PITCH = HD + 8;
    constexpr int TILE = 64 * PITCH * 2;
    extern __shared__ __align__(16) unsigned char psm[]; // [2 buffers][K tile | V tile] | float red[4][8]
    float* red = reinterpret_cast<float*>(psm + 4 * TILE);
    const uint32_t sbase = smem_u32(psm);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const uint32_t kvh = blockIdx.x, row = blockIdx.y, n_rep = p.H / p.KV;
    pdl_trigger();
    pdl_sync();
    // rpos: absolute position (RoPE); pos: cache row of the new key = min(rpos, S - 1) -- beyond the cache the sink roll has already
    // shifted the older rows (nn/cache.h:183-204)
    const uint32_t seq = uint32_t(p.row_seq[row]), rpos = uint32_t(p.row_pos[row]), pos = min(rpos, p.max_seq - 1);
    const uint32_t n_keys = pos + 1, n_tiles = (n_keys + 63) / 64;
    uint16_t* kbase = p.kc + (size_t(seq) * p.KV + kvh) * p.max_seq * HD;
    uint16_t* vbase = p.vc + (size_t(seq) * p.KV + kvh) * p.max_seq * HD;
    constexpr uint32_t half = HD / 2;

    // Q fragments: mma row g = query head kvh * n_rep + g (rows >= n_rep and rows 8..15 are zero padding)
    uint32_t qa[HD / 16][4];
    if (p.qkv) {
        const uint16_t* src = p.qkv + size_t(row) * (p.H + 2 * p.KV) * HD;
        // this KV head's new key, rotated, and value go to the cache first (this CTA is the only reader of that cache row in this launch)
        {
            const uint16_t* ks = src + size_t(p.H + kvh) * HD;
            for (uint32_t j = tid; j < half; j += 128) {
                const float a = bf16_bits_to_f32(ks[j]), b = bf16_bits_to_f32(ks[j + half]);
                const float cs = p.fcos[size_t(rpos) * half + j], sn = p.fsin[size_t(rpos) * half + j];
                kbase[size_t(pos) * HD + j] = f32_to_bf16_bits(__fsub_rn(__fmul_rn(cs, a), __fmul_rn(sn, b)));
                kbase[size_t(pos) * HD + j + half] = f32_to_bf16_bits(__fadd_rn(__fmul_rn(sn, a), __fmul_rn(cs, b)));
            }
            const uint16_t* vs = src + size_t(p.H + p.KV + kvh) * HD;
            for (uint32_t c = tid; c < HD / 8; c += 128) *reinterpret_cast<uint4*>(vbase + size_t(pos) * HD + c * 8) = *reinterpret_cast<const uint4*>(vs + c * 8);
        }
        // q rows rotated in registers: the partner of element d < hd/2 is d + hd/2, the same lane's fragment HD/32 k-steps further
        const bool valid = g < n_rep;
        const uint16_t* ql = src + size_t(kvh * n_rep + (valid ? g : 0)) * HD;
#pragma unroll
        for (int ks = 0; ks < HD / 32; ks++) {
#pragma unroll
            for (int h8 = 0; h8 < 2; h8++) {
                const uint32_t d = ks * 16 + h8 * 8 + 2 * t;
                const uint32_t lo = *reinterpret_cast<const uint32_t*>(ql + d), hi = *reinterpret_cast<const uint32_t*>(ql + d + half);
                const float2 cs = *reinterpret_cast<const float2*>(p.fcos + size_t(rpos) * half + d), sn = *reinterpret_cast<const float2*>(p.fsin + size_t(rpos) * half + d);
                const float a0 = bf_lo(lo), a1 = bf_hi(lo), b0 = bf_lo(hi), b1 = bf_hi(hi);
                const uint32_t r_lo = pack2(__fsub_rn(__fmul_rn(cs.x, a0), __fmul_rn(sn.x, b0)), __fsub_rn(__fmul_rn(cs.y, a1), __fmul_rn(sn.y, b1)));
                const uint32_t r_hi = pack2(__fadd_rn(__fmul_rn(sn.x, a0), __fmul_rn(cs.x, b0)), __fadd_rn(__fmul_rn(sn.y, a1), __fmul_rn(cs.y, b1)));
                qa[ks][h8 * 2] = valid ? r_lo : 0u;
                qa[ks + HD / 32][h8 * 2] = valid ? r_hi : 0u;
            }
            qa[ks][1] = qa[ks][3] = qa[ks + HD / 32][1] = qa[ks + HD / 32][3] = 0u;
        }
        __threadfence(); // the appended cache row is read back through cp.async (L2) below
        __syncthreads();
    } else {
        const bool valid = g < n_rep;
        const uint16_t* ql = p.q + size_t(row) * p.H * HD + size_t(kvh * n_rep + (valid ? g : 0)) * HD;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ks++) {
            qa[ks][0] = valid ? *reinterpret_cast<const uint32_t*>(ql + ks * 16 + 2 * t) : 0u;
            qa[ks][2] = valid ? *reinterpret_cast<const uint32_t*>(ql + ks * 16 + 8 + 2 * t) : 0u;
            qa[ks][1] = qa[ks][3] = 0u;
        }
    }
    auto issue = [&](uint32_t job) {
        const bool second = job >= n_tiles;
        const uint32_t kt = second ? job - n_tiles : job;
        const uint32_t buf = sbase + (job & 1u) * 2 * TILE;
        constexpr int CH = HD / 8;
        for (uint32_t c = tid; c < 64 * CH; c += 128) {
            const uint32_t r = c / CH, cc = c % CH, key = kt * 64 + r;
            const bool ok = key < n_keys;
            const size_t off = size_t(ok ? key : 0) * HD + cc * 8;
            pa_cp16(buf + (r * PITCH + cc * 8) * 2, kbase + off, ok);
            if (second) pa_cp16(buf + TILE + (r * PITCH + cc * 8) * 2, vbase + off, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    float sum = 0.0f, inv = 0.0f;
    float o[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; i++) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.0f;
    const float sl2e = p.scale * 1.4426950408889634f;
    const uint32_t n_jobs = 2 * n_tiles;
    issue(0);
    for (uint32_t job = 0; job < n_jobs; job++) {
        if (job == n_tiles) {
            // end of pass 1: quad partial -> warp partial (row g), published for the join below
            sum += __shfl_xor_sync(0xffffffffu, sum, 1), sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            if (t == 0) red[warp * 8 + g] = sum;
        }
        if (job + 1 < n_jobs) {
            issue(job + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const bool second = job >= n_tiles;
        const uint32_t kt = second ? job - n_tiles : job;
        const uint32_t kbuf = sbase + (job & 1u) * 2 * TILE, vbuf = kbuf + TILE;
        if (job == n_tiles) inv = 1.0f / (((red[g] + red[8 + g]) + red[16 + g]) + red[24 + g]);
        if (kt * 64 + warp * 16 < n_keys) {
            float s[2][4];
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
                s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f;
#pragma unroll
                for (int k2 = 0; k2 < HD / 32; k2++) {
                    const uint32_t addr = kbuf + ((warp * 16 + nt * 8 + (lane & 7)) * PITCH + k2 * 32 + (lane >> 3) * 8) * 2;
                    uint32_t b0, b1, b2, b3;
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(addr));
                    mma_bf16_16816(s[nt], qa[2 * k2], b0, b1);
                    mma_bf16_16816(s[nt], qa[2 * k2 + 1], b2, b3);
                }
            }
            uint32_t pa[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
                float e0, e1;
                if (POW2) {
                    e0 = exp2f_approx(rbf_alu(s[nt][0]) * sl2e), e1 = exp2f_approx(rbf_alu(s[nt][1]) * sl2e);
                } else {
                    const uint32_t lo2 = pack2_rn(__fmul_rn(rbf_alu(s[nt][0]), p.scale), __fmul_rn(rbf_alu(s[nt][1]), p.scale));
                    e0 = __expf(bf_lo(lo2)), e1 = __expf(bf_hi(lo2));
                }
                const uint32_t key = kt * 64 + warp * 16 + nt * 8 + 2 * t;
                e0 = key <= pos ? e0 : 0.0f, e1 = key + 1 <= pos ? e1 : 0.0f;
                if (!second) sum += e0 + e1;
                else pa[2 * nt] = pack2_rn(__fmul_rn(e0, inv), __fmul_rn(e1, inv)); // p = r(exp(s) * (1/total))
            }
            if (second) {
#pragma unroll
                for (int dp = 0; dp < HD / 16; dp++) {
                    const uint32_t mi = lane >> 3, mr = lane & 7;
                    const uint32_t addr = vbuf + ((warp * 16 + (mi & 1) * 8 + mr) * PITCH + dp * 16 + (mi >> 1) * 8) * 2;
                    uint32_t v0, v1, v2, v3;
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(addr));
                    mma_bf16_16816(o[2 * dp], pa, v0, v1);
                    mma_bf16_16816(o[2 * dp + 1], pa, v2, v3);
                }
            }
        }
        __syncthreads();
    }
    // join the four per-warp partial outputs in warp order, o = r(sum p.V)  (kernel/bmm.metal:76)
    float* part = reinterpret_cast<float*>(psm); // [4 warps][8 rows][HD]
    if (g < n_rep) {
#pragma unroll
        for (int nt = 0; nt < HD / 8; nt++) *reinterpret_cast<float2*>(part + (warp * 8 + g) * HD + nt * 8 + 2 * t) = make_float2(o[nt][0], o[nt][1]);
    }
    __syncthreads();
    for (uint32_t i = tid; i < n_rep * HD; i += 128) {
        const uint32_t r = i / HD, d = i % HD;
        const float v = ((part[r * HD + d] + part[(8 + r) * HD + d]) + part[(16 + r) * HD + d]) + part[(24 + r) * HD + d];
        p.out[size_t(row) * p.H * HD + size_t(kvh * n_rep + r) * HD + d] = f32_to_bf16_bits(v);
    }
}

// ---- QLoRA models on the tensor-core path --------------------------------------------------------------------------------------
// The reference dequantises the whole matrix to bf16 on every call, w = r(r(q) * r(s)) (quantization/lora.h:115, kernel/mul.metal:76-77),
// and quantization::linear caches its dequantised weight (quantization/linear.h:50-53).  For prompts and decode batches the engine
// keeps that bf16 image resident (a shadow of the packed int4 stream that batch-1 decode reads) and runs the GEMMs above on it; the
// adaptor term y = r(y + r(r(B . r(A . x)) * r(scale))) is applied by the two small kernels below with the reference's rounding points.
__global__ void __launch_bounds__(256) dequant_group_kernel(uint16_t* out, const int8_t* q, const float* scales, uint64_t n8, uint32_t K, uint32_t group)
{
    // 8 weights per thread (one group of 32 spans 4 threads); scales [N, K / group]
    for (uint64_t i = uint64_t(blockIdx.x) * 256 + threadIdx.x; i < n8; i += uint64_t(gridDim.x) * 256) {
        const uint64_t e = i * 8, row = e / K;
        const uint32_t k = uint32_t(e - row * K);
        const float s = rbf(scales[row * (K / group) + k / group]);
        const uint2 v = *reinterpret_cast<const uint2*>(q + e);
        const uint32_t w[2] = {v.x, v.y};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float a = float(int8_t((w[j >> 1] >> ((j & 1) * 16)) & 0xff)), b = float(int8_t((w[j >> 1] >> ((j & 1) * 16 + 8)) & 0xff));
            o[j] = pack2(__fmul_rn(a, s), __fmul_rn(b, s));
        }
        *reinterpret_cast<uint4*>(out + e) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}
// y2 = r(y + r(r(B . ax) * scale))  (quantization/lora.h:115-122), then the fused tail of the linear.  Thread = two adjacent columns
// (their B rows stay in registers) x 16 rows; ax = r(A . x) arrives as extra columns of the GEMM output (the stacked adaptor rows are
// appended to the bf16 image, so the same tcgen05 GEMM produces both products with the rounding of kernel/bmm.metal:76).
//   slices: 1 = one adaptor, 2 = w1|w3 row-interleaved (column parity picks the adaptor), 3 = q|k|v (column ranges pick it)
constexpr int kLoraRows = 16, kLoraMaxRank = 16;
struct lora_epi_params {
    const uint16_t* y;   // [rows, ldy]: columns [0, N) = r(x . Wd^T), columns [N, N + slices * rank) = r(x . A^T)
    uint16_t* out;       // EPI_NONE / EPI_RESIDUAL: [rows, ldo]; EPI_SWIGLU: [rows, ldo] with N/2 columns
    const uint16_t* res; // EPI_RESIDUAL
    const uint16_t* B;   // [N, rank]
    uint32_t rows, N, ldy, ldo, rank, slices, cols0, cols1;
    uint32_t rows_per_cta; // 16 for prompts (B rows amortised), 4 for decode batches (more CTAs)
    float scale;         // r(lora scale)
};
template <int EPI> __global__ void __launch_bounds__(256) lora_epilogue_kernel(const lora_epi_params p)
{
    pdl_trigger();
    pdl_sync();
    const uint32_t n = (blockIdx.x * 256 + threadIdx.x) * 2, row0 = blockIdx.y * p.rows_per_cta;
    if (n >= p.N) return;
    float bw[2][kLoraMaxRank];
    uint32_t slice[2];
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const uint32_t col = n + c;
        slice[c] = p.slices == 2 ? (col & 1u) : (p.slices == 3 ? (col < p.cols0 ? 0u : (col < p.cols1 ? 1u : 2u)) : 0u);
        const uint16_t* b = p.B + size_t(col) * p.rank;
#pragma unroll
        for (int j = 0; j < kLoraMaxRank; j++) bw[c][j] = uint32_t(j) < p.rank ? bf16_bits_to_f32(b[j]) : 0.0f;
    }
    const uint32_t row_end = min(row0 + p.rows_per_cta, p.rows);
    for (uint32_t row = row0; row < row_end; row++) {
        const uint16_t* yr = p.y + size_t(row) * p.ldy;
        const uint32_t yy = *reinterpret_cast<const uint32_t*>(yr + n);
        float v[2] = {bf_lo(yy), bf_hi(yy)};
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const uint16_t* a = yr + p.N + slice[c] * p.rank;
            float l = 0.0f;
#pragma unroll
            for (int j = 0; j < kLoraMaxRank; j += 2) {
                if (uint32_t(j) < p.rank) {
                    const uint32_t av = *reinterpret_cast<const uint32_t*>(a + j);
                    l = fmaf(bf_lo(av), bw[c][j], l), l = fmaf(bf_hi(av), bw[c][j + 1], l);
                }
            }
            v[c] = rbf(__fadd_rn(v[c], rbf(__fmul_rn(rbf(l), p.scale))));
        }
        if (EPI == EPI_SWIGLU) {
            p.out[size_t(row) * p.ldo + (n >> 1)] = f32_to_bf16_bits(__fmul_rn(silu_bf16(v[0]), v[1])); // (nn/transformer.h:57-59)
        } else if (EPI == EPI_RESIDUAL) {
            const uint32_t r = *reinterpret_cast<const uint32_t*>(p.res + size_t(row) * p.ldo + n);
            *reinterpret_cast<uint32_t*>(p.out + size_t(row) * p.ldo + n) = pack2(__fadd_rn(bf_lo(r), v[0]), __fadd_rn(bf_hi(r), v[1]));
        } else {
            *reinterpret_cast<uint32_t*>(p.out + size_t(row) * p.ldo + n) = pack2(v[0], v[1]);
        }
    }
}


// ---- tensor parallel: all-reduce of row-parallel partial sums on the tensor-core path (launcher: tc::tp_allreduce_rows) ------------------
__device__ __forceinline__ bool tp_rows_wait(const uint32_t* flag, uint32_t e, int* err)
{
    uint32_t v = 0;
    for (unsigned long long spins = 0;; spins++) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v >= e) return true;
        if (spins > (1ull << 24)) {
            atomicExch(err, 3);
            return false;
        }
    }
}
__global__ void __launch_bounds__(256) tp_allreduce_rows_kernel(const tp_rows_exchange x, const uint16_t* __restrict__ res, uint32_t rows, uint32_t D)
{
    __shared__ uint32_t s_e;
    pdl_trigger();
    pdl_sync(); // this rank's partial sums (the GEMM launched before) are complete
    const uint32_t W = x.world, me = x.rank;
    if (threadIdx.x == 0) {
        const uint32_t e = *reinterpret_cast<volatile unsigned*>(x.epoch) + 1; // (the epoch moves on only after every CTA of this launch has arrived below)
        if (blockIdx.x == 0) {
            __threadfence_system();
            for (uint32_t k = 0; k < W; k++) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(x.ready[k] + me), "r"(e) : "memory");
        }
        bool ok = true;
        for (uint32_t k = 0; k < W && ok; k++) ok = tp_rows_wait(x.ready[me] + k, e, x.err);
        s_e = ok ? e : 0u;
    }
    __syncthreads();
    const uint32_t e = s_e;
    if (e != 0u) {
        // slice `me` of the element range, 8 elements (two float4 per rank in, one uint4 out) per step
        const uint64_t n8 = uint64_t(rows) * D / 8, per = (n8 + W - 1) / W;
        const uint64_t b = min(n8, uint64_t(me) * per), en = min(n8, b + per);
        for (uint64_t i = b + uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < en; i += uint64_t(gridDim.x) * blockDim.x) {
            float4 lo[kTpRowsMaxWorld], hi[kTpRowsMaxWorld];
#pragma unroll
            for (uint32_t k = 0; k < uint32_t(kTpRowsMaxWorld); k++) {
                if (k < W) {
                    const float* src = x.partial[k] + i * 8;
                    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(lo[k].x), "=f"(lo[k].y), "=f"(lo[k].z), "=f"(lo[k].w) : "l"(src) : "memory");
                    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(hi[k].x), "=f"(hi[k].y), "=f"(hi[k].z), "=f"(hi[k].w) : "l"(src + 4) : "memory");
                }
            }
            float s[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (uint32_t k = 0; k < uint32_t(kTpRowsMaxWorld); k++) {
                if (k < W) s[0] += lo[k].x, s[1] += lo[k].y, s[2] += lo[k].z, s[3] += lo[k].w, s[4] += hi[k].x, s[5] += hi[k].y, s[6] += hi[k].z, s[7] += hi[k].w;
            }
            // h = r(x + r(sum))  (kernel/bmm.metal:76, nn/transformer.h:133,139)
            uint4 o;
            if (res != nullptr) {
                const uint4 r = *reinterpret_cast<const uint4*>(res + i * 8);
                const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
                o.x = pack2(__fadd_rn(bf_lo(rw[0]), rbf(s[0])), __fadd_rn(bf_hi(rw[0]), rbf(s[1])));
                o.y = pack2(__fadd_rn(bf_lo(rw[1]), rbf(s[2])), __fadd_rn(bf_hi(rw[1]), rbf(s[3])));
                o.z = pack2(__fadd_rn(bf_lo(rw[2]), rbf(s[4])), __fadd_rn(bf_hi(rw[2]), rbf(s[5])));
                o.w = pack2(__fadd_rn(bf_lo(rw[3]), rbf(s[6])), __fadd_rn(bf_hi(rw[3]), rbf(s[7])));
            } else {
                // no residual: the rounded sums themselves (QLoRA: r(x . Wd^T) | r(x . A^T), the adaptor epilogue follows)
                o.x = pack2(s[0], s[1]), o.y = pack2(s[2], s[3]), o.z = pack2(s[4], s[5]), o.w = pack2(s[6], s[7]);
            }
#pragma unroll
            for (uint32_t k = 0; k < uint32_t(kTpRowsMaxWorld); k++) {
                if (k < W) asm volatile("st.relaxed.sys.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(x.result[k] + i * 8), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
            }
        }
    }
    __syncthreads(); // the stores of every thread happen before thread 0's system-scope fence (cumulativity)
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned prev = atomicAdd(x.counter, 1u);
        if (prev == gridDim.x - 1) {
            // last CTA of this rank: the slice is in every rank's result buffer
            *x.counter = 0;
            if (e != 0u) {
                *x.epoch = e;
                __threadfence_system();
                for (uint32_t k = 0; k < W; k++) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(x.done[k] + me), "r"(e) : "memory");
            }
        }
        // nobody leaves before every rank's slice has arrived here: the kernels launched next read the whole result
        if (e != 0u)
            for (uint32_t k = 0; k < W; k++)
                if (!tp_rows_wait(x.done[me] + k, e, x.err)) break;
    }
    __syncthreads();
}

} // namespace tc
} // namespace mc
