// metalchat_b200/csrc/mc_stream_kernel.cuh — the streaming persistent decode kernel (sm_100a).
//
// One launch runs `steps` whole decode steps of nn::llama3::operator() (nn/llama.h:113-134) plus the greedy
// sampler tail.  Where the per-op path launches 5 kernels per transformer block and the HBM stream drains
// at every kernel boundary, this kernel keeps ONE weight stream running for the whole token:
//
//   * grid = one CTA per SM (cooperative), 8 consumer warps + 1 producer warp.
//   * The producer warp walks the static schedule of weight tiles of its CTA (every phase, every layer,
//     every step) and moves them HBM -> shared memory with cp.async.bulk (the TMA engine, one bulk copy
//     per weight row chunk) into a ring of 32 KiB stages guarded by full/empty mbarriers.  Weights do not
//     depend on activations, so the producer never waits for a phase barrier: while the consumers sit in
//     a grid barrier the ring (up to ~28 MB chip-wide) keeps filling and HBM stays busy.
//   * The consumer warps wait for a phase's input (grid barrier), stage the activation rows in shared
//     memory (embedding gather / RMSNorm / split-attention join fused in), then eat tiles from the ring:
//     a tile is 16 weight rows x KC k; each warp takes a k-slice and feeds it to mma.sync.m16n8k16
//     (bf16 x bf16 -> fp32) with the batch rows as the n dimension (up to 8 sequences cost one weight
//     pass).  Fragments are read straight from the row-major tile with a k permutation that is applied
//     to both operands (lane (g,t) reads 16 contiguous bytes of row g / row g+8 / activation row g), rows
//     are padded by 64 B so the 16-byte loads are bank-conflict free.
//   * Per block: QKV | attention | wo (+residual) | w1-w3 (+SiLU*mul) | w2 (+residual); then the vocab
//     projection with the greedy argmax fused.  RoPE and the KV-cache append happen at the head of the
//     attention phase (the QKV phase stores r(x.W^T) exactly like kernel/bmm.metal:76), attention is split
//     4 ways over the cached positions with a 4-CTA flag exchange of the exp-sums (softmax has no max
//     shift, kernel/softmax.metal:40-80), the partial outputs are joined in fixed order by the wo phase.
//
// Every bf16 rounding point r(.) of the reference chain is kept; fp32 sums are
// re-associated (mma k-blocks, 8 warp partials joined in warp order), which the stated tolerance covers.
#pragma once
#include "mc_quant_kernels.cuh"

namespace mc {

constexpr int kStWarps = 8;                           // consumer warps
constexpr int kStConsumers = kStWarps * 32;           // 256 consumer threads
constexpr int kStThreads = kStConsumers + 32;         // + one producer warp
constexpr int kStTileRows = 16;                       // weight rows per tile (the m of the mma)
constexpr int kStMaxKC = 1024;                        // k elements per tile row
constexpr int kStPad = 64;                            // bytes of padding per staged row
constexpr int kStStageBytes = kStTileRows * (kStMaxKC * 2 + kStPad);
constexpr int kStMaxStages = 8;
constexpr int kStSplits = 4;                          // CTAs per (row, head) in the attention phase
constexpr int kStMaxRows = 8;                         // activation rows (the n of the mma)
constexpr int kStHdrBytes = 512;                      // mbarriers + flags + reduce scratch
constexpr int kStRedBytes = 2 * kStWarps * 16 * 8 * 4; // double-buffered cross-warp partials

enum { ST_IN_ROWS = 0, ST_IN_EMBED = 1, ST_IN_ATTN = 2 };

struct st_gemv {
    const uint16_t* W;       // [N, K] bf16 row-major; layer l adds l * layer_stride bytes when `layered`
    const uint16_t* norm_w;  // PRO_RMSNORM: [K] (same layer stride)
    const uint16_t* x;       // ST_IN_ROWS: [rows, ldx]
    uint16_t* y;             // [rows, ldy]
    const uint16_t* res;     // EPI_RESIDUAL: [rows, ldy]
    uint32_t N, K, KC, ldx, ldy;
    int32_t pro, epi, in_kind, layered;
};

struct st_params {
    st_gemv g[5];            // qkv, wo, w1-w3, w2, head
    size_t layer_stride;     // bytes between consecutive layers in the weight arena
    size_t kv_layer_stride;  // elements between consecutive layers in the KV cache
    uint32_t n_layers, rows, steps, n_stages;
    uint32_t act_pitch;      // bytes between staged activation rows
    uint32_t act_bytes;
    float eps;
    // attention
    const uint16_t* qkv;     // [rows, (H + 2 KV) * hd]: r(x.W^T), not yet rotated
    uint16_t* kcache;        // [layer][n_seqs, KV, S, hd]
    uint16_t* vcache;
    const float* fcos;       // [2S, hd/2]
    const float* fsin;
    const int32_t* row_seq;  // [rows]
    int32_t* pos;            // [rows], advanced by the sampler tail
    int32_t* ids;            // [rows], next input ids
    uint32_t n_heads, n_kv_heads, head_dim, max_seq, vocab;
    float scale;             // r(1/sqrt(hd)) (nn/attention.h:88,115)
    float* xsum;             // [rows * H][4] exp-sums of the splits
    unsigned* acnt;          // [rows * H] arrival counters of the split groups
    float* opart;            // [rows][4][H * hd] fp32 partial attention outputs
    // embedding
    const uint16_t* embed_table;
    uint16_t* embed_out;     // x rows (the first residual)
    // grid synchronisation
    unsigned* bar;           // phase arrivals (monotonic inside a launch, reset by the last phase)
    unsigned* step_done;     // steps completed inside this launch
    int* err;
    // sampler tail
    float* am_val;           // [rows][G]
    int32_t* am_idx;
    int32_t* out_log;
    int32_t* step_counter;
    int32_t advance;
    unsigned long long* timing; // diagnostics (nullable): 4 globaltimer stamps of CTA 0 per phase
};

// ---- PTX helpers -------------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(mbar), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint4 lds128(uint32_t a)
{
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ uint4 ldcg128(const void* p)
{
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldcg_f4(const float* p)
{
    float4 r;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldcg_f32(const float* p)
{
    float r;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ int32_t ldcg_s32(const int32_t* p)
{
    int32_t r;
    asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint16_t ldcg_u16(const uint16_t* p)
{
    uint16_t r;
    asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// shared-memory carve-up of one CTA
struct st_ctx {
    uint32_t full0, empty0;   // shared addresses of full[kStMaxStages], empty[kStMaxStages]
    volatile int* dead;       // set when a wait timed out somewhere: every later wait falls through
    float* scr;               // [16] block-reduce scratch
    float* red;               // [2][8 warps][16 rows][8 cols]
    unsigned char* act;       // staged activation rows; attention scratch
    uint32_t act_addr;        // shared address of act
    uint32_t ring_addr;       // shared address of stage 0
    int* err;
};

// bounded waits: a lost arrival must end in an error code, never in a hung GPU
__device__ __forceinline__ void st_mbar_wait(const st_ctx& c, uint32_t a, uint32_t parity)
{
    if (mbar_try_wait(a, parity)) return;
    if (*c.dead) return;
    unsigned spins = 0;
    while (!mbar_try_wait(a, parity)) {
        if ((++spins & 255u) == 0) {
            if (*c.dead) return;
            if (*reinterpret_cast<volatile int*>(c.err) != 0) {
                *c.dead = 1;
                return;
            }
            if (spins > (1u << 22)) {
                atomicExch(c.err, 4);
                *c.dead = 1;
                return;
            }
        }
    }
}
__device__ __forceinline__ void st_spin_ge(const st_ctx& c, const unsigned* addr, unsigned target)
{
    if (*c.dead) return;
    unsigned spins = 0;
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
        if (int(v - target) >= 0) return;
        if ((++spins & 1023u) == 0) {
            if (*reinterpret_cast<volatile int*>(c.err) != 0) {
                *c.dead = 1;
                return;
            }
            if (spins > (1u << 22)) {
                atomicExch(c.err, 1);
                *c.dead = 1;
                return;
            }
        }
    }
}
// consumers: thread 0 waits for `target` arrivals, then everybody passes the CTA barrier
__device__ __forceinline__ void st_grid_wait(const st_ctx& c, const unsigned* counter, unsigned target)
{
    if (threadIdx.x == 0) st_spin_ge(c, counter, target);
    consumer_bar();
}
__device__ __forceinline__ void st_grid_arrive(unsigned* bar)
{
    consumer_bar();
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
}
__device__ __forceinline__ void st_stamp(unsigned long long* t, unsigned idx)
{
    if (t && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long v;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
        t[idx] = v;
    }
}
// sum over the 256 consumer threads (fixed partition: lane butterfly, then the 8 warp totals in order)
__device__ __forceinline__ float st_block_sum(float v, float* scr)
{
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) scr[threadIdx.x >> 5] = v;
    consumer_bar();
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < kStWarps; i++) t += scr[i];
    consumer_bar();
    return t;
}

// rows [rb, re) of an N-row matrix owned by this CTA (units of two rows, balanced to +-1 unit)
__device__ __forceinline__ void st_my_rows(uint32_t N, uint32_t& rb, uint32_t& re)
{
    const uint64_t units = N >> 1;
    rb = uint32_t(units * blockIdx.x / gridDim.x) * 2;
    re = uint32_t(units * (blockIdx.x + 1) / gridDim.x) * 2;
}

struct st_pipe {
    uint32_t stage, parity;
    __device__ __forceinline__ void advance(uint32_t n_stages)
    {
        if (++stage == n_stages) stage = 0, parity ^= 1u;
    }
};

// ---- producer: the weight tiles of one GEMV phase ---------------------------------------------------------------------
__device__ __forceinline__ void st_produce_gemv(const st_params& P, const st_gemv& g, uint32_t li, const st_ctx& c, st_pipe& pp, uint64_t policy)
{
    const uint32_t lane = threadIdx.x & 31;
    const char* W = reinterpret_cast<const char*>(g.W) + (g.layered ? size_t(li) * P.layer_stride : 0);
    const uint32_t pitch = g.KC * 2 + kStPad, row_bytes = g.KC * 2;
    uint32_t rb, re;
    st_my_rows(g.N, rb, re);
    for (uint32_t r0 = rb; r0 < re; r0 += kStTileRows) {
        const uint32_t nr = min(uint32_t(kStTileRows), re - r0);
        for (uint32_t kc = 0; kc < g.K; kc += g.KC) {
            st_mbar_wait(c, c.empty0 + pp.stage * 8, pp.parity ^ 1u);
            const uint32_t full = c.full0 + pp.stage * 8;
            if (lane == 0) mbar_expect_tx(full, nr * row_bytes);
            __syncwarp();
            if (lane < nr)
                bulk_g2s(c.ring_addr + pp.stage * kStStageBytes + lane * pitch, W + (size_t(r0 + lane) * g.K + kc) * 2, row_bytes, full, policy);
            pp.advance(P.n_stages);
        }
    }
}

// ---- consumers: stage the activation rows of a GEMV phase ----------------------------------------------------------------
__device__ __forceinline__ void st_stage_input(const st_params& P, const st_gemv& g, uint32_t li, const st_ctx& c)
{
    const uint32_t tid = threadIdx.x, K = g.K;
    const uint16_t* norm_w = g.pro == PRO_RMSNORM ? reinterpret_cast<const uint16_t*>(reinterpret_cast<const char*>(g.norm_w) + (g.layered ? size_t(li) * P.layer_stride : 0)) : nullptr;
    for (uint32_t m = 0; m < P.rows; m++) {
        uint16_t* dst = reinterpret_cast<uint16_t*>(c.act + size_t(m) * P.act_pitch);
        float part = 0.0f;
        if (g.in_kind == ST_IN_ATTN) {
            // o = r(sum_t p[t] V[t]): the four position splits are joined in split order (nn/attention.h:201-203)
            const float* src = P.opart + size_t(m) * kStSplits * K;
            for (uint32_t k = tid * 4; k < K; k += kStConsumers * 4) {
                float4 a = ldcg_f4(src + k);
#pragma unroll
                for (int s = 1; s < kStSplits; s++) {
                    const float4 b = ldcg_f4(src + size_t(s) * K + k);
                    a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
                }
                uint2 o;
                o.x = uint32_t(f32_to_bf16_bits(a.x)) | (uint32_t(f32_to_bf16_bits(a.y)) << 16);
                o.y = uint32_t(f32_to_bf16_bits(a.z)) | (uint32_t(f32_to_bf16_bits(a.w)) << 16);
                *reinterpret_cast<uint2*>(dst + k) = o;
            }
        } else {
            const uint16_t* xr;
            if (g.in_kind == ST_IN_EMBED) {
                // embedding gather fused into the first phase (kernel/embedding.metal:38-66)
                int32_t id = ldcg_s32(P.ids + m);
                if (id < 0 || uint32_t(id) >= P.vocab) id = 0;
                xr = P.embed_table + size_t(id) * K;
            } else {
                xr = g.x + size_t(m) * g.ldx;
            }
            for (uint32_t k = tid * 8; k < K; k += kStConsumers * 8) {
                const uint4 v = ldcg128(xr + k);
                *reinterpret_cast<uint4*>(dst + k) = v;
                if (g.in_kind == ST_IN_EMBED && blockIdx.x == 0) *reinterpret_cast<uint4*>(P.embed_out + size_t(m) * K + k) = v;
                if (g.pro == PRO_RMSNORM) {
                    float f;
                    f = bf_lo(v.x), part = fmaf(f, f, part);
                    f = bf_hi(v.x), part = fmaf(f, f, part);
                    f = bf_lo(v.y), part = fmaf(f, f, part);
                    f = bf_hi(v.y), part = fmaf(f, f, part);
                    f = bf_lo(v.z), part = fmaf(f, f, part);
                    f = bf_hi(v.z), part = fmaf(f, f, part);
                    f = bf_lo(v.w), part = fmaf(f, f, part);
                    f = bf_hi(v.w), part = fmaf(f, f, part);
                }
            }
        }
        if (g.pro == PRO_RMSNORM) {
            // n = r((0 + w) * x * rsqrt(mean(x^2) + eps))  (kernel/rmsnorm.metal:53-89); each thread re-reads its own elements
            const float total = st_block_sum(part, c.scr);
            const float inv = 1.0f / sqrtf(__fadd_rn(total / float(K), P.eps));
            for (uint32_t k = tid * 8; k < K; k += kStConsumers * 8) {
                const uint4 v = *reinterpret_cast<const uint4*>(dst + k);
                const uint4 gw = *reinterpret_cast<const uint4*>(norm_w + k);
                uint4 o;
#define MC_NORM2(d, vv, gg)                                                                          \
    d = uint32_t(f32_to_bf16_bits(__fmul_rn(__fmul_rn(bf_lo(gg), bf_lo(vv)), inv))) |                \
        (uint32_t(f32_to_bf16_bits(__fmul_rn(__fmul_rn(bf_hi(gg), bf_hi(vv)), inv))) << 16)
                MC_NORM2(o.x, v.x, gw.x);
                MC_NORM2(o.y, v.y, gw.y);
                MC_NORM2(o.z, v.z, gw.z);
                MC_NORM2(o.w, v.w, gw.w);
#undef MC_NORM2
                *reinterpret_cast<uint4*>(dst + k) = o;
            }
        }
    }
    consumer_bar();
}

// greedy argmax state of one epilogue thread (lowest index on ties)
struct st_best {
    float v;
    int32_t i;
};

// ---- consumers: one GEMV phase -----------------------------------------------------------------------------------------
__device__ __forceinline__ void st_consume_gemv(const st_params& P, const st_gemv& g, const st_ctx& c, st_pipe& cp, uint32_t& red_buf, st_best& best, bool track)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t gq = lane >> 2, t = lane & 3;
    const uint32_t pitch = g.KC * 2 + kStPad;
    const uint32_t kw = warp * (g.KC / kStWarps); // this warp's k-slice inside a tile
    const uint32_t ksteps = g.KC / kStWarps / 32;
    const uint32_t brow = gq < P.rows ? gq : 0;   // batch row of this lane's B fragment (unused columns read row 0)
    const uint32_t b_base = c.act_addr + brow * P.act_pitch + (kw + t * 8) * 2;
    const uint32_t a_off = gq * pitch + (kw + t * 8) * 2;
    // epilogue role: thread (row, col) of the 16 x 8 output block
    const uint32_t erow = tid >> 3, ecol = tid & 7;
    const bool etask = tid < 128 && ecol < P.rows;
    uint32_t rb, re;
    st_my_rows(g.N, rb, re);
    for (uint32_t r0 = rb; r0 < re; r0 += kStTileRows) {
        const uint32_t nr = min(uint32_t(kStTileRows), re - r0);
        float resv = 0.0f;
        if (g.epi == EPI_RESIDUAL && etask && erow < nr) resv = bf16_bits_to_f32(ldcg_u16(g.res + size_t(ecol) * g.ldy + r0 + erow));
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (uint32_t kc = 0; kc < g.K; kc += g.KC) {
            st_mbar_wait(c, c.full0 + cp.stage * 8, cp.parity);
            const uint32_t tile = c.ring_addr + cp.stage * kStStageBytes + a_off;
            const uint32_t bk = b_base + kc * 2;
#pragma unroll 4
            for (uint32_t s = 0; s < ksteps; s++) {
                const uint4 alo = lds128(tile + s * 64);
                const uint4 ahi = lds128(tile + 8 * pitch + s * 64);
                const uint4 b = lds128(bk + s * 64);
                mma_bf16_16816(acc, alo.x, ahi.x, alo.y, ahi.y, b.x, b.y);
                mma_bf16_16816(acc, alo.z, ahi.z, alo.w, ahi.w, b.z, b.w);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(c.empty0 + cp.stage * 8);
            cp.advance(P.n_stages);
        }
        // join the 8 k-slices in warp order
        float* rw = c.red + red_buf * (kStWarps * 128) + warp * 128;
        *reinterpret_cast<float2*>(rw + gq * 8 + 2 * t) = make_float2(acc[0], acc[1]);
        *reinterpret_cast<float2*>(rw + (gq + 8) * 8 + 2 * t) = make_float2(acc[2], acc[3]);
        consumer_bar();
        if (etask && erow < nr && !(g.epi == EPI_SWIGLU && (erow & 1u))) {
            const float* rr = c.red + red_buf * (kStWarps * 128) + erow * 8 + ecol;
            float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
            for (int w = 0; w < kStWarps; w++) s0 += rr[w * 128];
            const uint32_t R = r0 + erow;
            const float y0 = rbf(s0); // the bmm output buffer is T (kernel/bmm.metal:76)
            if (g.epi == EPI_SWIGLU) {
                // z = r(silu_T(g) * u), rows (2i, 2i+1) = (w1 row i, w3 row i)  (nn/transformer.h:57-59)
#pragma unroll
                for (int w = 0; w < kStWarps; w++) s1 += rr[w * 128 + 8];
                g.y[size_t(ecol) * g.ldy + (R >> 1)] = f32_to_bf16_bits(__fmul_rn(silu_bf16(y0), rbf(s1)));
            } else if (g.epi == EPI_RESIDUAL) {
                // h = r(x + a)  (nn/transformer.h:133,139)
                g.y[size_t(ecol) * g.ldy + R] = f32_to_bf16_bits(__fadd_rn(resv, y0));
            } else {
                g.y[size_t(ecol) * g.ldy + R] = f32_to_bf16_bits(y0);
                if (track && (y0 > best.v || (y0 == best.v && int32_t(R) < best.i))) best.v = y0, best.i = int32_t(R);
            }
        }
        red_buf ^= 1u;
    }
}

// ---- consumers: attention phase ---------------------------------------------------------------------------------------
// item = (row, head, split): rotate q and the new k (kernel/rope.metal:47-58), append k', v to the cache
// (nn/cache.h:207-214), s = r(r(q.K[t]) * scale), p = r(exp(s) / sum exp(s)) with the sum joined across the four splits,
// partial o = sum_t p[t] V[t] over this split's positions.
template <int HD>
__device__ __forceinline__ void st_attention(const st_params& P, uint32_t li, uint32_t gstep, const st_ctx& c)
{
    constexpr int LPP = HD / 8;        // lanes per cached position (16 bytes each)
    constexpr int SLOTS = kStConsumers / LPP;
    constexpr int IT = 5;              // position sweeps kept in flight
    const uint32_t tid = threadIdx.x;
    const uint32_t H = P.n_heads, KV = P.n_kv_heads, half = HD / 2, QO = H * HD, QKVN = (H + 2 * KV) * HD;
    float* sq = reinterpret_cast<float*>(c.act); // [HD] rotated q
    float* sk = sq + HD;                          // [HD] rotated new k
    float* sv = sk + HD;                          // [HD] new v
    float* spart = sv + HD;                       // [SLOTS][HD]
    float* sp = spart + SLOTS * HD;               // [chunk] scores / probabilities
    const uint32_t slot = tid / LPP, dl = tid % LPP;
    const size_t kv_off = size_t(li) * P.kv_layer_stride;
    const uint32_t n_items = P.rows * H * kStSplits;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t split = item & (kStSplits - 1), head = (item / kStSplits) % H, row = (item / kStSplits) / H;
        const int32_t seq = P.row_seq[row];
        const uint32_t pos = uint32_t(ldcg_s32(P.pos + row));
        const uint32_t np = pos + 1;
        const uint32_t kvh = head / (H / KV);
        const uint32_t chunk = (np + kStSplits - 1) / kStSplits;
        const uint32_t t0 = min(np, split * chunk), t1 = min(np, t0 + chunk);
        const uint32_t tc1 = min(t1, pos); // positions below `pos` come from the cache, `pos` itself is fresh
        const bool own = pos >= t0 && pos < t1;
        const size_t coff = kv_off + (size_t(seq) * KV + kvh) * P.max_seq * HD;
        const uint16_t* Kc = P.kcache + coff;
        const uint16_t* Vc = P.vcache + coff;

        // first block of K rows: requested before anything else is touched
        uint4 kreg[IT], vreg[IT];
#pragma unroll
        for (int i = 0; i < IT; i++) {
            const uint32_t tt = t0 + i * SLOTS + slot;
            if (tt < tc1) kreg[i] = ldcg128(Kc + size_t(tt) * HD + dl * 8);
        }
        const uint16_t* qrow = P.qkv + size_t(row) * QKVN;
        if (tid < half) {
            const float cs = P.fcos[size_t(pos) * half + tid], sn = P.fsin[size_t(pos) * half + tid];
            const float q0 = bf16_bits_to_f32(ldcg_u16(qrow + head * HD + tid)), q1 = bf16_bits_to_f32(ldcg_u16(qrow + head * HD + tid + half));
            sq[tid] = rbf(__fsub_rn(__fmul_rn(cs, q0), __fmul_rn(sn, q1)));
            sq[tid + half] = rbf(__fadd_rn(__fmul_rn(sn, q0), __fmul_rn(cs, q1)));
            if (own) {
                const uint16_t* kr = qrow + (H + kvh) * HD;
                const float k0 = bf16_bits_to_f32(ldcg_u16(kr + tid)), k1 = bf16_bits_to_f32(ldcg_u16(kr + tid + half));
                const float o0 = rbf(__fsub_rn(__fmul_rn(cs, k0), __fmul_rn(sn, k1)));
                const float o1 = rbf(__fadd_rn(__fmul_rn(sn, k0), __fmul_rn(cs, k1)));
                sk[tid] = o0, sk[tid + half] = o1;
                if (head % (H / KV) == 0) { // one writer per kv head
                    uint16_t* kd = P.kcache + coff + size_t(pos) * HD;
                    kd[tid] = f32_to_bf16_bits(o0), kd[tid + half] = f32_to_bf16_bits(o1);
                }
            }
        } else if (own && tid >= 64 && tid < 64 + HD) {
            const uint32_t d = tid - 64;
            const uint16_t vb = ldcg_u16(qrow + (H + KV + kvh) * HD + d);
            sv[d] = bf16_bits_to_f32(vb);
            if (head % (H / KV) == 0) P.vcache[coff + size_t(pos) * HD + d] = vb;
        }
        consumer_bar();
        float qv[8];
#pragma unroll
        for (int i = 0; i < 8; i++) qv[i] = sq[dl * 8 + i];
        // scores of the cached positions
        for (uint32_t tb = t0; tb < tc1; tb += IT * SLOTS) {
            if (tb != t0) {
#pragma unroll
                for (int i = 0; i < IT; i++) {
                    const uint32_t tt = tb + i * SLOTS + slot;
                    if (tt < tc1) kreg[i] = ldcg128(Kc + size_t(tt) * HD + dl * 8);
                }
            }
#pragma unroll
            for (int i = 0; i < IT; i++) {
                const uint32_t tt = tb + i * SLOTS + slot;
                float d = 0.0f;
                if (tt < tc1) {
                    const uint4 kv = kreg[i];
                    d = fmaf(qv[0], bf_lo(kv.x), d);
                    d = fmaf(qv[1], bf_hi(kv.x), d);
                    d = fmaf(qv[2], bf_lo(kv.y), d);
                    d = fmaf(qv[3], bf_hi(kv.y), d);
                    d = fmaf(qv[4], bf_lo(kv.z), d);
                    d = fmaf(qv[5], bf_hi(kv.z), d);
                    d = fmaf(qv[6], bf_lo(kv.w), d);
                    d = fmaf(qv[7], bf_hi(kv.w), d);
                }
#pragma unroll
                for (int off = LPP / 2; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
                if (tt < tc1 && dl == 0) sp[tt - t0] = rbf(__fmul_rn(rbf(d), P.scale));
            }
        }
        // first block of V rows goes in flight before the exchange
#pragma unroll
        for (int i = 0; i < IT; i++) {
            const uint32_t tt = t0 + i * SLOTS + slot;
            if (tt < tc1) vreg[i] = ldcg128(Vc + size_t(tt) * HD + dl * 8);
        }
        if (own && tid < 32) {
            float d = 0.0f;
            for (uint32_t i = tid; i < uint32_t(HD); i += 32) d = fmaf(sq[i], sk[i], d);
            d = warp_sum(d);
            if (tid == 0) sp[pos - t0] = rbf(__fmul_rn(rbf(d), P.scale));
        }
        consumer_bar();
        // softmax without max subtraction (kernel/softmax.metal:40-80); the exp-sum is joined across the 4 splits
        const uint32_t n_local = t1 - t0;
        float part = 0.0f;
        for (uint32_t i = tid; i < n_local; i += kStConsumers) part += expf(sp[i]);
        const float local_sum = st_block_sum(part, c.scr);
        const uint32_t group = item / kStSplits;
        if (tid == 0) {
            P.xsum[size_t(group) * kStSplits + split] = local_sum;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(P.acnt + group) : "memory");
            st_spin_ge(c, P.acnt + group, kStSplits * (gstep * P.n_layers + li + 1));
        }
        consumer_bar();
        float total = 0.0f;
#pragma unroll
        for (int s = 0; s < kStSplits; s++) total += ldcg_f32(P.xsum + size_t(group) * kStSplits + s);
        const float inv = 1.0f / total;
        for (uint32_t i = tid; i < n_local; i += kStConsumers) sp[i] = rbf(__fmul_rn(expf(sp[i]), inv));
        consumer_bar();
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = 0.0f;
        for (uint32_t tb = t0; tb < tc1; tb += IT * SLOTS) {
            if (tb != t0) {
#pragma unroll
                for (int i = 0; i < IT; i++) {
                    const uint32_t tt = tb + i * SLOTS + slot;
                    if (tt < tc1) vreg[i] = ldcg128(Vc + size_t(tt) * HD + dl * 8);
                }
            }
#pragma unroll
            for (int i = 0; i < IT; i++) {
                const uint32_t tt = tb + i * SLOTS + slot;
                if (tt < tc1) {
                    const uint4 vv = vreg[i];
                    const float pt = sp[tt - t0];
                    acc[0] = fmaf(pt, bf_lo(vv.x), acc[0]);
                    acc[1] = fmaf(pt, bf_hi(vv.x), acc[1]);
                    acc[2] = fmaf(pt, bf_lo(vv.y), acc[2]);
                    acc[3] = fmaf(pt, bf_hi(vv.y), acc[3]);
                    acc[4] = fmaf(pt, bf_lo(vv.z), acc[4]);
                    acc[5] = fmaf(pt, bf_hi(vv.z), acc[5]);
                    acc[6] = fmaf(pt, bf_lo(vv.w), acc[6]);
                    acc[7] = fmaf(pt, bf_hi(vv.w), acc[7]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) spart[slot * HD + dl * 8 + i] = acc[i];
        consumer_bar();
        if (tid < HD) {
            float o = 0.0f;
            for (int s = 0; s < SLOTS; s++) o += spart[s * HD + tid];
            if (own) o = fmaf(sp[pos - t0], sv[tid], o);
            P.opart[(size_t(row) * kStSplits + split) * QO + head * HD + tid] = o;
        }
        consumer_bar(); // the scratch is reused by the next item
    }
}

// ---- the kernel ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kStThreads, 1) decode_stream_kernel(const __grid_constant__ st_params P)
{
    extern __shared__ __align__(128) unsigned char smem[];
    st_ctx c;
    c.full0 = smem_u32(smem);
    c.empty0 = c.full0 + kStMaxStages * 8;
    c.dead = reinterpret_cast<volatile int*>(smem + 2 * kStMaxStages * 8);
    c.scr = reinterpret_cast<float*>(smem + 256);
    c.red = reinterpret_cast<float*>(smem + kStHdrBytes);
    c.act = smem + kStHdrBytes + kStRedBytes;
    c.act_addr = smem_u32(c.act);
    c.ring_addr = c.act_addr + P.act_bytes;
    c.err = P.err;
    const uint32_t tid = threadIdx.x;
    const unsigned G = gridDim.x;
    if (tid == 0) {
        for (uint32_t s = 0; s < P.n_stages; s++) {
            mbar_init(c.full0 + s * 8, 1);
            mbar_init(c.empty0 + s * 8, kStWarps);
        }
        *c.dead = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const uint32_t phases_per_step = P.n_layers * 5 + 1;

    if (tid >= kStConsumers) {
        // ---- producer warp: the whole launch's weight stream, never blocked by a phase barrier
        st_pipe pp{0, 0};
        const uint64_t policy = policy_evict_first();
        for (uint32_t step = 0; step < P.steps; step++) {
            for (uint32_t li = 0; li < P.n_layers; li++) {
                st_produce_gemv(P, P.g[0], li, c, pp, policy);
                st_produce_gemv(P, P.g[1], li, c, pp, policy);
                st_produce_gemv(P, P.g[2], li, c, pp, policy);
                st_produce_gemv(P, P.g[3], li, c, pp, policy);
            }
            st_produce_gemv(P, P.g[4], 0, c, pp, policy);
        }
        return;
    }

    // ---- consumer warps
    st_pipe cp{0, 0};
    uint32_t red_buf = 0;
    for (uint32_t step = 0; step < P.steps; step++) {
        unsigned gphase = step * phases_per_step; // phases completed before this one, launch-wide
        st_best best{-INFINITY, 0x7fffffff};
        for (uint32_t li = 0; li <= P.n_layers; li++) {
            const bool is_head = li == P.n_layers;
            for (uint32_t kind = 0; kind < (is_head ? 1u : 5u); kind++, gphase++) {
                unsigned long long* tm = P.timing ? P.timing + size_t(gphase) * 4 : nullptr;
                st_stamp(tm, 0);
                // wait for this phase's input: the previous phase of every CTA, or (first phase of a later step) the sampler tail
                if (gphase != 0) {
                    if (li == 0 && kind == 0) st_grid_wait(c, P.step_done, step);
                    else st_grid_wait(c, P.bar, gphase * G);
                }
                st_stamp(tm, 1);
                if (!is_head && kind == 1) {
                    if (P.head_dim == 64) st_attention<64>(P, li, step, c);
                    else st_attention<128>(P, li, step, c);
                    st_stamp(tm, 2);
                } else {
                    const st_gemv& g = P.g[is_head ? 4u : (kind == 0 ? 0u : kind - 1)];
                    st_stage_input(P, g, is_head ? 0 : li, c);
                    st_stamp(tm, 2);
                    st_consume_gemv(P, g, c, cp, red_buf, best, is_head);
                }
                st_stamp(tm, 3);
                if (is_head) {
                    // per-CTA argmax partial of every activation row
                    consumer_bar();
                    float* bv = c.red;
                    int32_t* bi = reinterpret_cast<int32_t*>(c.red + 128);
                    if (tid < 128) bv[tid] = best.v, bi[tid] = best.i;
                    consumer_bar();
                    if (tid < P.rows) {
                        float v = -INFINITY;
                        int32_t i = 0x7fffffff;
                        for (int r = 0; r < 16; r++) {
                            const float ov = bv[r * 8 + tid];
                            const int32_t oi = bi[r * 8 + tid];
                            if (ov > v || (ov == v && oi < i)) v = ov, i = oi;
                        }
                        P.am_val[tid * G + blockIdx.x] = v;
                        P.am_idx[tid * G + blockIdx.x] = i;
                    }
                }
                st_grid_arrive(P.bar);
            }
        }
        // ---- sampler tail (greedy): CTA 0 joins the per-CTA partials, feeds the id back and advances the position
        if (blockIdx.x == 0) {
            st_grid_wait(c, P.bar, gphase * G);
            const uint32_t warp = tid >> 5, lane = tid & 31;
            if (warp < P.rows) {
                float v = -INFINITY;
                int32_t i = 0x7fffffff;
                for (unsigned b = lane; b < G; b += 32) {
                    const float ov = ldcg_f32(P.am_val + warp * G + b);
                    const int32_t oi = ldcg_s32(P.am_idx + warp * G + b);
                    if (ov > v || (ov == v && oi < i)) v = ov, i = oi;
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, v, off);
                    const int32_t oi = __shfl_xor_sync(0xffffffffu, i, off);
                    if (ov > v || (ov == v && oi < i)) v = ov, i = oi;
                }
                if (lane == 0) {
                    if (i == 0x7fffffff) i = 0; // all-NaN / -inf row: argmax keeps index 0
                    const int32_t sc = ldcg_s32(P.step_counter);
                    P.out_log[size_t(sc) * P.rows + warp] = i;
                    if (P.advance) {
                        P.ids[warp] = i;
                        P.pos[warp] = ldcg_s32(P.pos + warp) + 1;
                    }
                }
            }
            consumer_bar();
            if (tid == 0) {
                *P.step_counter = ldcg_s32(P.step_counter) + 1;
                if (step + 1 == P.steps) {
                    // every CTA has made its last arrival and left its last wait: ready for the next launch
                    *P.bar = 0;
                    for (uint32_t i = 0; i < P.rows * P.n_heads; i++) P.acnt[i] = 0;
                    *P.step_done = 0;
                } else {
                    __threadfence();
                    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(P.step_done), "r"(step + 1) : "memory");
                }
            }
        }
    }
}

} // namespace mc
