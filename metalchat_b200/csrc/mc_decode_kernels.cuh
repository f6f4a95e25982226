// metalchat_b200/csrc/mc_decode_kernels.cuh — the fused sm_100a kernels of the decode engine.
//
// Where the reference runs ~30 Metal kernels per transformer block and rounds to bf16 after each
// (SURVEY.md §3.1, §8a "rounding chain"), the engine runs five kernels per block and keeps every
// one of those rounding points r(.) inside them:
//
//   K1  gemv<PRO_RMSNORM, EPI_QKV>     attention_norm + wq|wk|wv + rope(q,k) + KV-cache append
//   K5  attn_decode                    scores -> scale -> softmax(no max) -> PV, GQA by indexing
//   K1  gemv<PRO_NONE,    EPI_RESIDUAL> wo + residual add
//   K1  gemv<PRO_RMSNORM, EPI_SWIGLU>  ffn_norm + w1|w3 (row-interleaved) + silu*mul
//   K1  gemv<PRO_NONE,    EPI_RESIDUAL> w2 + residual add
//
// The GEMV streams each weight row exactly once with 128-bit L1-bypassing loads, two rows per
// warp, fp32 accumulation, warp-shuffle reduction; quantised weights are dequantised in
// registers with the reference's double rounding r(r(q)*r(s)) (kernel/mul.metal:76-77).
#pragma once
#include "mc_common.cuh"

// non-template kernels of the shared headers get internal linkage: several translation units include them
#define MC_KERNEL static __global__

namespace mc {

constexpr int kGemvThreads = 256;
constexpr int kGemvWarps = kGemvThreads / 32;
constexpr int kMaxMB = 4; // activation rows handled by one GEMV pass

enum { PRO_NONE = 0, PRO_RMSNORM = 1 };
enum { EPI_NONE = 0, EPI_QKV = 1, EPI_RESIDUAL = 2, EPI_SWIGLU = 3, EPI_PARTIAL_TP = 4 };
enum { WF_BF16 = 0, WF_W4 = 1, WF_W8ROW = 2, WF_W8G = 3 };

// ---- small device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream8(const void* p)
{
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ float rbf(float f) { return bf16_bits_to_f32(f32_to_bf16_bits(f)); } // r(.)
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
// PDL: wait for the producer grid's memory / let the consumer grid start its prologue
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// silu evaluated in bf16 steps: x / (T(1) + T(exp(-x)))  (kernel/activation.metal:34-35, quirk Q5)
__device__ __forceinline__ float silu_bf16(float g)
{
    const float e = rbf(expf(-g));
    const float d = rbf(__fadd_rn(1.0f, e));
    return rbf(__fdiv_rn(g, d));
}

// ---- TP exchange: all-reduce fused into the producing GEMV and the consuming GEMV -----------------------------------
// Row-parallel linears (wo, w2) end in an all-reduce of `rows x dim` partial sums.  Instead of a
// collective call between kernels, the PRODUCER GEMV's epilogue pushes its fp32 partials straight into every peer's
// exchange buffer over NVLink (P2P stores), the last CTA to finish publishes an epoch flag to all peers, and the
// CONSUMER GEMV's prologue waits for the flags and sums the `world` partials from LOCAL memory in rank order, adds the
// residual and rounds once:  h = r(x + r(sum_k partial_k)).  Buffers are double-buffered by epoch parity.
constexpr int kTpMaxWorld = 8;
struct tp_exchange {
    uint32_t world, rank;
    uint32_t rows_max, dim;            // buffer geometry: buf[parity][src_rank][rows_max][dim] fp32
    float* peer_buf[kTpMaxWorld];      // exchange buffer of every rank (this rank's own included), mapped on this GPU
    uint32_t* peer_flag[kTpMaxWorld];  // flags[src_rank] of every rank
    unsigned* done;                    // local: CTAs of the producer that have finished
    unsigned* epoch;                   // local: exchanges published so far
    int* err;                          // local: set when a flag wait times out
    uint32_t nowait;                   // diagnostics (MC_TP_NOWAIT): do not wait for the peers' flags -- results are wrong, the step time tells what the waits cost
};
__device__ __forceinline__ float* tp_slot(const tp_exchange& t, float* buf, uint32_t parity, uint32_t src)
{
    return buf + (size_t(parity) * t.world + src) * t.rows_max * t.dim;
}
// producer side, after the epilogue stores (which go to this rank's OWN slot): the last CTA to finish copies the finished partial
// vector to every peer with 16-byte stores over NVLink (scattered 4-byte remote stores from every warp cost ~30-50 us per exchange:
// profiles/bench_r01b_8b_bf16_tp*.json), then publishes the new epoch to every peer
__device__ __forceinline__ void tp_publish(const tp_exchange& t, uint32_t rows, uint32_t parity, unsigned* smem_word)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(t.done, 1u);
        *smem_word = prev == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (*smem_word == 0u) return;
    __threadfence(); // the other CTAs' partials (released by their fences before the counter) are visible from here on
    const uint32_t n4 = rows * t.dim / 4;
    const float4* src = reinterpret_cast<const float4*>(tp_slot(t, t.peer_buf[t.rank], parity, t.rank));
    for (uint32_t k = 0; k < t.world; k++) {
        if (k == t.rank) continue;
        float4* dst = reinterpret_cast<float4*>(tp_slot(t, t.peer_buf[k], parity, t.rank));
        for (uint32_t i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = __ldcg(src + i);
    }
    __syncthreads(); // the copies of every thread happen before thread 0's single system-scope fence (cumulativity)
    if (threadIdx.x == 0) {
        *t.done = 0;
        const unsigned e = *t.epoch + 1;
        *t.epoch = e;
        __threadfence_system();
        for (uint32_t k = 0; k < t.world; k++) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(t.peer_flag[k] + t.rank), "r"(e) : "memory");
    }
}
// consumer side: wait until every rank has published epoch e; returns e (0 on timeout)
__device__ __forceinline__ unsigned tp_wait(const tp_exchange& t, unsigned* smem_word)
{
    if (threadIdx.x == 0) {
        const unsigned e = *reinterpret_cast<volatile unsigned*>(t.epoch);
        unsigned ok = e;
        for (uint32_t k = 0; k < t.world && !t.nowait; k++) {
            unsigned v = 0;
            unsigned long long spins = 0;
            for (;;) {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(t.peer_flag[t.rank] + k) : "memory");
                if (v >= e) break;
                if (++spins > (1ull << 23)) {
                    atomicExch(t.err, 2);
                    ok = 0;
                    break;
                }
            }
        }
        *smem_word = ok;
    }
    __syncthreads();
    return *smem_word;
}

struct gemv_params {
    const void* W;         // weight rows, format WF_*
    const float* scales;   // WF_W4/WF_W8G: [N, K/group]; WF_W8ROW: [N]
    uint32_t N, K;         // local rows / reduction length
    uint32_t rows;         // activation rows (<= MB)
    const uint16_t* x;     // [rows, ldx] bf16 input
    uint32_t ldx;
    // PRO_RMSNORM
    const uint16_t* norm_w; // [K]
    float eps;
    // EPI_NONE / EPI_RESIDUAL / EPI_SWIGLU
    uint16_t* y;           // [rows, ldy]
    uint32_t ldy;
    const uint16_t* res;   // EPI_RESIDUAL: [rows, ldy]
    // EPI_QKV
    uint16_t* q;           // [rows, Hl*hd]
    uint16_t* kcache;      // this layer: [n_seqs, KVl, S, hd]
    uint16_t* vcache;
    const float* fcos;     // [2S, hd/2]
    const float* fsin;
    const int32_t* row_seq; // [rows]
    const int32_t* row_pos; // [rows]
    uint32_t n_heads, n_kv_heads, head_dim, max_seq;
    // LoRA epilogue (quantised layers): y += r(r(B . ax) * lora_scale) with ax = r(A . x) precomputed
    const uint16_t* lora_b; // [N, rank]
    const uint16_t* lora_ax; // [rows, rank] bf16
    uint32_t lora_rank;
    float lora_scale;      // already rounded to bf16
    uint32_t ksplit;       // warps per unit (1, 2, 4 or 8): intra-CTA split of the reduction
    int32_t pro, epi;      // PRO_* / EPI_* for the megakernel, which selects them at run time (template value -1)
    // megakernel only: the first phase gathers its input rows from the embedding table (and CTA 0 stores them to x)
    const uint16_t* embed_table;
    const int32_t* embed_ids;
    uint16_t* embed_out;
    // megakernel only: greedy argmax fused into the vocab projection; per-CTA partials [rows][gridDim.x]
    float* am_val;
    int32_t* am_idx;
    // tensor parallelism (see "TP exchange" below); tp.world == 0 when unused
    tp_exchange tp;
    int32_t tp_reduce;         // prologue: build the input rows from the peers' partial sums (+ residual)
    const uint16_t* tp_res;    // [rows, ldx] residual added to the reduced sum
    uint16_t* tp_out;          // [rows, ldx] where CTA 0 stores the reduced rows (the next residual)
    uint32_t tp_col0;          // EPI_PARTIAL_TP: first column of the exchange row this GEMV fills (quantised models: the adaptor's A . x goes behind the dim main sums)
    uint32_t tp_hold;          // EPI_PARTIAL_TP: fill the slot only; the GEMV launched next publishes the whole exchange row
};

// ---- grid-wide phase barrier of the persistent decode kernel ---------------------------------------------------
// Monotonic arrival counter in global memory: a phase is complete when every CTA has arrived once more.  The wait
// is placed AFTER a phase has requested its first weight chunk, so the weight stream never drains at a barrier.
struct mega_sync {
    unsigned* bar;       // arrival counter (reset to 0 by the last phase)
    int* err;            // set to 1 if a wait times out (never hang the GPU)
    unsigned target;     // arrivals that must be visible before this phase may read its input
    const void* pf_ptr;  // weights of a later phase to pull into L2 while this phase runs
    size_t pf_bytes;
    unsigned long long* timing; // diagnostics (nullable): CTA 0 stamps globaltimer at phase entry / after wait / at arrive
};
__device__ __forceinline__ void stamp(unsigned long long* t, unsigned idx)
{
    if (t && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long v;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
        t[idx] = v;
    }
}
__device__ __forceinline__ void grid_arrive(unsigned* bar)
{
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
}
__device__ __forceinline__ void grid_wait(const mega_sync& sy)
{
    if (threadIdx.x == 0) {
        unsigned v = 0;
        unsigned long long spins = 0;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(sy.bar) : "memory");
            if (v >= sy.target) break;
            ++spins;
            if ((spins & 1023) == 0 && *reinterpret_cast<volatile int*>(sy.err) != 0) break; // another CTA gave up
            if (spins > (1ull << 22)) { // ~a second: report instead of hanging the GPU
                atomicExch(sy.err, 1);
                break;
            }
        }
    }
    __syncthreads();
}
// every thread asks the L2 for one 4 KiB piece of [ptr, ptr+bytes), pieces dealt round-robin over CTAs
__device__ __forceinline__ void l2_prefetch_slice(const void* ptr, size_t bytes)
{
    constexpr size_t kPiece = 4096;
    const size_t pieces = bytes / kPiece;
    for (size_t c = size_t(threadIdx.x) * gridDim.x + blockIdx.x; c < pieces; c += size_t(blockDim.x) * gridDim.x) {
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(static_cast<const char*>(ptr) + c * kPiece), "r"(uint32_t(kPiece)) : "memory");
    }
}

// dot of 8 bf16 weights (uint4) with 8 bf16 activations (uint4), fp32 accumulate in ascending order
__device__ __forceinline__ float dot8(const uint4& w, const uint4& x, float acc)
{
    acc = fmaf(bf_lo(w.x), bf_lo(x.x), acc);
    acc = fmaf(bf_hi(w.x), bf_hi(x.x), acc);
    acc = fmaf(bf_lo(w.y), bf_lo(x.y), acc);
    acc = fmaf(bf_hi(w.y), bf_hi(x.y), acc);
    acc = fmaf(bf_lo(w.z), bf_lo(x.z), acc);
    acc = fmaf(bf_hi(w.z), bf_hi(x.z), acc);
    acc = fmaf(bf_lo(w.w), bf_lo(x.w), acc);
    acc = fmaf(bf_hi(w.w), bf_hi(x.w), acc);
    return acc;
}

// Block-wide sum with a fixed partition (warp butterfly, then 8 warp totals in order).
__device__ __forceinline__ float block_sum_256(float v, float* scratch /* [8] */)
{
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < kGemvWarps; i++) t += scratch[i];
    __syncthreads();
    return t;
}

// unit -> the two weight rows a warp group computes
__device__ __forceinline__ void unit_rows(int epi, const gemv_params& p, uint32_t u, uint32_t& r0, uint32_t& r1)
{
    if (epi == EPI_QKV) {
        const uint32_t half = p.head_dim >> 1;
        const uint32_t head = u / half, j = u - head * half;
        r0 = head * p.head_dim + j; // rope pair (j, j + hd/2) of one head (kernel/rope.metal:50-57)
        r1 = r0 + half;
    } else {
        r0 = 2 * u;
        r1 = r0 + 1;
    }
}

// K1: y = epilogue( prologue(x) . W^T ).  grid: any; CTA = 8 warps = (8/ksplit) units x ksplit k-slices.
// MEGA = false: a stand-alone kernel (programmatic dependent launch between kernels);
// MEGA = true : one phase of the persistent decode kernel (grid barrier between phases).
// PRO_T / EPI_T >= 0 fix the prologue / epilogue at compile time; -1 takes them from p.pro / p.epi.
struct phase_adj {
    size_t w_off;   // byte offset of this layer's weights / norm weights in the arena
    size_t kv_off;  // element offset of this layer's KV cache
    bool embed;     // gather the input rows from the embedding table
};
template <int MB, int PRO_T, int EPI_T, bool MEGA, int KS_T>
__device__ __forceinline__ void gemv_body(const gemv_params& p, unsigned char* smem, const mega_sync& sy, const phase_adj& adj)
{
    const int PRO = PRO_T >= 0 ? PRO_T : p.pro;
    const int EPI = EPI_T >= 0 ? EPI_T : p.epi;
    uint16_t* sx = reinterpret_cast<uint16_t*>(smem);                      // [MB][K] bf16
    float* sred = reinterpret_cast<float*>(smem + size_t(MB) * p.K * 2);   // [8 warps][2][MB]
    float* sscr = sred + kGemvWarps * 2 * MB;                              // [8]

    const uint32_t KSPLIT = KS_T > 0 ? uint32_t(KS_T) : p.ksplit; // compile-time in the stand-alone kernels
    const uint32_t UPC = kGemvWarps / KSPLIT; // units per CTA iteration
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t slot = warp / KSPLIT, ks = warp % KSPLIT;
    const uint32_t units = p.N >> 1;
    const uint32_t kslice = p.K / KSPLIT;
    const uint32_t kbeg = ks * kslice, kend = kbeg + kslice;
    const uint16_t* Wb = reinterpret_cast<const uint16_t*>(static_cast<const char*>(p.W) + adj.w_off);
    const uint16_t* norm_w = reinterpret_cast<const uint16_t*>(reinterpret_cast<const char*>(p.norm_w) + adj.w_off);

    uint32_t u = blockIdx.x * UPC + slot;
    const uint32_t ustride = gridDim.x * UPC;

    // -- prefetch the first weight chunk before touching the activations (they may still be in
    //    flight from the producer kernel under programmatic dependent launch)
    constexpr int U = MEGA ? 2 : 4; // 16-byte loads in flight per row per lane (megakernel phases stream L2-prefetched data)
    uint4 w0[U], w1[U];
    uint32_t r0 = 0, r1 = 0;
    if (u < units) {
        unit_rows(EPI, p, u, r0, r1);
#pragma unroll
        for (int i = 0; i < U; i++) {
            const uint32_t k = kbeg + (i * 32 + lane) * 8;
            if (k < kend) {
                w0[i] = ldg_stream(Wb + size_t(r0) * p.K + k);
                w1[i] = ldg_stream(Wb + size_t(r1) * p.K + k);
            }
        }
    }
    if (MEGA) {
        stamp(sy.timing, 0);
        if (sy.pf_bytes) l2_prefetch_slice(sy.pf_ptr, sy.pf_bytes);
        grid_wait(sy);
        stamp(sy.timing, 1);
    } else {
        pdl_launch_dependents();
        pdl_wait();
    }

    // -- tensor-parallel input: h = r(res + r(sum over ranks of the pushed fp32 partials)), summed in rank order
    const uint16_t* x_in = p.x;
    if (p.tp_reduce) {
        const unsigned e = tp_wait(p.tp, reinterpret_cast<unsigned*>(sscr));
        const uint32_t parity = (e - 1) & 1u;
        const float* mine = p.tp.peer_buf[p.tp.rank];
        // four consecutive k per thread, all loads of a row issued before the first use (a scalar loop serialised K / 256 L2 round trips:
        // ~10 us per exchange, profiles/r01b_tp2_per_kernel_profile_8b_v2.txt); ranks are still summed in rank order
        for (uint32_t m = 0; m < p.rows; m++) {
#pragma unroll 2
            for (uint32_t k = threadIdx.x * 4; k < p.K; k += kGemvThreads * 4) {
                float4 part[kTpMaxWorld];
#pragma unroll
                for (uint32_t src = 0; src < uint32_t(kTpMaxWorld); src++)
                    if (src < p.tp.world) part[src] = __ldcg(reinterpret_cast<const float4*>(tp_slot(p.tp, const_cast<float*>(mine), parity, src) + size_t(m) * p.tp.dim + k));
                const uint2 rv = *reinterpret_cast<const uint2*>(p.tp_res + size_t(m) * p.ldx + k);
                float4 sum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
                for (uint32_t src = 0; src < uint32_t(kTpMaxWorld); src++)
                    if (src < p.tp.world) sum.x += part[src].x, sum.y += part[src].y, sum.z += part[src].z, sum.w += part[src].w;
                uint2 hv;
                hv.x = uint32_t(f32_to_bf16_bits(__fadd_rn(bf_lo(rv.x), rbf(sum.x)))) | (uint32_t(f32_to_bf16_bits(__fadd_rn(bf_hi(rv.x), rbf(sum.y)))) << 16);
                hv.y = uint32_t(f32_to_bf16_bits(__fadd_rn(bf_lo(rv.y), rbf(sum.z)))) | (uint32_t(f32_to_bf16_bits(__fadd_rn(bf_hi(rv.y), rbf(sum.w)))) << 16);
                *reinterpret_cast<uint2*>(sx + size_t(m) * p.K + k) = hv;
                if (blockIdx.x == 0) *reinterpret_cast<uint2*>(p.tp_out + size_t(m) * p.ldx + k) = hv;
            }
        }
        __syncthreads();
        x_in = nullptr; // the rows are already staged in shared memory
    }

    // -- prologue: stage the activation rows in shared memory as bf16
    if (PRO == PRO_RMSNORM) {
        // n = r((0 + w) * x * rsqrt(mean(x^2) + eps))  (kernel/rmsnorm.metal:53-89)
        for (int m = 0; m < MB; m++) {
            if (uint32_t(m) < p.rows) {
                const uint16_t* xr = x_in ? x_in + size_t(m) * p.ldx : sx + size_t(m) * p.K;
                if (MEGA && adj.embed) {
                    // embedding gather fused into the first phase (kernel/embedding.metal:38-66)
                    xr = p.embed_table + size_t(p.embed_ids[m]) * p.K;
                    if (blockIdx.x == 0) {
                        for (uint32_t k = threadIdx.x * 8; k < p.K; k += kGemvThreads * 8)
                            *reinterpret_cast<uint4*>(p.embed_out + size_t(m) * p.ldx + k) = *reinterpret_cast<const uint4*>(xr + k);
                    }
                }
                float part = 0.0f;
                for (uint32_t k = threadIdx.x * 8; k < p.K; k += kGemvThreads * 8) {
                    const uint4 v = *reinterpret_cast<const uint4*>(xr + k);
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
                const float total = block_sum_256(part, sscr);
                const float inv = 1.0f / sqrtf(__fadd_rn(total / float(p.K), p.eps));
                for (uint32_t k = threadIdx.x * 8; k < p.K; k += kGemvThreads * 8) {
                    const uint4 v = *reinterpret_cast<const uint4*>(xr + k);
                    const uint4 g = *reinterpret_cast<const uint4*>(norm_w + k);
                    uint4 o;
#define MC_NORM2(dst, vv, gg)                                                                                   \
    dst = uint32_t(f32_to_bf16_bits(__fmul_rn(__fmul_rn(bf_lo(gg), bf_lo(vv)), inv))) |                         \
          (uint32_t(f32_to_bf16_bits(__fmul_rn(__fmul_rn(bf_hi(gg), bf_hi(vv)), inv))) << 16)
                    MC_NORM2(o.x, v.x, g.x);
                    MC_NORM2(o.y, v.y, g.y);
                    MC_NORM2(o.z, v.z, g.z);
                    MC_NORM2(o.w, v.w, g.w);
#undef MC_NORM2
                    *reinterpret_cast<uint4*>(sx + size_t(m) * p.K + k) = o;
                }
            }
        }
    } else if (x_in) {
        for (int m = 0; m < MB; m++) {
            if (uint32_t(m) < p.rows) {
                const uint16_t* xr = x_in + size_t(m) * p.ldx;
                for (uint32_t k = threadIdx.x * 8; k < p.K; k += kGemvThreads * 8)
                    *reinterpret_cast<uint4*>(sx + size_t(m) * p.K + k) = *reinterpret_cast<const uint4*>(xr + k);
            }
        }
    }
    // epilogue operands are requested before the weight stream is consumed so that their latency overlaps it
    const bool fin = ks == 0 && lane < p.rows && lane < uint32_t(MB); // this lane finalises activation row `lane`
    int32_t my_pos = 0, my_seq = 0;
    if (EPI == EPI_QKV && fin) my_pos = p.row_pos[lane], my_seq = p.row_seq[lane];
    uint32_t tp_parity = 0;
    if (EPI == EPI_PARTIAL_TP) tp_parity = *reinterpret_cast<volatile unsigned*>(p.tp.epoch) & 1u; // the epoch this launch will publish is +1
    float best_v = -INFINITY; // fused greedy argmax (lowest index on ties)
    int32_t best_i = 0x7fffffff;
    __syncthreads();

    for (; u - slot < units; u += ustride) { // loop bound is CTA-uniform (u - slot is the CTA's first unit)
        const bool active = u < units;
        float ep0 = 0.0f, ep1 = 0.0f; // EPI_RESIDUAL: residual values; EPI_QKV: cos, sin
        if (active && fin) {
            if (EPI == EPI_RESIDUAL) {
                ep0 = bf16_bits_to_f32(p.res[size_t(lane) * p.ldy + r0]);
                ep1 = bf16_bits_to_f32(p.res[size_t(lane) * p.ldy + r1]);
            } else if (EPI == EPI_QKV) {
                const uint32_t half = p.head_dim >> 1;
                const uint32_t j = r0 % p.head_dim;
                ep0 = p.fcos[size_t(my_pos) * half + j];
                ep1 = p.fsin[size_t(my_pos) * half + j];
            }
        }
        float acc0[MB], acc1[MB];
#pragma unroll
        for (int m = 0; m < MB; m++) acc0[m] = 0.0f, acc1[m] = 0.0f;
        uint32_t n0 = 0, n1 = 0;
        const uint32_t un = u + ustride;
        const bool next_active = un < units;
        if (active) {
            for (uint32_t kc = kbeg; kc < kend; kc += U * 256) {
                // issue the next chunk (or the first chunk of this warp's next unit) before consuming
                uint4 nw0[U], nw1[U];
                const uint32_t kn = kc + U * 256;
                const bool more = kn < kend;
                if (more) {
#pragma unroll
                    for (int i = 0; i < U; i++) {
                        const uint32_t k = kn + (i * 32 + lane) * 8;
                        if (k < kend) {
                            nw0[i] = ldg_stream(Wb + size_t(r0) * p.K + k);
                            nw1[i] = ldg_stream(Wb + size_t(r1) * p.K + k);
                        }
                    }
                } else if (next_active) {
                    unit_rows(EPI, p, un, n0, n1);
#pragma unroll
                    for (int i = 0; i < U; i++) {
                        const uint32_t k = kbeg + (i * 32 + lane) * 8;
                        if (k < kend) {
                            nw0[i] = ldg_stream(Wb + size_t(n0) * p.K + k);
                            nw1[i] = ldg_stream(Wb + size_t(n1) * p.K + k);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < U; i++) {
                    const uint32_t k = kc + (i * 32 + lane) * 8;
                    if (k < kend) {
#pragma unroll
                        for (int m = 0; m < MB; m++) {
                            const uint4 xv = *reinterpret_cast<const uint4*>(sx + size_t(m) * p.K + k);
                            acc0[m] = dot8(w0[i], xv, acc0[m]);
                            acc1[m] = dot8(w1[i], xv, acc1[m]);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < U; i++) w0[i] = nw0[i], w1[i] = nw1[i];
            }
        }
        // -- reduce over lanes, then over the KSPLIT k-slices of the unit
#pragma unroll
        for (int m = 0; m < MB; m++) {
            acc0[m] = warp_sum(acc0[m]);
            acc1[m] = warp_sum(acc1[m]);
        }
        if (KSPLIT > 1) {
            if (lane == 0) {
#pragma unroll
                for (int m = 0; m < MB; m++) {
                    sred[(warp * 2 + 0) * MB + m] = acc0[m];
                    sred[(warp * 2 + 1) * MB + m] = acc1[m];
                }
            }
            __syncthreads();
            if (ks == 0) {
#pragma unroll
                for (int m = 0; m < MB; m++) {
                    float a = 0.0f, b = 0.0f;
                    for (uint32_t s = 0; s < KSPLIT; s++) {
                        a += sred[((warp + s) * 2 + 0) * MB + m];
                        b += sred[((warp + s) * 2 + 1) * MB + m];
                    }
                    acc0[m] = a, acc1[m] = b;
                }
            }
            __syncthreads();
        }
        // -- epilogue: lane m finalises activation row m
        if (active && fin) {
            float a = 0.0f, b = 0.0f;
#pragma unroll
            for (int m = 0; m < MB; m++)
                if (lane == uint32_t(m)) a = acc0[m], b = acc1[m];
            const uint32_t m = lane;
            float y0 = rbf(a), y1 = rbf(b); // the bmm output buffer is T (kernel/bmm.metal:76)
            if (p.lora_b) {
                // lora_linear: y = r(y + r(r(B . ax) * scale))  (quantization/lora.h:115-122)
                float l0 = 0.0f, l1 = 0.0f;
                for (uint32_t j = 0; j < p.lora_rank; j++) {
                    const float axj = bf16_bits_to_f32(p.lora_ax[m * p.lora_rank + j]);
                    l0 = fmaf(axj, bf16_bits_to_f32(p.lora_b[size_t(r0) * p.lora_rank + j]), l0);
                    l1 = fmaf(axj, bf16_bits_to_f32(p.lora_b[size_t(r1) * p.lora_rank + j]), l1);
                }
                y0 = rbf(__fadd_rn(y0, rbf(__fmul_rn(rbf(l0), p.lora_scale))));
                y1 = rbf(__fadd_rn(y1, rbf(__fmul_rn(rbf(l1), p.lora_scale))));
            }
            if (EPI == EPI_PARTIAL_TP) {
                // unrounded fp32 partial sums go to this rank's own slot; tp_publish copies the finished vector to the peers
                float* dst = tp_slot(p.tp, p.tp.peer_buf[p.tp.rank], tp_parity, p.tp.rank) + size_t(m) * p.tp.dim + p.tp_col0;
                dst[r0] = a;
                dst[r1] = b;
            } else if (EPI == EPI_NONE) {
                p.y[size_t(m) * p.ldy + r0] = f32_to_bf16_bits(y0);
                p.y[size_t(m) * p.ldy + r1] = f32_to_bf16_bits(y1);
                if (MEGA) {
                    if (y0 > best_v || (y0 == best_v && int32_t(r0) < best_i)) best_v = y0, best_i = int32_t(r0);
                    if (y1 > best_v || (y1 == best_v && int32_t(r1) < best_i)) best_v = y1, best_i = int32_t(r1);
                }
            } else if (EPI == EPI_RESIDUAL) {
                // h = r(x + a)  (nn/transformer.h:133,139; kernel/arithmetic.metal:21-43)
                p.y[size_t(m) * p.ldy + r0] = f32_to_bf16_bits(__fadd_rn(ep0, y0));
                p.y[size_t(m) * p.ldy + r1] = f32_to_bf16_bits(__fadd_rn(ep1, y1));
            } else if (EPI == EPI_SWIGLU) {
                // z = r(silu_T(g) * u), rows (2i, 2i+1) = (w1 row i, w3 row i)  (nn/transformer.h:57-59)
                p.y[size_t(m) * p.ldy + (r0 >> 1)] = f32_to_bf16_bits(__fmul_rn(silu_bf16(y0), y1));
            } else { // EPI_QKV
                const uint32_t hd = p.head_dim, half = hd >> 1;
                const uint32_t head = r0 / hd, j = r0 - head * hd;
                const int32_t pos = my_pos, seq = my_seq;
                if (head < p.n_heads + p.n_kv_heads) {
                    // rope in fp32 with fp32 tables, one rounding (kernel/rope.metal:47-58)
                    const float c = ep0, s = ep1;
                    const float o0 = rbf(__fsub_rn(__fmul_rn(c, y0), __fmul_rn(s, y1)));
                    const float o1 = rbf(__fadd_rn(__fmul_rn(s, y0), __fmul_rn(c, y1)));
                    y0 = o0, y1 = o1;
                }
                if (head < p.n_heads) {
                    uint16_t* dst = p.q + size_t(m) * p.n_heads * hd + size_t(head) * hd + j;
                    dst[0] = f32_to_bf16_bits(y0);
                    dst[half] = f32_to_bf16_bits(y1);
                } else {
                    // sink_cache::update: bit copy into position pos (nn/cache.h:207-214)
                    const bool is_k = head < p.n_heads + p.n_kv_heads;
                    const uint32_t kvh = is_k ? head - p.n_heads : head - p.n_heads - p.n_kv_heads;
                    uint16_t* base = (is_k ? p.kcache : p.vcache) + adj.kv_off;
                    // a position beyond the cache writes the last row: the sink roll has shifted the others (nn/cache.h:183-204)
                    const uint32_t slot = min(uint32_t(pos), p.max_seq - 1);
                    uint16_t* dst = base + ((size_t(seq) * p.n_kv_heads + kvh) * p.max_seq + size_t(slot)) * hd + j;
                    dst[0] = f32_to_bf16_bits(y0);
                    dst[half] = f32_to_bf16_bits(y1);
                }
            }
        }
        r0 = n0, r1 = n1;
    }
    if (MEGA && EPI == EPI_NONE && p.am_val) {
        // per-CTA argmax partial of every activation row
        __syncthreads();
        if (fin) {
            sred[(warp * 2 + 0) * MB + lane] = best_v;
            reinterpret_cast<int32_t*>(sred)[(warp * 2 + 1) * MB + lane] = best_i;
        }
        __syncthreads();
        if (threadIdx.x < p.rows && threadIdx.x < uint32_t(MB)) {
            float bv = -INFINITY;
            int32_t bi = 0x7fffffff;
            for (uint32_t w = 0; w < uint32_t(kGemvWarps); w += KSPLIT) {
                const float v = sred[(w * 2 + 0) * MB + threadIdx.x];
                const int32_t i = reinterpret_cast<const int32_t*>(sred)[(w * 2 + 1) * MB + threadIdx.x];
                if (v > bv || (v == bv && i < bi)) bv = v, bi = i;
            }
            p.am_val[threadIdx.x * gridDim.x + blockIdx.x] = bv;
            p.am_idx[threadIdx.x * gridDim.x + blockIdx.x] = bi;
        }
    }
    if (EPI == EPI_PARTIAL_TP && !p.tp_hold) tp_publish(p.tp, p.rows, tp_parity, reinterpret_cast<unsigned*>(sscr));
    if (MEGA) {
        stamp(sy.timing, 2);
        grid_arrive(sy.bar);
    }
}

template <int MB, int PRO, int EPI, int KS>
__global__ void __launch_bounds__(kGemvThreads, 2) gemv_bf16_kernel(const gemv_params p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    gemv_body<MB, PRO, EPI, false, KS>(p, smem, mega_sync{}, phase_adj{0, 0, false});
}

// ---- K5: decode attention ------------------------------------------------------------------------------------
// One CTA per (head, activation row).  Reference chain per head (nn/attention.h:195-203):
//   s = r(q . K[t]); s = r(s * scale); p = r(exp(s) * (1 / sum exp(s))) (no max shift, quirk Q1);
//   o = r(sum_t p[t] * V[t]).  repeat_kv is replaced by kv = head / n_reps (no copies).
struct attn_params {
    const uint16_t* q;      // [rows, H*hd] (roped)
    const uint16_t* kcache; // this layer: [n_seqs, KV, S, hd]
    const uint16_t* vcache;
    uint16_t* out;          // [rows, H*hd]
    const int32_t* row_seq;
    const int32_t* row_pos;
    uint32_t n_heads, n_kv_heads, max_seq;
    float scale;            // r(1/sqrt(hd)) stored as T (nn/attention.h:88,115; quirk Q4)
    uint32_t key_begin;     // first visible cache position (0; the start of the chunk under the reference's chunk mask, quirk Q9)
};

constexpr int kAttnCluster = 4; // CTAs per (head, row): split of the cached positions, joined through DSMEM

// cluster helpers (raw PTX: barrier.cluster + mapa/ld.shared::cluster)
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(const float* local_smem_ptr, uint32_t rank)
{
    const uint32_t a = uint32_t(__cvta_generic_to_shared(local_smem_ptr));
    uint32_t ra;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
    return v;
}

// grid (n_heads * kAttnCluster, rows), cluster (kAttnCluster,1,1), 256 threads.
// Thread (slot = tid / LPP, dl = tid % LPP) owns positions slot, slot + SLOTS, ... of its CTA's chunk and the
// 8 head dims [8*dl, 8*dl+8): the same mapping serves q.K and P.V, so K and V rows are requested together.
template <int HD, int CL>
__device__ __forceinline__ void attn_body(const attn_params& p, unsigned char* smem, uint32_t head, uint32_t row, uint32_t crank, size_t kv_off)
{
    constexpr int LPP = HD / 8;       // lanes per position row (16 bytes each)
    constexpr int SLOTS = 256 / LPP;  // positions handled per sweep
    constexpr int IT = 4;             // sweeps per block of loads kept in flight
    float* sq = reinterpret_cast<float*>(smem);  // [HD]
    float* scr = sq + HD;                        // [8] block-sum scratch
    float* xch = scr + 8;                        // [4 + HD] exchanged through DSMEM: exp-sum, then partial o
    float* spart = xch + 4 + HD;                 // [SLOTS][HD]
    float* sp = spart + SLOTS * HD;              // [chunk] scores / probabilities of this CTA

    const int32_t seq = p.row_seq[row];
    const uint32_t P = min(uint32_t(p.row_pos[row]), p.max_seq - 1) + 1; // a full (rolled) cache shows all max_seq rows
    const uint32_t kvh = head / (p.n_heads / p.n_kv_heads);
    const size_t coff = kv_off + (size_t(seq) * p.n_kv_heads + kvh) * p.max_seq * HD;
    const uint16_t* Kc = p.kcache + coff;
    const uint16_t* Vc = p.vcache + coff;
    const uint32_t kb = min(p.key_begin, P - 1);                               // keys [kb, P) are visible
    const uint32_t chunk = ((P - kb + CL * SLOTS - 1) / (CL * SLOTS)) * SLOTS; // multiple of SLOTS
    const uint32_t t0 = min(P, kb + crank * chunk);
    const uint32_t t1 = min(P, t0 + chunk);
    const uint32_t slot = threadIdx.x / LPP, dl = threadIdx.x % LPP;
    const bool vpre = chunk <= IT * SLOTS; // the whole chunk fits one block of loads: fetch V together with K

    // first block of K (and V) rows: all requests are issued before anything is consumed
    uint4 kreg[IT], vreg[IT];
#pragma unroll
    for (int i = 0; i < IT; i++) {
        const uint32_t t = t0 + i * SLOTS + slot;
        if (t < t1) {
            kreg[i] = *reinterpret_cast<const uint4*>(Kc + size_t(t) * HD + dl * 8);
            if (vpre) vreg[i] = *reinterpret_cast<const uint4*>(Vc + size_t(t) * HD + dl * 8);
        }
    }
    if (threadIdx.x < HD) sq[threadIdx.x] = bf16_bits_to_f32(p.q[(size_t(row) * p.n_heads + head) * HD + threadIdx.x]);
    __syncthreads();
    float qv[8];
#pragma unroll
    for (int i = 0; i < 8; i++) qv[i] = sq[dl * 8 + i];

    // s = r(r(q . K[t]) * scale)
    for (uint32_t tb = t0; tb < t1; tb += IT * SLOTS) {
        if (tb != t0) {
#pragma unroll
            for (int i = 0; i < IT; i++) {
                const uint32_t t = tb + i * SLOTS + slot;
                if (t < t1) kreg[i] = *reinterpret_cast<const uint4*>(Kc + size_t(t) * HD + dl * 8);
            }
        }
#pragma unroll
        for (int i = 0; i < IT; i++) {
            const uint32_t t = tb + i * SLOTS + slot;
            float d = 0.0f;
            if (t < t1) {
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
            if (t < t1 && dl == 0) sp[t - t0] = rbf(__fmul_rn(rbf(d), p.scale));
        }
    }
    __syncthreads();
    // softmax without max subtraction (kernel/softmax.metal:40-80): the exp-sum is joined across the cluster
    const uint32_t n_local = t1 > t0 ? t1 - t0 : 0;
    float part = 0.0f;
    for (uint32_t t = threadIdx.x; t < n_local; t += 256) part += expf(sp[t]);
    const float local_sum = block_sum_256(part, scr);
    float total = local_sum;
    if (CL > 1) {
        if (threadIdx.x == 0) xch[0] = local_sum;
        cluster_sync_all();
        total = 0.0f;
#pragma unroll
        for (uint32_t r = 0; r < uint32_t(CL); r++) total += ld_dsmem_f32(&xch[0], r);
    }
    const float inv = 1.0f / total;
    for (uint32_t t = threadIdx.x; t < n_local; t += 256) sp[t] = rbf(__fmul_rn(expf(sp[t]), inv));
    __syncthreads();
    // o = sum_t p[t] * V[t]
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = 0.0f;
    for (uint32_t tb = t0; tb < t1; tb += IT * SLOTS) {
        if (!vpre) {
#pragma unroll
            for (int i = 0; i < IT; i++) {
                const uint32_t t = tb + i * SLOTS + slot;
                if (t < t1) vreg[i] = *reinterpret_cast<const uint4*>(Vc + size_t(t) * HD + dl * 8);
            }
        }
#pragma unroll
        for (int i = 0; i < IT; i++) {
            const uint32_t t = tb + i * SLOTS + slot;
            if (t < t1) {
                const uint4 vv = vreg[i];
                const float pt = sp[t - t0];
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
    __syncthreads();
    float o = 0.0f;
    if (threadIdx.x < HD) {
        for (int gg = 0; gg < SLOTS; gg++) o += spart[gg * HD + threadIdx.x];
        xch[4 + threadIdx.x] = o;
    }
    if (CL > 1) {
        cluster_sync_all();
        if (crank == 0 && threadIdx.x < HD) {
            o = 0.0f;
#pragma unroll
            for (uint32_t r = 0; r < uint32_t(CL); r++) o += ld_dsmem_f32(&xch[4 + threadIdx.x], r);
        }
    }
    if (crank == 0 && threadIdx.x < HD) p.out[(size_t(row) * p.n_heads + head) * HD + threadIdx.x] = f32_to_bf16_bits(o);
    if (CL > 1) cluster_sync_all(); // peers must not exit while rank 0 still reads their shared memory
    else __syncthreads();           // shared memory is reused by the next work item
}

template <int HD> __global__ void __launch_bounds__(256) attn_decode_kernel(const attn_params p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    pdl_launch_dependents();
    pdl_wait();
    attn_body<HD, kAttnCluster>(p, smem, blockIdx.x / kAttnCluster, blockIdx.y, cluster_ctarank(), 0);
}

// ---- sink-cache roll (nn/cache.h:183-204 + kernel/roll.metal:22-45) --------------------------------------------------------------
// A decode step at a position beyond the cache keeps the first `pre_len` rows (the sink tokens), moves rows (pre_len, S) one row to
// the left and writes the new row at S - 1.  The reference allocates a new cache and rolls into it; here the shift is in place:
// one CTA per (layer, kv head, K|V) stream and decode row, tiles of 2048 16-byte chunks, every tile loaded completely before it is
// stored one row lower (a tile's loads never touch what an earlier tile stored).  Rows whose position is inside the cache return.
MC_KERNEL void __launch_bounds__(256) kv_roll_kernel(uint16_t* kcache, uint16_t* vcache, size_t kv_layer_stride, const int32_t* row_seq, const int32_t* row_pos,
                                                      uint32_t n_kv_heads, uint32_t head_dim, uint32_t max_seq, uint32_t pre_len)
{
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t row = blockIdx.y;
    if (uint32_t(row_pos[row]) < max_seq) return;
    const uint32_t stream = blockIdx.x, which = stream & 1u, kvh = (stream >> 1) % n_kv_heads, layer = (stream >> 1) / n_kv_heads;
    uint16_t* base = (which ? vcache : kcache) + size_t(layer) * kv_layer_stride + (size_t(row_seq[row]) * n_kv_heads + kvh) * max_seq * head_dim;
    const uint32_t row_chunks = head_dim / 8;                          // 16-byte chunks per cache row
    const uint32_t n_chunks = (max_seq - pre_len - 1) * row_chunks;    // rows (pre_len, S) move to [pre_len, S - 1)
    const uint4* src = reinterpret_cast<const uint4*>(base + size_t(pre_len + 1) * head_dim);
    uint4* dst = reinterpret_cast<uint4*>(base + size_t(pre_len) * head_dim);
    constexpr uint32_t PER = 8;
    for (uint32_t c0 = 0; c0 < n_chunks; c0 += 256 * PER) {
        uint4 v[PER];
#pragma unroll
        for (uint32_t i = 0; i < PER; i++) {
            const uint32_t c = c0 + i * 256 + threadIdx.x;
            if (c < n_chunks) v[i] = src[c];
        }
        __syncthreads();
#pragma unroll
        for (uint32_t i = 0; i < PER; i++) {
            const uint32_t c = c0 + i * 256 + threadIdx.x;
            if (c < n_chunks) dst[c] = v[i];
        }
    }
}

// ---- K6: embedding gather (kernel/embedding.metal:38-66; lora_embedding quantization/lora.h:160-170) -----------
MC_KERNEL void embed_kernel(uint16_t* x, uint32_t ldx, const void* table, const float* row_scales, int fmt, uint32_t D,
                             uint32_t vocab, const int32_t* ids)
{
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t row = blockIdx.x;
    int32_t id = ids[row];
    if (id < 0 || uint32_t(id) >= vocab) id = 0; // validated on the host for host-provided ids
    if (fmt == WF_BF16) {
        const uint4* src = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(table) + size_t(id) * D);
        uint4* dst = reinterpret_cast<uint4*>(x + size_t(row) * ldx);
        for (uint32_t k = threadIdx.x; k < D / 8; k += blockDim.x) dst[k] = src[k];
    } else {
        const int8_t* src = static_cast<const int8_t*>(table) + size_t(id) * D;
        const float s = rbf(row_scales[id]);
        for (uint32_t k = threadIdx.x; k < D; k += blockDim.x)
            x[size_t(row) * ldx + k] = f32_to_bf16_bits(__fmul_rn(float(src[k]), s)); // r(r(q) * r(s)), kernel/mul.metal:76-77
    }
}

// ---- K7g: greedy argmax, lowest index on ties ---------------------------------------------------------------------
constexpr int kArgmaxBlocks = 64;
MC_KERNEL void __launch_bounds__(256) argmax_partial_kernel(const uint16_t* logits, uint32_t ld, uint32_t n, float* pval, int32_t* pidx)
{
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float sv[8];
    __shared__ int32_t si[8];
    const uint32_t row = blockIdx.y;
    const uint16_t* l = logits + size_t(row) * ld;
    const uint32_t per = (n + gridDim.x - 1) / gridDim.x;
    const uint32_t beg = blockIdx.x * per, end = min(n, beg + per);
    float bv = -INFINITY;
    int32_t bi = 0x7fffffff;
    for (uint32_t i = beg + threadIdx.x; i < end; i += 256) {
        const float f = bf16_bits_to_f32(l[i]);
        if (f > bv || (f == bv && int32_t(i) < bi)) bv = f, bi = int32_t(i);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
        const int32_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ov > bv || (ov == bv && oi < bi)) bv = ov, bi = oi;
    }
    if ((threadIdx.x & 31) == 0) sv[threadIdx.x >> 5] = bv, si[threadIdx.x >> 5] = bi;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++)
            if (sv[w] > bv || (sv[w] == bv && si[w] < bi)) bv = sv[w], bi = si[w];
        pval[row * gridDim.x + blockIdx.x] = bv;
        pidx[row * gridDim.x + blockIdx.x] = bi;
    }
}
// final stage + feedback: next id -> ids[row] (input of the next step), pos[row] += 1, log
MC_KERNEL void argmax_final_kernel(const float* pval, const int32_t* pidx, int nblk, int32_t* ids, int32_t* pos, int32_t* out_log,
                                    int32_t* step_counter, uint32_t rows, int advance)
{
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t row = threadIdx.x;
    if (row >= rows) return;
    float bv = -INFINITY;
    int32_t bi = 0x7fffffff;
    for (int b = 0; b < nblk; b++) {
        const float v = pval[row * nblk + b];
        const int32_t i = pidx[row * nblk + b];
        if (v > bv || (v == bv && i < bi)) bv = v, bi = i;
    }
    if (bi == 0x7fffffff) bi = 0; // all-NaN / -inf row: argmax keeps index 0
    const int32_t step = *step_counter;
    out_log[size_t(step) * rows + row] = bi;
    if (advance) {
        ids[row] = bi;
        pos[row] += 1;
    }
    __syncthreads();
    if (row == 0) *step_counter = step + 1;
}

// Vocabulary-sharded head: every rank pushes its local (value, global index) winner to all peers, waits for the
// others and picks the global winner in rank order (lowest index on ties) — all ranks arrive at the same token.
struct am_exchange {
    uint32_t world, rank, rows_max;
    float* peer_val[kTpMaxWorld];      // [parity][src_rank][rows_max]
    int32_t* peer_idx[kTpMaxWorld];
    uint32_t* peer_flag[kTpMaxWorld];  // flags[src_rank]
    unsigned* epoch;                   // local
    int* err;
};
MC_KERNEL void argmax_final_tp_kernel(const float* pval, const int32_t* pidx, int nblk, int32_t index_base, am_exchange x, int32_t* ids, int32_t* pos,
                                       int32_t* out_log, int32_t* step_counter, uint32_t rows, int advance)
{
    pdl_launch_dependents();
    pdl_wait();
    __shared__ unsigned ok;
    const uint32_t row = threadIdx.x;
    const unsigned e = *x.epoch + 1;
    const uint32_t parity = e & 1u;
    if (row < rows) {
        float bv = -INFINITY;
        int32_t bi = 0x7fffffff;
        for (int b = 0; b < nblk; b++) {
            const float v = pval[row * nblk + b];
            const int32_t i = pidx[row * nblk + b];
            if (v > bv || (v == bv && i < bi)) bv = v, bi = i;
        }
        if (bi != 0x7fffffff) bi += index_base;
        for (uint32_t k = 0; k < x.world; k++) {
            const size_t o = (size_t(parity) * x.world + x.rank) * x.rows_max + row;
            x.peer_val[k][o] = bv;
            x.peer_idx[k][o] = bi;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        for (uint32_t k = 0; k < x.world; k++) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(x.peer_flag[k] + x.rank), "r"(e) : "memory");
        unsigned good = 1;
        for (uint32_t k = 0; k < x.world; k++) {
            unsigned v = 0;
            unsigned long long spins = 0;
            for (;;) {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(x.peer_flag[x.rank] + k) : "memory");
                if (v >= e) break;
                if (++spins > (1ull << 23)) {
                    atomicExch(x.err, 3);
                    good = 0;
                    break;
                }
            }
        }
        ok = good;
        *x.epoch = e;
    }
    __syncthreads();
    if (row < rows) {
        float bv = -INFINITY;
        int32_t bi = 0x7fffffff;
        if (ok) {
            for (uint32_t k = 0; k < x.world; k++) {
                const size_t o = (size_t(parity) * x.world + k) * x.rows_max + row;
                const float v = x.peer_val[x.rank][o];
                const int32_t i = x.peer_idx[x.rank][o];
                if (v > bv || (v == bv && i < bi)) bv = v, bi = i;
            }
        }
        if (bi == 0x7fffffff) bi = 0;
        const int32_t step = *step_counter;
        out_log[size_t(step) * rows + row] = bi;
        if (advance) {
            ids[row] = bi;
            pos[row] += 1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *step_counter += 1;
}

// ---- synthetic weights (DESIGN.md "Synthetic data"; same hash as the test oracle) ----------------------------------
// dst[r * dst_ld + c] = value(src_row0 + r, src_col0 + c) of the [*, src_K] tensor `tid`
MC_KERNEL void gen_bf16_kernel(uint16_t* dst, size_t dst_ld, uint32_t rows, uint32_t cols, uint32_t src_row0, uint32_t src_col0,
                                uint32_t src_K, uint64_t seed, uint64_t tid, float scale, float bias)
{
    const uint64_t n = uint64_t(rows) * cols;
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t r = uint32_t(i / cols), c = uint32_t(i % cols);
        const uint64_t flat = uint64_t(src_row0 + r) * src_K + (src_col0 + c);
        const uint32_t u24 = uint32_t(hash3(seed, tid, flat) >> 40);
        const float u = __fsub_rn(__fmul_rn(float(u24), 1.0f / 8388608.0f), 1.0f);
        const float v = bias == 0.0f ? __fmul_rn(u, scale) : __fadd_rn(bias, __fmul_rn(scale, u));
        dst[size_t(r) * dst_ld + c] = f32_to_bf16_bits(v);
    }
}
MC_KERNEL void gen_f32_scales_kernel(float* dst, size_t dst_ld, uint32_t rows, uint32_t cols, uint32_t src_row0, uint32_t src_col0,
                                      uint32_t src_K, uint64_t seed, uint64_t tid, float c0)
{
    const uint64_t n = uint64_t(rows) * cols;
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t r = uint32_t(i / cols), c = uint32_t(i % cols);
        const uint64_t flat = uint64_t(src_row0 + r) * src_K + (src_col0 + c);
        const uint32_t u24 = uint32_t(hash3(seed, tid, flat) >> 40);
        const float u = __fsub_rn(__fmul_rn(float(u24), 1.0f / 8388608.0f), 1.0f);
        dst[size_t(r) * dst_ld + c] = __fmul_rn(__fadd_rn(1.0f, __fmul_rn(0.5f, u)), c0);
    }
}
MC_KERNEL void gen_i8_kernel(int8_t* dst, size_t dst_ld, uint32_t rows, uint32_t cols, uint32_t src_row0, uint32_t src_col0,
                              uint32_t src_K, uint64_t seed, uint64_t tid, int32_t lo, uint32_t range)
{
    const uint64_t n = uint64_t(rows) * cols;
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t r = uint32_t(i / cols), c = uint32_t(i % cols);
        const uint64_t flat = uint64_t(src_row0 + r) * src_K + (src_col0 + c);
        const uint32_t u32 = uint32_t(hash3(seed, tid, flat) >> 32);
        dst[size_t(r) * dst_ld + c] = int8_t(lo + int32_t((uint64_t(u32) * range) >> 32));
    }
}

} // namespace mc
