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
constexpr int kStEpiThreads = 64;                     // two epilogue warps: join the k-slices, finish and publish the outputs
#ifdef ST_USE_SETMAXNREG
constexpr int kStThreads = kStConsumers + kStEpiThreads + 64;
#else
constexpr int kStThreads = kStConsumers + kStEpiThreads + 32;
#endif
constexpr int kStTileRows = 16;                       // weight rows per tile (the m of the mma)
constexpr int kStMaxKC = 1024;                        // k elements per tile row
constexpr int kStPad = 64;                            // bytes of padding per staged row
constexpr int kStStageBytes = kStTileRows * (kStMaxKC * 2 + kStPad);
constexpr int kStMaxStages = 12;
constexpr int kStStageBytesPacked = 18432;            // quantised models: an int4 tile (16 KiB + 2 KiB of scales) per stage, more stages
constexpr int kStSplits = 4;                          // CTAs per (row, head) in the attention phase
constexpr int kStMaxRep = 8;                          // copies of every exchanged vector (spreads the pollers over L2 lines)
constexpr int kStMaxRows = 8;                         // activation rows (the n of the mma)
constexpr int kStHdrBytes = 896;                      // mbarriers + flags + reduce scratch
constexpr int kStRedBufs = 3;                         // cross-warp partial buffers in flight between the mma and the epilogue warps
constexpr int kStRedBytes = kStRedBufs * kStWarps * 16 * 8 * 4;
constexpr int kStTpBlocks = 8;                        // tensor parallel: 16-row blocks of a row-parallel phase one CTA may own (148 CTAs: dim <= 18 944)
constexpr int kStTpKeepBytes = kStTpBlocks * 16 * kStMaxRows * 4; // this rank's fp32 partial sums of those blocks, kept between the two epilogue passes
constexpr int kStTpMaxWorld = 8;
constexpr int kStTpAxCols = 64; // adaptor rank <= 64

// Activations travel between CTAs as TAGGED WORDS: one 8-byte word = two bf16 values (low half) + a 32-bit epoch tag
// (high half), written with one 8-byte store and read with 8/16-byte volatile loads.  The tag is unique per
// (launch, step, phase), so a consumer simply polls the data it needs until the tags match: no release fence, no
// arrival counter, no separate flag round trip (the scheme of low-latency collectives).  Every GEMV phase needs the whole
// input vector, i.e. data from every CTA, so the polling doubles as the phase barrier.
struct st_gemv {
    const void* W;           // WF_BF16: [N, K] bf16 row-major; WF_W4 / WF_W8ROW: fragment-packed (mc_quant_kernels.cuh).
                             // layer l adds l * layer_stride bytes when `layered` (also to norm_w, scales, lora_a, lora_b)
    const uint16_t* norm_w;  // PRO_RMSNORM: [K]
    const void* scales;      // WF_W4: packed bf16 group scales; WF_W8ROW: fp32 [N]
    const uint16_t* lora_a;  // quantised layers: stacked adaptor A rows [n_a, K] bf16, or null
    const uint16_t* lora_b;  // [N, rank] bf16
    const uint64_t* in_ll;   // tagged input rows [rows][K/2] (null: embedding gather)
    uint64_t* out_ll;        // tagged output rows: [rows][N/2], EPI_SWIGLU [rows][N/4]; null for the head
    const uint64_t* res_ll;  // EPI_RESIDUAL: [rows][N/2], written earlier by this same CTA
    uint64_t* ax_ll;         // tagged r(A . x) of this phase: [rows][n_a/2]
    uint16_t* y;             // head only: plain bf16 logits [rows][N]
    uint32_t N, K, KC, gran; // KC: k per tile; gran = row granularity of the CTA split (bf16: 2, 4 for gate/up pairs; packed: 16)
    uint32_t n_a;            // rows of lora_a (0: no adaptor)
    uint32_t ax_slices, slice_rows0, slice_rows1; // which rank-wide slice of ax a weight row uses (see lora_slice)
    int32_t fmt, pro, epi, layered;
    int32_t qkv_map;         // packed layouts: block rows are the rope pairs (j, j + hd/2) of EPI_QKV packing
};

struct st_params {
    st_gemv g[5];            // qkv, wo, w1-w3, w2, head
    size_t layer_stride;     // bytes between consecutive layers in the weight arena
    size_t kv_layer_stride;  // elements between consecutive layers in the KV cache
    uint32_t n_layers, rows, steps, n_stages;
    uint32_t stage_bytes;    // bytes of one ring stage
    uint32_t kc_a[4];        // k per tile of the LoRA-A rows of the four linears of a block
    uint32_t lora_rank;
    float lora_scale;        // r(scale) as fp32
    int32_t tok_fmt;         // WF_BF16, or WF_W8ROW: int8 embedding rows with one fp32 scale per row
    const float* tok_scales;
    uint32_t act_pitch;      // bytes between staged activation rows
    uint32_t act_bytes;
    uint32_t sax_off;        // quantised models: offset of the [8][n_a] r(A . x) scratch inside the activation region
    uint32_t tag_base;       // tags of this launch are tag_base + phase + 1
    uint32_t n_rep;          // every exchanged vector exists in n_rep copies (each polled by 1/n_rep of the CTAs), rep_stride words apart
    uint32_t rep_stride;
    uint32_t poll_ns;
    float eps;
    // attention
    const uint64_t* qkv_ll;  // [rows][(H + 2 KV) * hd / 2]: r(x.W^T), not yet rotated
    uint64_t* sc_ll;         // [rows * H][sc_words]: scores r(r(q.k) * scale) of all positions
    uint64_t* attn_ll;       // [rows][H * hd / 2]
    uint32_t sc_words;
    uint16_t* kcache;        // [layer][n_seqs, KV, S, hd]
    uint16_t* vcache;
    const float* fcos;       // [2S, hd/2]
    const float* fsin;
    const int32_t* row_seq;  // [rows]
    int32_t* pos;            // [rows]: positions of the first step; advanced by `steps` at the end when `advance`
    int32_t* ids;            // [rows]: ids of the first step; the last sampled ids at the end when `advance`
    uint32_t n_heads, n_kv_heads, head_dim, max_seq, vocab;
    float scale;             // r(1/sqrt(hd)) (nn/attention.h:88,115)
    // embedding
    const void* embed_table;
    // sampler tail
    uint64_t* am_ll;         // [rows][G][2]: per-CTA argmax partial {value bits, index}
    uint64_t* ids_ll;        // [rows]: sampled id of the step, tagged with the head phase
    int32_t* out_log;
    int32_t* step_counter;
    int32_t advance;
    int* err;
    unsigned long long* timing; // diagnostics (nullable): 4 globaltimer stamps per (CTA, phase)
    unsigned long long* dbg;    // diagnostics (nullable): per-block stamps of CTA 0 in the vocabulary projection: [block][8]
    // Tensor parallelism (tp_world > 1; the reference is single-device, nn/llama.h:86): this rank holds Hl = H / T heads, Fl = ffn / T
    // channels and Vl = vocab / T rows of the head.  The row-parallel linears (wo, w2) end in an all-reduce that is FUSED into their
    // epilogue: every rank's CTA c owns the same output rows, sends its fp32 partial sums as tagged words straight into the peers'
    // exchange regions over NVLink (peer stores), polls the peers' words in its own region, sums in rank order and goes on with
    // h = r(x + r(sum)) -- no collective call, no kernel boundary, no separate flag.  The greedy sampler joins (value, index) the same way.
    uint32_t tp_world, tp_rank;
    uint64_t* tp_part[kStTpMaxWorld]; // exchange region of every rank as mapped on THIS GPU: [kind 0 wo | 1 w2][src rank][row][dim] tagged words (fp32 payload)
    uint64_t* tp_am[kStTpMaxWorld];   // [src rank][row][2]: value bits, global index
    // arrival counters (opt-in experiment; the tags stay the proof): every CTA adds 1 to arrive[gphase & 7] when its part of phase gphase is
    // published; the staging of the next GEMV phase waits for the counter before its FIRST load, so the words are asked for once
    // instead of being polled by 256 threads x 148 CTAs while the producers' stores are still on their way (tools/hop_floor.cu:
    // 1.63 vs 2.30 us per exchange of a dim-2048 vector -- but the real step got 7 % SLOWER with it, see mc_engine.cu launch_stream).
    // arrive_base[k] = value of counter k before this launch.
    unsigned* arrive;
    uint32_t arrive_base[8];
    uint32_t arrive_on;
    uint64_t* tp_ax[kStTpMaxWorld];   // quantised models: [kind][src rank][row][kStTpAxCols] tagged fp32 partials of the row-parallel adaptors' A . x
    uint32_t tp_dim;                  // = dim (row pitch of the partial-sum words)
    uint32_t tp_index_base;           // first vocabulary row of this rank's head shard
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
    uint32_t rdy0, fre0;      // shared addresses of ready[kStRedBufs] (partials written), free[kStRedBufs] (partials consumed)
    uint32_t sbar;            // shared address of the end-of-step mbarrier
    float* escr;              // [128] scratch of the epilogue warps
    float* tpk;               // [kStTpBlocks][16][kStMaxRows] fp32 partial sums of a row-parallel phase (tensor parallel)
    volatile int* dead;       // set when a wait timed out somewhere: every later wait falls through
    float* scr;               // [16] block-reduce scratch
    float* red;               // [2][8 warps][16 rows][8 cols]
    unsigned char* act;       // staged activation rows; attention scratch
    uint32_t act_addr;        // shared address of act
    uint32_t ring_addr;       // shared address of stage 0
    int* err;
    uint32_t poll_ns;         // back-off between failed polls of a tagged word (all CTAs poll the same words)
};

// bounded waits: a lost arrival must end in an error code, never in a hung GPU
#ifdef ST_DEBUG_WHERE
__device__ int g_st_where[148 * 4];
__device__ int g_st_frozen;
__device__ __forceinline__ void st_where(int kind)
{
    const int role = threadIdx.x == 0 ? 0 : (threadIdx.x == kStConsumers ? 1 : (threadIdx.x == kStConsumers + kStEpiThreads ? 2 : -1));
    if (role >= 0 && *reinterpret_cast<volatile int*>(&g_st_frozen) == 0) g_st_where[blockIdx.x * 4 + role] = kind;
}
#else
__device__ __forceinline__ void st_where(int) {}
#endif
__device__ __forceinline__ void st_mbar_wait(const st_ctx& c, uint32_t a, uint32_t parity)
{
    if (mbar_try_wait(a, parity)) return;
    st_where(100 + int((a - c.full0) >> 3));
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
#ifdef ST_DEBUG_WHERE
                g_st_frozen = 1;
#endif
                atomicExch(c.err, 4);
                *c.dead = 1;
                return;
            }
        }
    }
}
// ---- tagged words ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_ll_store(uint64_t* p, uint32_t payload, uint32_t tag)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"((uint64_t(tag) << 32) | payload) : "memory");
}
__device__ __forceinline__ void st_ll_store_rep(const st_params& P, uint64_t* p, uint32_t payload, uint32_t tag)
{
    for (uint32_t r = 0; r < P.n_rep; r++) st_ll_store(p + size_t(r) * P.rep_stride, payload, tag);
}
__device__ __forceinline__ void st_ll_load2(const uint64_t* p, uint64_t& a, uint64_t& b)
{
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ uint64_t st_ll_load1(const uint64_t* p)
{
    uint64_t a;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    return a;
}
// words exchanged with other GPUs: system scope (the peer's store arrives in this GPU's L2 over NVLink)
__device__ __forceinline__ void st_ll_store_sys(uint64_t* p, uint32_t payload, uint32_t tag)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"((uint64_t(tag) << 32) | payload) : "memory");
}
__device__ __forceinline__ uint64_t st_ll_load1_sys(const uint64_t* p)
{
    uint64_t a;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    return a;
}
// one failed poll: returns true when the caller should give up (another wait timed out, or this one did)
__device__ __forceinline__ bool st_poll_backoff(const st_ctx& c, unsigned& spins, int where = 0)
{
    if (spins == 0) st_where(where);
    if (*c.dead) return true;
    if (c.poll_ns) __nanosleep(c.poll_ns);
    if ((++spins & 1023u) == 0) {
        if (*reinterpret_cast<volatile int*>(c.err) != 0) {
            *c.dead = 1;
            return true;
        }
        if (spins > (1u << 21)) {
#ifdef ST_DEBUG_WHERE
            g_st_frozen = 1;
#endif
            atomicCAS(c.err, 0, 1);
            atomicOr(c.err + 1, 1 << where); // which kinds of wait failed
            atomicMin(c.err + 2, int(blockIdx.x) * 1024 + int(threadIdx.x) + (where << 20)); // lowest (kind, CTA, thread)
            *c.dead = 1;
            return true;
        }
    }
    return false;
}
__device__ __forceinline__ uint32_t st_poll1(const st_ctx& c, const uint64_t* p, uint32_t tag, int where = 0)
{
    unsigned spins = 0;
    for (;;) {
        const uint64_t a = st_ll_load1(p);
        if (uint32_t(a >> 32) == tag) return uint32_t(a);
        if (st_poll_backoff(c, spins, where)) {
#ifdef ST_DEBUG_WHERE
            if (spins > (1u << 21)) printf("cta %d tid %d kind %d expects tag %u saw %u payload %u\n", blockIdx.x, threadIdx.x, where, tag, uint32_t(a >> 32), uint32_t(a));
#endif
            return 0;
        }
    }
}
__device__ __forceinline__ uint32_t st_poll1_sys(const st_ctx& c, const uint64_t* p, uint32_t tag, int where)
{
    unsigned spins = 0;
    for (;;) {
        const uint64_t a = st_ll_load1_sys(p);
        if (uint32_t(a >> 32) == tag) return uint32_t(a);
        if (st_poll_backoff(c, spins, where)) return 0;
    }
}
// one thread per CTA: this CTA's part of phase gphase is published
__device__ __forceinline__ void st_arrive(const st_params& P, uint32_t gphase)
{
    if (P.arrive_on) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(P.arrive + (gphase & 7u)) : "memory");
}
// one thread per CTA: every CTA has published phase gphase (bounded: after ~20 us the tags take over, which are always checked anyway)
__device__ __forceinline__ void st_arrived_wait(const st_params& P, uint32_t gphase)
{
    if (!P.arrive_on) return;
    const uint32_t k = gphase & 7u, want = P.arrive_base[k] + gridDim.x * ((gphase >> 3) + 1u);
    for (uint32_t spins = 0; spins < 4096u; spins++) {
        uint32_t v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(P.arrive + k) : "memory");
        if (int32_t(v - want) >= 0) return;
    }
}
__device__ __forceinline__ void st_stamp(unsigned long long* t, unsigned idx)
{
    if (t && (threadIdx.x == 0 || threadIdx.x == kStConsumers)) {
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
__device__ __forceinline__ uint32_t pack2(float lo, float hi) { return uint32_t(f32_to_bf16_bits(lo)) | (uint32_t(f32_to_bf16_bits(hi)) << 16); }

// rows [rb, re) of an N-row matrix owned by this CTA (units of `gran` rows, balanced to +-1 unit)
__device__ __forceinline__ void st_my_rows(uint32_t N, uint32_t gran, uint32_t& rb, uint32_t& re)
{
    const uint64_t units = N / gran;
    rb = uint32_t(units * blockIdx.x / gridDim.x) * gran;
    re = uint32_t(units * (blockIdx.x + 1) / gridDim.x) * gran;
}

struct st_pipe {
    uint32_t stage, parity;
    __device__ __forceinline__ void advance(uint32_t n_stages)
    {
        if (++stage == n_stages) stage = 0, parity ^= 1u;
    }
};

// ---- the static schedule of weight tiles of this CTA --------------------------------------------------------------------
// A GEMV phase of this CTA is a list of BLOCKS of up to 16 output rows; a block is a list of TILES (k-chunks) that travel
// through the ring.  Quantised layers put one extra block in front: two rows of the stacked LoRA adaptor A (bf16), owned
// by CTA c < n_a / 2, whose output r(A . x) every CTA needs in its epilogue (quantization/lora.h:115-122).
//   bf16 block  = rows [r0, r0 + nr) of a row-major matrix, tile = nr row chunks of KC k, one bulk copy per row,
//                 staged with a 64-byte row pad;
//   packed block = one super-unit of 16 rows in mma fragment order, tile = KC k = one contiguous run of weights
//                 (+ one run of group scales for int4).
enum { ST_T_ROWS = 0, ST_T_W4 = 1, ST_T_W8 = 2 };
constexpr uint32_t kStW4ScaleOff = 16384; // int4 tiles: group scales sit behind the (at most 16 KiB of) weights
struct st_block {
    uint32_t r0, nr;   // bf16: first row / rows; packed: super-unit index / 16
    bool is_a;         // the LoRA-A block
};
// the blocks of phase g for this CTA: the A block (if owned), then the main blocks [b0, b1)
struct st_range {
    bool has_a;
    uint32_t b0, b1;   // bf16: rows, step 16; packed: super-units, step 1
    __device__ __forceinline__ st_range(const st_gemv& g, bool quant)
    {
        has_a = quant && g.n_a != 0 && 2 * blockIdx.x < g.n_a;
        if (!quant || g.fmt == WF_BF16) st_my_rows(g.N, g.gran, b0, b1);
        else st_my_rows(g.N >> 4, 1, b0, b1);
    }
};
__device__ __forceinline__ const char* st_layer_ptr(const st_params& P, const st_gemv& g, const void* p, uint32_t li)
{
    return static_cast<const char*>(p) + (g.layered ? size_t(li) * P.layer_stride : 0);
}
struct st_tile {
    int kind;
    const char* src;      // ST_T_ROWS: first row chunk; packed: the weights
    const char* src2;     // ST_T_W4: the group scales
    size_t row_stride;    // ST_T_ROWS: bytes between consecutive rows in global memory
    uint32_t nr, bytes, bytes2, pitch;
};
template <bool Q> struct st_tile_iter {
    uint32_t step = 0, li = 0, gi = 0, blk = 0, b1 = 0, kc = 0;
    int stage = 0; // 0: phase not started, 1: A block, 2: main blocks
    __device__ __forceinline__ void next_phase(const st_params& P)
    {
        stage = 0;
        if (gi == 4) gi = 0, li = 0, step++;
        else if (gi == 3) {
            li++;
            gi = li == P.n_layers ? 4 : 0;
        } else gi++;
    }
    __device__ __forceinline__ bool next(const st_params& P, st_tile& t)
    {
        for (;;) {
            if (step >= P.steps) return false;
            const st_gemv& g = P.g[gi];
            if (stage == 0) {
                const st_range r(g, Q);
                blk = r.b0, b1 = r.b1, kc = 0;
                stage = r.has_a ? 1 : 2;
            }
            if (Q && stage == 1) {
                const uint32_t kca = P.kc_a[gi];
                t.kind = ST_T_ROWS, t.nr = 2;
                t.src = st_layer_ptr(P, g, g.lora_a, li) + (size_t(2 * blockIdx.x) * g.K + kc) * 2;
                t.row_stride = size_t(g.K) * 2, t.bytes = kca * 2, t.pitch = kca * 2 + kStPad;
                kc += kca;
                if (kc >= g.K) kc = 0, stage = 2;
                return true;
            }
            if (blk < b1) {
                const char* W = st_layer_ptr(P, g, g.W, li);
                uint32_t kc_this = g.KC;
                if (!Q || g.fmt == WF_BF16) {
                    // tiles are KC k wide; when K is not a multiple of KC the last tile of a block is shorter (still a multiple of 256)
                    kc_this = min(g.KC, g.K - kc);
                    t.kind = ST_T_ROWS, t.nr = min(uint32_t(kStTileRows), b1 - blk);
                    t.src = W + (size_t(blk) * g.K + kc) * 2;
                    t.row_stride = size_t(g.K) * 2, t.bytes = kc_this * 2, t.pitch = g.KC * 2 + kStPad;
                } else if (g.fmt == WF_W4) {
                    // [super][ktile of 64 k][lane][16 B] and scales [super][ktile][8][4] bf16
                    const size_t kt0 = size_t(blk) * (g.K >> 6) + (kc >> 6);
                    t.kind = ST_T_W4, t.src = W + kt0 * 512, t.bytes = (g.KC >> 6) * 512;
                    t.src2 = st_layer_ptr(P, g, g.scales, li) + kt0 * 64, t.bytes2 = (g.KC >> 6) * 64;
                } else {
                    // [super][ktile of 32 k][lane][16 B]
                    const size_t kt0 = size_t(blk) * (g.K >> 5) + (kc >> 5);
                    t.kind = ST_T_W8, t.src = W + kt0 * 512, t.bytes = (g.KC >> 5) * 512;
                }
                kc += kc_this;
                if (kc >= g.K) kc = 0, blk += (!Q || g.fmt == WF_BF16) ? uint32_t(kStTileRows) : 1u;
                return true;
            }
            next_phase(P);
        }
    }
};
template <bool Q> __device__ __forceinline__ void st_producer(const st_params& P, const st_ctx& c)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t policy = policy_evict_first();
    st_pipe pp{0, 0};
    st_tile_iter<Q> ld;
    st_tile t;
    while (ld.next(P, t)) {
        st_mbar_wait(c, c.empty0 + pp.stage * 8, pp.parity ^ 1u);
        const uint32_t full = c.full0 + pp.stage * 8, dst = c.ring_addr + pp.stage * P.stage_bytes;
        if (!Q || t.kind == ST_T_ROWS) {
            if (lane == 0) mbar_expect_tx(full, t.nr * t.bytes);
            __syncwarp();
            if (lane < t.nr) bulk_g2s(dst + lane * t.pitch, t.src + size_t(lane) * t.row_stride, t.bytes, full, policy);
        } else {
            // a packed tile is one contiguous run; it still goes as 2 KiB pieces (one per lane): many small bulk copies keep more
            // memory requests in flight than one large one
            const uint32_t piece = 2048, n_pieces = (t.bytes + piece - 1) / piece;
            if (lane == 0) mbar_expect_tx(full, t.bytes + (t.kind == ST_T_W4 ? t.bytes2 : 0u));
            __syncwarp();
            if (lane < n_pieces) bulk_g2s(dst + lane * piece, t.src + size_t(lane) * piece, min(piece, t.bytes - lane * piece), full, policy);
            else if (lane == n_pieces && t.kind == ST_T_W4) bulk_g2s(dst + kStW4ScaleOff, t.src2, t.bytes2, full, policy);
        }
        pp.advance(P.n_stages);
    }
}

// ---- consumers: stage the activation rows of a GEMV phase ----------------------------------------------------------------
// Polls the tagged input rows (or gathers the embedding rows), applies RMSNorm when the phase has one, and leaves the rows
// as bf16 in shared memory.  `xres_ll` (first phase of a step only): this CTA also publishes the raw embedding values of
// the rows it will later own in the wo phase - they are the first residual.
template <bool Q> __device__ __forceinline__ void st_stage_input(const st_params& P, const st_gemv& g, uint32_t li, const st_ctx& c, bool embed, uint32_t step, uint32_t tag_in,
                                               uint32_t tag_out, uint32_t gphase)
{
    constexpr int NB = 4; // tagged 16-byte loads in flight per thread
    const uint32_t tid = threadIdx.x, K = g.K, n_words = K >> 1;
    const uint16_t* norm_w = g.pro == PRO_RMSNORM ? reinterpret_cast<const uint16_t*>(reinterpret_cast<const char*>(g.norm_w) + (g.layered ? size_t(li) * P.layer_stride : 0)) : nullptr;
    // norm weights of this thread's elements: requested before anything is polled
    uint2 nw[NB];
    if (norm_w) {
#pragma unroll
        for (int j = 0; j < NB; j++) {
            const uint32_t w = tid * 2 + j * (kStConsumers * 2);
            if (w < n_words) nw[j] = *reinterpret_cast<const uint2*>(norm_w + w * 2);
        }
    }
    int32_t* sids = reinterpret_cast<int32_t*>(c.scr + 8);
    if (embed) {
        if (tid < P.rows) {
            int32_t id = step == 0 ? P.ids[tid] : int32_t(st_poll1(c, P.ids_ll + tid, tag_in, 2));
            if (id < 0 || uint32_t(id) >= P.vocab) id = 0;
            sids[tid] = id;
        }
        consumer_bar();
    }
    if (!embed && P.arrive_on) {
        // the producers of the phase before have all arrived: the loads below find their words at the first attempt
        if (tid == 0) st_arrived_wait(P, gphase - 1);
        consumer_bar();
    }
    for (uint32_t m = 0; m < P.rows; m++) {
        uint16_t* dst = reinterpret_cast<uint16_t*>(c.act + size_t(m) * P.act_pitch);
        float part = 0.0f;
        for (uint32_t w0 = tid * 2; w0 < n_words; w0 += NB * kStConsumers * 2) {
            uint64_t a[NB], b[NB];
            if (embed) {
                // embedding gather fused into the first phase (kernel/embedding.metal:38-66)
                if (!Q || P.tok_fmt == WF_BF16) {
                    const uint16_t* xr = static_cast<const uint16_t*>(P.embed_table) + size_t(sids[m]) * K;
#pragma unroll
                    for (int j = 0; j < NB; j++) {
                        const uint32_t w = w0 + j * (kStConsumers * 2);
                        if (w < n_words) {
                            const uint2 v = *reinterpret_cast<const uint2*>(xr + w * 2);
                            a[j] = v.x, b[j] = v.y;
                        }
                    }
                } else {
                    // lora_embedding: int8 row with one scale, x = r(r(q) * r(s))  (quantization/lora.h:160-170, kernel/mul.metal:76-77)
                    const int8_t* xr = static_cast<const int8_t*>(P.embed_table) + size_t(sids[m]) * K;
                    const float sc = rbf(P.tok_scales[sids[m]]);
#pragma unroll
                    for (int j = 0; j < NB; j++) {
                        const uint32_t w = w0 + j * (kStConsumers * 2);
                        if (w < n_words) {
                            const char4 q = *reinterpret_cast<const char4*>(xr + w * 2);
                            a[j] = pack2(__fmul_rn(float(q.x), sc), __fmul_rn(float(q.y), sc));
                            b[j] = pack2(__fmul_rn(float(q.z), sc), __fmul_rn(float(q.w), sc));
                        }
                    }
                }
            } else {
                const uint64_t* src = g.in_ll + size_t(blockIdx.x % P.n_rep) * P.rep_stride + size_t(m) * n_words;
#pragma unroll
                for (int j = 0; j < NB; j++) {
                    const uint32_t w = w0 + j * (kStConsumers * 2);
                    if (w < n_words) st_ll_load2(src + w, a[j], b[j]);
                }
#pragma unroll
                for (int j = 0; j < NB; j++) {
                    const uint32_t w = w0 + j * (kStConsumers * 2);
                    if (w < n_words) {
                        unsigned spins = 0;
                        while (uint32_t(a[j] >> 32) != tag_in || uint32_t(b[j] >> 32) != tag_in) {
                            if (st_poll_backoff(c, spins, 1)) break;
                            st_ll_load2(src + w, a[j], b[j]);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < NB; j++) {
                const uint32_t w = w0 + j * (kStConsumers * 2);
                if (w < n_words) {
                    const uint32_t lo = uint32_t(a[j]), hi = uint32_t(b[j]);
                    *reinterpret_cast<uint2*>(dst + w * 2) = make_uint2(lo, hi);
                    float f;
                    f = bf_lo(lo), part = fmaf(f, f, part);
                    f = bf_hi(lo), part = fmaf(f, f, part);
                    f = bf_lo(hi), part = fmaf(f, f, part);
                    f = bf_hi(hi), part = fmaf(f, f, part);
                }
            }
        }
        if (g.pro == PRO_RMSNORM) {
            // n = r((0 + w) * x * rsqrt(mean(x^2) + eps))  (kernel/rmsnorm.metal:53-89)
            const float total = st_block_sum(part, c.scr);
            const float inv = 1.0f / sqrtf(__fadd_rn(total / float(K), P.eps));
            if (embed) {
                // the raw rows are the residual of the first wo phase: publish the part this CTA will need there
                uint32_t rb, re;
                st_my_rows(P.g[1].N, P.g[1].gran, rb, re);
                for (uint32_t w = (rb >> 1) + tid; w < (re >> 1); w += kStConsumers)
                    st_ll_store(const_cast<uint64_t*>(P.g[1].res_ll) + size_t(m) * (P.g[1].N >> 1) + w, *reinterpret_cast<const uint32_t*>(dst + w * 2), tag_out);
                consumer_bar();
            }
#define MC_NORM2(d, vv, gg)                                                                          \
    d = uint32_t(f32_to_bf16_bits(__fmul_rn(__fmul_rn(bf_lo(gg), bf_lo(vv)), inv))) |                \
        (uint32_t(f32_to_bf16_bits(__fmul_rn(__fmul_rn(bf_hi(gg), bf_hi(vv)), inv))) << 16)
#pragma unroll
            for (int j = 0; j < NB; j++) {
                const uint32_t w = tid * 2 + j * (kStConsumers * 2);
                if (w < n_words) {
                    const uint2 v = *reinterpret_cast<const uint2*>(dst + w * 2);
                    uint2 o;
                    MC_NORM2(o.x, v.x, nw[j].x);
                    MC_NORM2(o.y, v.y, nw[j].y);
                    *reinterpret_cast<uint2*>(dst + w * 2) = o;
                }
            }
            // rows wider than the NB prefetched norm-weight words per thread (dim > 4096: the 70B shape): the rest reads its weights here
            for (uint32_t w = tid * 2 + NB * (kStConsumers * 2); w < n_words; w += kStConsumers * 2) {
                const uint2 v = *reinterpret_cast<const uint2*>(dst + w * 2), gw = *reinterpret_cast<const uint2*>(norm_w + w * 2);
                uint2 o;
                MC_NORM2(o.x, v.x, gw.x);
                MC_NORM2(o.y, v.y, gw.y);
                *reinterpret_cast<uint2*>(dst + w * 2) = o;
            }
#undef MC_NORM2
        }
    }
    consumer_bar();
}

// ---- one GEMV phase: the mma warps ---------------------------------------------------------------------------------------
// Each of the 8 mma warps multiplies its k-slice of every tile and hands the 16 x 8 partial block to the epilogue warps
// through one of kStRedBufs buffers (ready / free mbarriers): the mma warps never wait for an epilogue to finish.
// Partial block layout: [warp][block row][col], block row = position of the output row inside the block.
struct st_blk {
    uint32_t buf, parity; // next partial buffer and the parity of its current use
    __device__ __forceinline__ void advance()
    {
        if (++buf == uint32_t(kStRedBufs)) buf = 0, parity ^= 1u;
    }
};
__device__ __forceinline__ uint32_t lds32(uint32_t a)
{
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ uint2 lds64(uint32_t a)
{
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a));
    return r;
}
// hands acc (c0,c1 -> block row ra; c2,c3 -> block row rb) to the epilogue warps
__device__ __forceinline__ void st_hand_over(const st_ctx& c, st_blk& bl, const float (&acc)[4], uint32_t ra, uint32_t rb)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, t = lane & 3;
    st_mbar_wait(c, c.fre0 + bl.buf * 8, bl.parity ^ 1u); // the epilogue of the previous use of this buffer is done
    float* rw = c.red + bl.buf * (kStWarps * 128) + warp * 128;
    *reinterpret_cast<float2*>(rw + ra * 8 + 2 * t) = make_float2(acc[0], acc[1]);
    *reinterpret_cast<float2*>(rw + rb * 8 + 2 * t) = make_float2(acc[2], acc[3]);
    __syncwarp();
    if (lane == 0) mbar_arrive(c.rdy0 + bl.buf * 8);
    bl.advance();
}
// (x & mask) | magic in one LOP3 (the compiler emits two when both are immediates)
__device__ __forceinline__ uint32_t and_or(uint32_t x, uint32_t mask, uint32_t magic)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(x), "r"(mask), "r"(magic));
    return d;
}
// one block of bf16 rows: K / kc tiles of `kc` k, the 64-byte padded row-major staging
__device__ __forceinline__ void st_mma_rows(const st_params& P, const st_ctx& c, st_pipe& cp, st_blk& bl, uint32_t K, uint32_t kc_tile, uint32_t nr)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t gq = lane >> 2, t = lane & 3;
    const uint32_t pitch = kc_tile * 2 + kStPad;
    const uint32_t brow = gq < P.rows ? gq : 0;      // batch row of this lane's B fragment (unused columns read row 0)
    // rows beyond the block read a valid row instead (their results are never stored)
    const uint32_t a_row = min(gq, nr - 1) * pitch;
    const uint32_t a_hi = (min(gq + 8, nr - 1) - min(gq, nr - 1)) * pitch;
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (uint32_t kc = 0; kc < K; kc += kc_tile) {
        // the last tile of a block may be shorter (K not a multiple of the tile width; always a multiple of 256 = 8 warps x 32 k)
        const uint32_t kc_this = min(kc_tile, K - kc);
        const uint32_t kw = warp * (kc_this / kStWarps); // this warp's k-slice inside the tile
        const uint32_t ksteps = kc_this / kStWarps / 32;
        st_mbar_wait(c, c.full0 + cp.stage * 8, cp.parity);
        const uint32_t tile = c.ring_addr + cp.stage * P.stage_bytes + a_row + (kw + t * 8) * 2;
        const uint32_t bk = c.act_addr + brow * P.act_pitch + (kc + kw + t * 8) * 2;
        // the k permutation (lane t reads 8 consecutive k) is applied to both operands
#pragma unroll 4
        for (uint32_t s = 0; s < ksteps; s++) {
            const uint4 alo = lds128(tile + s * 64);
            const uint4 ahi = lds128(tile + a_hi + s * 64);
            const uint4 b = lds128(bk + s * 64);
            mma_bf16_16816(acc, alo.x, ahi.x, alo.y, ahi.y, b.x, b.y);
            mma_bf16_16816(acc, alo.z, ahi.z, alo.w, ahi.w, b.z, b.w);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(c.empty0 + cp.stage * 8);
        cp.advance(P.n_stages);
    }
    st_hand_over(c, bl, acc, gq, gq + 8);
}
// The two LoRA-A rows owned by this CTA against the staged activation rows: plain fp32 dot products over the whole consumer group
// (thread t takes the 16-byte chunks t, t + 256, ... of both rows).  An mma tile would spend 16 rows of tensor work on 2 and
// make the owners the slowest CTAs of the phase (measured: +1.3 us at K = 2048, +2.8 us at K = 8192, on the critical path of ax).
// Hands the per-warp partial sums over in the layout the epilogue's A-block code reads: [warp][adaptor row][batch row].
__device__ __forceinline__ void st_dot_a_rows(const st_params& P, const st_ctx& c, st_pipe& cp, st_blk& bl, uint32_t K, uint32_t kc_tile)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t pitch = kc_tile * 2 + kStPad;
    float a0s[kStMaxRows], a1s[kStMaxRows];
#pragma unroll
    for (int m = 0; m < kStMaxRows; m++) a0s[m] = 0.0f, a1s[m] = 0.0f;
    for (uint32_t kc = 0; kc < K; kc += kc_tile) {
        st_mbar_wait(c, c.full0 + cp.stage * 8, cp.parity);
        const uint32_t tile = c.ring_addr + cp.stage * P.stage_bytes;
        for (uint32_t k = threadIdx.x * 8; k < kc_tile; k += kStConsumers * 8) {
            const uint4 w0 = lds128(tile + k * 2), w1 = lds128(tile + pitch + k * 2);
#pragma unroll
            for (int m = 0; m < kStMaxRows; m++) {
                if (uint32_t(m) < P.rows) {
                    const uint4 x = lds128(c.act_addr + uint32_t(m) * P.act_pitch + (kc + k) * 2);
                    a0s[m] = dot8(w0, x, a0s[m]);
                    a1s[m] = dot8(w1, x, a1s[m]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(c.empty0 + cp.stage * 8);
        cp.advance(P.n_stages);
    }
#pragma unroll
    for (int m = 0; m < kStMaxRows; m++) {
        if (uint32_t(m) < P.rows) a0s[m] = warp_sum(a0s[m]), a1s[m] = warp_sum(a1s[m]);
    }
    st_mbar_wait(c, c.fre0 + bl.buf * 8, bl.parity ^ 1u); // the epilogue of the previous use of this buffer is done
    float* rw = c.red + bl.buf * (kStWarps * 128) + warp * 128;
    if (lane == 0) {
#pragma unroll
        for (int m = 0; m < kStMaxRows; m++)
            if (uint32_t(m) < P.rows) rw[m] = a0s[m], rw[8 + m] = a1s[m];
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(c.rdy0 + bl.buf * 8);
    bl.advance();
}
// one super-unit of packed weights: in-register dequant with the reference's two roundings r(r(q) * r(s))
// (kernel/mul.metal:76-77), then mma; see gemv_q_kernel for the fragment layout
__device__ __forceinline__ void st_dbg(unsigned long long* d, uint32_t blk, uint32_t slot, uint32_t who)
{
    if (d && blockIdx.x == 0 && threadIdx.x == who && blk < 64) {
        unsigned long long v;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
        d[blk * 8 + slot] = v;
    }
}
template <int FMT>
__device__ __forceinline__ void st_mma_packed(const st_params& P, const st_gemv& g, const st_ctx& c, st_pipe& cp, st_blk& bl, float rs0, float rs1, unsigned long long* dbg = nullptr, uint32_t dblk = 0)
{
    constexpr uint32_t KT = FMT == WF_W4 ? 64 : 32;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t gq = lane >> 2, t = lane & 3;
    const uint32_t kts = g.KC / KT;                                   // k-tiles per ring tile
    const uint32_t per = (kts + kStWarps - 1) / kStWarps;             // k-tiles of this warp
    const uint32_t kt_b = min(kts, warp * per), kt_e = min(kts, kt_b + per);
    const uint32_t brow = gq < P.rows ? gq : 0;
    const uint32_t xb = c.act_addr + brow * P.act_pitch + 4 * t;      // B fragment base of this lane
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f}, acc2[4] = {0.0f, 0.0f, 0.0f, 0.0f}; // independent chains: consecutive mma do not wait for each other
    float acc3[4] = {0.0f, 0.0f, 0.0f, 0.0f}, acc4[4] = {0.0f, 0.0f, 0.0f, 0.0f}; // (int4: four mma per k-tile, one chain each: 806 -> 763 us per step)
    st_dbg(dbg, dblk, 0, 0);
    // operands of one k-tile: the packed weights, (int4) the group scales, the B fragments.  The loads are volatile asm, i.e.
    // kept in program order, so the k-tile loop is software-pipelined by hand: the operands of k-tile i+1 are requested
    // before k-tile i is dequantised and multiplied.
    constexpr int NJ = FMT == WF_W4 ? 4 : 2;
    struct ops {
        uint4 wv;
        uint2 sv;
        uint32_t b0[NJ], b1[NJ];
    };
    auto load = [&](ops& o, uint32_t tile, uint32_t kc, uint32_t kt) {
        o.wv = lds128(tile + (kt * 32 + lane) * 16);
        if (FMT == WF_W4) o.sv = lds64(tile + kStW4ScaleOff + (kt * 8 + gq) * 8);
        const uint32_t xk = xb + (kc + kt * KT) * 2;
#pragma unroll
        for (int j = 0; j < NJ; j++) o.b0[j] = lds32(xk + j * 32), o.b1[j] = lds32(xk + j * 32 + 16);
    };
    // (kept in registers through volatile asm so that the compiler does not fold them back into immediates)
    uint32_t nib_mask = 0x000f000fu, nib_magic = 0x43004300u;
    asm volatile("" : "+r"(nib_mask), "+r"(nib_magic));
    // (tried in round 2, no effect on the step time -- 0.7635 / 0.7638 / 0.7602 ms on one box: a non-volatile mma so that ptxas may hoist the
    // next fragment's dequant above it, and the dequant of fragment j + 1 written before the mma of fragment j: the loop's issue
    // efficiency is not what bounds the phase)
    auto mma_pk = [&](float (&cc)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) { mma_bf16_16816(cc, a0, a1, a2, a3, b0, b1); };
    auto compute = [&](const ops& o) {
        const uint32_t words[4] = {o.wv.x, o.wv.y, o.wv.z, o.wv.w};
        if (FMT == WF_W4) {
            // scales of (row r0, row r1) for k-groups 0 and 1 of this k-tile, as (s,s) bf16x2
            const uint32_t s_g0 = __byte_perm(o.sv.x, 0, 0x1010), s_g1 = __byte_perm(o.sv.x, 0, 0x3232);
            const uint32_t s_h0 = __byte_perm(o.sv.y, 0, 0x1010), s_h1 = __byte_perm(o.sv.y, 0, 0x3232);
            auto deq = [&](int j, uint32_t (&a)[4]) {
                const uint32_t w = words[j];
                const uint32_t sg = j < 2 ? s_g0 : s_g1, sh = j < 2 ? s_h0 : s_h1;
                a[0] = hmul2_bf16(hsub2_bf16(and_or(w, nib_mask, nib_magic), 0x43084308u), sg);
                a[1] = hmul2_bf16(hsub2_bf16(and_or(w >> 4, nib_mask, nib_magic), 0x43084308u), sh);
                a[2] = hmul2_bf16(hsub2_bf16(and_or(w >> 8, nib_mask, nib_magic), 0x43084308u), sg);
                a[3] = hmul2_bf16(hsub2_bf16(and_or(w >> 12, nib_mask, nib_magic), 0x43084308u), sh);
            };
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t a[4];
                deq(j, a);
                if (j == 0) mma_pk(acc, a[0], a[1], a[2], a[3], o.b0[j], o.b1[j]);
                else if (j == 1) mma_pk(acc2, a[0], a[1], a[2], a[3], o.b0[j], o.b1[j]);
                else if (j == 2) mma_pk(acc3, a[0], a[1], a[2], a[3], o.b0[j], o.b1[j]);
                else mma_pk(acc4, a[0], a[1], a[2], a[3], o.b0[j], o.b1[j]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const uint32_t lo = words[2 * j] ^ 0x80808080u, hi = words[2 * j + 1] ^ 0x80808080u;
                const uint32_t a0 = pack_bf16x2(__fmul_rn(s8_to_f32(lo, 0), rs0), __fmul_rn(s8_to_f32(lo, 1), rs0));
                const uint32_t a1 = pack_bf16x2(__fmul_rn(s8_to_f32(lo, 2), rs1), __fmul_rn(s8_to_f32(lo, 3), rs1));
                const uint32_t a2 = pack_bf16x2(__fmul_rn(s8_to_f32(hi, 0), rs0), __fmul_rn(s8_to_f32(hi, 1), rs0));
                const uint32_t a3 = pack_bf16x2(__fmul_rn(s8_to_f32(hi, 2), rs1), __fmul_rn(s8_to_f32(hi, 3), rs1));
                if (j & 1) mma_pk(acc2, a0, a1, a2, a3, o.b0[j], o.b1[j]);
                else mma_pk(acc, a0, a1, a2, a3, o.b0[j], o.b1[j]);
            }
        }
    };
    for (uint32_t kc = 0; kc < g.K; kc += g.KC) {
        st_mbar_wait(c, c.full0 + cp.stage * 8, cp.parity);
        if (kc == 0) st_dbg(dbg, dblk, 1, 0);
        const uint32_t tile = c.ring_addr + cp.stage * P.stage_bytes;
        if (kt_b < kt_e) {
            ops p0, p1; // ping-pong: no register copies between iterations
            load(p0, tile, kc, kt_b);
            for (uint32_t kt = kt_b; kt < kt_e; kt += 2) {
                if (kt + 1 < kt_e) load(p1, tile, kc, kt + 1);
                compute(p0);
                if (kt + 1 < kt_e) {
                    if (kt + 2 < kt_e) load(p0, tile, kc, kt + 2);
                    compute(p1);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(c.empty0 + cp.stage * 8);
        cp.advance(P.n_stages);
    }
    st_dbg(dbg, dblk, 2, 0);
#pragma unroll
    for (int i = 0; i < 4; i++) acc[i] = (acc[i] + acc2[i]) + (acc3[i] + acc4[i]);
    // mma row g = unit su*8+g row r0, row g+8 = its row r1: rope pairs keep that order, adjacent-row units interleave
    if (g.qkv_map) st_hand_over(c, bl, acc, gq, gq + 8);
    else st_hand_over(c, bl, acc, 2 * gq, 2 * gq + 1);
    st_dbg(dbg, dblk, 3, 0);
}
template <bool Q> __device__ __forceinline__ void st_mma_gemv(const st_params& P, const st_gemv& g, const st_ctx& c, st_pipe& cp, st_blk& bl, uint32_t li)
{
    const st_range r(g, Q);
    if (Q && r.has_a) st_dot_a_rows(P, c, cp, bl, g.K, P.kc_a[&g - P.g]);
    if (!Q || g.fmt == WF_BF16) {
        for (uint32_t r0 = r.b0; r0 < r.b1; r0 += kStTileRows) st_mma_rows(P, c, cp, bl, g.K, g.KC, min(uint32_t(kStTileRows), r.b1 - r0));
    } else if (g.fmt == WF_W4) {
        for (uint32_t su = r.b0; su < r.b1; su++) st_mma_packed<WF_W4>(P, g, c, cp, bl, 0.0f, 0.0f);
    } else {
        // int8 rows with one fp32 scale per row: the scales of the next super-unit are requested one block ahead
        const uint32_t gq = (threadIdx.x & 31) >> 2;
        const float* sc = reinterpret_cast<const float*>(st_layer_ptr(P, g, g.scales, li));
        float s0 = 0.0f, s1 = 0.0f;
        if (r.b0 < r.b1) s0 = sc[r.b0 * 16 + 2 * gq], s1 = sc[r.b0 * 16 + 2 * gq + 1];
        for (uint32_t su = r.b0; su < r.b1; su++) {
            float n0 = 0.0f, n1 = 0.0f;
            if (su + 1 < r.b1) n0 = sc[(su + 1) * 16 + 2 * gq], n1 = sc[(su + 1) * 16 + 2 * gq + 1];
            st_mma_packed<WF_W8ROW>(P, g, c, cp, bl, rbf(s0), rbf(s1), P.dbg, su - r.b0);
            s0 = n0, s1 = n1;
        }
    }
}

// greedy argmax state of one storing epilogue lane, one slot per column (lowest index on ties)
struct st_best {
    float v[kStMaxRows];
    int32_t i[kStMaxRows];
};

// ---- one GEMV phase: the epilogue warps ------------------------------------------------------------------------------------
// Thread (row group, col) of the 16 x 8 output block joins the 8 k-slices in warp order and finishes `fin` consecutive rows
// (two; four for the gate/up pairs) so that every store is one tagged word.
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 2, 64;" ::: "memory"); }
__device__ __forceinline__ float st_lora_dot(const uint16_t* brow, const float* ax, uint32_t rank)
{
    float l = 0.0f;
    for (uint32_t j = 0; j < rank; j += 8) {
        const uint4 w = *reinterpret_cast<const uint4*>(brow + j);
        l = fmaf(ax[j + 0], bf_lo(w.x), l), l = fmaf(ax[j + 1], bf_hi(w.x), l);
        l = fmaf(ax[j + 2], bf_lo(w.y), l), l = fmaf(ax[j + 3], bf_hi(w.y), l);
        l = fmaf(ax[j + 4], bf_lo(w.z), l), l = fmaf(ax[j + 5], bf_hi(w.z), l);
        l = fmaf(ax[j + 6], bf_lo(w.w), l), l = fmaf(ax[j + 7], bf_hi(w.w), l);
    }
    return l;
}
template <bool Q, bool TP> __device__ __forceinline__ void st_epi_gemv(const st_params& P, const st_gemv& g, const st_ctx& c, st_blk& bl, st_best& best, bool is_head, uint32_t li,
                                            uint32_t tag_out, uint32_t res_tag)
{
    // Thread (block row, quarter): four consecutive lanes share one of the 16 block rows; each joins two of the eight
    // k-slices and a quarter of the LoRA dot product, the quarters meet through two shuffles.  Rows meet their neighbours
    // (pairs; gate/up quads) through shuffles too, so that every store is one tagged word.  Columns (batch rows) are looped.
    const uint32_t et = threadIdx.x - kStConsumers, lane = et & 31;
    const uint32_t erow = et >> 2, sub = et & 3;
    const uint32_t fin = g.epi == EPI_SWIGLU ? 4u : 2u;
    const uint32_t out_pitch = g.epi == EPI_SWIGLU ? g.N >> 2 : g.N >> 1;
    const bool storer = sub == 0 && (erow & (fin - 1)) == 0;
    const st_range r(g, Q);
    float* sax = reinterpret_cast<float*>(c.act + P.sax_off); // [8 rows][n_a] r(A . x) of this phase
    // Tensor parallel, row-parallel phase (wo, w2): see pass 1 / pass 2 below
    const bool tp_sum = TP && g.epi == EPI_RESIDUAL; // (a compile-time false without tensor parallelism: the single-GPU kernels carry none of this code)
    const uint32_t tp_kind = (&g - P.g) == 3 ? 1u : 0u;
    if (Q && r.has_a) {
        // the two adaptor rows owned by this CTA: ax = r(A . x) (quantization/lora.h:115), one tagged word per batch row
        st_mbar_wait(c, c.rdy0 + bl.buf * 8, bl.parity);
        bool done = false;
        if constexpr (TP) if (tp_sum) {
            // row-parallel linear: A holds this rank's k range only, so the two sums are partial.  They go to every peer as tagged fp32
            // words, the peers' arrive the same way; the ranks are summed in rank order and rounded once, like the single-GPU value.
            // item = (batch row, source rank, which of the two adaptor rows), dealt over the 64 epilogue threads
            const uint32_t W = P.tp_world, a0 = 2 * blockIdx.x, items = P.rows * W * 2;
            float* tmp = c.tpk; // [row][src][2] (pass 1 below has not started yet)
            auto own = [&](uint32_t row, uint32_t which) {
                const float* rr = c.red + bl.buf * (kStWarps * 128) + which * 8 + row;
                float v = 0.0f;
#pragma unroll
                for (int w = 0; w < kStWarps; w++) v += rr[w * 128];
                return v;
            };
            auto slot = [&](uint32_t at, uint32_t src, uint32_t row, uint32_t which) {
                return P.tp_ax[at] + (size_t(tp_kind * W + src) * kStMaxRows + row) * kStTpAxCols + a0 + which;
            };
            for (uint32_t i = et; i < items; i += kStEpiThreads) {
                const uint32_t row = i / (2 * W), k = (i >> 1) % W, which = i & 1u;
                const float v = own(row, which);
                if (k == P.tp_rank) tmp[i] = v;
                else st_ll_store_sys(slot(k, P.tp_rank, row, which), __float_as_uint(v), tag_out);
            }
            for (uint32_t i = et; i < items; i += kStEpiThreads) {
                const uint32_t row = i / (2 * W), k = (i >> 1) % W, which = i & 1u;
                if (k != P.tp_rank) tmp[i] = __uint_as_float(st_poll1_sys(c, slot(P.tp_rank, k, row, which), tag_out, 9));
            }
            epi_bar();
            if (et < P.rows) {
                float s0 = 0.0f, s1 = 0.0f;
                for (uint32_t k = 0; k < W; k++) s0 += tmp[(et * W + k) * 2], s1 += tmp[(et * W + k) * 2 + 1];
                st_ll_store(g.ax_ll + size_t(et) * (g.n_a >> 1) + blockIdx.x, pack2(s0, s1), tag_out);
            }
            epi_bar(); // tmp is pass 1's store from here on
            done = true;
        }
        if (!done && et < P.rows) {
            const float* rr = c.red + bl.buf * (kStWarps * 128) + et;
            float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
            for (int w = 0; w < kStWarps; w++) s0 += rr[w * 128], s1 += rr[w * 128 + 8];
            st_ll_store(g.ax_ll + size_t(et) * (g.n_a >> 1) + blockIdx.x, pack2(s0, s1), tag_out);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(c.fre0 + bl.buf * 8);
        bl.advance();
    }
    const uint32_t hd = P.head_dim, half = hd >> 1;
    const uint32_t rank = P.lora_rank, rq = rank >> 2; // adaptor columns of this quarter: [sub * rq, sub * rq + rq)
    const uint16_t* lora_b = Q && g.n_a ? reinterpret_cast<const uint16_t*>(st_layer_ptr(P, g, g.lora_b, li)) : nullptr;
    const bool fast_b = lora_b != nullptr && rank == 16; // the adaptor quarter of a block is requested one block ahead
    const uint32_t bstep = (!Q || g.fmt == WF_BF16) ? uint32_t(kStTileRows) : 1u;
    // output row of this thread's block row, and the rows of the block
    auto row_of = [&](uint32_t b, uint32_t& R, uint32_t& nr) {
        if (!Q || g.fmt == WF_BF16) R = b + erow, nr = min(uint32_t(kStTileRows), r.b1 - b);
        else if (!g.qkv_map) R = b * 16 + erow, nr = 16;
        else {
            const uint32_t unit0 = b * 8, head = unit0 / half, j0 = unit0 - head * half;
            R = head * hd + j0 + (erow & 7u) + (erow >> 3) * half, nr = 16;
        }
    };
    // epilogue operands of a block, requested early: the residual words (lane `sub` holds those of columns sub and sub + 4;
    // verified by their tag when used: the epilogue warps of a CTA without blocks run ahead of everybody else) and the
    // adaptor quarter
    uint64_t resw[2] = {0, 0}, n_resw[2] = {0, 0};
    uint2 bq = make_uint2(0, 0), n_bq = make_uint2(0, 0);
    auto fetch = [&](uint32_t b, uint64_t (&rw)[2], uint2& q) {
        uint32_t R, nr;
        row_of(b, R, nr);
        if (erow >= nr) return;
        if (g.epi == EPI_RESIDUAL && (erow & 1u) == 0) {
            if (sub < P.rows) rw[0] = st_ll_load1(g.res_ll + size_t(sub) * out_pitch + (R >> 1));
            if (sub + 4 < P.rows) rw[1] = st_ll_load1(g.res_ll + size_t(sub + 4) * out_pitch + (R >> 1));
        }
        if (fast_b) q = *reinterpret_cast<const uint2*>(lora_b + size_t(R) * 16 + 4 * sub);
    };
    // Tensor parallel, row-parallel phase (wo, w2): pass 1 joins the k-slices of every block of this CTA into fp32 partial sums, keeps
    // them in shared memory and sends them to the same rows' owner on every peer (tagged words, peer stores over NVLink); pass 2 -- the
    // block loop below -- polls the peers' words of a block, sums the ranks in rank order and finishes the rows.  The mma warps get
    // their partial buffers back in pass 1 already; NVLink latency is paid once per phase, not once per block.
    if constexpr (TP) if (tp_sum) {
        uint32_t bi = 0;
        for (uint32_t b = r.b0; b < r.b1; b += bstep, bi++) {
            uint32_t R, nr;
            row_of(b, R, nr);
            st_mbar_wait(c, c.rdy0 + bl.buf * 8, bl.parity);
            const float* rr = c.red + bl.buf * (kStWarps * 128) + (2 * sub) * 128 + erow * 8;
#pragma unroll
            for (uint32_t col = 0; col < uint32_t(kStMaxRows); col++) {
                if (col >= P.rows) break;
                float sum = rr[col] + rr[128 + col];
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                if (erow < nr && bi < uint32_t(kStTpBlocks)) {
                    if (sub == 0) c.tpk[(bi * 16 + erow) * kStMaxRows + col] = sum;
                    // lane `sub` of a row serves peers sub and sub + 4
                    for (uint32_t k = sub; k < P.tp_world; k += 4)
                        if (k != P.tp_rank)
                            st_ll_store_sys(P.tp_part[k] + (size_t(tp_kind * P.tp_world + P.tp_rank) * kStMaxRows + col) * P.tp_dim + R, __float_as_uint(sum), tag_out);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(c.fre0 + bl.buf * 8);
            bl.advance();
        }
        epi_bar(); // the partial sums kept by one epilogue warp are read by whichever warp owns the row in pass 2 (the same one; cheap)
    }
    if (r.b0 < r.b1) fetch(r.b0, resw, bq);
    if (Q && g.n_a && r.b0 < r.b1) {
        // every CTA needs all of ax for its LoRA-B epilogue (the barrier: the previous phase may still be reading its own)
        epi_bar();
        for (uint32_t i = et; i < P.rows * (g.n_a >> 1); i += kStEpiThreads) {
            const uint32_t m = i / (g.n_a >> 1), w = i - m * (g.n_a >> 1);
            const uint32_t v = st_poll1(c, g.ax_ll + size_t(m) * (g.n_a >> 1) + w, tag_out, 5);
            sax[m * g.n_a + 2 * w] = bf_lo(v), sax[m * g.n_a + 2 * w + 1] = bf_hi(v);
        }
        epi_bar();
    }
    for (uint32_t b = r.b0; b < r.b1; b += bstep) {
        uint32_t R, nr;
        row_of(b, R, nr);
        if (b + bstep < r.b1) fetch(b + bstep, n_resw, n_bq); // in flight while this block is finished
        uint32_t sl = 0; // which rank-wide slice of ax this weight row multiplies
        if (g.ax_slices == 2) sl = R & 1u;
        else if (g.ax_slices == 3) sl = R < g.slice_rows0 ? 0u : (R < g.slice_rows1 ? 1u : 2u);
        if (!tp_sum) st_mbar_wait(c, c.rdy0 + bl.buf * 8, bl.parity);
        if (is_head) st_dbg(P.dbg, (b - r.b0) / bstep, 4, kStConsumers);
        const float* rr = c.red + bl.buf * (kStWarps * 128) + (2 * sub) * 128 + erow * 8;
        const uint32_t tp_bi = (b - r.b0) / bstep;
#pragma unroll
        for (uint32_t col = 0; col < uint32_t(kStMaxRows); col++) {
            if (col >= P.rows) break; // (uniform) the loop is unrolled so that the per-column state stays in registers
            float sum = 0.0f;
            if (!tp_sum) {
                sum = rr[col] + rr[128 + col];
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            } else if constexpr (TP) {
                // all-reduce: the partial sums of the `world` ranks in rank order (this rank's from shared memory, the peers' from the words
                // they stored into this GPU's exchange region), then ONE rounding: h = r(x + r(sum)) like the single-GPU chain
                float pv[2] = {0.0f, 0.0f};
                const bool live = erow < nr && tp_bi < uint32_t(kStTpBlocks);
#pragma unroll
                for (uint32_t j = 0; j < 2; j++) {
                    const uint32_t src = sub + 4 * j;
                    if (live && src < P.tp_world && src != P.tp_rank)
                        pv[j] = __uint_as_float(st_poll1_sys(c, P.tp_part[P.tp_rank] + (size_t(tp_kind * P.tp_world + src) * kStMaxRows + col) * P.tp_dim + R, tag_out, 8));
                }
                const float own = live ? c.tpk[(tp_bi * 16 + erow) * kStMaxRows + col] : 0.0f;
                sum = 0.0f;
                for (uint32_t src = 0; src < P.tp_world; src++) {
                    const float v = __shfl_sync(0xffffffffu, src < 4 ? pv[0] : pv[1], (lane & ~3u) | (src & 3u));
                    sum += src == P.tp_rank ? own : v;
                }
            }
            float y = rbf(sum); // the bmm output buffer is T (kernel/bmm.metal:76)
            if (lora_b) {
                // y = r(y + r(r(B . ax) * scale))   (quantization/lora.h:115-122)
                const float* ax = sax + col * g.n_a + sl * rank + sub * rq;
                float l = 0.0f;
                if (fast_b) {
                    const float4 x4 = *reinterpret_cast<const float4*>(ax);
                    l = fmaf(x4.x, bf_lo(bq.x), l), l = fmaf(x4.y, bf_hi(bq.x), l), l = fmaf(x4.z, bf_lo(bq.y), l), l = fmaf(x4.w, bf_hi(bq.y), l);
                } else if (erow < nr) {
                    const uint16_t* br = lora_b + size_t(R) * rank + sub * rq;
                    for (uint32_t j = 0; j < rq; j++) l = fmaf(ax[j], bf16_bits_to_f32(br[j]), l);
                }
                l += __shfl_xor_sync(0xffffffffu, l, 1);
                l += __shfl_xor_sync(0xffffffffu, l, 2);
                y = rbf(__fadd_rn(y, rbf(__fmul_rn(rbf(l), P.lora_scale))));
            }
            const float y1 = __shfl_xor_sync(0xffffffffu, y, 4); // the next block row
            const bool st = storer && erow < nr;
            if (g.epi == EPI_SWIGLU) {
                // z = r(silu_T(g) * u), rows (2i, 2i+1) = (w1 row i, w3 row i)  (nn/transformer.h:57-59); two outputs per word
                const float z = __fmul_rn(silu_bf16(y), y1);
                const float z1 = __shfl_xor_sync(0xffffffffu, z, 8);
                if (st) st_ll_store_rep(P, g.out_ll + size_t(col) * out_pitch + (R >> 2), pack2(z, z1), tag_out);
            } else if (g.epi == EPI_RESIDUAL) {
                // h = r(x + a)  (nn/transformer.h:133,139); the residual word of column `col` sits in lane (col & 3) of this row
                const uint64_t rw = __shfl_sync(0xffffffffu, col < 4 ? resw[0] : resw[1], (lane & ~3u) | (col & 3u));
                if (st) {
                    uint32_t rv = uint32_t(rw);
                    if (uint32_t(rw >> 32) != res_tag) rv = st_poll1(c, g.res_ll + size_t(col) * out_pitch + (R >> 1), res_tag, 6);
                    st_ll_store_rep(P, g.out_ll + size_t(col) * out_pitch + (R >> 1), pack2(__fadd_rn(bf_lo(rv), y), __fadd_rn(bf_hi(rv), y1)), tag_out);
                }
            } else if (!is_head) {
                if (st) st_ll_store_rep(P, g.out_ll + size_t(col) * out_pitch + (R >> 1), pack2(y, y1), tag_out);
            } else if (st) {
                *reinterpret_cast<uint32_t*>(g.y + size_t(col) * g.N + R) = pack2(y, y1);
                // greedy argmax of column `col`: kept by the storing lanes, one slot per column
                if (y > best.v[col] || (y == best.v[col] && int32_t(R) < best.i[col])) best.v[col] = y, best.i[col] = int32_t(R);
                if (y1 > best.v[col] || (y1 == best.v[col] && int32_t(R + 1) < best.i[col])) best.v[col] = y1, best.i[col] = int32_t(R + 1);
            }
        }
        __syncwarp();
        if (!tp_sum) {
            if (lane == 0) mbar_arrive(c.fre0 + bl.buf * 8);
            bl.advance();
        }
        if (is_head) st_dbg(P.dbg, (b - r.b0) / bstep, 5, kStConsumers);
        resw[0] = n_resw[0], resw[1] = n_resw[1], bq = n_bq;
    }
}

// ---- consumers: attention phase ---------------------------------------------------------------------------------------
// item = (row, head, part).  SPLITS = 4: the four parts of a head split the cached positions for the scores and the head dimensions
// for the output (long histories: four SMs stream one head's K / V); SPLITS = 1: one CTA does a whole head -- no score exchange,
// i.e. one dependent hop less per block, which is what counts while a head's K / V (<= 1024 positions) is a few round trips of one SM.
//   1. rotate q and (position owner only) the new k (kernel/rope.metal:47-58); the owner of a kv head appends k', v to the
//      cache (nn/cache.h:207-214);
//   2. s[t] = r(r(q.K[t]) * scale) for this part's positions, (SPLITS > 1) published as tagged words;
//   3. (SPLITS > 1) every part polls ALL scores of the head; p = r(exp(s) / sum exp(s)) (no max shift, kernel/softmax.metal:40-80) in
//      one fixed order, identical on all parts;
//   4. o[d] = r(sum_t p[t] V[t][d]) for this part's share of the head dimensions, published as tagged words.
// Measured on B200 (1B, KV 512): one CTA per head is SLOWER than four (bf16 0.706 vs 0.649 ms per step, int4 0.845 vs 0.763): a single SM has
// too few loads in flight for a head's 128 KB of K / V, which costs more than the score-exchange hop saves.  The variant stays reachable
// through MC_ATTN_SINGLE_MAX (positions up to which one CTA takes a whole head); default 0 = never.
constexpr uint32_t kStAttnSingleMax = 0;
template <int HD, int SPLITS>
__device__ __forceinline__ void st_attention(const st_params& P, uint32_t li, uint32_t step, const st_ctx& c, uint32_t tag_in, uint32_t tag_out)
{
    constexpr int LPP = HD / 8;                 // lanes per cached K position (16 bytes each)
    constexpr int SLOTS = kStConsumers / LPP;   // K positions per sweep
    constexpr int IT = SPLITS == 1 ? 9 : 5;     // K sweeps kept in flight
    constexpr int DQ = HD / SPLITS;             // head dims finished by this part
    constexpr int VB = SPLITS == 1 ? 16 : 8;    // bytes of a V row per lane
    constexpr int VD = VB / 2;                  // head dims per lane
    constexpr int VL = DQ / VD;                 // lanes per cached V position
    constexpr int VSLOTS = kStConsumers / VL;   // V positions per sweep
    constexpr int VIT = SPLITS == 1 ? 8 : (HD == 64 ? 9 : 18); // V sweeps held in registers while the scores are finished
    const uint32_t tid = threadIdx.x;
    const uint32_t H = P.n_heads, KV = P.n_kv_heads, half = HD / 2, QKVW = (H + 2 * KV) * HD / 2;
    float* rawq = reinterpret_cast<float*>(c.act); // [HD] q as stored by the QKV phase
    float* rawk = rawq + HD;                        // [HD]
    float* sq = rawk + HD;                          // [HD] rotated q
    float* sk = sq + HD;                            // [HD] rotated new k
    float* sv = sk + HD;                            // [HD] new v
    float* spart = sv + HD;                         // [VSLOTS][DQ] (<= 2048 floats)
    float* sp = spart + VSLOTS * DQ;                // [np] scores, then probabilities
    const uint32_t slot = tid / LPP, dl = tid % LPP;
    const uint32_t vslot = tid / VL, vl = tid % VL;
    const size_t kv_off = size_t(li) * P.kv_layer_stride;
    const uint32_t n_items = P.rows * H * SPLITS;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t part = item % SPLITS, head = (item / SPLITS) % H, row = (item / SPLITS) / H;
        const int32_t seq = P.row_seq[row];
        const uint32_t pos = uint32_t(P.pos[row]) + step;
        const uint32_t np = pos + 1;
        const uint32_t kvh = head / (H / KV);
        const uint32_t chunk = SPLITS == 1 ? np : ((((np + SPLITS - 1) / SPLITS) + 1) & ~1u); // even: two scores per tagged word
        const uint32_t t0 = min(np, part * chunk), t1 = min(np, t0 + chunk);
        const uint32_t tc1 = min(t1, pos); // positions below `pos` come from the cache, `pos` itself is fresh
        const bool own = pos >= t0 && pos < t1;
        const bool writer = own && head % (H / KV) == 0;
        const size_t coff = kv_off + (size_t(seq) * KV + kvh) * P.max_seq * HD;
        const uint16_t* Kc = P.kcache + coff;
        const uint16_t* Vc = P.vcache + coff;

        // first block of K rows: requested before anything else is touched
        uint4 kreg[IT];
#pragma unroll
        for (int i = 0; i < IT; i++) {
            const uint32_t tt = t0 + i * SLOTS + slot;
            if (tt < tc1) kreg[i] = ldcg128(Kc + size_t(tt) * HD + dl * 8);
        }
        // q, v (and k for the owner) of this head from the QKV phase
        {
            const uint64_t* qrow = P.qkv_ll + size_t(blockIdx.x % P.n_rep) * P.rep_stride + size_t(row) * QKVW;
            const uint32_t grp = tid / half, i = tid % half; // 0: q, 1: v, 2: k
            if (grp < 2 || (grp == 2 && own)) {
                const uint32_t base = grp == 0 ? head * HD : (grp == 1 ? (H + KV + kvh) * HD : (H + kvh) * HD);
                const uint32_t w = st_poll1(c, qrow + (base >> 1) + i, tag_in, 3);
                float* d = grp == 0 ? rawq : (grp == 1 ? sv : rawk);
                d[2 * i] = bf_lo(w), d[2 * i + 1] = bf_hi(w);
                if (grp == 1 && writer) *reinterpret_cast<uint32_t*>(P.vcache + coff + size_t(pos) * HD + 2 * i) = w;
            }
        }
        consumer_bar();
        if (tid < half) {
            const float cs = P.fcos[size_t(pos) * half + tid], sn = P.fsin[size_t(pos) * half + tid];
            const float q0 = rawq[tid], q1 = rawq[tid + half];
            sq[tid] = rbf(__fsub_rn(__fmul_rn(cs, q0), __fmul_rn(sn, q1)));
            sq[tid + half] = rbf(__fadd_rn(__fmul_rn(sn, q0), __fmul_rn(cs, q1)));
            if (own) {
                const float k0 = rawk[tid], k1 = rawk[tid + half];
                const float o0 = rbf(__fsub_rn(__fmul_rn(cs, k0), __fmul_rn(sn, k1)));
                const float o1 = rbf(__fadd_rn(__fmul_rn(sn, k0), __fmul_rn(cs, k1)));
                sk[tid] = o0, sk[tid + half] = o1;
                if (writer) {
                    uint16_t* kd = P.kcache + coff + size_t(pos) * HD;
                    kd[tid] = f32_to_bf16_bits(o0), kd[tid + half] = f32_to_bf16_bits(o1);
                }
            }
        }
        if (writer) __threadfence(); // the appended row is read by other CTAs in the next step, long after the tagged words below
        consumer_bar();
        float qv[8];
#pragma unroll
        for (int i = 0; i < 8; i++) qv[i] = sq[dl * 8 + i];
        // V rows of the cached positions, this part's dims: requested before the scores are computed, in flight meanwhile
        auto load_v = [&](uint32_t tt, uint32_t (&dst)[VB / 4]) {
            const uint16_t* vp = Vc + size_t(tt) * HD + part * DQ + vl * VD;
            if (VB == 16) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(dst[0]), "=r"(dst[1]), "=r"(dst[VB / 4 - 2]), "=r"(dst[VB / 4 - 1]) : "l"(vp));
            else asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(dst[0]), "=r"(dst[1]) : "l"(vp));
        };
        // scores of this part's cached positions
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
                if (tt < tc1 && dl == 0) sp[tt] = rbf(__fmul_rn(rbf(d), P.scale));
            }
        }
        uint32_t vreg[VIT][VB / 4];
#pragma unroll
        for (int i = 0; i < VIT; i++) {
            const uint32_t tt = i * VSLOTS + vslot;
            if (tt < pos) load_v(tt, vreg[i]);
        }
        if (own && tid < 32) {
            float d = 0.0f;
            for (uint32_t i = tid; i < uint32_t(HD); i += 32) d = fmaf(sq[i], sk[i], d);
            d = warp_sum(d);
            if (tid == 0) sp[pos] = rbf(__fmul_rn(rbf(d), P.scale));
        }
        consumer_bar();
        if (SPLITS > 1) {
            // publish this part's scores, two per word, then collect everybody's
            uint64_t* sc = P.sc_ll + size_t(item / SPLITS) * P.sc_words;
            for (uint32_t tt = t0 + 2 * tid; tt < t1; tt += 2 * kStConsumers)
                st_ll_store(sc + (tt >> 1), pack2(sp[tt], tt + 1 < t1 ? sp[tt + 1] : 0.0f), tag_out);
            consumer_bar(); // sp is about to be overwritten with everybody's scores
            for (uint32_t w = tid; w < (np + 1) / 2; w += kStConsumers) {
                const uint32_t v = st_poll1(c, sc + w, tag_out, 4);
                sp[2 * w] = bf_lo(v), sp[2 * w + 1] = bf_hi(v);
            }
            consumer_bar();
        }
        float part_sum = 0.0f;
        for (uint32_t i = tid; i < np; i += kStConsumers) part_sum += expf(sp[i]);
        const float total = st_block_sum(part_sum, c.scr);
        const float inv = 1.0f / total;
        for (uint32_t i = tid; i < np; i += kStConsumers) sp[i] = rbf(__fmul_rn(expf(sp[i]), inv));
        consumer_bar();
        float acc[VD];
#pragma unroll
        for (int i = 0; i < VD; i++) acc[i] = 0.0f;
        auto fma_v = [&](float pt, const uint32_t (&v)[VB / 4]) {
#pragma unroll
            for (int j = 0; j < VB / 4; j++) {
                acc[2 * j] = fmaf(pt, bf_lo(v[j]), acc[2 * j]);
                acc[2 * j + 1] = fmaf(pt, bf_hi(v[j]), acc[2 * j + 1]);
            }
        };
#pragma unroll
        for (int i = 0; i < VIT; i++) {
            const uint32_t tt = i * VSLOTS + vslot;
            if (tt < pos) fma_v(sp[tt], vreg[i]);
        }
        // the rest of the history in batches of VIT sweeps (all loads of a batch in flight together)
        for (uint32_t tb = VIT * VSLOTS; tb < pos; tb += VIT * VSLOTS) {
#pragma unroll
            for (int i = 0; i < VIT; i++) {
                const uint32_t tt = tb + i * VSLOTS + vslot;
                if (tt < pos) load_v(tt, vreg[i]);
            }
#pragma unroll
            for (int i = 0; i < VIT; i++) {
                const uint32_t tt = tb + i * VSLOTS + vslot;
                if (tt < pos) fma_v(sp[tt], vreg[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < VD; i++) spart[vslot * DQ + vl * VD + i] = acc[i];
        consumer_bar();
        if (tid < DQ / 2) {
            float o0 = 0.0f, o1 = 0.0f;
            for (int s2 = 0; s2 < VSLOTS; s2++) o0 += spart[s2 * DQ + 2 * tid], o1 += spart[s2 * DQ + 2 * tid + 1];
            const float pp = sp[pos];
            o0 = fmaf(pp, sv[part * DQ + 2 * tid], o0);
            o1 = fmaf(pp, sv[part * DQ + 2 * tid + 1], o1);
            st_ll_store_rep(P, P.attn_ll + size_t(row) * (H * HD / 2) + ((head * HD + part * DQ) >> 1) + tid, pack2(o0, o1), tag_out);
        }
        consumer_bar(); // the scratch is reused by the next item
    }
}

// ---- the kernel ----------------------------------------------------------------------------------------------------------
// Q = the model has packed (int4 / int8) layers and LoRA adaptors; the bf16 instantiation carries none of that code
// TP = tensor-parallel shard (all-reduce fused into the wo / w2 epilogues, cross-rank argmax); ATTN_SPLITS = CTAs per attention head (1: every
// history of the launch is at most attn_single_max positions, decided by the host; kStSplits otherwise)
template <bool Q, int HD, bool TP, int ATTN_SPLITS> __global__ void __launch_bounds__(kStThreads, 1) decode_stream_kernel(const __grid_constant__ st_params P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    st_ctx c;
    c.full0 = smem_u32(smem);
    c.empty0 = c.full0 + kStMaxStages * 8;   // 96
    c.rdy0 = c.empty0 + kStMaxStages * 8;    // 192
    c.fre0 = c.rdy0 + kStRedBufs * 8;        // 216
    c.sbar = c.fre0 + kStRedBufs * 8;        // 240: the mma warps finished the last phase of a step
    c.dead = reinterpret_cast<volatile int*>(smem + 248);
    c.scr = reinterpret_cast<float*>(smem + 256);
    c.escr = reinterpret_cast<float*>(smem + 320);
    c.red = reinterpret_cast<float*>(smem + kStHdrBytes);
    c.tpk = reinterpret_cast<float*>(smem + kStHdrBytes + kStRedBytes);
    c.act = smem + kStHdrBytes + kStRedBytes + (TP ? kStTpKeepBytes : 0); // single-GPU kernels give those 4 KiB to the weight ring
    c.act_addr = smem_u32(c.act);
    c.ring_addr = c.act_addr + P.act_bytes;
    c.err = P.err;
    c.poll_ns = P.poll_ns;
    const uint32_t tid = threadIdx.x;
    const unsigned G = gridDim.x;
    if (tid == 0) {
        for (uint32_t s = 0; s < P.n_stages; s++) {
            mbar_init(c.full0 + s * 8, 1);
            mbar_init(c.empty0 + s * 8, kStWarps);
        }
        for (uint32_t s = 0; s < uint32_t(kStRedBufs); s++) {
            mbar_init(c.rdy0 + s * 8, kStWarps);
            mbar_init(c.fre0 + s * 8, kStEpiThreads / 32);
        }
        mbar_init(c.sbar, kStWarps);
        *c.dead = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const uint32_t phases_per_step = P.n_layers * 5 + 1;

#ifdef ST_USE_SETMAXNREG
    // register budget: 384 threads x 168 at launch; warpgroup 2 (epilogue + producer warps) hands registers to the two mma warpgroups
    if (tid >= kStConsumers) asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
#endif
    if (tid >= kStConsumers + kStEpiThreads) {
        // ---- producer warp: the whole launch's weight stream, never blocked by anything but a full ring
        if (tid < kStConsumers + kStEpiThreads + 32) st_producer<Q>(P, c);
        return;
    }
    if (tid >= kStConsumers) {
        // ---- epilogue warps: finish every 16-row block of every GEMV phase, then the sampler tail
        const uint32_t et = tid - kStConsumers, ewarp = et >> 5, lane = et & 31;
        st_blk bl{0, 0};
        const int32_t log0 = P.step_counter[0];
        for (uint32_t step = 0; step < P.steps; step++) {
            unsigned gphase = step * phases_per_step;
            st_best best;
#pragma unroll
            for (int k = 0; k < kStMaxRows; k++) best.v[k] = -INFINITY, best.i[k] = 0x7fffffff;
            for (uint32_t li = 0; li <= P.n_layers; li++) {
                const bool is_head = li == P.n_layers;
                for (uint32_t kind = 0; kind < (is_head ? 1u : 5u); kind++, gphase++) {
                    if (!is_head && kind == 1) continue;
                    // the residual of wo is the x of this layer (embedding: this step's first phase; else the w2 before), of w2 the h of wo
                    const uint32_t tag_out = P.tag_base + gphase + 1;
                    const uint32_t res_tag = kind == 2 ? (li == 0 ? tag_out - 2 : tag_out - 3) : tag_out - 2;
                    st_epi_gemv<Q, TP>(P, P.g[is_head ? 4u : (kind == 0 ? 0u : kind - 1)], c, bl, best, is_head, is_head ? 0 : li, tag_out, res_tag);
                    if (P.arrive_on) {
                        epi_bar(); // the stores of both epilogue warps precede the arrival (release, cumulative)
                        if (et == 0) st_arrive(P, gphase);
                    }
                    st_stamp(P.timing ? P.timing + (size_t(blockIdx.x) * (P.steps * phases_per_step) + gphase) * 4 : nullptr, 1);
                }
            }
            // sampler tail (greedy): every CTA publishes its argmax partial, CTA 0 joins them and publishes the next id.
            // The epilogue warps of a CTA without blocks run ahead of everybody: they may publish the partial of step s only when
            // the mma warps of this CTA have finished step s (else the partials of later steps would overwrite it unread).
            st_mbar_wait(c, c.sbar, step & 1u);
            const uint32_t tag_head = P.tag_base + gphase; // = tag_out of the head phase
            float* bv = c.escr;                                     // [8 cols][8 storing lanes]
            int32_t* bi = reinterpret_cast<int32_t*>(c.escr + 64);
            if ((et & 7u) == 0) {
#pragma unroll
                for (int k = 0; k < kStMaxRows; k++) bv[k * 8 + (et >> 3)] = best.v[k], bi[k * 8 + (et >> 3)] = best.i[k];
            }
            epi_bar();
            if (et < P.rows) {
                float v = -INFINITY;
                int32_t i = 0x7fffffff;
                for (int r = 0; r < 8; r++) {
                    const float ov = bv[et * 8 + r];
                    const int32_t oi = bi[et * 8 + r];
                    if (ov > v || (ov == v && oi < i)) v = ov, i = oi;
                }
                uint64_t* dst = P.am_ll + (size_t(et) * G + blockIdx.x) * 2;
                st_ll_store(dst, __float_as_uint(v), tag_head);
                st_ll_store(dst + 1, uint32_t(i), tag_head);
            }
            epi_bar();
            if (blockIdx.x == 0) {
                for (uint32_t row = ewarp; row < P.rows; row += kStEpiThreads / 32) {
                    float v = -INFINITY;
                    int32_t i = 0x7fffffff;
                    for (unsigned b = lane; b < G; b += 32) {
                        const uint64_t* src = P.am_ll + (size_t(row) * G + b) * 2;
                        const float ov = __uint_as_float(st_poll1(c, src, tag_head, 7));
                        const int32_t oi = int32_t(st_poll1(c, src + 1, tag_head));
                        if (ov > v || (ov == v && oi < i)) v = ov, i = oi;
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        const float ov = __shfl_xor_sync(0xffffffffu, v, off);
                        const int32_t oi = __shfl_xor_sync(0xffffffffu, i, off);
                        if (ov > v || (ov == v && oi < i)) v = ov, i = oi;
                    }
                    if constexpr (TP) {
                        // vocabulary-sharded head: every rank sends its (value, global index) winner to all ranks, then picks the global
                        // winner (lowest index on ties) from the `world` pairs in its own region -- all ranks arrive at the same token
                        if (i != 0x7fffffff) i += int32_t(P.tp_index_base);
                        for (uint32_t k = lane; k < P.tp_world; k += 32) {
                            uint64_t* dst = P.tp_am[k] + (size_t(P.tp_rank) * kStMaxRows + row) * 2;
                            st_ll_store_sys(dst, __float_as_uint(v), tag_head);
                            st_ll_store_sys(dst + 1, uint32_t(i), tag_head);
                        }
                        v = -INFINITY, i = 0x7fffffff;
                        for (uint32_t k = lane; k < P.tp_world; k += 32) {
                            const uint64_t* src = P.tp_am[P.tp_rank] + (size_t(k) * kStMaxRows + row) * 2;
                            const float ov = __uint_as_float(st_poll1_sys(c, src, tag_head, 9));
                            const int32_t oi = int32_t(st_poll1_sys(c, src + 1, tag_head, 9));
                            if (ov > v || (ov == v && oi < i)) v = ov, i = oi;
                        }
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) {
                            const float ov = __shfl_xor_sync(0xffffffffu, v, off);
                            const int32_t oi = __shfl_xor_sync(0xffffffffu, i, off);
                            if (ov > v || (ov == v && oi < i)) v = ov, i = oi;
                        }
                    }
                    if (lane == 0) {
                        if (i == 0x7fffffff) i = 0; // all-NaN / -inf row: argmax keeps index 0
                        st_ll_store(P.ids_ll + row, uint32_t(i), tag_head);
                        P.out_log[size_t(log0 + int32_t(step)) * P.rows + row] = i;
                        if (P.advance && step + 1 == P.steps) {
                            P.ids[row] = i;
                            P.pos[row] = P.pos[row] + int32_t(P.steps);
                        }
                    }
                }
                if (et == 0 && step + 1 == P.steps) P.step_counter[0] = log0 + int32_t(P.steps);
            }
        }
        return;
    }

    // ---- mma / staging / attention warps
    st_pipe cp{0, 0};
    st_blk bl{0, 0};
    for (uint32_t step = 0; step < P.steps; step++) {
        unsigned gphase = step * phases_per_step;
        for (uint32_t li = 0; li <= P.n_layers; li++) {
            const bool is_head = li == P.n_layers;
            for (uint32_t kind = 0; kind < (is_head ? 1u : 5u); kind++, gphase++) {
                // the output of phase k carries tag_base + k + 1; a phase consumes the output of the phase before it
                const uint32_t tag_in = P.tag_base + gphase, tag_out = tag_in + 1;
                unsigned long long* tm = P.timing ? P.timing + (size_t(blockIdx.x) * (P.steps * phases_per_step) + gphase) * 4 : nullptr;
                st_stamp(tm, 0);
                if (!is_head && kind == 1) {
                    st_attention<HD, ATTN_SPLITS>(P, li, step, c, tag_in, tag_out);
                    if (P.arrive_on) {
                        consumer_bar();
                        if (tid == 0) st_arrive(P, gphase);
                    }
                    st_stamp(tm, 2);
                } else {
                    const st_gemv& g = P.g[is_head ? 4u : (kind == 0 ? 0u : kind - 1)];
                    st_stage_input<Q>(P, g, is_head ? 0 : li, c, li == 0 && kind == 0 && !is_head, step, tag_in, tag_out, gphase);
                    st_stamp(tm, 2);
                    st_mma_gemv<Q>(P, g, c, cp, bl, is_head ? 0 : li);
                    if (is_head) {
                        __syncwarp();
                        if ((tid & 31) == 0) mbar_arrive(c.sbar);
                    }
                }
                st_stamp(tm, 3);
            }
        }
    }
}

// (packed layers, head_dim, tensor parallel, one CTA per attention head): each instantiation carries only the code it runs.
// Defined in mc_stream_bf16.cu / mc_stream_quant.cu / mc_stream_tp.cu.
using stream_kernel_fn = void (*)(const st_params);
stream_kernel_fn stream_kernel_bf16(uint32_t head_dim, bool single_cta_attention);
stream_kernel_fn stream_kernel_quant(uint32_t head_dim, bool single_cta_attention);
stream_kernel_fn stream_kernel_tp(uint32_t head_dim, bool single_cta_attention);
stream_kernel_fn stream_kernel_tp_quant(uint32_t head_dim, bool single_cta_attention);

} // namespace mc
