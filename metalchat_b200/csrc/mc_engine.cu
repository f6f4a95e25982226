// metalchat_b200/csrc/mc_engine.cu — the fused Llama-3 decode engine behind mc_llama_* (include/mc_cuda.h).
//
// Replaces the per-token host loop interpreter::read_until -> transformer::transform ->
// nn::llama3::operator() -> sampler (interpreter.h:360-374, transformer.h:357-364,
// nn/llama.h:113-134, nn/sampling.h:69-76) with one CUDA-graph replay per token: weights, KV cache
// and activations stay resident in HBM, the sampled id feeds the next step on the device.
//
// HBM layout (local = this tensor-parallel shard):
//   per layer   wqkv [(Hl+2KVl)*hd, D]   rows: q heads | k heads | v heads      (column-parallel)
//               wo   [D, Hl*hd]                                                   (row-parallel)
//               w13  [2*Fl, D]           row 2i = w1 row i (gate), 2i+1 = w3 row i (column-parallel)
//               w2   [D, Fl]                                                      (row-parallel)
//               attn_norm[D], ffn_norm[D]
//   global      tok [V, D], norm [D], out [Vl, D] (aliases tok when tied, huggingface/llama.h:103)
//   KV cache    [layer][seq][KVl][max_seq][hd] bf16 — one contiguous stream per (seq, kv head)
//   rope tables fcos/fsin [2*max_seq, hd/2] fp32 (nn/embedding.h:160-176)
#include "mc_quant_kernels.cuh"
#include "mc_sample_kernels.cuh"
#include "mc_stream_kernel.cuh"
#include "mc_prefill.h"

#include <dlfcn.h>

#include <cmath>
#include <map>
#include <memory>
#include <mutex>

using namespace mc;

namespace {

struct dbuf {
    void* p = nullptr;
    size_t bytes = 0;
    bool owned = true;
    void view(void* ptr, size_t n) // a slice of an arena owned by someone else
    {
        release();
        p = ptr, bytes = n, owned = false;
    }
    void alloc(size_t n)
    {
        release();
        owned = true;
        bytes = n;
        cudaError_t e = cudaMalloc(&p, n ? n : 1);
        if (e != cudaSuccess) {
            p = nullptr;
            cudaGetLastError();
            throw error(MC_ERR_ALLOC, "cuda: engine allocation of " + std::to_string(n) + " bytes failed: " + cudaGetErrorString(e));
        }
    }
    void release()
    {
        if (p && owned) cudaFree(p);
        p = nullptr;
        bytes = 0;
        owned = true;
    }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

// one linear layer in its device format
struct dlinear {
    uint32_t N = 0, K = 0; // local shape
    int fmt = WF_BF16;
    dbuf w;                // WF_BF16: bf16 [N,K]; WF_W8G/WF_W8ROW: int8 [N,K]; WF_W4: packed [N,K/2]
    dbuf scales;           // fp32 [N,K/group] or [N]
    dbuf lora_b;           // bf16 [N, rank]
    dbuf q8, s32;          // staging until finalize: int8 [N,K] and fp32 [N,K/32] (or [N]) in the reference layout
    dbuf wd;               // quantised models: resident bf16 image r(r(q) * r(s)) for the tensor-core prompt / batch path
    size_t stream_bytes() const { return w.bytes + scales.bytes + lora_b.bytes; }
};

struct dlayer {
    dbuf attn_norm, ffn_norm;
    dlinear wqkv, wo, w13, w2;
    dbuf lora_a_qkv, lora_a_o, lora_a_13, lora_a_2; // stacked bf16 [16*g, K]
};

} // namespace

namespace {
void nccl_comm_release(void* comm);
}
struct mc_llama {
    mc_device* dev = nullptr;
    mc_llama_config cfg{};
    uint32_t Hl = 0, KVl = 0, Fl = 0, Vl = 0; // local (sharded) dims
    bool finalized = false;
    bool tied = true;
    bool sink_roll = false;     // the decode call in flight reaches positions beyond the cache: every step starts with the sink roll
    uint32_t rope_rows = 0;     // rows of the RoPE tables (positions 0 .. rope_rows - 1)
    uint32_t key_begin = 0;     // first visible cache position of the prompt call in flight (MC_LLAMA_REF_CHUNK_MASK: its start_pos)
    bool image_a_dirty = false; // an adaptor A was replaced after the resident bf16 image had been built: its rows are re-copied by finalize
    std::vector<dlayer> layers;
    dbuf layer_arena;          // all per-layer weights, one fixed stride per layer (the streaming kernel indexes by layer)
    size_t layer_stride = 0;
    dbuf bar, errflag;         // (unused counter) / timeout flag of the persistent kernels' bounded waits
    // tensor parallelism: one exchange region per rank (partials | flags | argmax values | argmax flags), IPC-mapped on the peers
    dbuf tp_region, tp_local;  // tp_local: done counter, epoch, argmax epoch
    void* tp_peer_base[kTpMaxWorld] = {};
    bool tp_connected = false;
    // comparator (bench.py --tp-collective nccl): the two all-reduces of a block as ncclAllReduce calls between the per-op kernels
    void* nccl_comm = nullptr;
    uint32_t tp_host_epoch = 0; // row-parallel GEMVs enqueued so far: its parity is the half of the exchange buffer the next one fills
    size_t tp_off_flags = 0, tp_off_amval = 0, tp_off_amidx = 0, tp_off_amflags = 0;
    size_t tp_off_stpart = 0, tp_off_stam = 0, tp_stpart_gen = 0, tp_stam_gen = 0; // streaming kernel: tagged partial sums / argmax pairs, two generations each
    size_t tp_off_stax = 0, tp_stax_gen = 0;                                        // ... and the adaptors' partial A . x (quantised models)
    // tensor-core path under tensor parallelism (bf16 models): fp32 partial sums / bf16 results of a row-parallel GEMM, two halves each
    size_t tp_off_tc_partial = 0, tp_off_tc_result = 0, tp_off_tc_flags = 0, tp_tc_half_partial = 0, tp_tc_half_result = 0;
    uint32_t tp_tc_rows = 0; // rows one exchange holds (0: the path is off)
    uint32_t st_arrive_base[8] = {}; // streaming kernel: value of each arrival counter (m->bar) before the next launch
    // streaming persistent kernel (mc_stream_kernel.cuh): un-rotated q|k|v rows, split-attention exchange, step flag
    dbuf st_ll, st_timing;     // one arena of tagged words: x | h | z | qkv | attn | scores | argmax partials | ids
    size_t st_off[9] = {};
    uint32_t st_sc_words = 0, st_seq = 0;
    uint32_t st_last_pos = 0;  // last position any sequence of the decode call in flight reaches (picks the attention variant)
    size_t st_words = 0;       // words of one copy of the tagged-word arena (the arena holds kStMaxRep copies)
    bool st_timing_on = false;
    int st_ok = -1;            // -1 not probed yet, 0 not usable on this device / shape, 1 usable
    uint32_t st_grid = 0;
    dlinear tok, out;
    dbuf norm;
    dbuf fcos, fsin;
    dbuf kcache, vcache;
    // activations
    uint32_t max_rows = 0;
    dbuf x, h, q, attn, z, logits, logits_tmp, hidden_save;
    dbuf io_in, io_out;        // contiguous decode inputs / outputs (views below)
    // tensor-core prefill (mc_prefill.cu): activations of one prompt chunk, allocated on first use
    dbuf pf_ids, pf_x, pf_h, pf_n, pf_qkv, pf_q, pf_attn, pf_z;
    uint32_t pf_rows = 0;
    dbuf dt_n, dt_qkv;         // batched decode on the tensor cores: normed rows and un-rotated q|k|v rows of one step
    dbuf pf_t, pf_ax, dt_t;    // quantised models: r(x . Wd^T) before the adaptor term, r(A . x)
    dbuf ids, pos, row_seq, uniforms, out_log, step_counter, pval, pidx, lora_ax, pack_bad, cand;
    int32_t* pinned = nullptr; // host staging: ids | pos | out
    float scale_bf16 = 0.0f;
    std::map<uint64_t, cudaGraphExec_t> graphs;
    uint32_t launches_per_step = 0;
    std::vector<const uint16_t*> hidden_ptr; // per seq: where its last hidden row lives

    ~mc_llama()
    {
        for (auto& g : graphs) cudaGraphExecDestroy(g.second);
        if (nccl_comm) nccl_comm_release(nccl_comm);
        if (pinned) cudaFreeHost(pinned);
        for (auto& l : layers) {
            for (dbuf* b : {&l.attn_norm, &l.ffn_norm, &l.lora_a_qkv, &l.lora_a_o, &l.lora_a_13, &l.lora_a_2}) b->release();
            for (dlinear* d : {&l.wqkv, &l.wo, &l.w13, &l.w2}) d->w.release(), d->scales.release(), d->lora_b.release(), d->q8.release(), d->s32.release(), d->wd.release();
        }
        for (dlinear* d : {&tok, &out}) d->w.release(), d->scales.release(), d->lora_b.release(), d->q8.release(), d->s32.release(), d->wd.release();
        for (int k = 0; k < kTpMaxWorld; k++)
            if (tp_peer_base[k] && uint32_t(k) != cfg.tp_rank) cudaIpcCloseMemHandle(tp_peer_base[k]);
        for (dbuf* b : {&st_ll, &st_timing, &layer_arena, &bar, &errflag, &tp_region, &tp_local, &norm, &fcos, &fsin, &kcache, &vcache, &x, &h, &q, &attn, &z, &logits, &logits_tmp, &hidden_save, &io_in, &io_out, &ids, &pos, &row_seq,
                        &uniforms, &out_log, &step_counter, &pval, &pidx, &lora_ax, &pack_bad, &cand, &pf_ids, &pf_x, &pf_h, &pf_n, &pf_qkv, &pf_q, &pf_attn, &pf_z, &dt_n, &dt_qkv, &pf_t, &pf_ax, &dt_t})
            b->release();
    }
};

namespace {

constexpr uint32_t kMaxLogSteps = 4096;

void use(mc_llama* m)
{
    MC_REQUIRE(m != nullptr, "null model");
    MC_CUDA_CHECK(cudaSetDevice(m->dev->ordinal));
}

// RoPE tables on the host with the same libm calls as the scalar reference formula (kernel/rope.metal:93-97: 1/pow(theta, 2j/dim),
// cos/sin of float(pos) * freq).  The reference regenerates its 2 * max_seq_len rows from start_pos whenever a position leaves the
// table (nn/embedding.h:193-198) -- the value of a row depends on the absolute position only, so the engine keeps ONE table indexed by
// absolute position and grows it when decode runs past it.
void build_rope_tables(mc_llama* m, uint32_t rows)
{
    const mc_llama_config& c = m->cfg;
    const uint32_t half = c.head_dim / 2;
    std::vector<float> hc(size_t(rows) * half), hs(size_t(rows) * half);
    for (uint32_t i = 0; i < rows; i++) {
        for (uint32_t j = 0; j < half; j++) {
            const float freq = 1.0f / std::pow(c.rope_theta, 2.0f * float(j) / float(c.head_dim));
            const float angle = float(i) * freq;
            hc[size_t(i) * half + j] = std::cos(angle);
            hs[size_t(i) * half + j] = std::sin(angle);
        }
    }
    m->fcos.alloc(hc.size() * 4);
    m->fsin.alloc(hs.size() * 4);
    MC_CUDA_CHECK(cudaMemcpy(m->fcos.p, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice));
    MC_CUDA_CHECK(cudaMemcpy(m->fsin.p, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice));
    m->rope_rows = rows;
}

size_t kv_layer_elems(const mc_llama* m) { return size_t(m->cfg.n_seqs) * m->KVl * m->cfg.max_seq_len * m->cfg.head_dim; }

// ---- launch helpers -------------------------------------------------------------------------------------------
struct launcher {
    mc_llama* m;
    cudaStream_t s;
    bool pdl;
    uint32_t count = 0;
    std::vector<cudaEvent_t>* events = nullptr; // profiling: one event before every launch (+ one at the end by the caller)
    void mark()
    {
        if (events) {
            cudaEvent_t e;
            MC_CUDA_CHECK(cudaEventCreate(&e));
            MC_CUDA_CHECK(cudaEventRecord(e, s));
            events->push_back(e);
        }
    }
    template <typename... KArgs, typename... Args>
    void go_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, unsigned cluster_x, Args&&... args)
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid;
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cluster_x;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.numAttrs = 1;
        if (pdl) {
            attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.numAttrs = 2;
        }
        cfg.attrs = attr;
        mark();
        MC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
        count++;
        m->dev->launches.fetch_add(1);
    }
    template <typename... KArgs, typename... Args>
    void go(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args)
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid;
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        if (pdl) {
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
        }
        mark();
        MC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
        count++;
        m->dev->launches.fetch_add(1);
    }
};

size_t gemv_smem(uint32_t MB, uint32_t K) { return size_t(MB) * K * 2 + (kGemvWarps * 2 * MB + 8) * sizeof(float); }

// CTAs per SM of a stand-alone GEMV launch.  Measured on B200 (tools/sweep_gemv_ctas.sh): two CTAs per SM everywhere
// (16 warps x 8 x 16-byte loads in flight) beats one CTA per SM + earlier residency of the dependent kernel
// (709 vs 894 us per 1B decode step), so the default is 2; MC_GEMV_CTAS_PER_SM=1 keeps the experiment reachable.
uint32_t gemv_ctas_per_sm(uint32_t, uint32_t)
{
    static const int forced = [] {
        const char* e = getenv("MC_GEMV_CTAS_PER_SM");
        return e ? atoi(e) : 0;
    }();
    return forced == 1 ? 1u : 2u;
}

// k-split so that there is at least ~one unit per resident warp
uint32_t choose_ksplit(const mc_device* dev, uint32_t N, uint32_t K)
{
    const uint32_t units = N / 2;
    const uint32_t warps = uint32_t(dev->prop.multiProcessorCount) * gemv_ctas_per_sm(N, K) * kGemvWarps;
    uint32_t ksplit = 1;
    while (ksplit < 8 && units * ksplit < warps && (K / (ksplit * 2)) % 256 == 0 && K / (ksplit * 2) >= 512) ksplit *= 2;
    return ksplit;
}

template <int MB, int PRO, int EPI, int KS> void gemv_launch_ks(launcher& L, const gemv_params& p)
{
    auto kernel = gemv_bf16_kernel<MB, PRO, EPI, KS>;
    static bool configured[8] = {false};
    const int dev = L.m->dev->ordinal;
    if (!configured[dev & 7]) {
        MC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured[dev & 7] = true;
    }
    const size_t smem = gemv_smem(MB, p.K);
    MC_REQUIRE(smem <= 200 * 1024, "gemv: activation rows do not fit in shared memory");
    const uint32_t upc = kGemvWarps / p.ksplit;
    const uint32_t units = p.N / 2;
    const uint32_t ctas_needed = (units + upc - 1) / upc;
    const uint32_t cap = uint32_t(L.m->dev->prop.multiProcessorCount) * gemv_ctas_per_sm(p.N, p.K);
    const uint32_t grid = ctas_needed < cap ? ctas_needed : cap;
    L.go(kernel, dim3(grid), dim3(kGemvThreads), smem, p);
}
template <int MB, int PRO, int EPI> void gemv_launch_mb(launcher& L, gemv_params p)
{
    MC_REQUIRE(p.N % 2 == 0 && p.K % 256 == 0, "gemv: N must be even and K a multiple of 256");
    p.ksplit = choose_ksplit(L.m->dev, p.N, p.K);
    switch (p.ksplit) {
    case 1: gemv_launch_ks<MB, PRO, EPI, 1>(L, p); break;
    case 2: gemv_launch_ks<MB, PRO, EPI, 2>(L, p); break;
    case 4: gemv_launch_ks<MB, PRO, EPI, 4>(L, p); break;
    default: gemv_launch_ks<MB, PRO, EPI, 8>(L, p); break;
    }
}
template <int PRO, int EPI> void gemv_launch(launcher& L, const gemv_params& p)
{
    MC_REQUIRE(p.rows >= 1 && p.rows <= uint32_t(kMaxMB), "gemv: unsupported number of activation rows");
    switch (p.rows) {
    case 1: gemv_launch_mb<1, PRO, EPI>(L, p); break;
    case 2: gemv_launch_mb<2, PRO, EPI>(L, p); break;
    default: gemv_launch_mb<4, PRO, EPI>(L, p); break;
    }
}

// ---- quantised GEMV launch ---------------------------------------------------------------------------------------
size_t qgemv_smem(uint32_t rows, uint32_t K) { return size_t(rows + 1) * (K + kQPad) * 2 + (kGemvWarps * 32 * 4 + 8) * sizeof(float); }

uint32_t choose_qsplit(const mc_device* dev, uint32_t N, uint32_t K, uint32_t kt)
{
    const uint32_t supers = (N / 2 + 7) / 8, ktiles = K / kt;
    const uint32_t warps = uint32_t(dev->prop.multiProcessorCount) * 2 * kGemvWarps;
    uint32_t ks = 1;
    while (ks < 8 && supers * ks < warps && ktiles % (ks * 2) == 0 && ktiles / (ks * 2) >= 2) ks *= 2;
    return ks;
}
template <int FMT, int PRO, int EPI, int KS> void qgemv_launch_ks(launcher& L, const qgemv_params& q)
{
    auto kernel = gemv_q_kernel<FMT, PRO, EPI, KS>;
    static bool configured[8] = {false};
    const int dev = L.m->dev->ordinal;
    if (!configured[dev & 7]) {
        MC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured[dev & 7] = true;
    }
    const size_t smem = qgemv_smem(q.g.rows, q.g.K);
    MC_REQUIRE(smem <= 200 * 1024, "quantised gemv: activation rows do not fit in shared memory");
    const uint32_t supers = (q.g.N / 2 + 7) / 8, upc = kGemvWarps / KS;
    const uint32_t ctas_needed = (supers + upc - 1) / upc;
    const uint32_t cap = uint32_t(L.m->dev->prop.multiProcessorCount) * 2;
    L.go(kernel, dim3(ctas_needed < cap ? ctas_needed : cap), dim3(kGemvThreads), smem, q);
}
template <int FMT, int PRO, int EPI> void qgemv_launch(launcher& L, const qgemv_params& q)
{
    MC_REQUIRE(q.g.rows >= 1 && q.g.rows <= uint32_t(kQMaxMB), "quantised gemv: unsupported number of activation rows");
    MC_REQUIRE(q.g.N % 2 == 0 && q.g.K % 256 == 0, "quantised gemv: N must be even and K a multiple of 256");
    switch (choose_qsplit(L.m->dev, q.g.N, q.g.K, FMT == WF_W4 ? 64 : 32)) {
    case 1: qgemv_launch_ks<FMT, PRO, EPI, 1>(L, q); break;
    case 2: qgemv_launch_ks<FMT, PRO, EPI, 2>(L, q); break;
    case 4: qgemv_launch_ks<FMT, PRO, EPI, 4>(L, q); break;
    default: qgemv_launch_ks<FMT, PRO, EPI, 8>(L, q); break;
    }
}

// parameter blocks of the five GEMVs of layer `li` and of the head
gemv_params qkv_params(mc_llama* m, uint32_t li, uint32_t row0, uint32_t rows)
{
    const mc_llama_config& c = m->cfg;
    dlayer& ly = m->layers[li];
    gemv_params p{};
    p.W = ly.wqkv.w.p, p.N = ly.wqkv.N, p.K = c.dim, p.rows = rows;
    p.x = m->x.as<uint16_t>() + size_t(row0) * c.dim, p.ldx = c.dim, p.norm_w = ly.attn_norm.as<uint16_t>(), p.eps = c.norm_eps;
    p.q = m->q.as<uint16_t>() + size_t(row0) * m->Hl * c.head_dim;
    p.kcache = m->kcache.as<uint16_t>() + size_t(li) * kv_layer_elems(m);
    p.vcache = m->vcache.as<uint16_t>() + size_t(li) * kv_layer_elems(m);
    p.fcos = m->fcos.as<float>(), p.fsin = m->fsin.as<float>();
    p.row_seq = m->row_seq.as<int32_t>() + row0, p.row_pos = m->pos.as<int32_t>() + row0;
    p.n_heads = m->Hl, p.n_kv_heads = m->KVl, p.head_dim = c.head_dim, p.max_seq = c.max_seq_len;
    p.ksplit = choose_ksplit(m->dev, p.N, p.K);
    return p;
}
attn_params attn_params_of(mc_llama* m, uint32_t li, uint32_t row0)
{
    const mc_llama_config& c = m->cfg;
    attn_params a{};
    a.q = m->q.as<uint16_t>() + size_t(row0) * m->Hl * c.head_dim;
    a.kcache = m->kcache.as<uint16_t>() + size_t(li) * kv_layer_elems(m);
    a.vcache = m->vcache.as<uint16_t>() + size_t(li) * kv_layer_elems(m);
    a.out = m->attn.as<uint16_t>() + size_t(row0) * m->Hl * c.head_dim;
    a.row_seq = m->row_seq.as<int32_t>() + row0, a.row_pos = m->pos.as<int32_t>() + row0;
    a.n_heads = m->Hl, a.n_kv_heads = m->KVl, a.max_seq = c.max_seq_len, a.scale = m->scale_bf16;
    a.key_begin = m->key_begin;
    return a;
}
gemv_params wo_params(mc_llama* m, uint32_t li, uint32_t row0, uint32_t rows)
{
    const uint32_t D = m->cfg.dim, QO = m->Hl * m->cfg.head_dim;
    gemv_params p{};
    p.W = m->layers[li].wo.w.p, p.N = D, p.K = QO, p.rows = rows;
    p.x = m->attn.as<uint16_t>() + size_t(row0) * QO, p.ldx = QO;
    p.y = m->h.as<uint16_t>() + size_t(row0) * D, p.ldy = D, p.res = m->x.as<uint16_t>() + size_t(row0) * D;
    p.ksplit = choose_ksplit(m->dev, p.N, p.K);
    return p;
}
gemv_params w13_params(mc_llama* m, uint32_t li, uint32_t row0, uint32_t rows)
{
    const uint32_t D = m->cfg.dim;
    dlayer& ly = m->layers[li];
    gemv_params p{};
    p.W = ly.w13.w.p, p.N = 2 * m->Fl, p.K = D, p.rows = rows;
    p.x = m->h.as<uint16_t>() + size_t(row0) * D, p.ldx = D, p.norm_w = ly.ffn_norm.as<uint16_t>(), p.eps = m->cfg.norm_eps;
    p.y = m->z.as<uint16_t>() + size_t(row0) * m->Fl, p.ldy = m->Fl;
    p.ksplit = choose_ksplit(m->dev, p.N, p.K);
    return p;
}
gemv_params w2_params(mc_llama* m, uint32_t li, uint32_t row0, uint32_t rows)
{
    const uint32_t D = m->cfg.dim;
    gemv_params p{};
    p.W = m->layers[li].w2.w.p, p.N = D, p.K = m->Fl, p.rows = rows;
    p.x = m->z.as<uint16_t>() + size_t(row0) * m->Fl, p.ldx = m->Fl;
    p.y = m->x.as<uint16_t>() + size_t(row0) * D, p.ldy = D, p.res = m->h.as<uint16_t>() + size_t(row0) * D;
    p.ksplit = choose_ksplit(m->dev, p.N, p.K);
    return p;
}
// ---- NCCL comparator: libnccl.so.2 is resolved at run time (the library does not link against it; the product path never needs it) ----
struct nccl_uid {
    char internal[128];
};
struct nccl_api {
    void* lib = nullptr;
    int (*get_unique_id)(nccl_uid*) = nullptr;
    int (*comm_init_rank)(void**, int, nccl_uid, int) = nullptr;
    int (*all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*comm_destroy)(void*) = nullptr;
    const char* (*error_string)(int) = nullptr;
};
nccl_api& nccl()
{
    static nccl_api api;
    static std::once_flag once;
    std::call_once(once, [] {
        // the copy torch has already loaded when there is one (one NCCL per process), the system's otherwise
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return;
        api.lib = lib;
        api.get_unique_id = reinterpret_cast<decltype(api.get_unique_id)>(dlsym(lib, "ncclGetUniqueId"));
        api.comm_init_rank = reinterpret_cast<decltype(api.comm_init_rank)>(dlsym(lib, "ncclCommInitRank"));
        api.all_reduce = reinterpret_cast<decltype(api.all_reduce)>(dlsym(lib, "ncclAllReduce"));
        api.comm_destroy = reinterpret_cast<decltype(api.comm_destroy)>(dlsym(lib, "ncclCommDestroy"));
        api.error_string = reinterpret_cast<decltype(api.error_string)>(dlsym(lib, "ncclGetErrorString"));
    });
    if (!(api.lib && api.get_unique_id && api.comm_init_rank && api.all_reduce && api.comm_destroy)) throw ::mc::error(MC_ERR_RUNTIME, "libnccl.so.2 could not be loaded");
    return api;
}
void nccl_comm_release(void* comm) { nccl().comm_destroy(comm); }
void nccl_check(int rc, const char* what)
{
    if (rc == 0) return;
    const char* msg = nccl().error_string ? nccl().error_string(rc) : "?";
    throw std::runtime_error(std::string(what) + ": " + msg);
}
// floats per exchange row: the dim main sums, then (quantised models) the adaptor's partial A . x of the row-parallel linear
uint32_t tp_row_floats(const mc_llama* m) { return m->cfg.dim + (m->cfg.quant ? 64u : 0u); }
// the all-reduce of the partial sums the row-parallel GEMV enqueued just before has left in this rank's exchange buffer (in place)
void nccl_all_reduce_partials(mc_llama* m, launcher& L, uint32_t rows)
{
    const uint32_t parity = m->tp_host_epoch & 1u;
    m->tp_host_epoch++;
    float* buf = m->tp_region.as<float>() + size_t(parity) * kMaxMB * tp_row_floats(m);
    L.mark();
    nccl_check(nccl().all_reduce(buf, buf, size_t(rows) * m->cfg.dim, /*ncclFloat32*/ 7, /*ncclSum*/ 0, m->nccl_comm, L.s), "ncclAllReduce");
    L.count++;
}

tp_exchange tp_of(mc_llama* m)
{
    tp_exchange t{};
    t.world = m->cfg.tp_world, t.rank = m->cfg.tp_rank, t.rows_max = kMaxMB, t.dim = tp_row_floats(m);
    if (m->nccl_comm) t.world = 1, t.rank = 0; // the partial sums stay local: NCCL reduces them in place between the two kernels
    for (uint32_t k = 0; k < t.world; k++) {
        void* base = m->nccl_comm ? m->tp_region.p : m->tp_peer_base[k];
        t.peer_buf[k] = static_cast<float*>(base);
        t.peer_flag[k] = reinterpret_cast<uint32_t*>(static_cast<char*>(base) + m->tp_off_flags);
    }
    t.done = m->tp_local.as<unsigned>(), t.epoch = m->tp_local.as<unsigned>() + 1, t.err = m->errflag.as<int>();
    static const uint32_t nowait = getenv("MC_TP_NOWAIT") ? 1u : 0u;
    t.nowait = nowait;
    return t;
}
// makes `p` consume the previous row-parallel GEMV's partial sums: rows = r(res + r(sum)), stored to `out` by CTA 0
void tp_consume(mc_llama* m, gemv_params& p, const uint16_t* res, uint16_t* out)
{
    p.tp = tp_of(m), p.tp_reduce = 1, p.tp_res = res, p.tp_out = out;
}

gemv_params head_params(mc_llama* m, const uint16_t* x, uint32_t rows, uint16_t* logits_dst)
{
    const dlinear& hw = m->tied ? m->tok : m->out;
    gemv_params p{};
    // vocabulary-sharded head: this rank projects rows [rank*Vl, (rank+1)*Vl) of the (tied) table
    p.W = m->tied ? static_cast<const void*>(hw.w.as<uint16_t>() + size_t(m->cfg.tp_rank) * m->Vl * m->cfg.dim) : hw.w.p;
    p.N = m->Vl, p.K = m->cfg.dim, p.rows = rows;
    p.norm_w = m->norm.as<uint16_t>(), p.eps = m->cfg.norm_eps;
    p.x = x, p.ldx = m->cfg.dim, p.y = logits_dst, p.ldy = m->Vl;
    p.ksplit = choose_ksplit(m->dev, p.N, p.K);
    return p;
}

size_t attn_smem(const mc_llama* m, uint32_t cluster)
{
    const uint32_t hd = m->cfg.head_dim, slots = 256 / (hd / 8);
    const uint32_t chunk = ((m->cfg.max_seq_len + cluster * slots - 1) / (cluster * slots)) * slots;
    return (hd + 8 + 4 + hd + size_t(slots) * hd + chunk) * sizeof(float);
}

// QLoRA blocks (quantization/lora.h:94-122): every linear is  y = r(x . dq(W)^T) + r(2.0 * B(A(x)))  with
// ax = r(A . x) from a small bf16 GEMV over the stacked adaptors of the linears that share the input.
void enqueue_rows_quant(mc_llama* m, launcher& L, uint32_t row0, uint32_t rows, int head_mode, uint16_t* logits_dst)
{
    const mc_llama_config& c = m->cfg;
    const uint32_t D = c.dim, hd = c.head_dim, rank = c.lora_rank;
    const uint32_t ax_ld = 3 * rank;
    uint16_t* ax = m->lora_ax.as<uint16_t>() + size_t(row0) * ax_ld;
    const float lscale = bf16_bits_to_f32(f32_to_bf16_bits(c.lora_scale));
    auto lora_a = [&](const gemv_params& main, const dbuf& A, uint32_t n_rows_a, bool norm) {
        gemv_params p{};
        p.W = A.p, p.N = n_rows_a, p.K = main.K, p.rows = rows;
        p.x = main.x, p.ldx = main.ldx, p.norm_w = main.norm_w, p.eps = main.eps;
        p.y = ax, p.ldy = ax_ld;
        if (norm) gemv_launch<PRO_RMSNORM, EPI_NONE>(L, p);
        else gemv_launch<PRO_NONE, EPI_NONE>(L, p);
    };
    auto quant = [&](const gemv_params& g, const dlinear& d, uint32_t slices) {
        qgemv_params q{};
        q.g = g;
        q.scales = d.scales.p, q.lora_b = d.lora_b.as<uint16_t>(), q.lora_ax = ax, q.ax_ld = ax_ld, q.ax_slices = slices;
        q.slice_rows0 = m->Hl * hd, q.slice_rows1 = (m->Hl + m->KVl) * hd;
        q.rank = rank, q.lora_scale = lscale;
        return q;
    };
    const bool tp = c.tp_world > 1;
    if (tp) MC_REQUIRE(m->tp_connected, "tensor parallel model: mc_llama_tp_connect has not been called");
    // Row-parallel linear (wo / w2) under tensor parallelism: this rank's k range of the main sums AND of the adaptor's A . x go into
    // one exchange row [dim | rank] as unrounded fp32 (the adaptor GEMV fills its columns, the main GEMV publishes the row to the
    // peers); tp_finish_lora_kernel then sums the ranks and applies the roundings, the adaptor term and the residual for all rows.
    auto row_parallel = [&](const gemv_params& g, const dlinear& d, const dbuf& A) {
        gemv_params pa{};
        pa.W = A.p, pa.N = rank, pa.K = g.K, pa.rows = rows, pa.x = g.x, pa.ldx = g.ldx;
        pa.tp = tp_of(m), pa.tp_col0 = D, pa.tp_hold = 1;
        gemv_launch<PRO_NONE, EPI_PARTIAL_TP>(L, pa);
        qgemv_params q{};
        q.g = g, q.g.tp = tp_of(m);
        q.scales = d.scales.p, q.rank = rank, q.lora_scale = lscale;
        qgemv_launch<WF_W4, PRO_NONE, EPI_PARTIAL_TP>(L, q);
        tp_finish_params f{};
        f.tp = tp_of(m), f.res = g.res, f.out = g.y, f.lora_b = d.lora_b.as<uint16_t>();
        f.rows = rows, f.dim = D, f.ld = g.ldy, f.rank = rank, f.lora_scale = lscale;
        L.go(tp_finish_lora_kernel, dim3((D + 255) / 256), dim3(256), 0, f);
    };
    for (uint32_t li = 0; li < c.n_layers; li++) {
        dlayer& ly = m->layers[li];
        const gemv_params gq = qkv_params(m, li, row0, rows);
        lora_a(gq, ly.lora_a_qkv, 3 * rank, true);
        qgemv_launch<WF_W4, PRO_RMSNORM, EPI_QKV>(L, quant(gq, ly.wqkv, 3));
        const attn_params a = attn_params_of(m, li, row0);
        const size_t smem = attn_smem(m, kAttnCluster);
        if (hd == 64) L.go_cluster(attn_decode_kernel<64>, dim3(m->Hl * kAttnCluster, rows), dim3(256), smem, kAttnCluster, a);
        else L.go_cluster(attn_decode_kernel<128>, dim3(m->Hl * kAttnCluster, rows), dim3(256), smem, kAttnCluster, a);
        const gemv_params go = wo_params(m, li, row0, rows);
        if (tp) row_parallel(go, ly.wo, ly.lora_a_o);
        else {
            lora_a(go, ly.lora_a_o, rank, false);
            qgemv_launch<WF_W4, PRO_NONE, EPI_RESIDUAL>(L, quant(go, ly.wo, 1));
        }
        const gemv_params g13 = w13_params(m, li, row0, rows);
        lora_a(g13, ly.lora_a_13, 2 * rank, true);
        qgemv_launch<WF_W4, PRO_RMSNORM, EPI_SWIGLU>(L, quant(g13, ly.w13, 2));
        const gemv_params g2 = w2_params(m, li, row0, rows);
        if (tp) row_parallel(g2, ly.w2, ly.lora_a_2);
        else {
            lora_a(g2, ly.lora_a_2, rank, false);
            qgemv_launch<WF_W4, PRO_NONE, EPI_RESIDUAL>(L, quant(g2, ly.w2, 1));
        }
    }
    if (head_mode) {
        uint16_t* x = m->x.as<uint16_t>() + size_t(row0) * D;
        qgemv_params q{};
        q.g = head_mode == 2 ? head_params(m, x + size_t(rows - 1) * D, 1, logits_dst) : head_params(m, x, rows, logits_dst);
        q.scales = m->out.scales.p; // output is a quantization::linear: int8 with one scale per row, no adaptor (quantization/linear.h:17-64)
        qgemv_launch<WF_W8ROW, PRO_RMSNORM, EPI_NONE>(L, q);
    }
}

// Enqueue the forward pass of `rows` activation rows (<= kMaxMB) starting at row offset `row0` of the
// activation buffers, one kernel per fused op.  Row r reads ids[row0+r], pos[row0+r], row_seq[row0+r].
//   head_mode: 0 none, 1 logits for every row, 2 last row only
void enqueue_rows(mc_llama* m, launcher& L, uint32_t row0, uint32_t rows, int head_mode, uint16_t* logits_dst)
{
    const mc_llama_config& c = m->cfg;
    const uint32_t D = c.dim, hd = c.head_dim;
    uint16_t* x = m->x.as<uint16_t>() + size_t(row0) * D;
    const int32_t* ids = m->ids.as<int32_t>() + row0;

    L.go(embed_kernel, dim3(rows), dim3(256), 0, x, D, (const void*)m->tok.w.p, (const float*)m->tok.scales.p, m->tok.fmt, D, c.vocab, ids);
    if (c.quant) {
        enqueue_rows_quant(m, L, row0, rows, head_mode, logits_dst);
        return;
    }
    const bool tp = c.tp_world > 1;
    if (tp) MC_REQUIRE(m->tp_connected, "tensor parallel model: mc_llama_tp_connect has not been called");
    uint16_t* h = m->h.as<uint16_t>() + size_t(row0) * D;
    for (uint32_t li = 0; li < c.n_layers; li++) {
        gemv_params gq = qkv_params(m, li, row0, rows);
        if (tp && li > 0) tp_consume(m, gq, h, x); // x = r(h + all-reduced w2 output of the previous block)
        gemv_launch<PRO_RMSNORM, EPI_QKV>(L, gq);
        const attn_params a = attn_params_of(m, li, row0);
        const size_t smem = attn_smem(m, kAttnCluster);
        if (hd == 64) L.go_cluster(attn_decode_kernel<64>, dim3(m->Hl * kAttnCluster, rows), dim3(256), smem, kAttnCluster, a);
        else L.go_cluster(attn_decode_kernel<128>, dim3(m->Hl * kAttnCluster, rows), dim3(256), smem, kAttnCluster, a);
        gemv_params go = wo_params(m, li, row0, rows), g13 = w13_params(m, li, row0, rows), g2 = w2_params(m, li, row0, rows);
        if (!tp) {
            gemv_launch<PRO_NONE, EPI_RESIDUAL>(L, go);
            gemv_launch<PRO_RMSNORM, EPI_SWIGLU>(L, g13);
            gemv_launch<PRO_NONE, EPI_RESIDUAL>(L, g2);
        } else {
            // row-parallel wo / w2 push fp32 partials to every rank; the next GEMV's prologue finishes the all-reduce
            go.tp = tp_of(m), g2.tp = tp_of(m);
            gemv_launch<PRO_NONE, EPI_PARTIAL_TP>(L, go);
            if (m->nccl_comm) nccl_all_reduce_partials(m, L, rows);
            tp_consume(m, g13, x, h); // h = r(x + all-reduced wo output)
            gemv_launch<PRO_RMSNORM, EPI_SWIGLU>(L, g13);
            gemv_launch<PRO_NONE, EPI_PARTIAL_TP>(L, g2);
            if (m->nccl_comm) nccl_all_reduce_partials(m, L, rows);
        }
    }
    if (tp) {
        // the head's prologue completes the last block's all-reduce; without logits the reduction is not needed at all
        if (head_mode == 0) return;
        uint16_t* dst = head_mode == 2 ? m->logits_tmp.as<uint16_t>() : logits_dst;
        gemv_params gh = head_params(m, x, rows, dst);
        tp_consume(m, gh, h, x);
        gemv_launch<PRO_RMSNORM, EPI_NONE>(L, gh);
        if (head_mode == 2)
            MC_CUDA_CHECK(cudaMemcpyAsync(logits_dst, dst + size_t(rows - 1) * m->Vl, size_t(m->Vl) * 2, cudaMemcpyDeviceToDevice, L.s));
        return;
    }
    if (head_mode == 2) gemv_launch<PRO_RMSNORM, EPI_NONE>(L, head_params(m, x + size_t(rows - 1) * D, 1, logits_dst));
    else if (head_mode == 1) gemv_launch<PRO_RMSNORM, EPI_NONE>(L, head_params(m, x, rows, logits_dst));
}

// ---- the streaming persistent kernel (mc_stream_kernel.cuh) ---------------------------------------------------------------
// k width of the bf16 ring tiles: 1024 (the last tile of a block is shorter when K is not a multiple; K must be a multiple of 256 =
// 8 mma warps x 32 k), or the whole K when it is shorter than that
uint32_t stream_kc(uint32_t K)
{
    if (K == 0 || K % 256 != 0) return 0;
    return std::min<uint32_t>(K, 1024u);
}
// the kernel is specialised on (packed layers + adaptors, head_dim): each instantiation carries only the code it runs
using stream_kernel_t = void (*)(const st_params);
// the instantiations live in three translation units of their own (mc_stream_{bf16,quant,tp}.cu: they compile in parallel)
stream_kernel_t stream_kernel_of(bool quant, uint32_t head_dim, bool tp, bool single)
{
    if (tp) return quant ? mc::stream_kernel_tp_quant(head_dim, single) : mc::stream_kernel_tp(head_dim, single);
    return quant ? mc::stream_kernel_quant(head_dim, single) : mc::stream_kernel_bf16(head_dim, single);
}
struct stream_geom {
    uint32_t act_pitch, act_bytes, sax_off, n_stages, stage_bytes;
    size_t smem;
};
bool stream_geometry(const mc_llama* m, uint32_t rows, stream_geom& g)
{
    const mc_llama_config& c = m->cfg;
    const uint32_t kmax = std::max(std::max(c.dim, m->Hl * c.head_dim), m->Fl);
    // row pad: 64 bytes make the 16-byte fragment loads of the bf16 path conflict-free, 16 bytes the 4-byte loads of the packed paths
    g.act_pitch = kmax * 2 + (c.quant ? 16 : kStPad);
    const size_t attn_scratch = (size_t(5) * c.head_dim + 2048 + c.max_seq_len + 8) * sizeof(float); // q, k, q', k', v | per-slot partial outputs | scores
    g.sax_off = uint32_t((std::max(size_t(rows) * g.act_pitch, attn_scratch) + 127) & ~size_t(127));
    g.act_bytes = g.sax_off + (c.quant ? uint32_t(kStMaxRows * 3 * c.lora_rank * sizeof(float) + 127) & ~127u : 0u);
    const size_t fixed = kStHdrBytes + kStRedBytes + (c.tp_world > 1 ? kStTpKeepBytes : 0) + g.act_bytes;
    const size_t cap = 232448; // 227 KiB of dynamic shared memory per CTA on sm_100
    g.stage_bytes = c.quant ? kStStageBytesPacked : kStStageBytes;
    if (fixed + 2 * size_t(g.stage_bytes) > cap) return false;
    g.n_stages = uint32_t(std::min<size_t>(kStMaxStages, (cap - fixed) / g.stage_bytes));
    g.smem = fixed + size_t(g.n_stages) * g.stage_bytes;
    return true;
}
uint32_t stream_kc_packed(uint32_t K, uint32_t kmax)
{
    for (uint32_t kc : {2048u, 1024u, 512u, 256u})
        if (kc > kmax) continue;
        else
        if (K % kc == 0) return kc;
    return 0;
}
// the streaming kernel serves greedy bf16 decode of up to 8 sequences on one GPU; everything else takes the per-op path
bool stream_eligible(mc_llama* m, uint32_t n, const mc_sampler_config& sc)
{
    const mc_llama_config& c = m->cfg;
    static const bool env_off = getenv("MC_NO_STREAM") != nullptr;
    if (env_off || (c.flags & MC_LLAMA_NO_STREAM)) return false;
    // Measured on B200 (1B bf16, KV 512): 1 sequence 1595 vs 1431 tokens/s (streaming vs per-op), 2: 2500 vs 2552, 4: 3137 vs 3855,
    // 8: 3671 vs 3907 - the streaming kernel walks attention items, staged rows and epilogue columns one after the other, so
    // by default it serves single-sequence decode; MC_STREAM_MAX_ROWS raises the limit (the kernel itself handles up to 8).
    static const uint32_t max_rows = getenv("MC_STREAM_MAX_ROWS") ? uint32_t(atoi(getenv("MC_STREAM_MAX_ROWS"))) : 1u;
    if (sc.mode != 0 || n > std::min<uint32_t>(max_rows, kStMaxRows) || m->sink_roll) return false;
    // tensor parallel: bf16 models whose row-parallel phases fit the per-CTA partial-sum store; MC_TP_NO_STREAM keeps the per-op exchange
    static const bool tp_stream_off = getenv("MC_TP_NO_STREAM") != nullptr;
    if (c.tp_world != 1 && (!m->tp_connected || tp_stream_off || m->nccl_comm || c.tp_world > uint32_t(kStTpMaxWorld))) return false;
    if (m->st_ok < 0) {
        m->st_ok = 0;
        stream_geom g;
        // (geometry for the largest row count this model will be served with: 8 rows of an 8B / 70B-shard ffn vector do not fit next to the ring)
        bool shapes = stream_kc(c.dim) && stream_kc(m->Hl * c.head_dim) && stream_kc(m->Fl) && m->Fl % 2 == 0 && m->Vl % 2 == 0 &&
                      m->dev->prop.multiProcessorCount <= 256 && stream_geometry(m, std::min<uint32_t>(std::min<uint32_t>(max_rows, kStMaxRows), m->max_rows), g);
        // tensor parallel: a CTA keeps the partial sums of at most kStTpBlocks 16-row blocks of a row-parallel phase
        if (c.tp_world > 1) shapes = shapes && (c.dim + m->dev->prop.multiProcessorCount - 1) / m->dev->prop.multiProcessorCount + 16 <= uint32_t(kStTpBlocks) * 16;
        if (c.quant) // packed layouts: whole super-units of 16 rows, adaptor rows in pairs, rank in 16-byte steps
            shapes = shapes && ((m->Hl + 2 * m->KVl) * c.head_dim) % 16 == 0 && c.dim % 16 == 0 && (2 * m->Fl) % 16 == 0 && m->Vl % 16 == 0 && c.lora_rank % 8 == 0 &&
                     3 * c.lora_rank <= 128 && (c.head_dim / 2) % 8 == 0;
        if (shapes) {
            bool ok = true;
            int occ = 0, coop = 0;
            cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, m->dev->ordinal);
            for (bool single : {false, true}) {
                auto kernel = stream_kernel_of(c.quant != 0, c.head_dim, c.tp_world > 1, single);
                ok = ok && cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448) == cudaSuccess && coop &&
                     cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kStThreads, g.smem) == cudaSuccess && occ >= 1;
            }
            if (ok) {
                {
                    m->st_ok = 1;
                    m->st_grid = uint32_t(m->dev->prop.multiProcessorCount);
                }
            }
            cudaGetLastError();
        }
    }
    return m->st_ok == 1;
}
void launch_stream(mc_llama* m, launcher& L, uint32_t rows, int advance, uint32_t steps)
{
    const mc_llama_config& c = m->cfg;
    const uint32_t D = c.dim, hd = c.head_dim, QO = m->Hl * hd, QKVN = (m->Hl + 2 * m->KVl) * hd;
    const uint32_t phases = c.n_layers * 5 + 1;
    MC_REQUIRE(steps == 1 || advance, "stream kernel: several steps per launch need the sampled id fed back");
    MC_REQUIRE(uint64_t(steps) * phases + 2 < 65536, "stream kernel: too many steps for one launch");
    stream_geom geo;
    MC_REQUIRE(stream_geometry(m, rows, geo), "stream kernel: activation rows do not fit in shared memory");
    // tags of a launch are (sequence number << 16) + phase + 1: never 0, unique until the 16-bit sequence wraps
    if ((++m->st_seq & 0xffffu) == 0) {
        MC_CUDA_CHECK(cudaMemsetAsync(m->st_ll.p, 0, m->st_ll.bytes, L.s));
        if (c.tp_world > 1) {
            // the generation just left (see mc_llama_create): everything the peers sent into it has been consumed by this rank
            const size_t old_gen = ((m->st_seq - 1) >> 16) & 1u;
            MC_CUDA_CHECK(cudaMemsetAsync(m->tp_region.as<char>() + m->tp_off_stpart + old_gen * m->tp_stpart_gen, 0, m->tp_stpart_gen, L.s));
            MC_CUDA_CHECK(cudaMemsetAsync(m->tp_region.as<char>() + m->tp_off_stam + old_gen * m->tp_stam_gen, 0, m->tp_stam_gen, L.s));
            MC_CUDA_CHECK(cudaMemsetAsync(m->tp_region.as<char>() + m->tp_off_stax + old_gen * m->tp_stax_gen, 0, m->tp_stax_gen, L.s));
        }
        ++m->st_seq;
    }
    st_params P{};
    const dlayer& l0 = m->layers[0];
    uint64_t* ll = m->st_ll.as<uint64_t>();
    uint64_t *x_ll = ll + m->st_off[0], *h_ll = ll + m->st_off[1], *z_ll = ll + m->st_off[2], *qkv_ll = ll + m->st_off[3], *attn_ll = ll + m->st_off[4];
    uint64_t* ax_ll = ll + m->st_off[8];
    const bool Q = c.quant != 0;
    const uint32_t rank = c.lora_rank;
    auto gemv = [&](const dlinear& d, const dbuf* lora_a, uint32_t slices, int which, const dbuf* norm, const uint64_t* in_ll, uint64_t* out_ll,
                    const uint64_t* res_ll, uint32_t N, uint32_t K, int pro, int epi, int layered) {
        st_gemv g{};
        g.W = d.w.p, g.norm_w = norm ? norm->as<uint16_t>() : nullptr, g.in_ll = in_ll, g.out_ll = out_ll, g.res_ll = res_ll;
        g.N = N, g.K = K, g.fmt = d.fmt, g.pro = pro, g.epi = epi, g.layered = layered;
        if (d.fmt == WF_BF16) g.KC = stream_kc(K), g.gran = epi == EPI_SWIGLU ? 4 : 2;
        else g.KC = stream_kc_packed(K, d.fmt == WF_W4 ? 2048 : 1024), g.gran = 16, g.scales = d.scales.p; // 16 KiB tiles
        static const bool diag_no_lora = getenv("MC_STREAM_DIAG_NO_LORA") != nullptr; // diagnostics: what the adaptor path costs (results are wrong)
        if (lora_a && lora_a->p && !diag_no_lora) {
            g.lora_a = lora_a->as<uint16_t>(), g.lora_b = d.lora_b.as<uint16_t>(), g.n_a = slices * rank, g.ax_slices = slices;
            g.slice_rows0 = m->Hl * hd, g.slice_rows1 = (m->Hl + m->KVl) * hd;
            g.ax_ll = ax_ll + size_t(which) * kStMaxRows * 128;
        }
        return g;
    };
    const dlinear& head = m->tied ? m->tok : m->out;
    P.g[0] = gemv(l0.wqkv, &l0.lora_a_qkv, 3, 0, &l0.attn_norm, x_ll, qkv_ll, nullptr, QKVN, D, PRO_RMSNORM, EPI_NONE, 1);
    P.g[0].qkv_map = Q; // the int4 q|k|v rows were packed as rope pairs (mc_llama_finalize)
    P.g[1] = gemv(l0.wo, &l0.lora_a_o, 1, 1, nullptr, attn_ll, h_ll, x_ll, D, QO, PRO_NONE, EPI_RESIDUAL, 1);
    P.g[2] = gemv(l0.w13, &l0.lora_a_13, 2, 2, &l0.ffn_norm, h_ll, z_ll, nullptr, 2 * m->Fl, D, PRO_RMSNORM, EPI_SWIGLU, 1);
    P.g[3] = gemv(l0.w2, &l0.lora_a_2, 1, 3, nullptr, z_ll, x_ll, h_ll, D, m->Fl, PRO_NONE, EPI_RESIDUAL, 1);
    P.g[4] = gemv(head, nullptr, 0, 0, &m->norm, x_ll, nullptr, nullptr, m->Vl, D, PRO_RMSNORM, EPI_NONE, 0);
    if (m->tied) P.g[4].W = m->tok.w.as<uint16_t>() + size_t(c.tp_rank) * m->Vl * D;
    P.g[4].y = m->logits.as<uint16_t>();
    // LoRA-A rows travel two at a time: the largest k chunk whose two padded rows fit one stage
    for (int i = 0; i < 4; i++) {
        P.kc_a[i] = 256;
        for (uint32_t kc : {4096u, 2048u, 1024u, 512u})
            if (P.g[i].K % kc == 0 && 2 * (kc * 2 + kStPad) <= geo.stage_bytes) {
                P.kc_a[i] = kc;
                break;
            }
    }
    P.lora_rank = rank, P.lora_scale = bf16_bits_to_f32(f32_to_bf16_bits(c.lora_scale));
    P.tok_fmt = m->tok.fmt, P.tok_scales = m->tok.scales.as<float>();
    P.layer_stride = m->layer_stride, P.kv_layer_stride = kv_layer_elems(m);
    P.n_layers = c.n_layers, P.rows = rows, P.steps = steps, P.n_stages = geo.n_stages, P.stage_bytes = geo.stage_bytes, P.act_pitch = geo.act_pitch, P.act_bytes = geo.act_bytes, P.sax_off = geo.sax_off;
    P.tag_base = m->st_seq << 16;
    static const int env_rep = getenv("MC_STREAM_REP") ? atoi(getenv("MC_STREAM_REP")) : 1;
    P.n_rep = uint32_t(std::min(std::max(env_rep, 1), kStMaxRep)), P.rep_stride = uint32_t(m->st_words);
    static const int env_ns = getenv("MC_STREAM_POLL_NS") ? atoi(getenv("MC_STREAM_POLL_NS")) : 0;
    P.poll_ns = uint32_t(env_ns);
    P.eps = c.norm_eps;
    P.qkv_ll = qkv_ll, P.sc_ll = ll + m->st_off[5], P.attn_ll = attn_ll, P.sc_words = m->st_sc_words;
    P.kcache = m->kcache.as<uint16_t>(), P.vcache = m->vcache.as<uint16_t>(), P.fcos = m->fcos.as<float>(), P.fsin = m->fsin.as<float>();
    P.row_seq = m->row_seq.as<int32_t>(), P.pos = m->pos.as<int32_t>(), P.ids = m->ids.as<int32_t>();
    P.n_heads = m->Hl, P.n_kv_heads = m->KVl, P.head_dim = hd, P.max_seq = c.max_seq_len, P.vocab = c.vocab, P.scale = m->scale_bf16;
    P.embed_table = m->tok.w.p;
    P.am_ll = ll + m->st_off[6], P.ids_ll = ll + m->st_off[7];
    P.out_log = m->out_log.as<int32_t>(), P.step_counter = m->step_counter.as<int32_t>(), P.advance = advance;
    P.err = m->errflag.as<int>();
    P.tp_world = c.tp_world, P.tp_rank = c.tp_rank, P.tp_dim = D, P.tp_index_base = c.tp_rank * m->Vl;
    // one CTA per attention head while every history of the launch stays short (the host knows the last position of the launch)
    static const int env_single = getenv("MC_ATTN_SINGLE_MAX") ? atoi(getenv("MC_ATTN_SINGLE_MAX")) : int(kStAttnSingleMax);
    const bool attn_single = m->st_last_pos < uint32_t(std::max(env_single, 0));
    if (c.tp_world > 1) {
        const size_t gen = (m->st_seq >> 16) & 1u;
        for (uint32_t k = 0; k < c.tp_world; k++) {
            char* base = static_cast<char*>(m->tp_peer_base[k]);
            P.tp_part[k] = reinterpret_cast<uint64_t*>(base + m->tp_off_stpart + gen * m->tp_stpart_gen);
            P.tp_am[k] = reinterpret_cast<uint64_t*>(base + m->tp_off_stam + gen * m->tp_stam_gen);
            P.tp_ax[k] = reinterpret_cast<uint64_t*>(base + m->tp_off_stax + gen * m->tp_stax_gen);
        }
    }
    // arrival counters (opt-in experiment, MC_STREAM_ARRIVE=1): every CTA adds 1 to counter (gphase & 7) per phase and the staging of the next
    // phase waits for the count before its first load; the host keeps what the counters hold between launches.  Measured on B200 (1B, one box):
    // bf16 0.693 ms with vs 0.649 ms without, int4 0.833 vs 0.779 -- in the real step the producers arrive staggered, and counter -> barrier ->
    // load is a longer chain than polling the words themselves, although the synthetic exchange of tools/hop_floor.cu favours the counter.
    static const bool arrive_on = getenv("MC_STREAM_ARRIVE") != nullptr;
    P.arrive = m->bar.as<unsigned>(), P.arrive_on = arrive_on ? 1u : 0u;
    for (int k = 0; k < 8; k++) P.arrive_base[k] = m->st_arrive_base[k];
    if (P.arrive_on)
        for (uint64_t gp = 0, total = uint64_t(steps) * phases; gp < total; gp++) m->st_arrive_base[gp & 7u] += m->st_grid;
    P.timing = m->st_timing_on ? m->st_timing.as<unsigned long long>() : nullptr;
    P.dbg = m->st_timing_on ? m->st_timing.as<unsigned long long>() + size_t(m->st_grid) * (c.n_layers * 5 + 1) * 4 : nullptr;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(m->st_grid), cfg.blockDim = dim3(kStThreads), cfg.dynamicSmemBytes = geo.smem, cfg.stream = L.s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative; // the CTAs poll each other's output: all of them must be co-resident
    attr[0].val.cooperative = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    L.mark();
    MC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, stream_kernel_of(Q, hd, c.tp_world > 1, attn_single), P));
    L.count++;
    m->dev->launches.fetch_add(1);
}

// launches the two sampling kernels over logits [rows, vocab] (bf16, row pitch ld)
void launch_sampler(launcher& L, sample_params sp, const mc_sampler_config& sc, uint32_t rows, uint32_t vocab, unsigned long long* cand)
{
    MC_REQUIRE(sc.top_k >= 1 && sc.top_k <= uint32_t(kSampleMaxK), "sampler: top_k must be in [1, 64]");
    MC_REQUIRE(sc.top_k <= vocab, "sampler: top_k exceeds the vocabulary");
    MC_REQUIRE(sc.temperature > 0.0f, "sampler: temperature must be positive");
    const uint32_t blocks = (vocab + kSampleSlice - 1) / kSampleSlice;
    L.go(sample_select_kernel, dim3(blocks, rows), dim3(256), 0, sp.logits, sp.ld, vocab, sp.index_base, cand);
    sp.cand = cand, sp.n_cand = blocks * kSampleKeep;
    sp.sort_n = 64;
    while (sp.sort_n < sp.n_cand) sp.sort_n <<= 1;
    sp.top_k = sc.top_k, sp.intended = sc.intended, sp.rows = rows;
    const float temp_t = bf16_bits_to_f32(f32_to_bf16_bits(sc.temperature));       // T(temperature)
    sp.inv_t = bf16_bits_to_f32(f32_to_bf16_bits(1.0f / temp_t));                   // T(1 / T(temperature))
    sp.top_p = bf16_bits_to_f32(f32_to_bf16_bits(sc.top_p));
    if (!L.m->dev->sampler_configured) { // function attributes are per device
        MC_CUDA_CHECK(cudaFuncSetAttribute(sample_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        L.m->dev->sampler_configured = true;
    }
    MC_REQUIRE(size_t(sp.sort_n) * 8 <= 160 * 1024, "sampler: vocabulary shard too large for the candidate sort");
    L.go(sample_finish_kernel, dim3(rows), dim3(1024), size_t(sp.sort_n) * 8, sp);
}

void enqueue_sample(mc_llama* m, launcher& L, uint32_t rows, const mc_sampler_config& sc, int advance)
{
    if (sc.mode == 1) {
        MC_REQUIRE(m->cfg.tp_world == 1, "the multinomial sampler is not available under tensor parallelism (greedy only)");
        // make_default_sampler (nn/sampling.h:306-316): top-k -> nucleus -> multinomial, all on the device
        sample_params sp{};
        sp.logits = m->logits.as<uint16_t>(), sp.ld = m->Vl, sp.index_base = 0;
        sp.uniforms = m->uniforms.as<float>();
        sp.ids = m->ids.as<int32_t>(), sp.pos = m->pos.as<int32_t>(), sp.out_log = m->out_log.as<int32_t>();
        sp.step_counter = m->step_counter.as<int32_t>(), sp.advance = advance;
        launch_sampler(L, sp, sc, rows, m->Vl, m->cand.as<unsigned long long>());
        L.go(sample_step_kernel, dim3(1), dim3(32), 0, m->step_counter.as<int32_t>());
        return;
    }
    MC_REQUIRE(sc.mode == 0, "unknown sampler mode");
    if (m->cfg.tp_world > 1) {
        L.go(argmax_partial_kernel, dim3(kArgmaxBlocks, rows), dim3(256), 0, (const uint16_t*)m->logits.p, m->Vl, m->Vl, m->pval.as<float>(),
             m->pidx.as<int32_t>());
        am_exchange x{};
        x.world = m->cfg.tp_world, x.rank = m->cfg.tp_rank, x.rows_max = m->max_rows; // every decode row has its own slot (was kMaxMB: rows >= 4 aliased the next rank's slots)
        for (uint32_t k = 0; k < x.world; k++) {
            char* base = static_cast<char*>(m->tp_peer_base[k]);
            x.peer_val[k] = reinterpret_cast<float*>(base + m->tp_off_amval);
            x.peer_idx[k] = reinterpret_cast<int32_t*>(base + m->tp_off_amidx);
            x.peer_flag[k] = reinterpret_cast<uint32_t*>(base + m->tp_off_amflags);
        }
        x.epoch = m->tp_local.as<unsigned>() + 2, x.err = m->errflag.as<int>();
        L.go(argmax_final_tp_kernel, dim3(1), dim3(((rows + 31) / 32) * 32), 0, (const float*)m->pval.p, (const int32_t*)m->pidx.p, kArgmaxBlocks,
             int32_t(m->cfg.tp_rank * m->Vl), x, m->ids.as<int32_t>(), m->pos.as<int32_t>(), m->out_log.as<int32_t>(), m->step_counter.as<int32_t>(), rows,
             advance);
        return;
    }
    L.go(argmax_partial_kernel, dim3(kArgmaxBlocks, rows), dim3(256), 0, (const uint16_t*)m->logits.p, m->Vl, m->Vl,
         m->pval.as<float>(), m->pidx.as<int32_t>());
    L.go(argmax_final_kernel, dim3(1), dim3(((rows + 31) / 32) * 32), 0, (const float*)m->pval.p, (const int32_t*)m->pidx.p,
         kArgmaxBlocks, m->ids.as<int32_t>(), m->pos.as<int32_t>(), m->out_log.as<int32_t>(), m->step_counter.as<int32_t>(), rows, advance);
}

// Tensor parallel on the tensor-core path (bf16 models): a row-parallel linear (wo, w2) stores its unrounded fp32 sums into this rank's
// half `half` (0: wo, 1: w2) of the exchange region, tc::tp_allreduce_rows sums the ranks, rounds, adds the residual and leaves the
// bf16 rows in every rank's result half.  Returns where the rows are.
bool tp_tc_ready(const mc_llama* m) { return m->cfg.tp_world > 1 && m->tp_tc_rows != 0 && m->tp_connected && !m->nccl_comm; }
uint16_t* tp_tc_result(const mc_llama* m, uint32_t half)
{
    return reinterpret_cast<uint16_t*>(m->tp_region.as<char>() + m->tp_off_tc_result + half * m->tp_tc_half_result);
}
uint32_t tp_tc_row_parallel(mc_llama* m, cudaStream_t s, const uint16_t* X, uint32_t K, const uint16_t* W, uint32_t width, uint32_t half, const uint16_t* res,
                            uint32_t rows)
{
    const mc_llama_config& c = m->cfg;
    const int sms = m->dev->prop.multiProcessorCount;
    MC_REQUIRE(rows <= m->tp_tc_rows, "tensor parallel: too many rows for the exchange region");
    float* mine = reinterpret_cast<float*>(m->tp_region.as<char>() + m->tp_off_tc_partial + half * m->tp_tc_half_partial);
    uint32_t n = uint32_t(tc::gemm(s, sms, tc::GEMM_PARTIAL_F32, X, K, W, reinterpret_cast<uint16_t*>(mine), nullptr, rows, width, K, width, m->errflag.as<int>()));
    tc::tp_rows_exchange x{};
    x.world = c.tp_world, x.rank = c.tp_rank;
    for (uint32_t k = 0; k < c.tp_world; k++) {
        char* base = static_cast<char*>(m->tp_peer_base[k]);
        x.partial[k] = reinterpret_cast<const float*>(base + m->tp_off_tc_partial + half * m->tp_tc_half_partial);
        x.result[k] = reinterpret_cast<uint16_t*>(base + m->tp_off_tc_result + half * m->tp_tc_half_result);
        x.ready[k] = reinterpret_cast<uint32_t*>(base + m->tp_off_tc_flags);
        x.done[k] = reinterpret_cast<uint32_t*>(base + m->tp_off_tc_flags + 64);
    }
    x.counter = m->tp_local.as<unsigned>() + 8, x.epoch = m->tp_local.as<unsigned>() + 9, x.err = m->errflag.as<int>();
    n += uint32_t(tc::tp_allreduce_rows(s, sms, x, res, rows, width));
    return n;
}
uint32_t tc_pad_rows(uint32_t a_rows);
// wo / w2 of a tensor-parallel shard on the tensor-core path.  bf16: out = r(res + r(sum of the ranks' x . W^T)), left in the exchange
// region's result half.  QLoRA: the GEMM runs on the resident image [W | A] of this rank's k range, the all-reduce yields
// r(x . Wd^T) | r(x . A^T) over the full k, the adaptor epilogue (quantization/lora.h:115-122) and the residual follow into Y.
// Returns where the rows are.
uint16_t* tp_tc_linear(mc_llama* m, cudaStream_t s, const uint16_t* X, uint32_t K, const dlinear& d, uint32_t half, const uint16_t* res, uint16_t* Y, uint32_t rows,
                       uint32_t& launches)
{
    const mc_llama_config& c = m->cfg;
    if (!c.quant) {
        launches += tp_tc_row_parallel(m, s, X, K, d.w.as<uint16_t>(), c.dim, half, res, rows);
        return tp_tc_result(m, half);
    }
    const uint32_t width = c.dim + tc_pad_rows(c.lora_rank);
    launches += tp_tc_row_parallel(m, s, X, K, d.wd.as<uint16_t>(), width, half, nullptr, rows);
    launches += uint32_t(tc::lora_epilogue(s, tc::GEMM_RESIDUAL, tp_tc_result(m, half), width, Y, res, d.lora_b.as<uint16_t>(), rows, c.dim, c.dim, c.lora_rank, 1,
                                           m->Hl * c.head_dim, (m->Hl + m->KVl) * c.head_dim, bf16_bits_to_f32(f32_to_bf16_bits(c.lora_scale))));
    return Y;
}

// One linear of a block on the tensor-core path.  bf16 models: a single GEMM with the fused tail.  QLoRA models: the GEMM runs on the
// resident bf16 image and stores r(x . Wd^T); the adaptor term and the tail follow (quantization/lora.h:115-122):
//   y = r(r(x . Wd^T) + r(r(B . r(A . x)) * r(scale))), then residual add / SiLU*mul / plain store.
struct tc_scratch {
    uint16_t* t;   // [rows, max (N + padded adaptor rows)]: r(x . Wd^T) | r(x . A^T)
};
uint32_t tc_pad_rows(uint32_t a_rows) { return (a_rows + 31) & ~31u; }
uint32_t tc_linear(mc_llama* m, cudaStream_t s, int mode, const uint16_t* X, uint32_t ldx, const dlinear& d, uint32_t a_rows, uint32_t slices, uint16_t* Y,
                   const uint16_t* res, uint32_t rows, uint32_t N, uint32_t K, uint32_t ldy, const tc_scratch& sc)
{
    const mc_llama_config& c = m->cfg;
    const int sms = m->dev->prop.multiProcessorCount;
    int* err = m->errflag.as<int>();
    if (!c.quant) return uint32_t(tc::gemm(s, sms, mode, X, ldx, d.w.as<uint16_t>(), Y, res, rows, N, K, ldy, err));
    uint32_t n = 0;
    const uint32_t Next = N + tc_pad_rows(a_rows);
    n += tc::gemm(s, sms, tc::GEMM_STORE, X, ldx, d.wd.as<uint16_t>(), sc.t, nullptr, rows, Next, K, Next, err);
    n += tc::lora_epilogue(s, mode, sc.t, Next, Y, res, d.lora_b.as<uint16_t>(), rows, N, ldy, c.lora_rank, slices, m->Hl * c.head_dim,
                           (m->Hl + m->KVl) * c.head_dim, bf16_bits_to_f32(f32_to_bf16_bits(c.lora_scale)));
    return n;
}

// ---- batched decode on the tensor cores -------------------------------------------------------------------------------------
// Many sequences per step (BASELINE.json "batch 32"): every linear of the step is ONE tcgen05 GEMM over all rows, so the
// weights are streamed once per step instead of once per 4 rows; attention stays the per-(row, head) cluster kernel.
uint32_t decode_tc_min_rows()
{
    static const uint32_t v = [] {
        const char* e = getenv("MC_TC_DECODE_MIN");
        return e ? uint32_t(atoi(e)) : 5u; // measured (1B, B200): 4 rows = one GEMV pass ties with the GEMM path; from 5 rows on the GEMV needs two passes and loses 2x
    }();
    return v;
}
bool decode_tc_eligible(const mc_llama* m, uint32_t n)
{
    const mc_llama_config& c = m->cfg;
    if ((c.flags & MC_LLAMA_NO_TC_PREFILL) || (c.tp_world != 1 && !tp_tc_ready(m)) || n < decode_tc_min_rows() || !m->dt_n.p) return false;
    if (c.quant && (!m->layers[0].wqkv.wd.p || !m->out.wd.p || c.lora_rank % 2 != 0 || c.lora_rank > 16)) return false;
    const uint32_t D = c.dim, QO = m->Hl * c.head_dim, QKV = (m->Hl + 2 * m->KVl) * c.head_dim, F = m->Fl;
    return tc::gemm_supported(QKV, D, D, QKV) && tc::gemm_supported(D, QO, QO, D) && tc::gemm_supported(2 * F, D, D, F) && tc::gemm_supported(D, F, F, D) &&
           tc::gemm_supported(m->Vl, D, D, m->Vl);
}
void enqueue_rows_tc(mc_llama* m, launcher& L, uint32_t rows)
{
    const mc_llama_config& c = m->cfg;
    const uint32_t D = c.dim, hd = c.head_dim, H = m->Hl, KV = m->KVl, QO = H * hd, QKV = (H + 2 * KV) * hd, F = m->Fl;
    cudaStream_t s = L.s;
    const int sms = m->dev->prop.multiProcessorCount;
    uint16_t *x = m->x.as<uint16_t>(), *h = m->h.as<uint16_t>(), *n = m->dt_n.as<uint16_t>(), *qkv = m->dt_qkv.as<uint16_t>();
    uint16_t *q = m->q.as<uint16_t>(), *attn = m->attn.as<uint16_t>(), *z = m->z.as<uint16_t>();
    const tc_scratch sc{m->dt_t.as<uint16_t>()};
    const uint32_t rank = c.lora_rank;
    const bool tp = c.tp_world > 1;
    int* err = m->errflag.as<int>();
    tc::set_pdl(L.pdl);
    auto count = [&](int k) { L.count += uint32_t(k), m->dev->launches.fetch_add(uint64_t(k)); };
    L.go(embed_kernel, dim3(rows), dim3(256), 0, x, D, (const void*)m->tok.w.p, (const float*)m->tok.scales.p, m->tok.fmt, D, c.vocab, m->ids.as<int32_t>());
    for (uint32_t li = 0; li < c.n_layers; li++) {
        dlayer& ly = m->layers[li];
        uint16_t* kc = m->kcache.as<uint16_t>() + size_t(li) * kv_layer_elems(m);
        uint16_t* vc = m->vcache.as<uint16_t>() + size_t(li) * kv_layer_elems(m);
        count(tc::rmsnorm_rows(s, n, x, ly.attn_norm.as<uint16_t>(), rows, D, c.norm_eps));
        count(tc_linear(m, s, tc::GEMM_STORE, n, D, ly.wqkv, 3 * rank, 3, qkv, nullptr, rows, QKV, D, QKV, sc));
        if (tc::decode_attn_gqa_supported(H, KV, hd)) {
            // RoPE, the KV append and the attention of the step in one launch
            count(tc::decode_attn_gqa(s, nullptr, kc, vc, attn, rows, m->row_seq.as<int32_t>(), m->pos.as<int32_t>(), H, KV, hd, c.max_seq_len, m->scale_bf16, qkv,
                                      m->fcos.as<float>(), m->fsin.as<float>()));
        } else {
            count(tc::rope_append(s, qkv, q, kc, vc, m->fcos.as<float>(), m->fsin.as<float>(), rows, 0, 0, H, KV, hd, c.max_seq_len, m->row_seq.as<int32_t>(),
                                  m->pos.as<int32_t>()));
            const attn_params a = attn_params_of(m, li, 0);
            const size_t smem = attn_smem(m, kAttnCluster);
            if (hd == 64) L.go_cluster(attn_decode_kernel<64>, dim3(H * kAttnCluster, rows), dim3(256), smem, kAttnCluster, a);
            else L.go_cluster(attn_decode_kernel<128>, dim3(H * kAttnCluster, rows), dim3(256), smem, kAttnCluster, a);
        }
        if (tp) {
            // row-parallel wo / w2: fp32 partial sums + all-reduce over NVLink (the residual stream then lives in the exchange region's result halves)
            uint32_t k = 0;
            h = tp_tc_linear(m, s, attn, QO, ly.wo, 0, x, m->h.as<uint16_t>(), rows, k);
            count(int(k));
        } else count(tc_linear(m, s, tc::GEMM_RESIDUAL, attn, QO, ly.wo, rank, 1, h, x, rows, D, QO, D, sc));
        count(tc::rmsnorm_rows(s, n, h, ly.ffn_norm.as<uint16_t>(), rows, D, c.norm_eps));
        count(tc_linear(m, s, tc::GEMM_SWIGLU, n, D, ly.w13, 2 * rank, 2, z, nullptr, rows, 2 * F, D, F, sc));
        if (tp) {
            uint32_t k = 0;
            x = tp_tc_linear(m, s, z, F, ly.w2, 1, h, m->x.as<uint16_t>(), rows, k);
            count(int(k));
        } else count(tc_linear(m, s, tc::GEMM_RESIDUAL, z, F, ly.w2, rank, 1, x, h, rows, D, F, D, sc));
    }
    if (tp && x != m->x.as<uint16_t>()) {
        // the last hidden rows go back to where the rest of the engine expects them
        MC_CUDA_CHECK(cudaMemcpyAsync(m->x.p, x, size_t(rows) * D * 2, cudaMemcpyDeviceToDevice, s));
        x = m->x.as<uint16_t>();
    }
    count(tc::rmsnorm_rows(s, n, x, m->norm.as<uint16_t>(), rows, D, c.norm_eps));
    // vocabulary projection: the (tied) bf16 table (this rank's vocabulary slice), or the cached bf16 image of the int8 output matrix (quantization/linear.h:50-53)
    const uint16_t* head_w = c.quant ? m->out.wd.as<uint16_t>() : (m->tied ? m->tok.w.as<uint16_t>() + size_t(c.tp_rank) * m->Vl * D : m->out.w.as<uint16_t>());
    count(tc::gemm(s, sms, tc::GEMM_STORE, n, D, head_w, m->logits.as<uint16_t>(), nullptr, rows, m->Vl, D, m->Vl, err));
}

// one decode step for rows [0, n): forward in chunks of kMaxMB rows, then sample
void enqueue_decode_step(mc_llama* m, launcher& L, uint32_t n, const mc_sampler_config& sc, int advance)
{
    if (m->sink_roll) {
        // nn::sink_cache (nn/cache.h:123-126,183-204): pre_len = bit_width(max_seq_len) - 1 sink rows stay, the rest moves one row left
        uint32_t pre_len = 0;
        while ((2u << pre_len) <= m->cfg.max_seq_len) pre_len++;
        L.go(kv_roll_kernel, dim3(m->cfg.n_layers * m->KVl * 2, n), dim3(256), 0, m->kcache.as<uint16_t>(), m->vcache.as<uint16_t>(), kv_layer_elems(m),
             (const int32_t*)m->row_seq.as<int32_t>(), (const int32_t*)m->pos.as<int32_t>(), m->KVl, m->cfg.head_dim, m->cfg.max_seq_len, pre_len);
    }
    if (stream_eligible(m, n, sc)) {
        launch_stream(m, L, n, advance, 1);
        return;
    }
    if (decode_tc_eligible(m, n)) {
        enqueue_rows_tc(m, L, n);
        enqueue_sample(m, L, n, sc, advance);
        return;
    }
    for (uint32_t r0 = 0; r0 < n; r0 += kMaxMB) {
        const uint32_t rows = std::min<uint32_t>(kMaxMB, n - r0);
        enqueue_rows(m, L, r0, rows, 1, m->logits.as<uint16_t>() + size_t(r0) * m->Vl);
    }
    enqueue_sample(m, L, n, sc, advance);
}

cudaGraphExec_t decode_graph(mc_llama* m, uint32_t n, const mc_sampler_config& sc, int advance)
{
    uint32_t tbits, pbits;
    memcpy(&tbits, &sc.temperature, 4), memcpy(&pbits, &sc.top_p, 4);
    const uint64_t key = mix64((uint64_t(n) << 40) ^ (uint64_t(m->sink_roll) << 37) ^ (uint64_t(sc.mode) << 36) ^ (uint64_t(sc.intended) << 35) ^ (uint64_t(advance) << 34) ^
                               (uint64_t(sc.top_k) << 24) ^ (uint64_t(tbits) * 0x9E3779B97F4A7C15ull) ^ (uint64_t(pbits) << 1));
    auto it = m->graphs.find(key);
    if (it != m->graphs.end()) return it->second;
    cudaStream_t s = m->dev->stream;
    launcher L{m, s, !(m->cfg.flags & MC_LLAMA_NO_PDL)};
    MC_CUDA_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    cudaGraph_t g = nullptr;
    try {
        enqueue_decode_step(m, L, n, sc, advance);
    } catch (...) {
        cudaStreamEndCapture(s, &g);
        if (g) cudaGraphDestroy(g);
        throw;
    }
    MC_CUDA_CHECK(cudaStreamEndCapture(s, &g));
    cudaGraphExec_t ex = nullptr;
    cudaError_t e = cudaGraphInstantiate(&ex, g, 0);
    cudaGraphDestroy(g);
    MC_CUDA_CHECK(e);
    m->graphs[key] = ex;
    m->launches_per_step = L.count;
    return ex;
}

void run_decode_step(mc_llama* m, uint32_t n, const mc_sampler_config& sc, int advance)
{
    if ((m->cfg.flags & MC_LLAMA_NO_GRAPH) || stream_eligible(m, n, sc)) {
        launcher L{m, m->dev->stream, !(m->cfg.flags & MC_LLAMA_NO_PDL)};
        enqueue_decode_step(m, L, n, sc, advance);
        m->launches_per_step = L.count;
    } else {
        cudaGraphExec_t ex = decode_graph(m, n, sc, advance);
        MC_CUDA_CHECK(cudaGraphLaunch(ex, m->dev->stream));
        m->dev->launches.fetch_add(m->launches_per_step);
    }
}

// ---- parameter names ---------------------------------------------------------------------------------------------
// A reference parameter (full, unsharded, host) maps to a 2-D slice of a fused device tensor.
enum { GEN_BF16 = 0, GEN_I8 = 1, GEN_SCALE = 2 };
struct slice2d {
    void* dst;          // device destination of element (0,0) of the slice
    size_t dst_ld;      // destination row pitch in elements
    uint32_t rows, cols;
    uint32_t src_row0, src_col0, src_K; // slice origin and row length of the full tensor
    size_t elem;        // bytes per element
    uint64_t full_elems;
    int gen;            // which synthetic generator fills it
};

bool parse_layer_name(const std::string& name, uint32_t& layer, std::string& rest)
{
    if (name.rfind("layers.", 0) != 0) return false;
    const size_t dot = name.find('.', 7);
    if (dot == std::string::npos) return false;
    for (size_t i = 7; i < dot; i++)
        if (name[i] < '0' || name[i] > '9') return false;
    if (dot == 7 || dot - 7 > 6) return false;
    layer = uint32_t(std::stoul(name.substr(7, dot - 7)));
    rest = name.substr(dot + 1);
    return true;
}

// where one of the seven reference linears of a block lives inside the fused device tensors
struct linspec {
    dlinear* d = nullptr;
    dbuf* lora_a = nullptr;
    uint32_t a_row0 = 0;       // first row of its adaptor A inside the stacked A matrix
    uint32_t dst_row0 = 0;     // first fused row
    uint32_t dst_step = 1;     // fused row step (2 = interleaved w1/w3)
    uint32_t rows = 0;         // local rows
    uint32_t src_row0 = 0;     // first row of the full tensor held by this shard
    uint32_t Kl = 0;           // local reduction length
    uint32_t src_col0 = 0;     // first column of the full tensor held by this shard
    uint32_t K_full = 0, N_full = 0;
};
bool linear_spec(mc_llama* m, dlayer& ly, const std::string& lin, linspec& o)
{
    const mc_llama_config& c = m->cfg;
    const uint32_t D = c.dim, hd = c.head_dim, r = c.tp_rank, rank = c.lora_rank;
    const uint32_t QO = c.n_heads * hd, KO = c.n_kv_heads * hd;
    if (lin == "attention.wq") o = {&ly.wqkv, &ly.lora_a_qkv, 0, 0, 1, m->Hl * hd, r * m->Hl * hd, D, 0, D, QO};
    else if (lin == "attention.wk") o = {&ly.wqkv, &ly.lora_a_qkv, rank, m->Hl * hd, 1, m->KVl * hd, r * m->KVl * hd, D, 0, D, KO};
    else if (lin == "attention.wv") o = {&ly.wqkv, &ly.lora_a_qkv, 2 * rank, (m->Hl + m->KVl) * hd, 1, m->KVl * hd, r * m->KVl * hd, D, 0, D, KO};
    else if (lin == "attention.wo") o = {&ly.wo, &ly.lora_a_o, 0, 0, 1, D, 0, m->Hl * hd, r * m->Hl * hd, QO, D};
    else if (lin == "feed_forward.w1") o = {&ly.w13, &ly.lora_a_13, 0, 0, 2, m->Fl, r * m->Fl, D, 0, D, c.ffn_dim};
    else if (lin == "feed_forward.w3") o = {&ly.w13, &ly.lora_a_13, rank, 1, 2, m->Fl, r * m->Fl, D, 0, D, c.ffn_dim};
    else if (lin == "feed_forward.w2") o = {&ly.w2, &ly.lora_a_2, 0, 0, 1, D, 0, m->Fl, r * m->Fl, c.ffn_dim, D};
    else return false;
    return true;
}

bool ends_with(const std::string& s, const std::string& suf, std::string& stem)
{
    if (s.size() <= suf.size() || s.compare(s.size() - suf.size(), suf.size(), suf) != 0) return false;
    stem = s.substr(0, s.size() - suf.size());
    return true;
}

// Resolves a reference parameter path to its destination slice(s) in the fused device layout.
std::vector<slice2d> resolve(mc_llama* m, const std::string& name)
{
    const mc_llama_config& c = m->cfg;
    const uint32_t D = c.dim, r = c.tp_rank, G = c.group_size, rank = c.lora_rank;
    const bool Q = c.quant != 0;
    std::vector<slice2d> out;
    auto vec = [&](dbuf& b, uint32_t n) { out.push_back({b.p, n, 1, n, 0, 0, n, 2, n, GEN_BF16}); };
    auto staged = [&](dbuf& b, const char* what) {
        if (!b.p) throw error(MC_ERR_INVALID, std::string("parameter ") + name + ": " + what + " were already packed by mc_llama_finalize");
        return b.p;
    };
    uint32_t li = 0;
    std::string rest, stem;
    if (name == "tok_embeddings.weight") {
        if (Q) out.push_back({m->tok.w.p, D, c.vocab, D, 0, 0, D, 1, uint64_t(c.vocab) * D, GEN_I8});
        else out.push_back({m->tok.w.p, D, c.vocab, D, 0, 0, D, 2, uint64_t(c.vocab) * D, GEN_BF16});
    } else if (name == "tok_embeddings.scales" && Q) {
        out.push_back({m->tok.scales.p, 1, c.vocab, 1, 0, 0, 1, 4, uint64_t(c.vocab), GEN_SCALE});
    } else if (name == "output.weight") {
        if (!Q) {
            // the reference always registers an `output` linear (nn/llama.h:79); its HF adaptor aliases it to tok_embeddings
            // (huggingface/llama.h:103).  Setting it explicitly unties the head: it gets its own [Vl, D] shard from here on.
            if (!m->out.w.p) {
                m->out.N = m->Vl, m->out.K = D, m->out.fmt = WF_BF16;
                m->out.w.alloc(size_t(m->Vl) * D * 2);
                m->tied = false;
            }
            out.push_back({m->out.w.p, D, m->Vl, D, r * m->Vl, 0, D, 2, uint64_t(c.vocab) * D, GEN_BF16});
        } else {
            out.push_back({staged(m->out.q8, "the weights"), D, m->Vl, D, r * m->Vl, 0, D, 1, uint64_t(c.vocab) * D, GEN_I8});
        }
    } else if (name == "output.scales" && Q) {
        // the cached bf16 image of the output projection (quantization/linear.h:50-53) was built from these scales
        if (m->out.wd.p) throw error(MC_ERR_INVALID, "parameter output.scales: the dequantised output projection was already built by mc_llama_finalize");
        out.push_back({m->out.scales.p, 1, m->Vl, 1, r * m->Vl, 0, 1, 4, uint64_t(c.vocab), GEN_SCALE});
    } else if (name == "norm.weight") {
        vec(m->norm, D);
    } else if (parse_layer_name(name, li, rest)) {
        if (li >= c.n_layers) throw error(MC_ERR_NOT_FOUND, "layer index out of range: " + name);
        dlayer& ly = m->layers[li];
        linspec sp;
        if (rest == "attention_norm.weight") vec(ly.attn_norm, D);
        else if (rest == "ffn_norm.weight") vec(ly.ffn_norm, D);
        else if (ends_with(rest, ".adaptor.A.weight", stem) && Q && linear_spec(m, ly, stem, sp)) {
            if (sp.d->wd.p) m->image_a_dirty = true; // rows [N, N + R) of the bf16 image hold a copy of A
            out.push_back({sp.lora_a->as<uint16_t>() + size_t(sp.a_row0) * sp.Kl, sp.Kl, rank, sp.Kl, 0, sp.src_col0, sp.K_full, 2, uint64_t(rank) * sp.K_full, GEN_BF16});
        } else if (ends_with(rest, ".adaptor.B.weight", stem) && Q && linear_spec(m, ly, stem, sp))
            out.push_back({sp.d->lora_b.as<uint16_t>() + size_t(sp.dst_row0) * rank, size_t(rank) * sp.dst_step, sp.rows, rank, sp.src_row0, 0, rank, 2,
                           uint64_t(sp.N_full) * rank, GEN_BF16});
        else if (ends_with(rest, ".scales", stem) && Q && linear_spec(m, ly, stem, sp))
            out.push_back({static_cast<float*>(staged(sp.d->s32, "the scales")) + size_t(sp.dst_row0) * (sp.Kl / G), size_t(sp.Kl / G) * sp.dst_step, sp.rows,
                           sp.Kl / G, sp.src_row0, sp.src_col0 / G, sp.K_full / G, 4, uint64_t(sp.N_full) * (sp.K_full / G), GEN_SCALE});
        else if (ends_with(rest, ".weight", stem) && linear_spec(m, ly, stem, sp)) {
            if (Q)
                out.push_back({static_cast<int8_t*>(staged(sp.d->q8, "the weights")) + size_t(sp.dst_row0) * sp.Kl, size_t(sp.Kl) * sp.dst_step, sp.rows, sp.Kl,
                               sp.src_row0, sp.src_col0, sp.K_full, 1, uint64_t(sp.N_full) * sp.K_full, GEN_I8});
            else
                out.push_back({sp.d->w.as<uint16_t>() + size_t(sp.dst_row0) * sp.Kl, size_t(sp.Kl) * sp.dst_step, sp.rows, sp.Kl, sp.src_row0, sp.src_col0,
                               sp.K_full, 2, uint64_t(sp.N_full) * sp.K_full, GEN_BF16});
        } else throw error(MC_ERR_NOT_FOUND, "unknown parameter: " + name);
    } else {
        throw error(MC_ERR_NOT_FOUND, "unknown parameter: " + name);
    }
    return out;
}

unsigned gen_blocks(uint64_t n) { return unsigned(std::min<uint64_t>((n + 255) / 256, 148 * 32)); }
void gen_bf16(mc_llama* m, const slice2d& s, uint64_t seed, uint64_t tid, float scale, float bias)
{
    gen_bf16_kernel<<<gen_blocks(uint64_t(s.rows) * s.cols), 256, 0, m->dev->stream>>>(static_cast<uint16_t*>(s.dst), s.dst_ld, s.rows, s.cols, s.src_row0,
                                                                                        s.src_col0, s.src_K, seed, tid, scale, bias);
    MC_CUDA_CHECK(cudaGetLastError());
}
void gen_i8(mc_llama* m, const slice2d& s, uint64_t seed, uint64_t tid, int32_t lo, uint32_t range)
{
    gen_i8_kernel<<<gen_blocks(uint64_t(s.rows) * s.cols), 256, 0, m->dev->stream>>>(static_cast<int8_t*>(s.dst), s.dst_ld, s.rows, s.cols, s.src_row0, s.src_col0,
                                                                                      s.src_K, seed, tid, lo, range);
    MC_CUDA_CHECK(cudaGetLastError());
}
void gen_scales(mc_llama* m, const slice2d& s, uint64_t seed, uint64_t tid, float c0)
{
    gen_f32_scales_kernel<<<gen_blocks(uint64_t(s.rows) * s.cols), 256, 0, m->dev->stream>>>(static_cast<float*>(s.dst), s.dst_ld, s.rows, s.cols, s.src_row0,
                                                                                              s.src_col0, s.src_K, seed, tid, c0);
    MC_CUDA_CHECK(cudaGetLastError());
}

// generator kinds shared with the test oracle (DESIGN.md "Synthetic data")
enum : uint32_t { K_ATTN_NORM = 0, K_FFN_NORM = 1, K_WQ = 2, K_WK = 3, K_WV = 4, K_WO = 5, K_W1 = 6, K_W2 = 7, K_W3 = 8, K_SCALES = 16, K_LORA_A = 32,
                  K_LORA_B = 48, G_TOK = 0, G_NORM = 1, G_OUT = 2 };
uint64_t tid_layer(uint32_t layer, uint32_t kind) { return uint64_t(layer + 1) * 256 + kind; }

size_t w4_bytes(uint32_t N, uint32_t K) { return size_t((N / 2 + 7) / 8) * (K / 64) * 512; }
size_t w4_scale_bytes(uint32_t N, uint32_t K) { return size_t((N / 2 + 7) / 8) * (K / 64) * 64; }
size_t w8_bytes(uint32_t N, uint32_t K) { return size_t((N / 2 + 7) / 8) * (K / 32) * 512; }

gemv_params shape_only(uint32_t N, uint32_t K, uint32_t head_dim)
{
    gemv_params p{};
    p.N = N, p.K = K, p.head_dim = head_dim;
    return p;
}

} // namespace

extern "C" {

mc_status mc_llama_create(mc_device* dev, const mc_llama_config* cfg, mc_llama** out)
{
    MC_API_BEGIN
    MC_REQUIRE(dev && cfg && out, "bad arguments");
    MC_CUDA_CHECK(cudaSetDevice(dev->ordinal));
    mc_llama_config c = *cfg;
    if (c.n_seqs == 0) c.n_seqs = 1;
    if (c.tp_world == 0) c.tp_world = 1;
    MC_REQUIRE(c.tp_world <= uint32_t(kTpMaxWorld), "tensor-parallel degree must be <= 8");
    MC_REQUIRE(c.tp_rank < c.tp_world, "tp_rank out of range");
    MC_REQUIRE(c.head_dim == 64 || c.head_dim == 128, "head_dim must be 64 or 128");
    MC_REQUIRE(c.n_heads % c.n_kv_heads == 0, "n_heads must be a multiple of n_kv_heads");
    MC_REQUIRE(c.n_heads % c.tp_world == 0 && c.n_kv_heads % c.tp_world == 0 && c.ffn_dim % c.tp_world == 0 && c.vocab % c.tp_world == 0,
               "dimensions are not divisible by the tensor-parallel degree");
    MC_REQUIRE(c.dim % 256 == 0 && c.ffn_dim % 256 == 0 && (c.n_heads * c.head_dim) % 256 == 0, "dim, ffn_dim and n_heads*head_dim must be multiples of 256");
    MC_REQUIRE(c.vocab % 2 == 0, "vocab must be even");
    MC_REQUIRE(c.max_seq_len >= 1 && c.max_seq_len <= 16384, "max_seq_len out of range");
    if (c.quant) {
        MC_REQUIRE(c.group_size == 32, "quantised layout: group_size must be 32 (huggingface/llama.h:167)");
        MC_REQUIRE(c.lora_rank >= 1 && c.lora_rank <= 64, "quantised layout: lora_rank out of range");
        MC_REQUIRE(c.tp_world == 1 || ((c.n_heads / c.tp_world * c.head_dim) % 256 == 0 && (c.ffn_dim / c.tp_world) % 256 == 0),
                   "quantised layout under tensor parallelism: every rank's slice of n_heads*head_dim and of ffn_dim must be a multiple of 256");
    }
    auto m = std::make_unique<mc_llama>();
    m->dev = dev;
    m->cfg = c;
    m->Hl = c.n_heads / c.tp_world, m->KVl = c.n_kv_heads / c.tp_world, m->Fl = c.ffn_dim / c.tp_world, m->Vl = c.vocab / c.tp_world;
    m->tied = c.quant == 0;
    const uint32_t D = c.dim, hd = c.head_dim;
    m->layers.resize(c.n_layers);
    const uint32_t QKVN = (m->Hl + 2 * m->KVl) * hd, QOl = m->Hl * hd;
    if (!c.quant) {
        // one arena, fixed stride per layer: attn_norm | ffn_norm | wqkv | wo | w13 | w2 (256-byte aligned pieces)
        auto al = [](size_t n) { return (n + 255) & ~size_t(255); };
        const size_t sz[6] = {al(size_t(D) * 2), al(size_t(D) * 2), al(size_t(QKVN) * D * 2), al(size_t(D) * QOl * 2),
                              al(size_t(2) * m->Fl * D * 2), al(size_t(D) * m->Fl * 2)};
        m->layer_stride = sz[0] + sz[1] + sz[2] + sz[3] + sz[4] + sz[5];
        m->layer_arena.alloc(m->layer_stride * c.n_layers);
        for (uint32_t li = 0; li < c.n_layers; li++) {
            dlayer& ly = m->layers[li];
            char* base = m->layer_arena.as<char>() + size_t(li) * m->layer_stride;
            ly.attn_norm.view(base, size_t(D) * 2), base += sz[0];
            ly.ffn_norm.view(base, size_t(D) * 2), base += sz[1];
            ly.wqkv.N = QKVN, ly.wqkv.K = D;
            ly.wqkv.w.view(base, size_t(QKVN) * D * 2), base += sz[2];
            ly.wo.N = D, ly.wo.K = QOl;
            ly.wo.w.view(base, size_t(D) * QOl * 2), base += sz[3];
            ly.w13.N = 2 * m->Fl, ly.w13.K = D;
            ly.w13.w.view(base, size_t(2) * m->Fl * D * 2), base += sz[4];
            ly.w2.N = D, ly.w2.K = m->Fl;
            ly.w2.w.view(base, size_t(D) * m->Fl * 2);
        }
        m->tok.N = c.vocab, m->tok.K = D;
        m->tok.w.alloc(size_t(c.vocab) * D * 2);
    } else {
        // QLoRA layout (huggingface/llama.h:152-171): int4-range weights packed two per byte in mma fragment order +
        // bf16 r(scale) per 32 weights + LoRA A/B in bf16; int8 per-row tables for tok_embeddings / output.
        const uint32_t rank = c.lora_rank;
        // one arena with a fixed stride per layer (the streaming kernel indexes by layer): norms | per linear: packed weights,
        // packed scales, LoRA B, stacked LoRA A.  The reference-layout staging (q8, s32) is separate and released by
        // mc_llama_finalize.
        auto al = [](size_t n) { return (n + 255) & ~size_t(255); };
        struct shape {
            uint32_t N, K;
        } const sh[4] = {{QKVN, D}, {D, QOl}, {2 * m->Fl, D}, {D, m->Fl}};
        const uint32_t a_rows[4] = {3 * rank, rank, 2 * rank, rank};
        size_t stride = 2 * al(size_t(D) * 2);
        for (int i = 0; i < 4; i++) {
            MC_REQUIRE(sh[i].K % 256 == 0 && sh[i].N % 2 == 0, "quantised layout: K must be a multiple of 256");
            stride += al(w4_bytes(sh[i].N, sh[i].K)) + al(w4_scale_bytes(sh[i].N, sh[i].K)) + al(size_t(sh[i].N) * rank * 2) + al(size_t(a_rows[i]) * sh[i].K * 2);
        }
        m->layer_stride = stride;
        m->layer_arena.alloc(stride * c.n_layers);
        for (uint32_t li = 0; li < c.n_layers; li++) {
            dlayer& ly = m->layers[li];
            char* base = m->layer_arena.as<char>() + size_t(li) * stride;
            ly.attn_norm.view(base, size_t(D) * 2), base += al(size_t(D) * 2);
            ly.ffn_norm.view(base, size_t(D) * 2), base += al(size_t(D) * 2);
            dlinear* lin[4] = {&ly.wqkv, &ly.wo, &ly.w13, &ly.w2};
            dbuf* la[4] = {&ly.lora_a_qkv, &ly.lora_a_o, &ly.lora_a_13, &ly.lora_a_2};
            for (int i = 0; i < 4; i++) {
                dlinear* d = lin[i];
                d->N = sh[i].N, d->K = sh[i].K, d->fmt = WF_W4;
                d->w.view(base, w4_bytes(d->N, d->K)), base += al(w4_bytes(d->N, d->K));
                d->scales.view(base, w4_scale_bytes(d->N, d->K)), base += al(w4_scale_bytes(d->N, d->K));
                d->lora_b.view(base, size_t(d->N) * rank * 2), base += al(size_t(d->N) * rank * 2);
                la[i]->view(base, size_t(a_rows[i]) * d->K * 2), base += al(size_t(a_rows[i]) * d->K * 2);
                d->q8.alloc(size_t(d->N) * d->K);
                d->s32.alloc(size_t(d->N) * (d->K / 32) * 4);
            }
        }
        m->tok.N = c.vocab, m->tok.K = D, m->tok.fmt = WF_W8ROW;
        m->tok.w.alloc(size_t(c.vocab) * D);      // natural row-major int8: the embedding gather reads whole rows
        m->tok.scales.alloc(size_t(c.vocab) * 4);
        m->out.N = m->Vl, m->out.K = D, m->out.fmt = WF_W8ROW;
        m->out.w.alloc(w8_bytes(m->Vl, D));
        m->out.scales.alloc(size_t(m->Vl) * 4);
        m->out.q8.alloc(size_t(m->Vl) * D);
        m->lora_ax.alloc(size_t(std::max<uint32_t>(c.n_seqs, kMaxMB)) * 3 * rank * 2);
        m->pack_bad.alloc(4);
        MC_CUDA_CHECK(cudaMemset(m->pack_bad.p, 0, 4));
    }
    m->bar.alloc(256);
    MC_CUDA_CHECK(cudaMemset(m->bar.p, 0, 256));
    if (c.tp_world > 1) {
        // partial sums travel in passes of kMaxMB rows; the argmax exchange carries every sequence of a step at once
        const size_t rows_max = kMaxMB, T = c.tp_world, am_rows = std::max<uint32_t>(c.n_seqs, kMaxMB);
        const size_t part = 2 * T * rows_max * tp_row_floats(m.get()) * sizeof(float);
        m->tp_off_flags = part;
        m->tp_off_amval = m->tp_off_flags + 256;
        m->tp_off_amidx = m->tp_off_amval + ((2 * T * am_rows * 4 + 255) & ~size_t(255));
        m->tp_off_amflags = m->tp_off_amidx + ((2 * T * am_rows * 4 + 255) & ~size_t(255));
        // streaming kernel under tensor parallelism: [generation][kind wo | w2][src rank][row][dim] tagged fp32 words and
        // [generation][src rank][row][2] argmax words.  Tags repeat when the 16-bit launch counter wraps; the generation (bit 16 of
        // the counter) alternates the region then, and a rank clears the region it has just left: every word a peer sent into it has
        // been consumed by then, and no peer writes there again before this rank has moved on twice.
        m->tp_off_stpart = m->tp_off_amflags + 256;
        m->tp_stpart_gen = size_t(2) * T * kStMaxRows * D * 8;
        m->tp_off_stam = m->tp_off_stpart + 2 * m->tp_stpart_gen;
        m->tp_stam_gen = (T * kStMaxRows * 2 * 8 + 255) & ~size_t(255);
        m->tp_off_stax = m->tp_off_stam + 2 * m->tp_stam_gen;
        m->tp_stax_gen = size_t(2) * T * kStMaxRows * kStTpAxCols * 8;
        size_t region_bytes = m->tp_off_stax + 2 * m->tp_stax_gen;
        if (!(c.flags & MC_LLAMA_NO_TC_PREFILL) && !(c.quant && (c.flags & MC_LLAMA_NO_SHADOW))) {
            // prompts and decode batches on the tcgen05 path: a chunk of rows is all-reduced at once (tc::tp_allreduce_rows); quantised
            // models carry the adaptor's A . x columns behind the dim main sums
            m->tp_tc_rows = std::max<uint32_t>(std::min<uint32_t>(2048u, c.max_seq_len), std::max<uint32_t>(c.n_seqs, kMaxMB));
            const size_t width = D + (c.quant ? ((c.lora_rank + 31) & ~31u) : 0u);
            m->tp_tc_half_partial = (size_t(m->tp_tc_rows) * width * 4 + 255) & ~size_t(255);
            m->tp_tc_half_result = (size_t(m->tp_tc_rows) * width * 2 + 255) & ~size_t(255);
            m->tp_off_tc_partial = (region_bytes + 255) & ~size_t(255);
            m->tp_off_tc_result = m->tp_off_tc_partial + 2 * m->tp_tc_half_partial;
            m->tp_off_tc_flags = m->tp_off_tc_result + 2 * m->tp_tc_half_result;
            region_bytes = m->tp_off_tc_flags + 256;
        }
        m->tp_region.alloc(region_bytes);
        MC_CUDA_CHECK(cudaMemset(m->tp_region.p, 0, m->tp_region.bytes));
        m->tp_local.alloc(256);
        MC_CUDA_CHECK(cudaMemset(m->tp_local.p, 0, 256));
        m->tp_peer_base[c.tp_rank] = m->tp_region.p;
    }
    m->norm.alloc(size_t(D) * 2);
    build_rope_tables(m.get(), 2 * c.max_seq_len); // nn/embedding.h:171: 2 * max_seq_len positions
    const size_t kv_bytes = size_t(c.n_layers) * kv_layer_elems(m.get()) * 2;
    m->kcache.alloc(kv_bytes);
    m->vcache.alloc(kv_bytes);
    MC_CUDA_CHECK(cudaMemset(m->kcache.p, 0, kv_bytes));
    MC_CUDA_CHECK(cudaMemset(m->vcache.p, 0, kv_bytes));
    m->max_rows = std::max<uint32_t>(c.n_seqs, kMaxMB);
    const uint32_t R = m->max_rows;
    m->x.alloc(size_t(R) * D * 2), m->h.alloc(size_t(R) * D * 2);
    m->q.alloc(size_t(R) * m->Hl * hd * 2), m->attn.alloc(size_t(R) * m->Hl * hd * 2);
    m->z.alloc(size_t(R) * m->Fl * 2);
    if ((c.tp_world == 1 || m->tp_tc_rows) && c.n_seqs >= decode_tc_min_rows()) {
        m->dt_n.alloc(size_t(R) * D * 2);
        m->dt_qkv.alloc(size_t(R) * (m->Hl + 2 * m->KVl) * hd * 2);
        if (c.quant) m->dt_t.alloc(size_t(R) * (std::max(std::max((m->Hl + 2 * m->KVl) * hd, 2 * m->Fl), D) + 128) * 2);
    }
    {
        // tagged-word exchange buffers of the streaming kernel (8 bytes per word = two bf16 + tag)
        const size_t R8 = kStMaxRows;
        m->st_sc_words = (c.max_seq_len + 1) / 2 + 1;
        const size_t words[9] = {R8 * D / 2, R8 * D / 2, R8 * m->Fl / 2, R8 * QKVN / 2, R8 * QOl / 2, R8 * m->Hl * m->st_sc_words, R8 * 256 * 2, 64, 4 * R8 * 128};
        size_t total = 0;
        for (int i = 0; i < 9; i++) m->st_off[i] = total, total += (words[i] + 31) & ~size_t(31);
        m->st_words = total;
        m->st_ll.alloc(total * 8 * kStMaxRep);
        MC_CUDA_CHECK(cudaMemset(m->st_ll.p, 0, m->st_ll.bytes));
    }
    m->logits.alloc(size_t(R) * m->Vl * 2);
    if (c.tp_world > 1) m->logits_tmp.alloc(size_t(kMaxMB) * m->Vl * 2);
    m->hidden_save.alloc(size_t(c.n_seqs) * D * 2);
    // decode inputs and outputs are contiguous so that a per-token call moves them with ONE copy each way:
    //   io_in  = ids[R] | pos[R] | row_seq[R] | step_counter      io_out = error flag (16 B) | out_log
    m->io_in.alloc((size_t(3) * R + 4) * 4);
    m->ids.view(m->io_in.p, R * 4), m->pos.view(m->io_in.as<int32_t>() + R, R * 4), m->row_seq.view(m->io_in.as<int32_t>() + 2 * R, R * 4);
    m->step_counter.view(m->io_in.as<int32_t>() + 3 * R, 4);
    m->io_out.alloc(16 + size_t(kMaxLogSteps) * R * 4);
    MC_CUDA_CHECK(cudaMemset(m->io_out.p, 0, 16));
    MC_CUDA_CHECK(cudaMemset(m->io_out.as<char>() + 8, 0x7f, 4)); // (diagnostics: lowest failing wait, see check_device_error)
    m->errflag.view(m->io_out.p, 16), m->out_log.view(m->io_out.as<char>() + 16, size_t(kMaxLogSteps) * R * 4);
    m->uniforms.alloc(size_t(kMaxLogSteps) * R * 4);
    m->cand.alloc(size_t(R) * ((m->Vl + kSampleSlice - 1) / kSampleSlice) * kSampleKeep * 8);
    m->pval.alloc(size_t(R) * 1024 * 4), m->pidx.alloc(size_t(R) * 1024 * 4);
    MC_CUDA_CHECK(cudaMemset(m->logits.p, 0, m->logits.bytes));
    MC_CUDA_CHECK(cudaMemset(m->hidden_save.p, 0, m->hidden_save.bytes));
    MC_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&m->pinned), (size_t(16) * R + 512) * 4, cudaHostAllocDefault));
    m->scale_bf16 = bf16_bits_to_f32(f32_to_bf16_bits(1.0f / std::sqrt(float(hd))));
    m->hidden_ptr.assign(c.n_seqs, nullptr);
    *out = m.release();
    MC_API_END
}

mc_status mc_llama_destroy(mc_llama* m)
{
    MC_API_BEGIN
    if (m) {
        cudaSetDevice(m->dev->ordinal);
        cudaStreamSynchronize(m->dev->stream);
        delete m;
    }
    MC_API_END
}

mc_status mc_llama_set_tensor(mc_llama* m, const char* name, const void* host, size_t nbytes)
{
    MC_API_BEGIN
    nvtx_range nvtx_("mc_llama_set_tensor");
    use(m);
    MC_REQUIRE(name && host, "bad arguments");
    for (const slice2d& s : resolve(m, name)) {
        if (nbytes != s.full_elems * s.elem) {
            throw error(MC_ERR_INVALID, std::string("parameter ") + name + ": expected " + std::to_string(s.full_elems * s.elem) + " bytes, got " + std::to_string(nbytes));
        }
        const char* src = static_cast<const char*>(host) + (size_t(s.src_row0) * s.src_K + s.src_col0) * s.elem;
        MC_CUDA_CHECK(cudaMemcpy2DAsync(s.dst, s.dst_ld * s.elem, src, size_t(s.src_K) * s.elem, size_t(s.cols) * s.elem, s.rows,
                                        cudaMemcpyHostToDevice, m->dev->stream));
    }
    MC_CUDA_CHECK(cudaStreamSynchronize(m->dev->stream));
    m->finalized = false;
    MC_API_END
}

mc_status mc_llama_get_config(mc_llama* m, mc_llama_config* cfg)
{
    MC_API_BEGIN
    MC_REQUIRE(m && cfg, "bad arguments");
    *cfg = m->cfg;
    MC_API_END
}

mc_status mc_llama_init_random(mc_llama* m, uint64_t seed)
{
    MC_API_BEGIN
    nvtx_range nvtx_("mc_llama_init_random");
    use(m);
    const mc_llama_config& c = m->cfg;
    const bool Q = c.quant != 0;
    auto bf = [&](const std::string& name, uint64_t tid, float scale, float bias) {
        for (const slice2d& s : resolve(m, name)) gen_bf16(m, s, seed, tid, scale, bias);
    };
    auto i8 = [&](const std::string& name, uint64_t tid, int32_t lo, uint32_t range) {
        for (const slice2d& s : resolve(m, name)) gen_i8(m, s, seed, tid, lo, range);
    };
    auto sc = [&](const std::string& name, uint64_t tid, float c0) {
        for (const slice2d& s : resolve(m, name)) gen_scales(m, s, seed, tid, c0);
    };
    // one reference linear [N, K]: bf16 U/sqrt(K), or int4-range q + group scales + LoRA A/B
    auto linear = [&](const std::string& prefix, uint64_t tid, uint32_t K) {
        const float inv_sqrt_k = 1.0f / std::sqrt(float(K));
        if (!Q) {
            bf(prefix + ".weight", tid, inv_sqrt_k, 0.0f);
        } else {
            i8(prefix + ".weight", tid, -7, 15); // zero-mean int4 range (DESIGN.md "Synthetic data": [-8, 7] has mean -0.5 and collapses the random-init model)
            sc(prefix + ".scales", tid + K_SCALES, inv_sqrt_k * 0.125f);
            bf(prefix + ".adaptor.A.weight", tid + K_LORA_A, inv_sqrt_k, 0.0f);
            bf(prefix + ".adaptor.B.weight", tid + K_LORA_B, 1.0f / std::sqrt(float(c.lora_rank)), 0.0f);
        }
    };
    for (uint32_t i = 0; i < c.n_layers; i++) {
        const std::string p = "layers." + std::to_string(i) + ".";
        bf(p + "attention_norm.weight", tid_layer(i, K_ATTN_NORM), 0.1f, 1.0f);
        bf(p + "ffn_norm.weight", tid_layer(i, K_FFN_NORM), 0.1f, 1.0f);
        linear(p + "attention.wq", tid_layer(i, K_WQ), c.dim);
        linear(p + "attention.wk", tid_layer(i, K_WK), c.dim);
        linear(p + "attention.wv", tid_layer(i, K_WV), c.dim);
        linear(p + "attention.wo", tid_layer(i, K_WO), c.n_heads * c.head_dim);
        linear(p + "feed_forward.w1", tid_layer(i, K_W1), c.dim);
        linear(p + "feed_forward.w2", tid_layer(i, K_W2), c.ffn_dim);
        linear(p + "feed_forward.w3", tid_layer(i, K_W3), c.dim);
    }
    bf("norm.weight", G_NORM, 0.1f, 1.0f);
    if (!Q) {
        bf("tok_embeddings.weight", G_TOK, 0.0625f, 0.0f);
    } else {
        i8("tok_embeddings.weight", G_TOK, -127, 255);
        sc("tok_embeddings.scales", G_TOK + K_SCALES, 0.0625f / 127.0f);
        i8("output.weight", G_OUT, -127, 255);
        sc("output.scales", G_OUT + K_SCALES, (1.0f / std::sqrt(float(c.dim))) / 127.0f);
    }
    MC_CUDA_CHECK(cudaStreamSynchronize(m->dev->stream));
    m->finalized = false;
    MC_API_END
}

// Packs the staged reference-layout tensors into the streaming layouts (bit-exact, see mc_unpack_w4) and frees the staging.
mc_status mc_llama_finalize(mc_llama* m)
{
    MC_API_BEGIN
    nvtx_range nvtx_("mc_llama_finalize");
    use(m);
    if (m->cfg.quant && !m->finalized) {
        cudaStream_t s = m->dev->stream;
        auto pack = [&](dlinear& d, int epi) {
            if (!d.q8.p) return; // already packed
            const gemv_params sh = shape_only(d.N, d.K, m->cfg.head_dim);
            const uint64_t words = uint64_t(d.w.bytes / 4);
            pack_w4_kernel<<<gen_blocks(words), 256, 0, s>>>(d.w.as<uint32_t>(), d.q8.as<int8_t>(), sh, epi, m->pack_bad.as<int>());
            pack_w4_scales_kernel<<<gen_blocks(d.scales.bytes / 2), 256, 0, s>>>(d.scales.as<uint16_t>(), d.s32.as<float>(), sh, epi);
            MC_CUDA_CHECK(cudaGetLastError());
        };
        for (dlayer& ly : m->layers) {
            pack(ly.wqkv, EPI_QKV), pack(ly.wo, EPI_RESIDUAL), pack(ly.w13, EPI_SWIGLU), pack(ly.w2, EPI_RESIDUAL);
        }
        if (m->out.q8.p) {
            pack_w8_kernel<<<gen_blocks(m->out.w.bytes / 4), 256, 0, s>>>(m->out.w.as<uint32_t>(), m->out.q8.as<int8_t>(), shape_only(m->out.N, m->out.K, 0),
                                                                           EPI_NONE);
            MC_CUDA_CHECK(cudaGetLastError());
        }
        int bad = 0;
        MC_CUDA_CHECK(cudaMemcpyAsync(&bad, m->pack_bad.p, 4, cudaMemcpyDeviceToHost, s));
        MC_CUDA_CHECK(cudaStreamSynchronize(s));
        if (bad) {
            MC_CUDA_CHECK(cudaMemset(m->pack_bad.p, 0, 4));
            throw error(MC_ERR_INVALID, "quantised weights outside the int4 range [-8, 7] cannot be packed (quantization/lora.h stores int4-range values in int8)");
        }
        // resident bf16 image of every quantised matrix for the tensor-core prompt / batch path (the reference dequantises the whole
        // matrix on every call, quantization/lora.h:115, and caches the output projection, quantization/linear.h:50-53)
        if (!(m->cfg.flags & MC_LLAMA_NO_SHADOW)) {
            for (dlayer& ly : m->layers)
                for (dlinear* d : {&ly.wqkv, &ly.wo, &ly.w13, &ly.w2}) {
                    // rows [0, N): r(r(q) * r(s)); rows [N, N + R): the stacked adaptor A of this linear; zero rows up to a multiple of 32
                    const dbuf& A = d == &ly.wqkv ? ly.lora_a_qkv : (d == &ly.wo ? ly.lora_a_o : (d == &ly.w13 ? ly.lora_a_13 : ly.lora_a_2));
                    const uint32_t R = uint32_t(A.bytes / (size_t(d->K) * 2)), Rpad = (R + 31) & ~31u;
                    if (!d->q8.p) {
                        // image built by an earlier finalize: only the adaptor rows can have changed since (the packed weights are final)
                        if (d->wd.p && m->image_a_dirty)
                            MC_CUDA_CHECK(cudaMemcpyAsync(d->wd.as<uint16_t>() + size_t(d->N) * d->K, A.p, size_t(R) * d->K * 2, cudaMemcpyDeviceToDevice, s));
                        continue;
                    }
                    d->wd.alloc(size_t(d->N + Rpad) * d->K * 2);
                    tc::dequant_group(s, d->wd.as<uint16_t>(), d->q8.as<int8_t>(), d->s32.as<float>(), d->N, d->K, m->cfg.group_size);
                    MC_CUDA_CHECK(cudaMemsetAsync(d->wd.as<uint16_t>() + size_t(d->N) * d->K, 0, size_t(Rpad) * d->K * 2, s));
                    MC_CUDA_CHECK(cudaMemcpyAsync(d->wd.as<uint16_t>() + size_t(d->N) * d->K, A.p, size_t(R) * d->K * 2, cudaMemcpyDeviceToDevice, s));
                }
            if (m->out.q8.p) {
                m->out.wd.alloc(size_t(m->out.N) * m->out.K * 2);
                tc::dequant_group(s, m->out.wd.as<uint16_t>(), m->out.q8.as<int8_t>(), m->out.scales.as<float>(), m->out.N, m->out.K, m->out.K);
            }
            MC_CUDA_CHECK(cudaStreamSynchronize(s));
            m->image_a_dirty = false;
        }
        for (dlayer& ly : m->layers)
            for (dlinear* d : {&ly.wqkv, &ly.wo, &ly.w13, &ly.w2}) d->q8.release(), d->s32.release();
        m->out.q8.release();
    }
    m->finalized = true;
    MC_API_END
}

mc_status mc_llama_weight_bytes(mc_llama* m, uint64_t* streamed_per_step, uint64_t* resident)
{
    MC_API_BEGIN
    MC_REQUIRE(m, "null model");
    uint64_t stream = 0, res = 0;
    for (auto& ly : m->layers) {
        for (const dlinear* d : {&ly.wqkv, &ly.wo, &ly.w13, &ly.w2}) stream += d->stream_bytes();
        stream += ly.attn_norm.bytes + ly.ffn_norm.bytes + ly.lora_a_qkv.bytes + ly.lora_a_o.bytes + ly.lora_a_13.bytes + ly.lora_a_2.bytes;
    }
    const dlinear& hw = m->tied ? m->tok : m->out;
    stream += size_t(m->Vl) * m->cfg.dim * (hw.fmt == WF_BF16 ? 2 : 1) + hw.scales.bytes + m->norm.bytes;
    stream += size_t(m->cfg.dim) * (m->tok.fmt == WF_BF16 ? 2 : 1); // one embedding row
    res = stream + m->kcache.bytes + m->vcache.bytes;
    if (!m->tied) res += m->tok.stream_bytes();
    else res += m->tok.w.bytes - size_t(m->Vl) * m->cfg.dim * 2;
    // quantised models: the resident bf16 image used by the tensor-core prompt / batch path (not part of the batch-1 stream)
    for (auto& ly : m->layers)
        for (const dlinear* d : {&ly.wqkv, &ly.wo, &ly.w13, &ly.w2}) res += d->wd.bytes;
    res += m->out.wd.bytes;
    if (streamed_per_step) *streamed_per_step = stream;
    if (resident) *resident = res;
    MC_API_END
}

namespace {

// ---- tensor-core prefill (mc_prefill.cu) ---------------------------------------------------------------------------------
// Prompts of bf16 models on one GPU take the tcgen05 GEMM path: a chunk of up to kPfChunk positions goes through every
// block as ONE [rows, K] x [N, K]^T product per linear (weights read once per chunk, not once per 4 rows).
constexpr uint32_t kPfChunk = 2048;
uint32_t prefill_tc_min_rows()
{
    static const uint32_t v = [] {
        const char* e = getenv("MC_TC_PREFILL_MIN");
        return e ? uint32_t(atoi(e)) : 8u;
    }();
    return v;
}
bool prefill_tc_eligible(const mc_llama* m, uint32_t len)
{
    const mc_llama_config& c = m->cfg;
    static const bool env_off = getenv("MC_NO_TC_PREFILL") != nullptr;
    if (env_off || (c.flags & MC_LLAMA_NO_TC_PREFILL) || (c.tp_world != 1 && !tp_tc_ready(m)) || len < prefill_tc_min_rows()) return false;
    if (c.quant && (!m->layers[0].wqkv.wd.p || c.lora_rank % 2 != 0 || c.lora_rank > 16)) return false; // quantised: needs the bf16 image (mc_llama_finalize)
    if (c.head_dim != 64 && c.head_dim != 128) return false;
    const uint32_t D = c.dim, QO = m->Hl * c.head_dim, QKV = (m->Hl + 2 * m->KVl) * c.head_dim, F = m->Fl;
    return tc::gemm_supported(QKV, D, D, QKV) && tc::gemm_supported(D, QO, QO, D) && tc::gemm_supported(2 * F, D, D, F) && tc::gemm_supported(D, F, F, D);
}
void prefill_tc(mc_llama* m, uint32_t seq, const int32_t* ids, uint32_t len, uint32_t start_pos)
{
    const mc_llama_config& c = m->cfg;
    const uint32_t D = c.dim, hd = c.head_dim, H = m->Hl, KV = m->KVl, QO = H * hd, QKV = (H + 2 * KV) * hd, F = m->Fl;
    cudaStream_t s = m->dev->stream;
    const uint32_t cap = std::min<uint32_t>(kPfChunk, c.max_seq_len);
    const uint32_t rank = c.lora_rank, maxN = std::max(std::max(QKV, 2 * F), D);
    if (m->pf_rows < cap) {
        m->pf_ids.alloc(size_t(cap) * 4);
        m->pf_x.alloc(size_t(cap) * D * 2), m->pf_h.alloc(size_t(cap) * D * 2), m->pf_n.alloc(size_t(cap) * D * 2);
        m->pf_qkv.alloc(size_t(cap) * QKV * 2), m->pf_q.alloc(size_t(cap) * QO * 2), m->pf_attn.alloc(size_t(cap) * QO * 2);
        m->pf_z.alloc(size_t(cap) * F * 2);
        if (c.quant) m->pf_t.alloc(size_t(cap) * (maxN + tc_pad_rows(3 * rank)) * 2);
        m->pf_rows = cap;
    }
    uint16_t *x = m->pf_x.as<uint16_t>(), *h = m->pf_h.as<uint16_t>(), *n = m->pf_n.as<uint16_t>(), *qkv = m->pf_qkv.as<uint16_t>();
    uint16_t *q = m->pf_q.as<uint16_t>(), *attn = m->pf_attn.as<uint16_t>(), *z = m->pf_z.as<uint16_t>();
    const tc_scratch sc{m->pf_t.as<uint16_t>()};
    const bool tp = c.tp_world > 1;
    uint16_t* const x0 = x;
    int* err = m->errflag.as<int>();
    uint32_t launches = 0;
    launcher L{m, s, false};
    tc::set_pdl(!(c.flags & MC_LLAMA_NO_PDL));
    for (uint32_t t0 = 0; t0 < len; t0 += cap) {
        const uint32_t rows = std::min(cap, len - t0), pos0 = start_pos + t0;
        MC_CUDA_CHECK(cudaMemcpyAsync(m->pf_ids.p, ids + t0, size_t(rows) * 4, cudaMemcpyHostToDevice, s));
        // embedding gather (bf16 rows, or int8 rows with one scale: quantization/lora.h:160-170)
        x = x0;
        L.go(embed_kernel, dim3(rows), dim3(256), 0, x, D, (const void*)m->tok.w.p, (const float*)m->tok.scales.p, m->tok.fmt, D, c.vocab,
             (const int32_t*)m->pf_ids.as<int32_t>());
        for (uint32_t li = 0; li < c.n_layers; li++) {
            dlayer& ly = m->layers[li];
            uint16_t* kc = m->kcache.as<uint16_t>() + size_t(li) * kv_layer_elems(m);
            uint16_t* vc = m->vcache.as<uint16_t>() + size_t(li) * kv_layer_elems(m);
            launches += tc::rmsnorm_rows(s, n, x, ly.attn_norm.as<uint16_t>(), rows, D, c.norm_eps);
            launches += tc_linear(m, s, tc::GEMM_STORE, n, D, ly.wqkv, 3 * rank, 3, qkv, nullptr, rows, QKV, D, QKV, sc);
            launches += tc::rope_append(s, qkv, q, kc, vc, m->fcos.as<float>(), m->fsin.as<float>(), rows, seq, pos0, H, KV, hd, c.max_seq_len);
            launches += tc::prefill_attn(s, q, kc, vc, attn, rows, seq, pos0, H, KV, hd, c.max_seq_len, m->scale_bf16, m->key_begin);
            if (tp) {
                // row-parallel wo / w2: fp32 partial sums + all-reduce over NVLink (see tp_tc_row_parallel)
                h = tp_tc_linear(m, s, attn, QO, ly.wo, 0, x, m->pf_h.as<uint16_t>(), rows, launches);
            } else launches += tc_linear(m, s, tc::GEMM_RESIDUAL, attn, QO, ly.wo, rank, 1, h, x, rows, D, QO, D, sc);
            launches += tc::rmsnorm_rows(s, n, h, ly.ffn_norm.as<uint16_t>(), rows, D, c.norm_eps);
            launches += tc_linear(m, s, tc::GEMM_SWIGLU, n, D, ly.w13, 2 * rank, 2, z, nullptr, rows, 2 * F, D, F, sc);
            if (tp) {
                x = tp_tc_linear(m, s, z, F, ly.w2, 1, h, x0, rows, launches);
            } else launches += tc_linear(m, s, tc::GEMM_RESIDUAL, z, F, ly.w2, rank, 1, x, h, rows, D, F, D, sc);
        }
        if (t0 + rows >= len) {
            // only the last position is projected (nn/llama.h:128-133)
            const uint16_t* last = x + size_t(rows - 1) * D;
            uint16_t* dst = m->logits.as<uint16_t>() + size_t(seq) * m->Vl;
            if (c.quant) {
                qgemv_params qh{};
                qh.g = head_params(m, last, 1, dst);
                qh.scales = m->out.scales.p; // int8 rows with one scale per row, no adaptor (quantization/linear.h:17-64)
                qgemv_launch<WF_W8ROW, PRO_RMSNORM, EPI_NONE>(L, qh);
            } else {
                gemv_launch<PRO_RMSNORM, EPI_NONE>(L, head_params(m, last, 1, dst));
            }
            MC_CUDA_CHECK(cudaMemcpyAsync(m->hidden_save.as<uint16_t>() + size_t(seq) * D, last, size_t(D) * 2, cudaMemcpyDeviceToDevice, s));
        }
        MC_CUDA_CHECK(cudaStreamSynchronize(s)); // pf_ids is reused by the next chunk; `ids` may be pageable
    }
    (void)err;
    m->dev->launches.fetch_add(launches);
    int flag = 0;
    MC_CUDA_CHECK(cudaMemcpy(&flag, err, 4, cudaMemcpyDeviceToHost));
    if (flag) {
        cudaMemset(err, 0, 4);
        throw error(MC_ERR_RUNTIME, "prefill: a bounded wait of the tensor-core GEMM timed out (code " + std::to_string(flag) + ")");
    }
}

} // namespace

mc_status mc_llama_prefill(mc_llama* m, uint32_t seq, const int32_t* ids, uint32_t len, uint32_t start_pos)
{
    MC_API_BEGIN
    nvtx_range nvtx_("mc_llama_prefill");
    use(m);
    MC_REQUIRE(m->finalized, "mc_llama_finalize must be called before prefill");
    MC_REQUIRE(ids && len > 0, "prefill: empty input");
    MC_REQUIRE(seq < m->cfg.n_seqs, "prefill: sequence index out of range");
    // Prompts must fit the cache.  (Decode steps beyond it roll the cache like nn::sink_cache, nn/cache.h:183-204; a multi-token call at
    // start_pos >= max_seq_len would roll by `len` and, with the reference's chunk mask, see only itself -- feed such input token by token.)
    MC_REQUIRE(uint64_t(start_pos) + len <= m->cfg.max_seq_len, "prefill: start_pos + len exceeds max_seq_len (decode token by token beyond the cache)");
    for (uint32_t i = 0; i < len; i++) MC_REQUIRE(ids[i] >= 0 && uint32_t(ids[i]) < m->cfg.vocab, "prefill: token id out of range");
    // Quirk Q9 (nn/attention.h:283-299): make_causal_mask leaves the columns of the cached prefix at -inf whenever len > 1, so a
    // prompt chunk at start_pos > 0 attends only to itself in the reference.  The engine's default is the intended reading (the
    // prefix is visible: long prompts are chunked internally and multi-turn prompts see their history); MC_LLAMA_REF_CHUNK_MASK
    // reproduces the reference bit for bit.  Both are checked against the oracle (tests/test_gpu_prefill.py).
    struct key_begin_scope {
        mc_llama* m;
        ~key_begin_scope() { m->key_begin = 0; }
    } scope{m};
    m->key_begin = ((m->cfg.flags & MC_LLAMA_REF_CHUNK_MASK) && len > 1) ? start_pos : 0;
    if (prefill_tc_eligible(m, len)) {
        prefill_tc(m, seq, ids, len, start_pos);
        return MC_OK;
    }
    cudaStream_t s = m->dev->stream;
    launcher L{m, s, false};
    for (uint32_t t0 = 0; t0 < len; t0 += kMaxMB) {
        const uint32_t rows = std::min<uint32_t>(kMaxMB, len - t0);
        int32_t* st = m->pinned;
        for (uint32_t r = 0; r < rows; r++) {
            st[r] = ids[t0 + r];
            st[kMaxMB + r] = int32_t(start_pos + t0 + r);
            st[2 * kMaxMB + r] = int32_t(seq);
        }
        MC_CUDA_CHECK(cudaMemcpyAsync(m->ids.p, st, rows * 4, cudaMemcpyHostToDevice, s));
        MC_CUDA_CHECK(cudaMemcpyAsync(m->pos.p, st + kMaxMB, rows * 4, cudaMemcpyHostToDevice, s));
        MC_CUDA_CHECK(cudaMemcpyAsync(m->row_seq.p, st + 2 * kMaxMB, rows * 4, cudaMemcpyHostToDevice, s));
        const bool last = t0 + rows >= len;
        enqueue_rows(m, L, 0, rows, last ? 2 : 0, m->logits.as<uint16_t>() + size_t(seq) * m->Vl);
        if (last) {
            MC_CUDA_CHECK(cudaMemcpyAsync(m->hidden_save.as<uint16_t>() + size_t(seq) * m->cfg.dim,
                                          m->x.as<uint16_t>() + size_t(rows - 1) * m->cfg.dim, size_t(m->cfg.dim) * 2, cudaMemcpyDeviceToDevice, s));
        }
        MC_CUDA_CHECK(cudaStreamSynchronize(s)); // the pinned staging area is reused by the next chunk
    }
    MC_API_END
}

static void check_device_error(mc_llama* m, int flag)
{
    if (flag) {
        // an aborted launch leaves the arrival counters of the streaming kernel short of what the host has booked: start them over
        cudaMemset(m->bar.p, 0, m->bar.bytes);
        for (uint32_t& b : m->st_arrive_base) b = 0;
#ifdef ST_DEBUG_WHERE
        int w[148 * 4];
        cudaMemcpyFromSymbol(w, g_st_where, sizeof(w));
        for (int i = 0; i < 148; i++) fprintf(stderr, "CTA %d: mma %d epi %d prod %d\n", i, w[i * 4], w[i * 4 + 1], w[i * 4 + 2]);
        {
            std::vector<uint64_t> am(148 * 2);
            cudaMemcpy(am.data(), m->st_ll.as<uint64_t>() + m->st_off[6], am.size() * 8, cudaMemcpyDeviceToHost);
            uint64_t idw = 0;
            cudaMemcpy(&idw, m->st_ll.as<uint64_t>() + m->st_off[7], 8, cudaMemcpyDeviceToHost);
            fprintf(stderr, "tag_base %u ids tag %u\n", m->st_seq << 16, unsigned(idw >> 32));
            for (int i = 0; i < 148; i++) fprintf(stderr, "am[%d] tags %u %u\n", i, unsigned(am[2 * i] >> 32) - (m->st_seq << 16), unsigned(am[2 * i + 1] >> 32) - (m->st_seq << 16));
        }
#endif
        int info[4] = {0, 0, 0, 0};
        cudaMemcpy(info, m->errflag.p, 16, cudaMemcpyDeviceToHost);
        cudaMemset(m->errflag.p, 0, 16);
        cudaMemset(m->bar.p, 0, 4);
        cudaMemset(static_cast<char*>(m->errflag.p) + 8, 0x7f, 4);
        throw error(MC_ERR_RUNTIME, "persistent decode kernel: a bounded wait timed out (code " + std::to_string(flag) + ", wait kinds mask " + std::to_string(info[1]) +
                                        ", first: kind " + std::to_string(info[2] >> 20) + " CTA " + std::to_string((info[2] >> 10) & 1023) + " thread " +
                                        std::to_string(info[2] & 1023) + "; CTAs not co-resident?)");
    }
}

static void stage_decode_inputs(mc_llama* m, uint32_t n, const int32_t* ids, const int32_t* pos, uint32_t steps)
{
    MC_REQUIRE(n >= 1 && n <= m->cfg.n_seqs, "decode: number of sequences out of range");
    MC_REQUIRE(ids && pos, "decode: null ids/pos");
    const uint32_t R = m->max_rows;
    int32_t* st = m->pinned + 4 * R + 64; // staging image of io_in: ids | pos | row_seq | step counter
    uint64_t last = 0;
    for (uint32_t r = 0; r < n; r++) {
        MC_REQUIRE(ids[r] >= 0 && uint32_t(ids[r]) < m->cfg.vocab, "decode: token id out of range");
        MC_REQUIRE(pos[r] >= 0 && uint64_t(pos[r]) + steps < (1u << 30), "decode: position out of range");
        st[r] = ids[r], st[R + r] = pos[r], st[2 * R + r] = int32_t(r);
        last = std::max<uint64_t>(last, uint64_t(pos[r]) + steps - 1);
    }
    // positions beyond the cache: the sink-cache roll runs at the head of every step of this call (nn/cache.h:183-204); RoPE keeps the
    // absolute position, so the tables must reach it
    m->sink_roll = last >= m->cfg.max_seq_len;
    m->st_last_pos = uint32_t(last);
    if (m->sink_roll) MC_REQUIRE(m->cfg.tp_world == 1, "decode: positions beyond max_seq_len are not supported under tensor parallelism");
    if (last >= m->rope_rows) {
        MC_CUDA_CHECK(cudaStreamSynchronize(m->dev->stream));
        uint32_t rows = m->rope_rows;
        while (rows <= last) rows *= 2;
        build_rope_tables(m, rows);
        for (auto& g : m->graphs) cudaGraphExecDestroy(g.second); // the captured launches hold the old table pointers
        m->graphs.clear();
    }
    st[3 * R] = 0; // step counter
    MC_CUDA_CHECK(cudaMemcpyAsync(m->io_in.p, st, (size_t(3) * R + 1) * 4, cudaMemcpyHostToDevice, m->dev->stream));
}

mc_status mc_llama_decode(mc_llama* m, uint32_t n, const int32_t* ids, const int32_t* pos, const float* uniforms,
                          const mc_sampler_config* sampler, int32_t* out_ids)
{
    MC_API_BEGIN
    nvtx_range nvtx_("mc_llama_decode");
    use(m);
    MC_REQUIRE(m->finalized, "mc_llama_finalize must be called before decode");
    MC_REQUIRE(out_ids, "decode: null output");
    mc_sampler_config sc{};
    if (sampler) sc = *sampler;
    stage_decode_inputs(m, n, ids, pos, 1);
    if (sc.mode == 1) {
        MC_REQUIRE(uniforms, "decode: the multinomial sampler needs one injected uniform per sequence");
        MC_CUDA_CHECK(cudaMemcpyAsync(m->uniforms.p, uniforms, n * 4, cudaMemcpyHostToDevice, m->dev->stream));
    }
    run_decode_step(m, n, sc, 0);
    cudaStream_t s = m->dev->stream;
    int32_t* st_out = m->pinned + 8 * m->max_rows + 128; // image of io_out: error flag (16 B) | first n log entries
    MC_CUDA_CHECK(cudaMemcpyAsync(st_out, m->io_out.p, 16 + n * 4, cudaMemcpyDeviceToHost, s));
    MC_CUDA_CHECK(cudaStreamSynchronize(s));
    check_device_error(m, st_out[0]);
    memcpy(out_ids, st_out + 4, n * 4);
    MC_API_END
}

mc_status mc_llama_decode_loop(mc_llama* m, uint32_t n, const int32_t* first_ids, const int32_t* first_pos, uint32_t steps,
                               const float* uniforms, const mc_sampler_config* sampler, int32_t* out_ids, float* elapsed_ms)
{
    MC_API_BEGIN
    nvtx_range nvtx_("mc_llama_decode_loop");
    use(m);
    MC_REQUIRE(m->finalized, "mc_llama_finalize must be called before decode");
    MC_REQUIRE(steps >= 1 && steps <= kMaxLogSteps, "decode_loop: steps out of range");
    mc_sampler_config sc{};
    if (sampler) sc = *sampler;
    stage_decode_inputs(m, n, first_ids, first_pos, steps);
    if (sc.mode == 1) {
        MC_REQUIRE(uniforms, "decode_loop: the multinomial sampler needs steps*n injected uniforms");
        MC_CUDA_CHECK(cudaMemcpyAsync(m->uniforms.p, uniforms, size_t(steps) * n * 4, cudaMemcpyHostToDevice, m->dev->stream));
    }
    cudaStream_t s = m->dev->stream;
    struct event_pair {
        cudaEvent_t a = nullptr, b = nullptr;
        ~event_pair()
        {
            if (a) cudaEventDestroy(a);
            if (b) cudaEventDestroy(b);
        }
    } ev;
    MC_CUDA_CHECK(cudaEventCreate(&ev.a));
    MC_CUDA_CHECK(cudaEventCreate(&ev.b));
    const cudaEvent_t e0 = ev.a, e1 = ev.b;
    const bool stream = stream_eligible(m, n, sc);
    if (!stream && !(m->cfg.flags & MC_LLAMA_NO_GRAPH)) decode_graph(m, n, sc, 1); // instantiate outside the timed region
    MC_CUDA_CHECK(cudaEventRecord(e0, s));
    if (stream) {
        // the persistent kernel loops over the steps itself: the weight stream of step i+1 starts under the sampler tail of step i
        launcher L{m, s, false};
        const uint32_t per_launch = std::max<uint32_t>(1, std::min<uint32_t>(512, 60000 / (m->cfg.n_layers * 5 + 1)));
        for (uint32_t done = 0; done < steps; done += per_launch) launch_stream(m, L, n, 1, std::min(per_launch, steps - done));
        m->launches_per_step = 1;
    } else {
        for (uint32_t i = 0; i < steps; i++) run_decode_step(m, n, sc, 1);
    }
    MC_CUDA_CHECK(cudaEventRecord(e1, s));
    MC_CUDA_CHECK(cudaStreamSynchronize(s));
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (elapsed_ms) *elapsed_ms = ms;
    int errv = 0;
    MC_CUDA_CHECK(cudaMemcpy(&errv, m->errflag.p, 4, cudaMemcpyDeviceToHost));
    check_device_error(m, errv);
    if (out_ids) MC_CUDA_CHECK(cudaMemcpy(out_ids, m->out_log.p, size_t(steps) * n * 4, cudaMemcpyDeviceToHost));
    MC_API_END
}

mc_status mc_llama_logits(mc_llama* m, uint32_t seq, void* host_bf16, size_t nbytes)
{
    MC_API_BEGIN
    use(m);
    MC_REQUIRE(seq < m->cfg.n_seqs && host_bf16, "bad arguments");
    MC_REQUIRE(nbytes == size_t(m->Vl) * 2, "logits: expected vocab*2 bytes");
    MC_CUDA_CHECK(cudaStreamSynchronize(m->dev->stream));
    MC_CUDA_CHECK(cudaMemcpy(host_bf16, m->logits.as<uint16_t>() + size_t(seq) * m->Vl, nbytes, cudaMemcpyDeviceToHost));
    MC_API_END
}

mc_status mc_llama_hidden(mc_llama* m, uint32_t seq, void* host_bf16, size_t nbytes)
{
    MC_API_BEGIN
    use(m);
    MC_REQUIRE(seq < m->cfg.n_seqs && host_bf16, "bad arguments");
    MC_REQUIRE(nbytes == size_t(m->cfg.dim) * 2, "hidden: expected dim*2 bytes");
    MC_CUDA_CHECK(cudaStreamSynchronize(m->dev->stream));
    MC_CUDA_CHECK(cudaMemcpy(host_bf16, m->hidden_save.as<uint16_t>() + size_t(seq) * m->cfg.dim, nbytes, cudaMemcpyDeviceToHost));
    MC_API_END
}

// Copies rows [0, n_pos) of the (seq, layer) cache in the reference layout [pos, n_kv_heads, head_dim]
// (nn/cache.h:150-160); the device layout is [kv head][pos][hd].
mc_status mc_llama_cache(mc_llama* m, uint32_t seq, uint32_t layer, int which, uint32_t n_pos, void* host_bf16, size_t nbytes)
{
    MC_API_BEGIN
    use(m);
    const mc_llama_config& c = m->cfg;
    MC_REQUIRE(seq < c.n_seqs && layer < c.n_layers && host_bf16 && n_pos <= c.max_seq_len, "bad arguments");
    MC_REQUIRE(nbytes == size_t(n_pos) * m->KVl * c.head_dim * 2, "cache: unexpected size");
    MC_CUDA_CHECK(cudaStreamSynchronize(m->dev->stream));
    const uint16_t* base = (which ? m->vcache : m->kcache).as<uint16_t>() + size_t(layer) * kv_layer_elems(m) +
                           size_t(seq) * m->KVl * c.max_seq_len * c.head_dim;
    for (uint32_t kvh = 0; kvh < m->KVl; kvh++) {
        MC_CUDA_CHECK(cudaMemcpy2D(static_cast<uint16_t*>(host_bf16) + size_t(kvh) * c.head_dim, size_t(m->KVl) * c.head_dim * 2,
                                   base + size_t(kvh) * c.max_seq_len * c.head_dim, size_t(c.head_dim) * 2, size_t(c.head_dim) * 2, n_pos,
                                   cudaMemcpyDeviceToHost));
    }
    MC_API_END
}

// One ungraphed decode step of sequences [0,n) at their current device-side ids/pos with a CUDA event before
// every launch: us[i] = time from launch i to launch i+1 (the last entry ends at step completion).
mc_status mc_llama_profile_step(mc_llama* m, uint32_t n, float* us, uint32_t cap, uint32_t* count)
{
    const uint32_t n_rows = n;
    MC_API_BEGIN
    use(m);
    MC_REQUIRE(m->finalized && us && count, "bad arguments");
    MC_REQUIRE(n >= 1 && n <= m->cfg.n_seqs, "profile_step: number of sequences out of range");
    mc_sampler_config sc{};
    if (stream_eligible(m, n, sc)) {
        // streaming kernel: four globaltimer stamps per (CTA, phase) = phase entry, (unused), input staged, tiles consumed;
        // us[(cta * phases + k) * 4 + j] = microseconds since the earliest stamp of the launch
        const uint32_t phases = m->cfg.n_layers * 5 + 1, G = m->st_grid;
        const size_t n = size_t(G) * phases * 4 + 512; // + per-block stamps of CTA 0 in the vocabulary projection
        if (!m->st_timing.p) m->st_timing.alloc(n * 8);
        MC_CUDA_CHECK(cudaMemsetAsync(m->st_timing.p, 0, m->st_timing.bytes, m->dev->stream));
        MC_CUDA_CHECK(cudaMemsetAsync(m->step_counter.p, 0, 4, m->dev->stream));
        m->st_timing_on = true;
        launcher L{m, m->dev->stream, false};
        try {
            launch_stream(m, L, n_rows, 0, 1);
        } catch (...) {
            m->st_timing_on = false;
            throw;
        }
        m->st_timing_on = false;
        MC_CUDA_CHECK(cudaStreamSynchronize(m->dev->stream));
        std::vector<unsigned long long> t(n);
        MC_CUDA_CHECK(cudaMemcpy(t.data(), m->st_timing.p, n * 8, cudaMemcpyDeviceToHost));
        unsigned long long t0 = ~0ull;
        for (auto v : t)
            if (v && v < t0) t0 = v;
        *count = uint32_t(std::min<size_t>(n, cap));
        for (size_t i = 0; i < n && i < cap; i++) us[i] = t[i] ? float(t[i] - t0) * 1e-3f : 0.0f;
        return MC_OK;
    }
    std::vector<cudaEvent_t> ev;
    launcher L{m, m->dev->stream, false};
    L.events = &ev;
    MC_CUDA_CHECK(cudaMemsetAsync(m->step_counter.p, 0, 4, m->dev->stream));
    enqueue_decode_step(m, L, n, sc, 0);
    L.mark();
    MC_CUDA_CHECK(cudaStreamSynchronize(m->dev->stream));
    *count = uint32_t(ev.size() - 1);
    for (size_t i = 0; i + 1 < ev.size(); i++) {
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
        if (i < cap) us[i] = ms * 1e3f;
    }
    for (auto e : ev) cudaEventDestroy(e);
    MC_API_END
}

// ---- tensor parallel wiring: every rank exports one IPC handle, all ranks import all handles ------------------------------------
mc_status mc_llama_tp_export(mc_llama* m, void* handle, size_t cap)
{
    MC_API_BEGIN
    use(m);
    MC_REQUIRE(handle && cap >= sizeof(cudaIpcMemHandle_t), "tp_export: handle buffer must hold 64 bytes");
    MC_REQUIRE(m->cfg.tp_world > 1, "tp_export: the model is not tensor parallel");
    cudaIpcMemHandle_t h;
    MC_CUDA_CHECK(cudaIpcGetMemHandle(&h, m->tp_region.p));
    memcpy(handle, &h, sizeof(h));
    MC_API_END
}

mc_status mc_llama_tp_connect(mc_llama* m, const void* handles, size_t nbytes)
{
    MC_API_BEGIN
    use(m);
    const uint32_t T = m->cfg.tp_world;
    MC_REQUIRE(T > 1, "tp_connect: the model is not tensor parallel");
    MC_REQUIRE(handles && nbytes == size_t(T) * sizeof(cudaIpcMemHandle_t), "tp_connect: expected world x 64 bytes of handles");
    for (uint32_t k = 0; k < T; k++) {
        if (k == m->cfg.tp_rank || m->tp_peer_base[k]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + size_t(k) * sizeof(h), sizeof(h));
        void* p = nullptr;
        MC_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        m->tp_peer_base[k] = p;
    }
    m->tp_connected = true;
    MC_API_END
}

mc_status mc_nccl_unique_id(void* id, size_t cap)
{
    MC_API_BEGIN
    MC_REQUIRE(id && cap >= sizeof(nccl_uid), "nccl_unique_id: the buffer must hold 128 bytes");
    nccl_uid u;
    nccl_check(nccl().get_unique_id(&u), "ncclGetUniqueId");
    memcpy(id, &u, sizeof(u));
    MC_API_END
}

mc_status mc_llama_tp_use_nccl(mc_llama* m, const void* id, size_t nbytes)
{
    MC_API_BEGIN
    use(m);
    MC_REQUIRE(m->cfg.tp_world > 1, "tp_use_nccl: the model is not tensor parallel");
    MC_REQUIRE(m->cfg.quant == 0, "tp_use_nccl: the comparator covers bf16 models only");
    MC_REQUIRE(id && nbytes == sizeof(nccl_uid), "tp_use_nccl: expected the 128 bytes of mc_nccl_unique_id");
    MC_REQUIRE(!m->nccl_comm, "tp_use_nccl: already set");
    MC_REQUIRE(m->tp_host_epoch == 0 && m->graphs.empty(), "tp_use_nccl: call it before the first step");
    nccl_uid u;
    memcpy(&u, id, sizeof(u));
    void* comm = nullptr;
    nccl_check(nccl().comm_init_rank(&comm, int(m->cfg.tp_world), u, int(m->cfg.tp_rank)), "ncclCommInitRank"); // collective: every rank calls it
    m->nccl_comm = comm;
    MC_API_END
}

mc_status mc_llama_launches_per_step(mc_llama* m, uint32_t* kernels)
{
    MC_API_BEGIN
    MC_REQUIRE(m && kernels, "bad arguments");
    *kernels = m->launches_per_step;
    MC_API_END
}

// ---- stand-alone sampler (parity tests against the oracle's sample_default) ------------------------------------------------
mc_status mc_sample_default(mc_device* dev, mc_buffer* logits_bf16, uint32_t rows, uint32_t vocab, const mc_sampler_config* cfg, const float* uniforms,
                            int32_t* topk_idx, uint16_t* probs_sorted, int32_t* probs_idx, int32_t* choice, int32_t* token)
{
    MC_API_BEGIN
    MC_REQUIRE(dev && logits_bf16 && cfg && uniforms && rows >= 1, "bad arguments");
    MC_CUDA_CHECK(cudaSetDevice(dev->ordinal));
    MC_REQUIRE(logits_bf16->size >= size_t(rows) * vocab * 2, "sample_default: logits buffer too small");
    const uint32_t k = cfg->top_k, blocks = (vocab + kSampleSlice - 1) / kSampleSlice;
    dbuf cand, du, d_topk, d_ps, d_pi, d_ch, d_tok;
    struct guard {
        std::vector<dbuf*> b;
        ~guard()
        {
            for (auto* x : b) x->release();
        }
    } g{{&cand, &du, &d_topk, &d_ps, &d_pi, &d_ch, &d_tok}};
    cand.alloc(size_t(rows) * blocks * kSampleKeep * 8), du.alloc(rows * 4);
    d_topk.alloc(size_t(rows) * kSampleMaxK * 4), d_ps.alloc(size_t(rows) * kSampleMaxK * 2), d_pi.alloc(size_t(rows) * kSampleMaxK * 4);
    d_ch.alloc(rows * 4), d_tok.alloc(rows * 4);
    MC_CUDA_CHECK(cudaMemcpyAsync(du.p, uniforms, rows * 4, cudaMemcpyHostToDevice, dev->stream));
    mc_llama shim;
    shim.dev = dev;
    launcher L{&shim, dev->stream, false};
    sample_params sp{};
    sp.logits = static_cast<const uint16_t*>(logits_bf16->dptr), sp.ld = vocab, sp.index_base = 0, sp.uniforms = du.as<float>();
    sp.topk_idx = d_topk.as<int32_t>(), sp.probs_sorted = d_ps.as<uint16_t>(), sp.probs_idx = d_pi.as<int32_t>();
    sp.choice = d_ch.as<int32_t>(), sp.token = d_tok.as<int32_t>();
    launch_sampler(L, sp, *cfg, rows, vocab, cand.as<unsigned long long>());
    MC_CUDA_CHECK(cudaStreamSynchronize(dev->stream));
    if (topk_idx) MC_CUDA_CHECK(cudaMemcpy(topk_idx, d_topk.p, size_t(rows) * k * 4, cudaMemcpyDeviceToHost));
    if (probs_sorted) MC_CUDA_CHECK(cudaMemcpy(probs_sorted, d_ps.p, size_t(rows) * k * 2, cudaMemcpyDeviceToHost));
    if (probs_idx) MC_CUDA_CHECK(cudaMemcpy(probs_idx, d_pi.p, size_t(rows) * k * 4, cudaMemcpyDeviceToHost));
    if (choice) MC_CUDA_CHECK(cudaMemcpy(choice, d_ch.p, rows * 4, cudaMemcpyDeviceToHost));
    if (token) MC_CUDA_CHECK(cudaMemcpy(token, d_tok.p, rows * 4, cudaMemcpyDeviceToHost));
    MC_API_END
}

// ---- stand-alone linear (roofline measurement, parity tests) ------------------------------------------------------------
mc_status mc_gemm_bf16(mc_device* dev, mc_buffer* y, mc_buffer* x, mc_buffer* w, mc_buffer* res, uint32_t M, uint32_t N, uint32_t K, int mode, uint32_t iters,
                       float* elapsed_ms)
{
    MC_API_BEGIN
    MC_REQUIRE(dev && y && x && w, "bad arguments");
    MC_CUDA_CHECK(cudaSetDevice(dev->ordinal));
    MC_REQUIRE(mode == tc::GEMM_STORE || mode == tc::GEMM_RESIDUAL || mode == tc::GEMM_SWIGLU, "gemm_bf16: mode must be 0 (store), 2 (residual) or 3 (swiglu)");
    MC_REQUIRE(M >= 1 && iters >= 1 && tc::gemm_supported(N, K, K, N), "gemm_bf16: K must be a multiple of 64 and N a multiple of 32");
    const uint32_t ncols = mode == tc::GEMM_SWIGLU ? N / 2 : N;
    MC_REQUIRE(w->size >= size_t(N) * K * 2 && x->size >= size_t(M) * K * 2 && y->size >= size_t(M) * ncols * 2, "gemm_bf16: buffer too small");
    MC_REQUIRE(mode != tc::GEMM_RESIDUAL || (res && res->size >= size_t(M) * N * 2), "gemm_bf16: residual buffer missing or too small");
    tc::set_pdl(false);
    dbuf err;
    err.alloc(4);
    cudaStream_t s = dev->stream;
    MC_CUDA_CHECK(cudaMemsetAsync(err.p, 0, 4, s));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    MC_CUDA_CHECK(cudaEventCreate(&e0));
    MC_CUDA_CHECK(cudaEventCreate(&e1));
    MC_CUDA_CHECK(cudaEventRecord(e0, s));
    for (uint32_t i = 0; i < iters; i++)
        tc::gemm(s, dev->prop.multiProcessorCount, mode, static_cast<const uint16_t*>(x->dptr), K, static_cast<const uint16_t*>(w->dptr), static_cast<uint16_t*>(y->dptr),
                 res ? static_cast<const uint16_t*>(res->dptr) : nullptr, M, N, K, ncols, err.as<int>());
    MC_CUDA_CHECK(cudaEventRecord(e1, s));
    int flag = 0;
    cudaError_t ce = cudaMemcpyAsync(&flag, err.p, 4, cudaMemcpyDeviceToHost, s);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    float ms = 0.0f;
    if (ce == cudaSuccess) ce = cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    err.release();
    MC_CUDA_CHECK(ce);
    dev->launches.fetch_add(iters);
    if (flag) throw error(MC_ERR_RUNTIME, "gemm_bf16: a bounded wait of the tensor-core GEMM timed out (code " + std::to_string(flag) + ")");
    if (elapsed_ms) *elapsed_ms = ms;
    MC_API_END
}

mc_status mc_linear_bf16(mc_device* dev, mc_buffer* y, mc_buffer* x, mc_buffer* w, uint32_t M, uint32_t N, uint32_t K)
{
    MC_API_BEGIN
    MC_REQUIRE(dev && y && x && w, "bad arguments");
    MC_CUDA_CHECK(cudaSetDevice(dev->ordinal));
    MC_REQUIRE(w->size >= size_t(N) * K * 2 && x->size >= size_t(M) * K * 2 && y->size >= size_t(M) * N * 2, "linear_bf16: buffer too small");
    if (M > uint32_t(kMaxMB)) return mc_gemm_bf16(dev, y, x, w, nullptr, M, N, K, 0, 1, nullptr); // many rows: the tensor-core GEMM
    MC_REQUIRE(M >= 1, "linear_bf16: M must be positive");
    MC_REQUIRE(N % 2 == 0 && K % 256 == 0, "linear_bf16: N must be even and K a multiple of 256");
    mc_llama shim;
    shim.dev = dev;
    launcher L{&shim, dev->stream, false};
    gemv_params p{};
    p.W = w->dptr, p.N = N, p.K = K, p.rows = M;
    p.x = static_cast<const uint16_t*>(x->dptr), p.ldx = K;
    p.y = static_cast<uint16_t*>(y->dptr), p.ldy = N;
    gemv_launch<PRO_NONE, EPI_NONE>(L, p);
    MC_API_END
}

// ---- stand-alone attention / embedding kernels (isolated parity tests against the oracle) --------------------------------------
namespace {
struct row_arrays {
    dbuf seq, pos;
    row_arrays(cudaStream_t s, uint32_t rows, const int32_t* row_seq, const int32_t* row_pos)
    {
        seq.alloc(size_t(rows) * 4), pos.alloc(size_t(rows) * 4);
        MC_CUDA_CHECK(cudaMemcpyAsync(seq.p, row_seq, size_t(rows) * 4, cudaMemcpyHostToDevice, s));
        MC_CUDA_CHECK(cudaMemcpyAsync(pos.p, row_pos, size_t(rows) * 4, cudaMemcpyHostToDevice, s));
    }
    ~row_arrays() { seq.release(), pos.release(); }
};
float stored_scale(uint32_t hd) { return bf16_bits_to_f32(f32_to_bf16_bits(1.0f / std::sqrt(float(hd)))); }
} // namespace

mc_status mc_attn_decode(mc_device* dev, mc_buffer* out, mc_buffer* q, mc_buffer* kcache, mc_buffer* vcache, uint32_t rows, const int32_t* row_seq,
                         const int32_t* row_pos, uint32_t n_seqs, uint32_t H, uint32_t KV, uint32_t hd, uint32_t max_seq, int kernel)
{
    MC_API_BEGIN
    MC_REQUIRE(dev && out && q && kcache && vcache && row_seq && row_pos && rows >= 1, "bad arguments");
    MC_CUDA_CHECK(cudaSetDevice(dev->ordinal));
    MC_REQUIRE(hd == 64 || hd == 128, "attn_decode: head_dim must be 64 or 128");
    MC_REQUIRE(KV >= 1 && H % KV == 0 && max_seq >= 1 && n_seqs >= 1, "attn_decode: bad head / cache geometry");
    const size_t cache_bytes = size_t(n_seqs) * KV * max_seq * hd * 2, act_bytes = size_t(rows) * H * hd * 2;
    MC_REQUIRE(kcache->size >= cache_bytes && vcache->size >= cache_bytes && q->size >= act_bytes && out->size >= act_bytes, "attn_decode: buffer too small");
    for (uint32_t r = 0; r < rows; r++)
        MC_REQUIRE(row_seq[r] >= 0 && uint32_t(row_seq[r]) < n_seqs && row_pos[r] >= 0 && uint32_t(row_pos[r]) < max_seq, "attn_decode: row sequence / position out of range");
    cudaStream_t s = dev->stream;
    row_arrays ra(s, rows, row_seq, row_pos);
    const float scale = stored_scale(hd);
    if (kernel == 1) {
        tc::set_pdl(false);
        tc::decode_attn_gqa(s, static_cast<const uint16_t*>(q->dptr), static_cast<uint16_t*>(kcache->dptr), static_cast<uint16_t*>(vcache->dptr),
                            static_cast<uint16_t*>(out->dptr), rows, ra.seq.as<int32_t>(), ra.pos.as<int32_t>(), H, KV, hd, max_seq, scale);
    } else {
        MC_REQUIRE(kernel == 0, "attn_decode: kernel must be 0 (cluster-split) or 1 (grouped-query)");
        mc_llama shim;
        shim.dev = dev;
        shim.cfg.head_dim = hd, shim.cfg.max_seq_len = max_seq;
        launcher L{&shim, s, false};
        attn_params a{};
        a.q = static_cast<const uint16_t*>(q->dptr), a.kcache = static_cast<const uint16_t*>(kcache->dptr), a.vcache = static_cast<const uint16_t*>(vcache->dptr);
        a.out = static_cast<uint16_t*>(out->dptr), a.row_seq = ra.seq.as<int32_t>(), a.row_pos = ra.pos.as<int32_t>();
        a.n_heads = H, a.n_kv_heads = KV, a.max_seq = max_seq, a.scale = scale;
        const size_t smem = attn_smem(&shim, kAttnCluster);
        if (smem > 48 * 1024) {
            if (hd == 64) MC_CUDA_CHECK(cudaFuncSetAttribute(attn_decode_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
            else MC_CUDA_CHECK(cudaFuncSetAttribute(attn_decode_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        }
        if (hd == 64) L.go_cluster(attn_decode_kernel<64>, dim3(H * kAttnCluster, rows), dim3(256), smem, kAttnCluster, a);
        else L.go_cluster(attn_decode_kernel<128>, dim3(H * kAttnCluster, rows), dim3(256), smem, kAttnCluster, a);
    }
    MC_CUDA_CHECK(cudaStreamSynchronize(s));
    dev->launches.fetch_add(kernel == 1 ? 1 : 0);
    MC_API_END
}

mc_status mc_attn_prefill(mc_device* dev, mc_buffer* out, mc_buffer* q, mc_buffer* kcache, mc_buffer* vcache, uint32_t rows, uint32_t seq, uint32_t start_pos,
                          uint32_t key_begin, uint32_t n_seqs, uint32_t H, uint32_t KV, uint32_t hd, uint32_t max_seq)
{
    MC_API_BEGIN
    MC_REQUIRE(dev && out && q && kcache && vcache && rows >= 1, "bad arguments");
    MC_CUDA_CHECK(cudaSetDevice(dev->ordinal));
    MC_REQUIRE(hd == 64 || hd == 128, "attn_prefill: head_dim must be 64 or 128");
    MC_REQUIRE(KV >= 1 && H % KV == 0 && seq < n_seqs && uint64_t(start_pos) + rows <= max_seq && key_begin <= start_pos, "attn_prefill: bad geometry");
    const size_t cache_bytes = size_t(n_seqs) * KV * max_seq * hd * 2, act_bytes = size_t(rows) * H * hd * 2;
    MC_REQUIRE(kcache->size >= cache_bytes && vcache->size >= cache_bytes && q->size >= act_bytes && out->size >= act_bytes, "attn_prefill: buffer too small");
    tc::set_pdl(false);
    tc::prefill_attn(dev->stream, static_cast<const uint16_t*>(q->dptr), static_cast<const uint16_t*>(kcache->dptr), static_cast<const uint16_t*>(vcache->dptr),
                     static_cast<uint16_t*>(out->dptr), rows, seq, start_pos, H, KV, hd, max_seq, stored_scale(hd), key_begin);
    MC_CUDA_CHECK(cudaStreamSynchronize(dev->stream));
    dev->launches.fetch_add(1);
    MC_API_END
}

mc_status mc_embed_rows(mc_device* dev, mc_buffer* out, mc_buffer* table, mc_buffer* row_scales, const int32_t* ids, uint32_t rows, uint32_t D, uint32_t vocab)
{
    MC_API_BEGIN
    MC_REQUIRE(dev && out && table && ids && rows >= 1 && D % 8 == 0, "bad arguments");
    MC_CUDA_CHECK(cudaSetDevice(dev->ordinal));
    const bool q8 = row_scales != nullptr;
    MC_REQUIRE(table->size >= size_t(vocab) * D * (q8 ? 1 : 2) && out->size >= size_t(rows) * D * 2, "embed_rows: buffer too small");
    MC_REQUIRE(!q8 || row_scales->size >= size_t(vocab) * 4, "embed_rows: one fp32 scale per table row expected");
    for (uint32_t r = 0; r < rows; r++) MC_REQUIRE(ids[r] >= 0 && uint32_t(ids[r]) < vocab, "embed_rows: token id out of range");
    dbuf dids;
    dids.alloc(size_t(rows) * 4);
    struct guard {
        dbuf& b;
        ~guard() { b.release(); }
    } g{dids};
    MC_CUDA_CHECK(cudaMemcpyAsync(dids.p, ids, size_t(rows) * 4, cudaMemcpyHostToDevice, dev->stream));
    embed_kernel<<<rows, 256, 0, dev->stream>>>(static_cast<uint16_t*>(out->dptr), D, table->dptr, q8 ? static_cast<const float*>(row_scales->dptr) : nullptr,
                                                 q8 ? WF_W8ROW : WF_BF16, D, vocab, dids.as<int32_t>());
    MC_CUDA_CHECK(cudaGetLastError());
    MC_CUDA_CHECK(cudaStreamSynchronize(dev->stream));
    dev->launches.fetch_add(1);
    MC_API_END
}

mc_status mc_pack_w4(mc_device* dev, mc_buffer* w4, mc_buffer* scales_packed, mc_buffer* q8, mc_buffer* scales_f32, uint32_t N, uint32_t K)
{
    MC_API_BEGIN
    MC_REQUIRE(dev && w4 && q8, "bad arguments");
    MC_CUDA_CHECK(cudaSetDevice(dev->ordinal));
    MC_REQUIRE(N % 2 == 0 && K % 256 == 0, "pack_w4: N must be even and K a multiple of 256");
    MC_REQUIRE(q8->size >= size_t(N) * K && w4->size >= w4_bytes(N, K), "pack_w4: buffer too small");
    int* bad = nullptr;
    MC_CUDA_CHECK(cudaMalloc(&bad, 4));
    MC_CUDA_CHECK(cudaMemsetAsync(bad, 0, 4, dev->stream));
    const gemv_params sh = shape_only(N, K, 0);
    pack_w4_kernel<<<gen_blocks(w4_bytes(N, K) / 4), 256, 0, dev->stream>>>(static_cast<uint32_t*>(w4->dptr), static_cast<const int8_t*>(q8->dptr), sh, EPI_NONE, bad);
    if (scales_packed && scales_f32) {
        MC_REQUIRE(scales_f32->size >= size_t(N) * (K / 32) * 4 && scales_packed->size >= w4_scale_bytes(N, K), "pack_w4: scale buffer too small");
        pack_w4_scales_kernel<<<gen_blocks(w4_scale_bytes(N, K) / 2), 256, 0, dev->stream>>>(static_cast<uint16_t*>(scales_packed->dptr),
                                                                                             static_cast<const float*>(scales_f32->dptr), sh, EPI_NONE);
    }
    int h = 0;
    cudaError_t e = cudaMemcpyAsync(&h, bad, 4, cudaMemcpyDeviceToHost, dev->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(dev->stream);
    cudaFree(bad);
    MC_CUDA_CHECK(e);
    MC_CUDA_CHECK(cudaGetLastError());
    if (h) throw error(MC_ERR_INVALID, "pack_w4: weight outside the int4 range [-8, 7]");
    dev->launches.fetch_add(2);
    MC_API_END
}

mc_status mc_unpack_w4(mc_device* dev, mc_buffer* q8, mc_buffer* w4, uint32_t N, uint32_t K)
{
    MC_API_BEGIN
    MC_REQUIRE(dev && w4 && q8, "bad arguments");
    MC_CUDA_CHECK(cudaSetDevice(dev->ordinal));
    MC_REQUIRE(N % 2 == 0 && K % 256 == 0, "unpack_w4: N must be even and K a multiple of 256");
    MC_REQUIRE(q8->size >= size_t(N) * K && w4->size >= w4_bytes(N, K), "unpack_w4: buffer too small");
    unpack_w4_kernel<<<gen_blocks(w4_bytes(N, K) / 4), 256, 0, dev->stream>>>(static_cast<int8_t*>(q8->dptr), static_cast<const uint32_t*>(w4->dptr),
                                                                               shape_only(N, K, 0), EPI_NONE);
    MC_CUDA_CHECK(cudaGetLastError());
    dev->launches.fetch_add(1);
    MC_API_END
}

mc_status mc_w4_sizes(uint32_t N, uint32_t K, size_t* w4_nbytes, size_t* scales_nbytes)
{
    MC_API_BEGIN
    MC_REQUIRE(N % 2 == 0 && K % 256 == 0, "w4_sizes: N must be even and K a multiple of 256");
    if (w4_nbytes) *w4_nbytes = w4_bytes(N, K);
    if (scales_nbytes) *scales_nbytes = w4_scale_bytes(N, K);
    MC_API_END
}

mc_status mc_linear_w4(mc_device* dev, mc_buffer* y, mc_buffer* x, mc_buffer* w4, mc_buffer* scales_packed, uint32_t M, uint32_t N, uint32_t K)
{
    MC_API_BEGIN
    MC_REQUIRE(dev && y && x && w4 && scales_packed, "bad arguments");
    MC_CUDA_CHECK(cudaSetDevice(dev->ordinal));
    MC_REQUIRE(M >= 1 && M <= uint32_t(kQMaxMB), "linear_w4: M must be in [1,8] for the streaming GEMV path");
    MC_REQUIRE(N % 2 == 0 && K % 256 == 0, "linear_w4: N must be even and K a multiple of 256");
    MC_REQUIRE(w4->size >= w4_bytes(N, K) && scales_packed->size >= w4_scale_bytes(N, K) && x->size >= size_t(M) * K * 2 && y->size >= size_t(M) * N * 2,
               "linear_w4: buffer too small");
    mc_llama shim;
    shim.dev = dev;
    launcher L{&shim, dev->stream, false};
    qgemv_params q{};
    q.g.W = w4->dptr, q.g.N = N, q.g.K = K, q.g.rows = M;
    q.g.x = static_cast<const uint16_t*>(x->dptr), q.g.ldx = K;
    q.g.y = static_cast<uint16_t*>(y->dptr), q.g.ldy = N;
    q.scales = scales_packed->dptr;
    qgemv_launch<WF_W4, PRO_NONE, EPI_NONE>(L, q);
    MC_API_END
}

} // extern "C"
