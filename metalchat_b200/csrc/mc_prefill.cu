// metalchat_b200/csrc/mc_prefill.cu — launchers of the tensor-core prefill kernels (mc_gemm_tc.cuh).
//
// Replaces, for prompts, the chain nn::linear -> kernel::bmm (nn/linear.h:70-81, kernel/bmm.h:27-89) and the attention of
// nn/attention.h:161-206 when many positions are processed at once: tcgen05/TMEM GEMMs fed by TMA, a two-pass causal
// attention on mma.sync, row-wise RMSNorm / RoPE / KV append around them.
#include "mc_gemm_tc.cuh"
#include "mc_prefill.h"

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>

namespace mc {
namespace tc {

namespace {

thread_local bool g_pdl = false; // programmatic dependent launch for the kernels launched by this thread (tc::set_pdl)

// every kernel of this file goes through here: optional cluster dimension, optional programmatic stream serialization
template <typename... KArgs, typename... Args>
void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, unsigned cluster, Args&&... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = s;
    cudaLaunchAttribute attr[2];
    unsigned n = 0;
    if (cluster > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster, attr[n].val.clusterDim.y = 1, attr[n].val.clusterDim.z = 1;
        n++;
    }
    if (g_pdl) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        n++;
    }
    cfg.attrs = attr, cfg.numAttrs = n;
    MC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
}

using encode_fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                               CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime (libcuda is not linked)
encode_fn tensor_map_encoder()
{
    static encode_fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
            cudaGetLastError();
            throw error(MC_ERR_RUNTIME, "cuda: cuTensorMapEncodeTiled is not available from this driver (the prefill GEMM needs TMA)");
        }
        return reinterpret_cast<encode_fn>(p);
    }();
    return fn;
}

// K-major bf16 matrix [rows, K] with row pitch ld elements; box = box_rows x 64 elements, 128-byte swizzle, zero fill
CUtensorMap make_map(const uint16_t* base, uint32_t rows, uint32_t K, uint32_t ld, uint32_t box_rows)
{
    using key_t = std::tuple<const void*, uint32_t, uint32_t, uint32_t, uint32_t>;
    static std::mutex mu;
    static std::map<key_t, CUtensorMap> cache;
    const key_t key{base, rows, K, ld, box_rows};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    CUtensorMap m;
    const cuuint64_t dims[2] = {K, rows};
    const cuuint64_t strides[1] = {cuuint64_t(ld) * 2};
    const cuuint32_t box[2] = {uint32_t(kTcBK), box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = tensor_map_encoder()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<uint16_t*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw error(MC_ERR_RUNTIME, "cuda: cuTensorMapEncodeTiled failed with code " + std::to_string(int(r)));
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, m);
    return m;
}

template <int BN, int EPI, int AROWS> void launch_gemm(cudaStream_t s, int sm_count, const CUtensorMap& mx, const CUtensorMap& mw, const gemm_tc_params& p)
{
    auto kernel = gemm_tc_kernel<BN, EPI, AROWS>;
    static bool configured[8] = {false};
    int dev = 0;
    MC_CUDA_CHECK(cudaGetDevice(&dev));
    if (!configured[dev & 7]) {
        MC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(BN, AROWS)));
        configured[dev & 7] = true;
    }
    const uint32_t tiles = ((p.M + kTcBM - 1) / kTcBM) * ((p.N + BN - 1) / BN);
    const uint32_t grid = tiles < uint32_t(sm_count) ? tiles : uint32_t(sm_count);
    launch_k(kernel, dim3(grid), dim3(kTcThreads), tc_smem_bytes(BN, AROWS), s, 1, mx, mw, p);
}
template <int BN, int AROWS> void launch_gemm_epi(cudaStream_t s, int sm_count, int mode, const CUtensorMap& mx, const CUtensorMap& mw, const gemm_tc_params& p)
{
    switch (mode) {
    case GEMM_STORE: launch_gemm<BN, EPI_NONE, AROWS>(s, sm_count, mx, mw, p); break;
    case GEMM_RESIDUAL: launch_gemm<BN, EPI_RESIDUAL, AROWS>(s, sm_count, mx, mw, p); break;
    case GEMM_SWIGLU: launch_gemm<BN, EPI_SWIGLU, AROWS>(s, sm_count, mx, mw, p); break;
    case GEMM_PARTIAL_F32: launch_gemm<BN, EPI_PARTIAL, AROWS>(s, sm_count, mx, mw, p); break;
    default: throw error(MC_ERR_INVALID, "prefill gemm: unknown epilogue");
    }
}
template <int AROWS>
void launch_gemm_bn(cudaStream_t stream, int sm_count, int mode, const uint16_t* X, uint32_t ldx, const uint16_t* W, const gemm_tc_params& p)
{
    const uint32_t M = p.M, N = p.N, K = p.K;
    const uint32_t m_blocks = (M + kTcBM - 1) / kTcBM;
    // the widest tile that still gives every SM a tile: 256 for prompts (highest flop per byte staged), down to 32 weight rows
    // per CTA for a decode batch, where the weights are streamed once and the point is to keep all SMs loading
    const CUtensorMap mx = make_map(X, M, K, ldx, AROWS);
    static const int forced = [] {
        const char* e = getenv("MC_TC_BN"); // experiments: force the tile width
        return e ? atoi(e) : 0;
    }();
    // (an M = 128, K = 16 MMA occupies the tensor core for >= 128 cycles whatever its N -- the A operand is read at 32 bytes per
    // cycle -- so a 128-wide tile runs at half rate: 256 wins as soon as it fills half of the SMs)
    if (forced == 256 || (!forced && m_blocks * ((N + 255) / 256) * 2 >= uint32_t(sm_count))) {
        launch_gemm_epi<256, AROWS>(stream, sm_count, mode, mx, make_map(W, N, K, K, 256), p);
    } else if (forced == 128 || (!forced && (m_blocks * ((N + 127) / 128) >= uint32_t(sm_count) || M > 256))) {
        launch_gemm_epi<128, AROWS>(stream, sm_count, mode, mx, make_map(W, N, K, K, 128), p);
    } else if (forced == 64 || (!forced && m_blocks * ((N + 63) / 64) >= uint32_t(sm_count))) {
        launch_gemm_epi<64, AROWS>(stream, sm_count, mode, mx, make_map(W, N, K, K, 64), p);
    } else {
        launch_gemm_epi<32, AROWS>(stream, sm_count, mode, mx, make_map(W, N, K, K, 32), p);
    }
}

// decode-batch GEMM split over k inside a cluster (gemm_tc_splitk_kernel): the smallest power-of-two split that gives every SM a CTA
template <int EPI> void launch_splitk(cudaStream_t s, uint32_t ks, const CUtensorMap& mx, const CUtensorMap& mw, const gemm_tc_params& p)
{
    auto kernel = gemm_tc_splitk_kernel<EPI>;
    static bool configured[8] = {false};
    int dev = 0;
    MC_CUDA_CHECK(cudaGetDevice(&dev));
    if (!configured[dev & 7]) {
        MC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSkSmem));
        configured[dev & 7] = true;
    }
    launch_k(kernel, dim3(((p.N + kSkBN - 1) / kSkBN) * ks), dim3(kTcThreads), kSkSmem, s, ks, mx, mw, p);
}
uint32_t splitk_factor(uint32_t N, uint32_t K, int sm_count)
{
    static const int off = getenv("MC_TC_NO_SPLITK") != nullptr;
    const uint32_t tiles = (N + kSkBN - 1) / kSkBN, k_blocks = K / kTcBK;
    if (off || tiles * 2 >= 3 * uint32_t(sm_count)) return 0; // enough 256-row tiles for the persistent kernel
    static const uint32_t ks_max = [] {
        const char* e = getenv("MC_TC_KS_MAX"); // experiments: cap the cluster size of the k split
        return e ? uint32_t(atoi(e)) : 8u;
    }();
    uint32_t ks = 1;
    while (ks < ks_max && tiles * ks < uint32_t(sm_count) && k_blocks % (ks * 2) == 0 && k_blocks / (ks * 2) >= 2) ks *= 2;
    return ks;
}

} // namespace

void set_pdl(bool on) { g_pdl = on; }

bool gemm_supported(uint32_t N, uint32_t K, uint32_t ldx, uint32_t ldy)
{
    return K >= uint32_t(kTcBK) && K % kTcBK == 0 && N % 32 == 0 && ldx % 8 == 0 && ldy % 8 == 0;
}

int gemm(cudaStream_t stream, int sm_count, int mode, const uint16_t* X, uint32_t ldx, const uint16_t* W, uint16_t* Y, const uint16_t* res, uint32_t M, uint32_t N,
         uint32_t K, uint32_t ldy, int* err)
{
    MC_REQUIRE(M >= 1 && gemm_supported(N, K, ldx, ldy), "prefill gemm: unsupported shape (K % 64, N % 32, pitches % 8)");
    MC_REQUIRE((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(res)) % 16 == 0,
               "prefill gemm: operands must be 16-byte aligned");
    gemm_tc_params p{};
    p.Y = Y, p.res = res, p.M = M, p.N = N, p.K = K, p.ldy = ldy, p.err = err;
    // a decode batch of <= 32 rows stages only 32 rows of X per k block (see tc_stages); small matrices are split over k
    const uint32_t ks = M <= 32 && mode != GEMM_PARTIAL_F32 ? splitk_factor(N, K, sm_count) : 0; // (the partial-sum mode lives in the persistent kernel only)
    if (ks) {
        const CUtensorMap mx = make_map(X, M, K, ldx, kSkRows), mw = make_map(W, N, K, K, kSkBN);
        switch (mode) {
        case GEMM_STORE: launch_splitk<EPI_NONE>(stream, ks, mx, mw, p); break;
        case GEMM_RESIDUAL: launch_splitk<EPI_RESIDUAL>(stream, ks, mx, mw, p); break;
        case GEMM_SWIGLU: launch_splitk<EPI_SWIGLU>(stream, ks, mx, mw, p); break;
        default: throw error(MC_ERR_INVALID, "prefill gemm: unknown epilogue");
        }
    } else if (M <= 32) launch_gemm_bn<32>(stream, sm_count, mode, X, ldx, W, p);
    else launch_gemm_bn<128>(stream, sm_count, mode, X, ldx, W, p);
    return 1;
}

int tp_allreduce_rows(cudaStream_t stream, int sm_count, const tp_rows_exchange& x, const uint16_t* res, uint32_t rows, uint32_t D)
{
    MC_REQUIRE(x.world >= 2 && x.world <= uint32_t(kTpRowsMaxWorld) && x.rank < x.world, "tp all-reduce: bad world / rank");
    MC_REQUIRE(D % 8 == 0 && rows >= 1, "tp all-reduce: dim must be a multiple of 8");
    const uint64_t n8 = uint64_t(rows) * D / 8, mine = (n8 + x.world - 1) / x.world;
    const uint32_t grid = uint32_t(std::min<uint64_t>(uint64_t(sm_count), std::max<uint64_t>(1, (mine + 255) / 256))); // co-resident: the CTAs wait for each other's peers
    launch_k(tp_allreduce_rows_kernel, dim3(grid), dim3(256), 0, stream, 1, x, res, rows, D);
    return 1;
}

int embed_rows(cudaStream_t stream, uint16_t* out, const uint16_t* table, const int32_t* ids, uint32_t rows, uint32_t D)
{
    MC_REQUIRE(D % 8 == 0, "prefill embedding: dim must be a multiple of 8");
    launch_k(embed_rows_kernel, dim3(rows), dim3(256), 0, stream, 1, out, table, ids, D);
    return 1;
}
int rmsnorm_rows(cudaStream_t stream, uint16_t* out, const uint16_t* x, const uint16_t* w, uint32_t rows, uint32_t D, float eps)
{
    MC_REQUIRE(D % 8 == 0, "prefill rmsnorm: dim must be a multiple of 8");
    launch_k(rmsnorm_rows_kernel, dim3(rows), dim3(256), 0, stream, 1, out, x, w, D, eps);
    return 1;
}
int rope_append(cudaStream_t stream, const uint16_t* qkv, uint16_t* q, uint16_t* kcache_layer, uint16_t* vcache_layer, const float* fcos, const float* fsin,
                uint32_t rows, uint32_t seq, uint32_t start_pos, uint32_t H, uint32_t KV, uint32_t hd, uint32_t max_seq, const int32_t* row_seq, const int32_t* row_pos)
{
    launch_k(rope_append_kernel, dim3(rows), dim3(256), 0, stream, 1, qkv, (H + 2 * KV) * hd, q, kcache_layer, vcache_layer, fcos, fsin, row_seq, row_pos, seq,
             start_pos, H, KV, hd, max_seq);
    return 1;
}
int prefill_attn(cudaStream_t stream, const uint16_t* q, const uint16_t* kcache_layer, const uint16_t* vcache_layer, uint16_t* out, uint32_t rows, uint32_t seq,
                 uint32_t start_pos, uint32_t H, uint32_t KV, uint32_t hd, uint32_t max_seq, float scale, uint32_t key_begin)
{
    MC_REQUIRE(hd == 64 || hd == 128, "prefill attention: head_dim must be 64 or 128");
    MC_REQUIRE(key_begin <= start_pos, "prefill attention: the first visible key lies beyond the first row");
    pattn_params p{};
    p.q = q, p.out = out, p.rows = rows, p.start_pos = start_pos, p.H = H, p.KV = KV, p.max_seq = max_seq, p.scale = scale, p.key_begin = key_begin;
    p.kc = kcache_layer + size_t(seq) * KV * max_seq * hd;
    p.vc = vcache_layer + size_t(seq) * KV * max_seq * hd;
    const uint32_t n_rep = H / KV;
    p.rep = (n_rep == 2 || n_rep == 4 || n_rep == 8) ? n_rep : 1;
    const uint32_t qt = 128 / p.rep;
    const dim3 grid(H / p.rep, (rows + qt - 1) / qt);
    const size_t smem = size_t(4) * 64 * (hd + 8) * 2;
    uint32_t sbits;
    memcpy(&sbits, &scale, 4);
    const bool pow2 = (sbits & 0x007fffffu) == 0 && scale > 1e-30f; // exact power of two (and far from the subnormal range)
    auto launch = [&](auto kernel) {
        if (smem > 48 * 1024) {
            // once per device would do; the call is cheap and idempotent
            MC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        }
        launch_k(kernel, grid, dim3(256), smem, stream, 1, p);
    };
    if (hd == 64) {
        if (pow2) launch(prefill_attn_kernel<64, true>);
        else launch(prefill_attn_kernel<64, false>);
    } else {
        if (pow2) launch(prefill_attn_kernel<128, true>);
        else launch(prefill_attn_kernel<128, false>);
    }
    MC_CUDA_CHECK(cudaGetLastError());
    return 1;
}

int decode_attn_gqa(cudaStream_t stream, const uint16_t* q, uint16_t* kcache_layer, uint16_t* vcache_layer, uint16_t* out, uint32_t rows,
                    const int32_t* row_seq, const int32_t* row_pos, uint32_t H, uint32_t KV, uint32_t hd, uint32_t max_seq, float scale, const uint16_t* qkv,
                    const float* fcos, const float* fsin)
{
    MC_REQUIRE(decode_attn_gqa_supported(H, KV, hd), "batched decode attention: head_dim must be 64 or 128 and at most 8 query heads per KV head");
    MC_REQUIRE(q || (qkv && fcos && fsin), "batched decode attention: rotated q rows, or un-rotated q|k|v rows with the rope tables");
    dattn_params p{};
    p.qkv = qkv, p.fcos = fcos, p.fsin = fsin;
    p.q = q, p.kc = kcache_layer, p.vc = vcache_layer, p.out = out, p.row_seq = row_seq, p.row_pos = row_pos, p.H = H, p.KV = KV, p.max_seq = max_seq, p.scale = scale;
    const dim3 grid(KV, rows);
    const size_t smem = size_t(4) * 64 * (hd + 8) * 2 + 32 * sizeof(float);
    uint32_t sbits;
    memcpy(&sbits, &scale, 4);
    const bool pow2 = (sbits & 0x007fffffu) == 0 && scale > 1e-30f;
    auto launch = [&](auto kernel) {
        if (smem > 48 * 1024) MC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        launch_k(kernel, grid, dim3(128), smem, stream, 1, p);
    };
    if (hd == 64) {
        if (pow2) launch(decode_attn_gqa_kernel<64, true>);
        else launch(decode_attn_gqa_kernel<64, false>);
    } else {
        if (pow2) launch(decode_attn_gqa_kernel<128, true>);
        else launch(decode_attn_gqa_kernel<128, false>);
    }
    MC_CUDA_CHECK(cudaGetLastError());
    return 1;
}
bool decode_attn_gqa_supported(uint32_t H, uint32_t KV, uint32_t hd) { return (hd == 64 || hd == 128) && KV > 0 && H % KV == 0 && H / KV <= 8; }

int dequant_group(cudaStream_t stream, uint16_t* out, const int8_t* q, const float* scales, uint32_t N, uint32_t K, uint32_t group)
{
    MC_REQUIRE(K % 8 == 0 && group % 8 == 0 && K % group == 0, "dequant: K and the group size must be multiples of 8");
    const uint64_t n8 = uint64_t(N) * K / 8;
    const unsigned blocks = unsigned(std::min<uint64_t>((n8 + 255) / 256, 148 * 16));
    dequant_group_kernel<<<blocks, 256, 0, stream>>>(out, q, scales, n8, K, group);
    MC_CUDA_CHECK(cudaGetLastError());
    return 1;
}
int lora_epilogue(cudaStream_t stream, int mode, const uint16_t* y, uint32_t ldy, uint16_t* out, const uint16_t* res, const uint16_t* B, uint32_t rows, uint32_t N,
                  uint32_t ldo, uint32_t rank, uint32_t slices, uint32_t cols0, uint32_t cols1, float scale)
{
    MC_REQUIRE(N % 2 == 0 && rank % 2 == 0 && rank <= uint32_t(kLoraMaxRank) && ldo % 2 == 0 && ldy % 2 == 0, "lora epilogue: even N / pitches and an even rank <= 16");
    lora_epi_params p{};
    p.y = y, p.out = out, p.res = res, p.B = B, p.rows = rows, p.N = N, p.ldy = ldy, p.ldo = ldo, p.rank = rank, p.slices = slices, p.cols0 = cols0, p.cols1 = cols1,
    p.scale = scale;
    p.rows_per_cta = rows <= 64 ? 4u : uint32_t(kLoraRows);
    const dim3 grid((N / 2 + 255) / 256, (rows + p.rows_per_cta - 1) / p.rows_per_cta);
    switch (mode) {
    case GEMM_STORE: launch_k(lora_epilogue_kernel<EPI_NONE>, grid, dim3(256), 0, stream, 1, p); break;
    case GEMM_RESIDUAL: launch_k(lora_epilogue_kernel<EPI_RESIDUAL>, grid, dim3(256), 0, stream, 1, p); break;
    case GEMM_SWIGLU: launch_k(lora_epilogue_kernel<EPI_SWIGLU>, grid, dim3(256), 0, stream, 1, p); break;
    default: throw error(MC_ERR_INVALID, "lora epilogue: unknown mode");
    }
    return 1;
}

} // namespace tc
} // namespace mc
