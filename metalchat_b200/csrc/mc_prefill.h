// metalchat_b200/csrc/mc_prefill.h — host interface of the tensor-core prefill kernels (mc_prefill.cu / mc_gemm_tc.cuh).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace mc {
namespace tc {

enum { GEMM_STORE = 0, GEMM_RESIDUAL = 2, GEMM_SWIGLU = 3, GEMM_PARTIAL_F32 = 4 }; // GEMM_PARTIAL_F32: Y is a float matrix of unrounded sums, ldy in floats

// Kernels launched by the calling thread from now on carry the programmatic-stream-serialization attribute: every kernel of this
// file starts with griddepcontrol.launch_dependents and waits (griddepcontrol.wait) before it touches the previous kernel's results,
// so its prologue -- barrier setup, TMEM allocation, the first ring-full of weight tiles -- overlaps the previous kernel's tail.
void set_pdl(bool on);

// shapes the tcgen05 GEMM accepts (TMA needs 16-byte row pitches; tiles are 64 wide in k and 32 wide in the epilogue)
bool gemm_supported(uint32_t N, uint32_t K, uint32_t ldx, uint32_t ldy);

// Y[M, ldy] = epilogue(X[M, ldx(:K)] . W[N, K]^T) on `stream`; `err` is the device error flag of bounded waits.
// mode GEMM_SWIGLU writes N/2 columns.  Returns the number of kernels launched (1).
int gemm(cudaStream_t stream, int sm_count, int mode, const uint16_t* X, uint32_t ldx, const uint16_t* W, uint16_t* Y, const uint16_t* res, uint32_t M, uint32_t N,
         uint32_t K, uint32_t ldy, int* err);

// Tensor parallel, tensor-core path: all-reduce of the fp32 partial sums [rows, D] of a row-parallel linear (GEMM_PARTIAL_F32) over the
// ranks of one box, fused with the rounding and the residual add:  out = r(res + r(sum over ranks, in rank order)).
// Reduce-scatter by peer reads + all-gather by peer writes over NVLink: rank r sums slice r of every rank's partial buffer and
// stores the finished bf16 slice into every rank's result buffer.  Epoch flags (release / acquire at system scope) order the phases;
// every wait is bounded and raises *err.  The buffers of an exchange alternate between two halves (parity): see mc_engine.cu.
constexpr int kTpRowsMaxWorld = 8;
struct tp_rows_exchange {
    uint32_t world, rank;
    const float* partial[kTpRowsMaxWorld]; // partial buffer (this parity) of every rank as mapped on this GPU: [rows][D] fp32
    uint16_t* result[kTpRowsMaxWorld];     // result buffer (this parity) of every rank: [rows][D] bf16
    uint32_t* ready[kTpRowsMaxWorld];      // flag array [world] of every rank: ready[k][src] = epochs for which src's partial sums are complete
    uint32_t* done[kTpRowsMaxWorld];       // flag array [world] of every rank: done[k][src] = epochs for which src's slice has arrived in k's result
    unsigned* counter;                     // local: CTAs of this launch that have stored their part
    unsigned* epoch;                       // local: exchanges completed so far
    int* err;
};
int tp_allreduce_rows(cudaStream_t stream, int sm_count, const tp_rows_exchange& x, const uint16_t* res, uint32_t rows, uint32_t D);

int embed_rows(cudaStream_t stream, uint16_t* out, const uint16_t* table, const int32_t* ids, uint32_t rows, uint32_t D);
int rmsnorm_rows(cudaStream_t stream, uint16_t* out, const uint16_t* x, const uint16_t* w, uint32_t rows, uint32_t D, float eps);
// qkv [rows, (H + 2 KV) hd] -> q [rows, H hd] rotated, K/V cache rows [seq][kv][start_pos + row] of one layer;
// with row_seq / row_pos (device arrays) every row carries its own sequence and position (batched decode)
int rope_append(cudaStream_t stream, const uint16_t* qkv, uint16_t* q, uint16_t* kcache_layer, uint16_t* vcache_layer, const float* fcos, const float* fsin,
                uint32_t rows, uint32_t seq, uint32_t start_pos, uint32_t H, uint32_t KV, uint32_t hd, uint32_t max_seq, const int32_t* row_seq = nullptr,
                const int32_t* row_pos = nullptr);
// causal attention of `rows` consecutive positions of sequence `seq` against its cache (positions key_begin .. start_pos + row);
// key_begin = start_pos reproduces the reference's chunk mask, which hides the cached prefix (quirk Q9, nn/attention.h:283-299)
int prefill_attn(cudaStream_t stream, const uint16_t* q, const uint16_t* kcache_layer, const uint16_t* vcache_layer, uint16_t* out, uint32_t rows, uint32_t seq,
                 uint32_t start_pos, uint32_t H, uint32_t KV, uint32_t hd, uint32_t max_seq, float scale, uint32_t key_begin = 0);

// one decode step of `rows` sequences: row r attends keys 0 .. row_pos[r] of sequence row_seq[r]; the H / KV query heads of a KV head
// share one pass over its cache (grouped-query attention)
bool decode_attn_gqa_supported(uint32_t H, uint32_t KV, uint32_t hd);
// With qkv (un-rotated q|k|v rows of the step, [rows, (H + 2 KV) hd]) the kernel rotates q and k and appends k', v itself (q may be null).
int decode_attn_gqa(cudaStream_t stream, const uint16_t* q, uint16_t* kcache_layer, uint16_t* vcache_layer, uint16_t* out, uint32_t rows,
                    const int32_t* row_seq, const int32_t* row_pos, uint32_t H, uint32_t KV, uint32_t hd, uint32_t max_seq, float scale,
                    const uint16_t* qkv = nullptr, const float* fcos = nullptr, const float* fsin = nullptr);

// QLoRA on the tensor-core path: bf16 image of a group-quantised matrix, w = r(r(q) * r(s)) (kernel/mul.metal:76-77; group == K gives the
// per-row scale of quantization::linear), and y2 = r(y + r(r(B . ax) * scale)) + the fused tail
int dequant_group(cudaStream_t stream, uint16_t* out, const int8_t* q, const float* scales, uint32_t N, uint32_t K, uint32_t group);
// y [rows, ldy]: columns [0, N) = r(x . Wd^T), columns [N, ...) = r(x . A^T) (the stacked adaptor rows are appended to the bf16 image)
int lora_epilogue(cudaStream_t stream, int mode, const uint16_t* y, uint32_t ldy, uint16_t* out, const uint16_t* res, const uint16_t* B, uint32_t rows, uint32_t N,
                  uint32_t ldo, uint32_t rank, uint32_t slices, uint32_t cols0, uint32_t cols1, float scale);

} // namespace tc
} // namespace mc
