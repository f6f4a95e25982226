/* include/mc_cuda.h — C ABI of libmc_cuda.so, the B200 (sm_100a) accelerator backend.
 *
 * This is the drop-in boundary for the transformer decode hot path of
 * ybubnov/metalchat.  The reference has no C ABI: its seam is five Metal-bound
 * translation units behind opaque pimpl handles (include/metalchat/metal.h:14-34,
 * src/metal_impl.h:19-112).  Every entry point below names the reference interface
 * it replaces; a maintainer re-targets src/{metal,accelerator,allocator,kernel,
 * kernel_thread}.cc onto these calls (see INTEGRATION.md for the binding).
 *
 * Conventions: plain C, plain pointers and sizes, no CUDA/torch types.  Every call
 * returns an mc_status (0 = ok); the message of the last failure on the calling
 * thread is available from mc_last_error().  Handles are reference counted where the
 * reference uses std::shared_ptr.  A device owns ONE in-order CUDA stream (the
 * reference: one MTL command queue, src/kernel_thread.cc:13-56); calls on one device
 * must come from one producer thread at a time (kernel_thread is not synchronised
 * either, kernel_thread.h:57-294).
 */
#ifndef MC_CUDA_H
#define MC_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MC_API __attribute__((visibility("default")))

typedef enum mc_status {
    MC_OK = 0,
    MC_ERR_INVALID = 1,   /* std::invalid_argument in the reference (kernel.h:126-140)      */
    MC_ERR_RUNTIME = 2,   /* std::runtime_error (src/accelerator.cc:117-158, kernel_thread) */
    MC_ERR_ALLOC = 3,     /* alloc_error : std::bad_alloc (allocator.h:20-34)               */
    MC_ERR_NOT_FOUND = 4, /* "kernel not found" from hardware_accelerator::load            */
    MC_ERR_FULL = 5       /* command buffer at capacity (kernel_thread.h:222-243)           */
} mc_status;

typedef struct mc_device mc_device;
typedef struct mc_buffer mc_buffer;
typedef struct mc_kernel mc_kernel;
typedef struct mc_cmdbuf mc_cmdbuf;
typedef struct mc_llama mc_llama;

MC_API const char* mc_last_error(void);
MC_API const char* mc_version(void);

/* ---- device: metal::device / hardware_accelerator ctor ---------------------------
 * replaces MTL::CreateSystemDefaultDevice + newCommandQueue + newLibrary
 * (src/accelerator.cc:73-113, src/metal.cc:51-55).                                  */
MC_API mc_status mc_device_count(int* count);
MC_API mc_status mc_device_create(int ordinal, mc_device** out);
MC_API mc_status mc_device_destroy(mc_device* dev);
MC_API mc_status mc_device_name(mc_device* dev, char* out, size_t cap);           /* accelerator.h:108-113 name()             */
MC_API mc_status mc_device_max_buffer(mc_device* dev, size_t* bytes);             /* MTL::Device::maxBufferLength             */
MC_API mc_status mc_device_sm_count(mc_device* dev, int* count);
MC_API mc_status mc_device_synchronize(mc_device* dev);
/* raw cudaStream_t of the device (for event timing by a harness that shares the stream) */
MC_API mc_status mc_device_stream(mc_device* dev, void** cuda_stream);

/* ---- memory: hardware_memory_allocator / heap / nocopy (src/allocator.cc:46-275) -- */
enum {
    MC_MEM_DEVICE = 0, /* cudaMalloc: device resident, not host dereferenceable          */
    MC_MEM_SHARED = 1, /* cudaMallocManaged: host dereferenceable like MTL Shared storage
                          (src/metal.cc:21-25 contents()); valid on the host after mc_wait */
    MC_MEM_PINNED = 2  /* cudaHostAlloc(mapped): host memory the GPU reads over PCIe     */
};
MC_API mc_status mc_alloc(mc_device* dev, size_t size, int flags, mc_buffer** out);                 /* newBuffer(size, Shared)        src/allocator.cc:46-60   */
MC_API mc_status mc_alloc_copy(mc_device* dev, const void* src, size_t size, int flags, mc_buffer** out); /* newBuffer(ptr, size, ...)  src/allocator.cc:62-80   */
MC_API mc_status mc_wrap_host(mc_device* dev, void* host, size_t size, mc_buffer** out);            /* nocopy_allocator newBuffer(noCopy) src/allocator.cc:127-173: cudaHostRegister */
MC_API mc_status mc_buffer_retain(mc_buffer* buf);
MC_API mc_status mc_buffer_release(mc_buffer* buf);
MC_API mc_status mc_buffer_host_ptr(mc_buffer* buf, void** out);  /* metal::data(buffer)  metal.h:17-22 */
MC_API mc_status mc_buffer_dev_ptr(mc_buffer* buf, void** out);
MC_API mc_status mc_buffer_size(mc_buffer* buf, size_t* out);     /* metal::size(buffer)  metal.h:17-22 */
/* Explicit copies on the device stream (no unified memory on B200). */
MC_API mc_status mc_memcpy_h2d(mc_device* dev, mc_buffer* dst, size_t dst_off, const void* src, size_t size);
MC_API mc_status mc_memcpy_d2h(mc_device* dev, void* dst, mc_buffer* src, size_t src_off, size_t size);
MC_API mc_status mc_memset(mc_device* dev, mc_buffer* dst, size_t dst_off, int value, size_t size);
/* hardware_heap_allocator: one fixed arena, bump sub-allocation (src/allocator.cc:82-125) */
typedef struct mc_heap mc_heap;
MC_API mc_status mc_heap_create(mc_device* dev, size_t capacity, mc_heap** out);
MC_API mc_status mc_heap_alloc(mc_heap* heap, size_t size, mc_buffer** out);
MC_API mc_status mc_heap_reset(mc_heap* heap);
MC_API mc_status mc_heap_destroy(mc_heap* heap);

/* ---- kernels: hardware_accelerator::load (src/accelerator.cc:117-158) ------------
 * `name` is the reference's mangled host name, e.g. "softmax_bfloat", "bmm_8_float",
 * "cumsum_2_bfloat", "hadamard_broadcast_bfloat_int8_t_float" (accelerator.h:175-218).
 * All 71 names of metalchat.metallib resolve.                  */
MC_API mc_status mc_kernel_lookup(mc_device* dev, const char* name, mc_kernel** out);
MC_API mc_status mc_kernel_release(mc_kernel* k);                  /* the shared_kernel handle goes out of scope (metal.h:27-28) */
MC_API mc_status mc_kernel_name(mc_kernel* k, const char** out);
MC_API mc_status mc_kernel_max_threads(mc_kernel* k, size_t* out); /* 1024, src/kernel.cc:75-79 */
MC_API mc_status mc_kernel_count(int* count);
MC_API mc_status mc_kernel_name_at(int index, const char** out);

/* ---- encode / dispatch: hardware_function_encoder + kernel_thread -----------------
 * (src/kernel_thread.cc:69-120,134-144,185-199; kernel_thread.h:104-137).
 * Arguments are bound by slot index in the reference's bind order: a tensor occupies
 * two consecutive slots — tensor_layout<N> bytes {sizes[N],strides[N],offsets[N]} as
 * uint32 (tensor/concept.h:24-33), then the buffer + byte offset; a scalar occupies one
 * slot of raw bytes.  mc_dispatch consumes the bound slots and enqueues the kernel on
 * the device stream; `grid` is total THREADS (Metal dispatchThreads) and is only
 * validated (group <= 1024 threads, grid >= group per dimension; kernel.h:126-140) —
 * the CUDA launch shape is chosen by the backend.                                     */
MC_API mc_status mc_stream_begin(mc_device* dev, size_t capacity, mc_cmdbuf** out);
MC_API mc_status mc_set_bytes(mc_cmdbuf* cb, uint32_t index, const void* bytes, size_t size);  /* setBytes   */
MC_API mc_status mc_set_buffer(mc_cmdbuf* cb, uint32_t index, mc_buffer* buf, size_t offset);   /* setBuffer  */
MC_API mc_status mc_barrier(mc_cmdbuf* cb, mc_buffer* buf);                                      /* memoryBarrier: no-op on an in-order stream */
MC_API mc_status mc_dispatch(mc_cmdbuf* cb, mc_kernel* k, const uint32_t grid[3], const uint32_t group[3]);
MC_API mc_status mc_on_completed(mc_cmdbuf* cb, void (*fn)(void* user, int status), void* user); /* addCompletedHandler */
MC_API mc_status mc_commit(mc_cmdbuf* cb);                                                       /* commit()            */
MC_API mc_status mc_wait(mc_cmdbuf* cb, char* err, size_t cap);                                  /* waitUntilCompleted + error() */
MC_API mc_status mc_cmdbuf_size(mc_cmdbuf* cb, size_t* n);
MC_API mc_status mc_cmdbuf_release(mc_cmdbuf* cb);
MC_API mc_status mc_launch_count(mc_device* dev, uint64_t* launches); /* kernels launched by this backend so far */

/* ---- fused decode engine: transformer<llama3>::transform --------------------------
 * (transformer.h:357-364 -> nn/llama.h:113-134 -> nn/sampling.h:69-76).               */
typedef struct mc_llama_config {
    uint32_t dim, n_layers, n_heads, n_kv_heads, head_dim, ffn_dim, vocab, max_seq_len;
    float rope_theta, norm_eps;
    uint32_t quant;      /* 0: bf16 weights; 1: QLoRA layout (huggingface/llama.h:152-171)    */
    uint32_t lora_rank;  /* 16                                                                  */
    float lora_scale;    /* 2.0                                                                 */
    uint32_t group_size; /* 32                                                                  */
    uint32_t n_seqs;     /* independent bs=1 sequences decoded together (nn/llama.h:86 is 1)    */
    uint32_t tp_rank, tp_world; /* tensor-parallel shard (0,1 = single GPU)                     */
    uint32_t flags;      /* MC_LLAMA_* below                                                    */
} mc_llama_config;
enum {
    MC_LLAMA_W4_PACKED = 1u << 0, /* store QLoRA int8-in-int4-range weights two per byte       */
    MC_LLAMA_NO_GRAPH = 1u << 1,  /* launch kernels directly instead of replaying a CUDA graph  */
    MC_LLAMA_NO_PDL = 1u << 2,    /* no programmatic dependent launch between decode kernels    */
    /* 1u << 3 was an experimental grid-barrier megakernel, superseded by the streaming persistent kernel and removed */
    MC_LLAMA_NO_STREAM = 1u << 4,  /* do not use the streaming persistent kernel (TMA weight ring): per-op kernels under a CUDA graph */
    MC_LLAMA_NO_TC_PREFILL = 1u << 5, /* prompts go through the 4-row GEMV kernels instead of the tcgen05 GEMM path */
    MC_LLAMA_NO_SHADOW = 1u << 6,     /* quantised models: do not keep the resident bf16 image (2 bytes per weight) that the tensor-core
                                         prompt / batch path multiplies; prompts and batches then take the packed GEMV kernels */
    MC_LLAMA_REF_CHUNK_MASK = 1u << 7 /* prompts of len > 1 at start_pos > 0 do NOT see the cached prefix, exactly like make_causal_mask
                                         (nn/attention.h:283-299 leaves those columns at -inf); default: the prefix is visible */
};
typedef struct mc_sampler_config {
    uint32_t mode;       /* 0 greedy argmax (lowest index on ties); 1 top-k -> nucleus -> multinomial (nn/sampling.h:306-316) */
    uint32_t top_k;      /* 50 (nn/sampling.h:309)                                              */
    float temperature;   /* 0.6 (nn/sampling.h:179-181)                                         */
    float top_p;         /* 0.9                                                                 */
    uint32_t intended;   /* multinomial: 0 reference-exact a=input[row,S-1], 1 a=input[row,N-1] (kernel/multinomial.metal:107) */
} mc_sampler_config;

MC_API mc_status mc_llama_create(mc_device* dev, const mc_llama_config* cfg, mc_llama** out);
MC_API mc_status mc_llama_destroy(mc_llama* m);
/* Parameters by the reference's registered layer paths, e.g.
 * "layers.3.attention.wq.weight", "tok_embeddings.weight", "layers.0.feed_forward.w1.scales",
 * "layers.0.attention.wo.adaptor.A.weight".  Data is the FULL (unsharded) tensor in host memory. */
MC_API mc_status mc_llama_set_tensor(mc_llama* m, const char* name, const void* host, size_t nbytes);
/* The configuration the model was created with (defaults filled in). */
MC_API mc_status mc_llama_get_config(mc_llama* m, mc_llama_config* cfg);

/* ---- safetensors -> device (replaces safetensor_document::open / load, src/safetensor.cc:83-153,237-253 and
 * include/metalchat/safetensor.h:689-747, for the decode path).  `path` is one file or a directory of *.safetensors shards.
 * The file is mapped read-only; entries point into the mapping and stay valid until mc_safetensors_close. */
typedef struct mc_safetensors mc_safetensors;
typedef struct mc_safetensors_entry {
    const char* name;    /* as written in the file                                              */
    const char* dtype;   /* BOOL,I8,U8,I16,U16,F16,BF16,I32,U32,F32,F64,I64,U64 (safetensor.h:251-264) */
    uint32_t rank;
    uint64_t shape[8];
    const void* data;    /* host pointer into the mapping                                       */
    uint64_t nbytes;
} mc_safetensors_entry;
MC_API mc_status mc_safetensors_open(const char* path, mc_safetensors** out);
MC_API mc_status mc_safetensors_close(mc_safetensors* st);
MC_API mc_status mc_safetensors_count(mc_safetensors* st, uint32_t* n);
MC_API mc_status mc_safetensors_entry_at(mc_safetensors* st, uint32_t index, mc_safetensors_entry* out);
MC_API mc_status mc_safetensors_find(mc_safetensors* st, const char* name, mc_safetensors_entry* out);
MC_API mc_status mc_safetensors_metadata(mc_safetensors* st, const char* key, const char** value); /* NULL when absent */
/* Loads every parameter of `m` found in `st` (each one the FULL tensor: tensor-parallel ranks take their slices).
 * MC_LOAD_HF_NAMES: the file uses HuggingFace names (renamed like huggingface/llama.h:88-103; lm_head -> output);
 * MC_LOAD_META_PERMUTE: Meta-format checkpoint, rows of wq / wk are permuted to the rotate-half order (reference.h:73-94);
 * MC_LOAD_STRICT: a parameter of the model missing from the file is an error (output.weight of a bf16 model excepted:
 * it stays tied to the embedding, huggingface/llama.h:103).  fp32 / fp16 tensors of a bf16 model are rounded to bf16. */
#define MC_LOAD_HF_NAMES 1u
#define MC_LOAD_META_PERMUTE 2u
#define MC_LOAD_STRICT 4u
MC_API mc_status mc_llama_load_safetensors(mc_llama* m, mc_safetensors* st, uint32_t flags, uint32_t* n_loaded);
/* Device-side synthetic weights: counter-hash generator keyed by (seed, tensor id, index),
 * distributions in DESIGN.md "Synthetic data"; bit-identical to the test oracle's generator. */
MC_API mc_status mc_llama_init_random(mc_llama* m, uint64_t seed);
/* Must be called after the weights are set and before prefill/decode (packs fused layouts). */
MC_API mc_status mc_llama_finalize(mc_llama* m);
MC_API mc_status mc_llama_weight_bytes(mc_llama* m, uint64_t* streamed_per_step, uint64_t* resident);
/* ids[len] of sequence `seq` at positions [start_pos, start_pos+len): fills the KV cache and
 * leaves the logits of the LAST position (nn/llama.h:128-133).                              */
MC_API mc_status mc_llama_prefill(mc_llama* m, uint32_t seq, const int32_t* ids, uint32_t len, uint32_t start_pos);
/* One decode step for sequences [0, n): ids[n] (host) at pos[n] -> out_ids[n] (host).
 * uniforms[n] (host, may be NULL for greedy) are the injected multinomial draws.
 * Includes the H2D of ids/pos/uniforms and the D2H of out_ids.                             */
MC_API mc_status mc_llama_decode(mc_llama* m, uint32_t n, const int32_t* ids, const int32_t* pos, const float* uniforms,
                                 const mc_sampler_config* sampler, int32_t* out_ids);
/* `steps` greedy/sampled decode steps entirely on the device: the sampled id feeds the next
 * step without a host round trip; out_ids[steps*n] (host, may be NULL) receives every id.
 * elapsed_ms (may be NULL) receives the CUDA-event time of the loop.                        */
MC_API mc_status mc_llama_decode_loop(mc_llama* m, uint32_t n, const int32_t* first_ids, const int32_t* first_pos,
                                      uint32_t steps, const float* uniforms, const mc_sampler_config* sampler,
                                      int32_t* out_ids, float* elapsed_ms);
/* Copy out state for parity checks: logits [vocab] bf16, last hidden [dim] bf16, cache rows. */
MC_API mc_status mc_llama_logits(mc_llama* m, uint32_t seq, void* host_bf16, size_t nbytes);
MC_API mc_status mc_llama_hidden(mc_llama* m, uint32_t seq, void* host_bf16, size_t nbytes);
MC_API mc_status mc_llama_cache(mc_llama* m, uint32_t seq, uint32_t layer, int which, uint32_t n_pos, void* host_bf16, size_t nbytes);
MC_API mc_status mc_llama_launches_per_step(mc_llama* m, uint32_t* kernels);
/* Tensor parallelism (no reference counterpart: the reference is single-device, nn/llama.h:86).  One process per GPU;
 * every rank creates the model with (tp_rank, tp_world), exports a 64-byte IPC handle of its exchange region, the host
 * gathers all handles (torch.distributed / MPI / files) and hands the world x 64 bytes to every rank.  bf16 and QLoRA models
 * shard (column-split wq/wk/wv/w1/w3, row-split wo/w2, vocabulary-split head); every rank makes the same prefill / decode calls
 * and receives the same token ids.  The all-reduce after wo and w2 never leaves the kernels' own code: tagged words stored into
 * the peers' memory by the streaming decode kernel, partial-sum pushes by the per-op GEMV kernels, peer reads / writes by the
 * all-reduce of the tensor-core prompt / batch path -- all over NVLink peer memory, no NCCL (DESIGN.md "Multi-GPU"). */
MC_API mc_status mc_llama_tp_export(mc_llama* m, void* handle, size_t cap);
MC_API mc_status mc_llama_tp_connect(mc_llama* m, const void* handles, size_t nbytes);
/* Comparator for measurements only (bench.py --tp-collective nccl): the two all-reduces of a block become ncclAllReduce calls
 * between the per-op kernels instead of the exchange fused into the kernels.  libnccl.so.2 is resolved with dlopen at this
 * point (the library does not link against it).  Rank 0 makes the 128-byte id, the host hands it to every rank, every rank
 * calls mc_llama_tp_use_nccl (a collective call) after mc_llama_tp_connect and before its first step. */
MC_API mc_status mc_nccl_unique_id(void* id, size_t cap);
MC_API mc_status mc_llama_tp_use_nccl(mc_llama* m, const void* id, size_t nbytes);
/* Diagnostics: one ungraphed decode step with a CUDA event before every launch; us[i] = device time of launch i. */
MC_API mc_status mc_llama_profile_step(mc_llama* m, uint32_t n, float* us, uint32_t cap, uint32_t* count);

/* The default sampler chain (make_default_sampler, nn/sampling.h:306-316: top-k -> nucleus -> multinomial) over
 * logits [rows, vocab] bf16 on the device, with injected uniforms[rows]; host outputs (any may be NULL):
 * topk_idx[rows,k], probs_sorted[rows,k] (bf16 bits after the top-p mask), probs_idx[rows,k], choice[rows], token[rows]. */
MC_API mc_status mc_sample_default(mc_device* dev, mc_buffer* logits_bf16, uint32_t rows, uint32_t vocab, const mc_sampler_config* cfg,
                                   const float* uniforms, int32_t* topk_idx, uint16_t* probs_sorted, int32_t* probs_idx, int32_t* choice,
                                   int32_t* token);

/* ---- stand-alone hot kernels (for roofline measurement and parity tests) -----------
 * y[M,N] = x[M,K] * W[N,K]^T with fp32 accumulation and one RNE rounding to bf16
 * (kernel/bmm.metal:24-82 through nn/linear.h:70-81).  M <= 4 takes the streaming GEMV path, larger M mc_gemm_bf16. */
MC_API mc_status mc_linear_bf16(mc_device* dev, mc_buffer* y, mc_buffer* x, mc_buffer* w, uint32_t M, uint32_t N, uint32_t K);
/* The prefill linear on the tcgen05 tensor cores (TMA-fed, fp32 accumulators in tensor memory), any M:
 *   mode 0: y[M,N]   = r(x . W^T)                               (nn/linear.h:70-81, kernel/bmm.metal:24-82)
 *   mode 2: y[M,N]   = r(res + r(x . W^T))                      (nn/transformer.h:133,139 residual add fused)
 *   mode 3: y[M,N/2] = r(silu_T(g) * u), (g,u) = columns (2i, 2i+1) of r(x . W^T), W = w1|w3 row-interleaved (nn/transformer.h:57-59)
 * K % 64 == 0, N % 32 == 0.  Runs `iters` times; elapsed_ms (may be NULL) receives the CUDA-event time of all of them. */
MC_API mc_status mc_gemm_bf16(mc_device* dev, mc_buffer* y, mc_buffer* x, mc_buffer* w, mc_buffer* res, uint32_t M, uint32_t N, uint32_t K, int mode,
                              uint32_t iters, float* elapsed_ms);
/* Decode attention of `rows` single-position queries (nn/attention.h:161-206 from the scores on: bmm, scalar_mul by T(1/sqrt(hd)),
 * softmax without max shift, bmm; repeat_kv replaced by kv = head / (H / KV)): q, out [rows, H*hd] bf16 (q already rotated), caches
 * [n_seqs][KV][max_seq][hd] bf16, row r attends positions 0 .. row_pos[r] of sequence row_seq[r] (host arrays).
 * kernel 0: attn_decode_kernel (per (row, head), cached positions split over a 4-CTA cluster); 1: decode_attn_gqa_kernel (per
 * (row, KV head), tensor-core tiles, the query heads of a KV head share one pass over its cache). */
MC_API mc_status mc_attn_decode(mc_device* dev, mc_buffer* out, mc_buffer* q, mc_buffer* kcache, mc_buffer* vcache, uint32_t rows, const int32_t* row_seq,
                                const int32_t* row_pos, uint32_t n_seqs, uint32_t H, uint32_t KV, uint32_t hd, uint32_t max_seq, int kernel);
/* Causal prompt attention (the same chain with make_causal_mask, nn/attention.h:283-299): `rows` consecutive positions of sequence `seq`
 * starting at start_pos; row i sees cache positions key_begin .. start_pos + i (key_begin = start_pos is the reference's chunk mask, Q9). */
MC_API mc_status mc_attn_prefill(mc_device* dev, mc_buffer* out, mc_buffer* q, mc_buffer* kcache, mc_buffer* vcache, uint32_t rows, uint32_t seq, uint32_t start_pos,
                                 uint32_t key_begin, uint32_t n_seqs, uint32_t H, uint32_t KV, uint32_t hd, uint32_t max_seq);
/* Embedding gather of the engine (kernel/embedding.metal:25-70): out[r, :] = table[ids[r], :]; with row_scales (fp32 [vocab]) the table is
 * int8 and the row is dequantised like lora_embedding, r(r(q) * r(s)) (quantization/lora.h:160-170, kernel/mul.metal:76-77). */
MC_API mc_status mc_embed_rows(mc_device* dev, mc_buffer* out, mc_buffer* table, mc_buffer* row_scales, const int32_t* ids, uint32_t rows, uint32_t D, uint32_t vocab);
/* QLoRA base linear over packed int4 (quantization/lora.h:94-122 + kernel/mul.metal:59-85 fused, without the adaptor):
 * y[M,N] = r( x[M,K] . r(r(q) * r(s))^T ), M <= 8.  w4 / scales_packed come from mc_pack_w4. */
MC_API mc_status mc_linear_w4(mc_device* dev, mc_buffer* y, mc_buffer* x, mc_buffer* w4, mc_buffer* scales_packed, uint32_t M, uint32_t N, uint32_t K);
/* Packs int8 [N,K] (values in [-8,7]; anything else is MC_ERR_INVALID) and fp32 scales [N,K/32] into the engine's
 * streaming layout (two weights per byte in tensor-core fragment order, bf16 r(scale)); mc_unpack_w4 is the exact inverse. */
MC_API mc_status mc_pack_w4(mc_device* dev, mc_buffer* w4, mc_buffer* scales_packed, mc_buffer* q8, mc_buffer* scales_f32, uint32_t N, uint32_t K);
MC_API mc_status mc_unpack_w4(mc_device* dev, mc_buffer* q8, mc_buffer* w4, uint32_t N, uint32_t K);
MC_API mc_status mc_w4_sizes(uint32_t N, uint32_t K, size_t* w4_nbytes, size_t* scales_nbytes);

#ifdef __cplusplus
}
#endif
#endif /* MC_CUDA_H */
