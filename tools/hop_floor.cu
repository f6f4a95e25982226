// Measures the floor of one grid-wide dependent exchange ("hop") on this GPU: what the streaming decode kernel pays between
// two phases of a transformer block even when the phase itself moves no weights.  One persistent CTA per SM; in hop h every
// CTA publishes its 1/148th of a W-word vector as tagged 8-byte words (payload, tag) and then polls the whole vector, the
// way metalchat_b200/csrc/mc_stream_kernel.cuh stages a phase input.  Variants: how many words travel, a plain atomic-counter
// grid barrier for comparison.   Build + run (GPU box):
//     nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/hop_floor tools/hop_floor.cu && /tmp/hop_floor
// DESIGN.md ("why batch-1 decode of a 1B model stops well below the HBM roofline") quotes its output.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                                     \
    do {                                                                                          \
        cudaError_t e_ = (x);                                                                     \
        if (e_ != cudaSuccess) {                                                                  \
            std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));       \
            std::exit(1);                                                                         \
        }                                                                                         \
    } while (0)

__device__ __forceinline__ void ll_store(uint64_t* p, uint32_t payload, uint32_t tag)
{
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(payload), "r"(tag) : "memory");
}
__device__ __forceinline__ void ll_load2(const uint64_t* p, uint64_t& a, uint64_t& b)
{
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

// words: tagged words per hop; every CTA stores words / gridDim.x of them (rounded up) and polls all of them
__global__ void __launch_bounds__(256, 1) hop_tagged(uint64_t* buf, uint32_t words, uint32_t hops, uint32_t tag0, float* sink)
{
    const uint32_t per = (words + gridDim.x - 1) / gridDim.x, w_b = min(words, blockIdx.x * per), w_e = min(words, w_b + per);
    float acc = 0.0f;
    for (uint32_t h = 0; h < hops; h++) {
        uint64_t* v = buf + size_t(h & 1u) * words;
        const uint32_t tag = tag0 + h + 1;
        for (uint32_t w = w_b + threadIdx.x; w < w_e; w += blockDim.x) ll_store(v + w, __float_as_uint(acc) + w, tag);
        for (uint32_t w = threadIdx.x * 2; w < words; w += blockDim.x * 2) {
            uint64_t a, b;
            ll_load2(v + w, a, b);
            while (uint32_t(a >> 32) != tag || uint32_t(b >> 32) != tag) ll_load2(v + w, a, b);
            acc += __uint_as_float(uint32_t(a)) * 1e-30f + __uint_as_float(uint32_t(b)) * 1e-30f;
        }
        __syncthreads(); // the staged vector is complete in this CTA (the engine's consumer barrier)
    }
    if (acc == 123.0f) *sink = acc;
}

// tagged words, but a thread that finds a word missing sleeps before it asks again (fewer polls in the way of the stores)
__global__ void __launch_bounds__(256, 1) hop_tagged_sleep(uint64_t* buf, uint32_t words, uint32_t hops, uint32_t tag0, float* sink, uint32_t ns)
{
    const uint32_t per = (words + gridDim.x - 1) / gridDim.x, w_b = min(words, blockIdx.x * per), w_e = min(words, w_b + per);
    float acc = 0.0f;
    for (uint32_t h = 0; h < hops; h++) {
        uint64_t* v = buf + size_t(h & 1u) * words;
        const uint32_t tag = tag0 + h + 1;
        for (uint32_t w = w_b + threadIdx.x; w < w_e; w += blockDim.x) ll_store(v + w, __float_as_uint(acc) + w, tag);
        for (uint32_t w = threadIdx.x * 2; w < words; w += blockDim.x * 2) {
            uint64_t a, b;
            ll_load2(v + w, a, b);
            while (uint32_t(a >> 32) != tag || uint32_t(b >> 32) != tag) {
                __nanosleep(ns);
                ll_load2(v + w, a, b);
            }
            acc += __uint_as_float(uint32_t(a)) * 1e-30f + __uint_as_float(uint32_t(b)) * 1e-30f;
        }
        __syncthreads();
    }
    if (acc == 123.0f) *sink = acc;
}

// data as plain 4-byte words + arrival counters: every CTA stores its slice, one thread adds 1 (release) to counter
// [blockIdx % n_ctr] of the hop; one thread per CTA polls the n_ctr counters (acquire), then everybody loads the vector from L2
__global__ void __launch_bounds__(256, 1) hop_counted(uint32_t* data, unsigned* ctr, uint32_t words, uint32_t hops, uint32_t n_ctr, unsigned base, float* sink)
{
    const uint32_t per = (words + gridDim.x - 1) / gridDim.x, w_b = min(words, blockIdx.x * per), w_e = min(words, w_b + per);
    float acc = 0.0f;
    for (uint32_t h = 0; h < hops; h++) {
        uint32_t* v = data + size_t(h & 1u) * words;
        unsigned* cc = ctr + (h & 1u) * 32;
        for (uint32_t w = w_b + threadIdx.x; w < w_e; w += blockDim.x) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(v + w), "r"(__float_as_uint(acc) + w) : "memory");
        __syncthreads();
        if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(cc + blockIdx.x % n_ctr) : "memory");
        if (threadIdx.x < n_ctr) {
            // counter i collects the CTAs with blockIdx % n_ctr == i; each is reused every other hop
            const unsigned members = (gridDim.x - threadIdx.x + n_ctr - 1) / n_ctr, want = base + (h / 2 + 1) * members;
            unsigned got;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(got) : "l"(cc + threadIdx.x) : "memory");
            } while (got < want);
        }
        __syncthreads();
        for (uint32_t w = threadIdx.x * 4; w < words; w += blockDim.x * 4) {
            uint32_t a, b, c, d;
            asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(v + w) : "memory");
            acc += __uint_as_float(a) * 1e-30f + __uint_as_float(b + c + d) * 1e-30f;
        }
        __syncthreads();
    }
    if (acc == 123.0f) *sink = acc;
}

// the same chain through an atomic counter: arrive, spin until everybody arrived
__global__ void __launch_bounds__(256, 1) hop_counter(unsigned* ctr, uint32_t hops, unsigned base)
{
    for (uint32_t h = 0; h < hops; h++) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(ctr, 1u);
            const unsigned want = base + (h + 1) * gridDim.x;
            unsigned v;
            do {
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            } while (v < want);
        }
        __syncthreads();
    }
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int ctas = prop.multiProcessorCount;
    const uint32_t hops = 4000;
    uint64_t* buf;
    float* sink;
    unsigned* ctr;
    CK(cudaMalloc(&buf, 2 * 16384 * sizeof(uint64_t)));
    CK(cudaMemset(buf, 0, 2 * 16384 * sizeof(uint64_t)));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMalloc(&ctr, 4));
    CK(cudaMemset(ctr, 0, 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    std::printf("{\"sms\": %d, \"hops\": %u", ctas, hops);
    uint32_t tag0 = 0;
    for (uint32_t words : {148u, 1024u, 2048u, 4096u, 8192u}) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaEventRecord(e0));
            hop_tagged<<<ctas, 256>>>(buf, words, hops, tag0, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best) best = ms;
            tag0 += hops;
        }
        std::printf(", \"tagged_%u_words_us\": %.3f", words, best * 1000.0f / hops);
    }
    for (uint32_t ns : {32u, 128u}) {
        for (uint32_t words : {1024u, 4096u}) {
            float best = 1e30f;
            for (int rep = 0; rep < 4; rep++) {
                CK(cudaEventRecord(e0));
                hop_tagged_sleep<<<ctas, 256>>>(buf, words, hops, tag0, sink, ns);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (rep && ms < best) best = ms;
                tag0 += hops;
            }
            std::printf(", \"tagged_sleep%u_%u_words_us\": %.3f", ns, words, best * 1000.0f / hops);
        }
    }
    {
        uint32_t* data;
        unsigned* cc;
        CK(cudaMalloc(&data, 2 * 16384 * 4));
        CK(cudaMemset(data, 0, 2 * 16384 * 4));
        CK(cudaMalloc(&cc, 64 * 4));
        for (uint32_t n_ctr : {1u, 4u, 16u}) {
            for (uint32_t words : {1024u, 4096u}) {
                CK(cudaMemset(cc, 0, 64 * 4));
                unsigned base = 0;
                float best = 1e30f;
                for (int rep = 0; rep < 4; rep++) {
                    CK(cudaEventRecord(e0));
                    hop_counted<<<ctas, 256>>>(data, cc, words, hops, n_ctr, base, sink);
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    float ms;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (rep && ms < best) best = ms;
                    CK(cudaMemset(cc, 0, 64 * 4));
                }
                std::printf(", \"counted%u_%u_words_us\": %.3f", n_ctr, words, best * 1000.0f / hops);
            }
        }
    }
    {
        float best = 1e30f;
        unsigned base = 0;
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaEventRecord(e0));
            hop_counter<<<ctas, 256>>>(ctr, hops, base);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best) best = ms;
            base += hops * ctas;
        }
        std::printf(", \"atomic_counter_us\": %.3f", best * 1000.0f / hops);
    }
    std::printf("}\n");
    return 0;
}
