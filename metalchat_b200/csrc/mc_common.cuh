// metalchat_b200/csrc/mc_common.cuh — internal types shared by the translation units of
// libmc_cuda.so (runtime, op-level kernels, decode engine).  Not part of the C ABI.
#pragma once
#include "../../include/mc_cuda.h"

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace mc {

// ---- errors ---------------------------------------------------------------------------
// The reference throws C++ exceptions (SURVEY §8b "Errors"); the C ABI turns them into an
// mc_status plus a thread-local message.
struct error : std::runtime_error {
    mc_status code;
    error(mc_status c, const std::string& m) : std::runtime_error(m), code(c) {}
};
void set_last_error(const std::string& m);
mc_status fail(mc_status code, const std::string& m);

#define MC_CUDA_CHECK(expr)                                                                       \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            throw ::mc::error(                                                                    \
                (_e == cudaErrorMemoryAllocation) ? MC_ERR_ALLOC : MC_ERR_RUNTIME,                \
                std::string("cuda: ") + cudaGetErrorString(_e) + " at " #expr                     \
            );                                                                                    \
        }                                                                                         \
    } while (0)

#define MC_REQUIRE(cond, msg)                                                                     \
    do {                                                                                          \
        if (!(cond)) throw ::mc::error(MC_ERR_INVALID, std::string(msg));                         \
    } while (0)

// NVTX range for the span of a C-ABI call (header-only NVTX v3: a no-op unless a tool is attached; shows up in nsys / ncu --nvtx)
struct nvtx_range {
    explicit nvtx_range(const char* name) { nvtxRangePushA(name); }
    ~nvtx_range() { nvtxRangePop(); }
};

#define MC_API_BEGIN try {
#define MC_API_END                                                                                \
    return MC_OK;                                                                                 \
    }                                                                                             \
    catch (const ::mc::error& e) { return ::mc::fail(e.code, e.what()); }                         \
    catch (const std::bad_alloc&) { return ::mc::fail(MC_ERR_ALLOC, "out of host memory"); }      \
    catch (const std::exception& e) { return ::mc::fail(MC_ERR_RUNTIME, e.what()); }

// ---- tensor_layout<N> (tensor/concept.h:24-33 == kernel/tensor.h:11-15) ----------------
template <int N> struct layout {
    uint32_t sizes[N];
    uint32_t strides[N];
    uint32_t offsets[N];
};

template <typename T, int N> struct tview;
template <typename T> struct tview<T, 1> {
    T* data;
    layout<1> l;
    __device__ __forceinline__ T& at(uint32_t i) const { return data[l.strides[0] * i + l.offsets[0]]; }
    __host__ __device__ uint32_t size(int d) const { return l.sizes[d]; }
};
template <typename T> struct tview<T, 2> {
    T* data;
    layout<2> l;
    __device__ __forceinline__ T& at(uint32_t i, uint32_t j) const
    {
        return data[l.strides[0] * i + l.offsets[0] + l.strides[1] * j + l.offsets[1]];
    }
    __host__ __device__ uint32_t size(int d) const { return l.sizes[d]; }
};
template <typename T> struct tview<T, 3> {
    T* data;
    layout<3> l;
    __device__ __forceinline__ T& at(uint32_t i, uint32_t j, uint32_t k) const
    {
        return data
            [l.strides[0] * i + l.offsets[0] + l.strides[1] * j + l.offsets[1] + l.strides[2] * k +
             l.offsets[2]];
    }
    __host__ __device__ uint32_t size(int d) const { return l.sizes[d]; }
};

// ---- bf16 ------------------------------------------------------------------------------
// Storage type of the "bfloat" kernels.  Conversion is round-to-nearest-even with the
// reference's NaN quieting (include/metalchat/dtype.h:26-57): identical bits to the host type.
struct bf16 {
    uint16_t bits;
};
__host__ __device__ __forceinline__ uint16_t f32_to_bf16_bits(float f)
{
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
#endif
    if ((u & 0x7fffffffu) > 0x7f800000u) return uint16_t((u >> 16) | 0x40u);
    return uint16_t((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}
__host__ __device__ __forceinline__ float bf16_bits_to_f32(uint16_t b)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(uint32_t(b) << 16);
#else
    uint32_t u = uint32_t(b) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
// value <-> fp32 for the two instantiated element types
__host__ __device__ __forceinline__ float to_f32(float v) { return v; }
__host__ __device__ __forceinline__ float to_f32(bf16 v) { return bf16_bits_to_f32(v.bits); }
template <typename T> __host__ __device__ __forceinline__ T from_f32(float f);
template <> __host__ __device__ __forceinline__ float from_f32<float>(float f) { return f; }
template <> __host__ __device__ __forceinline__ bf16 from_f32<bf16>(float f) { return bf16{f32_to_bf16_bits(f)}; }
// round-trip through T (the r(.) of SURVEY §8a)
template <typename T> __host__ __device__ __forceinline__ float round_to(float f) { return to_f32(from_f32<T>(f)); }

// ---- synthetic data generator (same integer hash as the test oracle) ----------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
__host__ __device__ __forceinline__ uint64_t hash3(uint64_t seed, uint64_t tensor_id, uint64_t idx)
{
    return mix64(
        seed * 0x9E3779B97F4A7C15ull + tensor_id * 0xBF58476D1CE4E5B9ull + idx * 0x94D049BB133111EBull +
        0x2545F4914F6CDD1Dull
    );
}

// ---- handles -----------------------------------------------------------------------------
} // namespace mc

struct mc_device {
    int ordinal = 0;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop{};
    std::atomic<uint64_t> launches{0};
    // per-device state of the op-level kernels (one process may drive several GPUs)
    int* bad_ids = nullptr;      // device view of the flag below
    int* bad_ids_host = nullptr; // mapped pinned int: an embedding kernel saw an out-of-range id (read back by mc_wait)
    bool sampler_configured = false;
};

struct mc_buffer {
    std::atomic<int> refs{1};
    mc_device* dev = nullptr;
    void* dptr = nullptr;  // device-visible address
    void* hptr = nullptr;  // host-visible address (null for MC_MEM_DEVICE)
    size_t size = 0;
    int kind = 0;          // MC_MEM_* or 100 = registered host wrap, 101 = heap slice
    mc_buffer* parent = nullptr;
};

struct mc_heap {
    mc_device* dev = nullptr;
    mc_buffer* arena = nullptr;
    size_t used = 0;
};

namespace mc {

struct arg_slot {
    enum kind_t : uint8_t { none = 0, bytes = 1, buffer = 2 } kind = none;
    uint8_t nbytes = 0;
    alignas(8) unsigned char data[40];
    mc_buffer* buf = nullptr;
    size_t offset = 0;
};
constexpr int kMaxSlots = 16;

struct arg_pack {
    arg_slot slot[kMaxSlots];
    mc_device* dev = nullptr; // the device the kernel is dispatched on (per-device scratch lives there)
    // typed accessors used by the launchers; throw invalid_argument on a mismatch
    template <int N> layout<N> lay(int i) const
    {
        MC_REQUIRE(i < kMaxSlots && slot[i].kind == arg_slot::bytes && slot[i].nbytes == sizeof(layout<N>),
                   "kernel argument " + std::to_string(i) + ": expected tensor_layout<" + std::to_string(N) + "> bytes");
        layout<N> l;
        memcpy(&l, slot[i].data, sizeof(l));
        return l;
    }
    template <typename T> T* ptr(int i) const
    {
        MC_REQUIRE(i < kMaxSlots && slot[i].kind == arg_slot::buffer && slot[i].buf, "kernel argument " + std::to_string(i) + ": expected a buffer");
        return reinterpret_cast<T*>(static_cast<char*>(slot[i].buf->dptr) + slot[i].offset);
    }
    template <typename T> T scalar(int i) const
    {
        MC_REQUIRE(i < kMaxSlots && slot[i].kind == arg_slot::bytes && slot[i].nbytes >= sizeof(T),
                   "kernel argument " + std::to_string(i) + ": expected " + std::to_string(sizeof(T)) + " scalar bytes");
        T v;
        memcpy(&v, slot[i].data, sizeof(T));
        return v;
    }
    // tensor argument occupying slots (i, i+1)
    template <typename T, int N> tview<T, N> tensor(int i) const { return tview<T, N>{ptr<T>(i + 1), lay<N>(i)}; }
};

using launcher_fn = void (*)(const arg_pack&, cudaStream_t);

struct kernel_entry {
    const char* name;
    launcher_fn launch;
};
const std::vector<kernel_entry>& kernel_registry();

} // namespace mc

struct mc_kernel {
    const mc::kernel_entry* entry = nullptr;
    mc_device* dev = nullptr;
};

struct mc_cmdbuf {
    mc_device* dev = nullptr;
    size_t capacity = 64;
    size_t size = 0;
    bool committed = false;
    mc::arg_pack args;
    cudaEvent_t done = nullptr;
    std::vector<std::pair<void (*)(void*, int), void*>> handlers;
};
