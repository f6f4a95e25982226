// oracle/orc_common.h — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Shared primitives of the CPU oracle: bf16 conversion, tensor_layout PODs and
// the counter-hash generator used for synthetic weights.  The oracle restates
// the *semantics* of ybubnov/metalchat's Metal kernels (kernel/*.metal) and of
// the layer composition in include/metalchat/nn/*.h in scalar C++; every
// function cites the reference file:line it follows.  Nothing here is copied
// from the reference (GPL-3.0); it is a from-scratch restatement.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

// ---- bf16 -------------------------------------------------------------------
// Device-side `T(x)` conversion in the Metal kernels is round-to-nearest-even
// (kernel/bmm.metal:76, kernel/rmsnorm.metal:89 ...).  NaN is quieted the way the
// reference host type does it (include/metalchat/dtype.h:44-47).
static inline uint16_t f32_to_bf16(float f)
{
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) {
        return uint16_t((u >> 16) | 0x40u);
    }
    const uint32_t bias = 0x7fffu + ((u >> 16) & 1u);
    return uint16_t((u + bias) >> 16);
}

// Host-side bf16 assignment of the reference flushes fp32 subnormals and zero to
// signed zero (include/metalchat/dtype.h:36-42, quirk Q17).  Used only for
// host-generated data (weights, scalars).
static inline uint16_t f32_to_bf16_host(float f)
{
    uint32_t u;
    std::memcpy(&u, &f, 4);
    const int c = std::fpclassify(f);
    if (c == FP_SUBNORMAL || c == FP_ZERO) {
        return uint16_t((u >> 16) & 0x8000u);
    }
    return f32_to_bf16(f);
}

static inline float bf16_to_f32(uint16_t b)
{
    uint32_t u = uint32_t(b) << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

// Value types the templated kernels are instantiated with ("bfloat" / "float",
// include/metalchat/dtype.h:83-118).
struct bf16_t {
    uint16_t bits;
    bf16_t() : bits(0) {}
    bf16_t(float f) : bits(f32_to_bf16(f)) {}
    operator float() const { return bf16_to_f32(bits); }
};
static_assert(sizeof(bf16_t) == 2, "bf16_t must be 2 bytes");

// ---- tensor_layout<N> ----------------------------------------------------------
// The POD every kernel receives per tensor argument: element-unit sizes, strides
// and per-dimension offsets that are *added* to the address
// (include/metalchat/tensor/concept.h:24-33 == kernel/tensor.h:11-15,125-132).
template <int N> struct layout {
    uint32_t sizes[N];
    uint32_t strides[N];
    uint32_t offsets[N];
};

template <typename T> struct view1 {
    T* data;
    const layout<1>* l;
    T& at(uint32_t i) const { return data[l->strides[0] * i + l->offsets[0]]; }
    uint32_t size(int d) const { return l->sizes[d]; }
};
template <typename T> struct view2 {
    T* data;
    const layout<2>* l;
    T& at(uint32_t i, uint32_t j) const
    {
        return data[l->strides[0] * i + l->offsets[0] + l->strides[1] * j + l->offsets[1]];
    }
    uint32_t size(int d) const { return l->sizes[d]; }
};
template <typename T> struct view3 {
    T* data;
    const layout<3>* l;
    T& at(uint32_t i, uint32_t j, uint32_t k) const
    {
        return data
            [l->strides[0] * i + l->offsets[0] + l->strides[1] * j + l->offsets[1] +
             l->strides[2] * k + l->offsets[2]];
    }
    uint32_t size(int d) const { return l->sizes[d]; }
};

// ---- synthetic data generator (spec in DESIGN.md "Synthetic data") -----------
// Counter-based: value = f(seed, tensor_id, flat_index); no state, so the CPU
// oracle and the CUDA initialiser produce identical bits independently.
static inline uint64_t mix64(uint64_t z)
{
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
static inline uint64_t hash3(uint64_t seed, uint64_t tensor_id, uint64_t idx)
{
    return mix64(
        seed * 0x9E3779B97F4A7C15ull + tensor_id * 0xBF58476D1CE4E5B9ull +
        idx * 0x94D049BB133111EBull + 0x2545F4914F6CDD1Dull
    );
}
// U[-1, 1): exact in fp32 (24 random bits).
static inline float hash_uniform(uint64_t seed, uint64_t tensor_id, uint64_t idx)
{
    const uint32_t u24 = uint32_t(hash3(seed, tensor_id, idx) >> 40);
    return float(u24) * (1.0f / 8388608.0f) - 1.0f;
}
// integer in [lo, lo+range)
static inline int32_t hash_int(uint64_t seed, uint64_t tensor_id, uint64_t idx, int32_t lo, uint32_t range)
{
    const uint32_t u32 = uint32_t(hash3(seed, tensor_id, idx) >> 32);
    return lo + int32_t((uint64_t(u32) * range) >> 32);
}

} // namespace orc
