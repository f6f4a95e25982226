// oracle/ref/orc_mc_abi.cc — TEST INFRASTRUCTURE ONLY (never linked into the product, never shipped).
//
// A CPU implementation of the device / memory / kernel-lookup / encode-dispatch groups of include/mc_cuda.h whose
// kernels are the scalar oracle functions of oracle/orc_ops.h.  The façade translation units (metalchat_b200/facade/*.cc:
// the CUDA replacements of the reference's five Metal-bound TUs) are written against that C ABI only, so linking them
// with THIS library instead of libmc_cuda.so gives "the reference's own, unmodified header code (nn::llama3, attention,
// sink_cache, samplers, kernel wrappers) running over the oracle's op kernels on the host".  That is how the composition
// in oracle/orc_model.h is pinned: oracle/_ref/ref_driver_cpu and tests/test_oracle_ref.py require bit-equality between
// the two (VERDICT r01 "Next" #3).
//
// Execution model: "device memory" is host memory, a dispatch runs the kernel synchronously on the calling thread,
// commit runs the completion handlers, wait returns at once.  Argument slots follow kernel_thread.h:104-137: a tensor is
// tensor_layout<N> bytes followed by its buffer (+ byte offset), a scalar is raw bytes.
#include "../../include/mc_cuda.h"
#include "../orc_ops.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

using namespace orc;

struct mc_device {
    int ordinal = 0;
    std::atomic<uint64_t> launches{0};
};
struct mc_buffer {
    std::atomic<int> refs{1};
    void* ptr = nullptr;
    size_t size = 0;
    bool owned = true;
    mc_buffer* parent = nullptr;
};
struct mc_heap {
    mc_buffer* arena = nullptr;
    size_t used = 0;
};

namespace {

thread_local std::string g_err;
mc_status fail(mc_status c, const std::string& m)
{
    g_err = m;
    return c;
}

struct slot {
    int kind = 0; // 0 none, 1 bytes, 2 buffer
    unsigned char data[40];
    size_t nbytes = 0;
    mc_buffer* buf = nullptr;
    size_t offset = 0;
};
constexpr int kSlots = 16;
struct args {
    slot s[kSlots];
    template <int N> const layout<N>& lay(int i) const
    {
        if (s[i].kind != 1 || s[i].nbytes != sizeof(layout<N>)) throw std::invalid_argument("kernel argument " + std::to_string(i) + ": expected tensor_layout bytes");
        return *reinterpret_cast<const layout<N>*>(s[i].data);
    }
    template <typename T> T* ptr(int i) const
    {
        if (s[i].kind != 2 || !s[i].buf) throw std::invalid_argument("kernel argument " + std::to_string(i) + ": expected a buffer");
        return reinterpret_cast<T*>(static_cast<char*>(s[i].buf->ptr) + s[i].offset);
    }
    template <typename T> T scalar(int i) const
    {
        if (s[i].kind != 1 || s[i].nbytes < sizeof(T)) throw std::invalid_argument("kernel argument " + std::to_string(i) + ": expected scalar bytes");
        T v;
        std::memcpy(&v, s[i].data, sizeof(T));
        return v;
    }
};
using kernel_fn = std::function<void(const args&)>;
using bf = bf16_t;

// the 71 host names of metalchat.metallib, argument order = the reference's bind order
template <typename T> void reg_typed(std::map<std::string, kernel_fn>& r, const std::string& sfx)
{
    r["bmm_8_" + sfx] = [](const args& a) { bmm(a.ptr<T>(1), a.lay<3>(0), a.ptr<const T>(3), a.lay<3>(2), a.ptr<const T>(5), a.lay<3>(4)); };
    r["rmsnorm_" + sfx] = [](const args& a) {
        rmsnorm(a.ptr<T>(1), a.lay<2>(0), a.ptr<const T>(3), a.lay<2>(2), a.ptr<const T>(5), a.lay<1>(4), a.scalar<float>(6), a.scalar<float>(7), a.scalar<uint32_t>(8));
    };
    r["softmax_" + sfx] = [](const args& a) { softmax(a.ptr<T>(1), a.lay<2>(0), a.ptr<const T>(3), a.lay<2>(2), a.scalar<uint32_t>(4)); };
    r["sum_" + sfx] = [](const args& a) { sum(a.ptr<T>(1), a.lay<1>(0), a.ptr<const T>(3), a.lay<2>(2), a.scalar<uint32_t>(4)); };
    r["rope_" + sfx] = [](const args& a) {
        rope(a.ptr<T>(1), a.lay<2>(0), a.ptr<const T>(3), a.lay<2>(2), a.ptr<const float>(5), a.lay<2>(4), a.ptr<const float>(7), a.lay<2>(6), a.scalar<uint32_t>(8),
             a.scalar<uint32_t>(9), a.scalar<uint32_t>(10));
    };
    r["embedding_" + sfx] = [](const args& a) { embedding(a.ptr<T>(1), a.lay<3>(0), a.ptr<const int32_t>(3), a.lay<2>(2), a.ptr<const T>(5), a.lay<2>(4)); };
    r["sort_" + sfx] = [](const args& a) { sort(a.ptr<T>(1), a.lay<2>(0), a.ptr<int32_t>(3), a.lay<2>(2), a.ptr<const T>(5), a.lay<2>(4)); };
    for (uint32_t b = 2; b <= 1024; b *= 2)
        r["cumsum_" + std::to_string(b) + "_" + sfx] = [b](const args& a) { cumsum(a.ptr<T>(1), a.lay<2>(0), a.ptr<const T>(3), a.lay<2>(2), b); };
    r["multinomial_" + sfx] = [](const args& a) {
        multinomial(a.ptr<int32_t>(1), a.lay<2>(0), a.ptr<const T>(3), a.lay<2>(2), a.scalar<uint64_t>(4), a.scalar<uint64_t>(5), nullptr, 0,
                    (a.s[3].buf->size - a.s[3].offset) / sizeof(T));
    };
    r["add_" + sfx] = [](const args& a) { binary2(a.ptr<T>(1), a.lay<2>(0), a.ptr<const T>(3), a.lay<2>(2), a.ptr<const T>(5), a.lay<2>(4), [](float x, float y) { return x + y; }); };
    r["sub_" + sfx] = [](const args& a) { binary2(a.ptr<T>(1), a.lay<2>(0), a.ptr<const T>(3), a.lay<2>(2), a.ptr<const T>(5), a.lay<2>(4), [](float x, float y) { return x - y; }); };
    r["div_" + sfx] = [](const args& a) { binary2(a.ptr<T>(1), a.lay<2>(0), a.ptr<const T>(3), a.lay<2>(2), a.ptr<const T>(5), a.lay<2>(4), [](float x, float y) { return x / y; }); };
    r["hadamard_" + sfx] = [](const args& a) { binary2(a.ptr<T>(1), a.lay<2>(0), a.ptr<const T>(3), a.lay<2>(2), a.ptr<const T>(5), a.lay<2>(4), [](float x, float y) { return x * y; }); };
    r["add_broadcast_" + sfx] = [](const args& a) { add_broadcast(a.ptr<T>(1), a.lay<2>(0), a.ptr<const T>(3), a.lay<2>(2), a.ptr<const T>(5), a.lay<1>(4)); };
    r["scalar_mul_" + sfx] = [](const args& a) { scalar_mul(a.ptr<T>(1), a.lay<2>(0), a.ptr<const T>(3), a.lay<2>(2), a.scalar<T>(4)); };
    r["silu_" + sfx] = [](const args& a) { silu(a.ptr<T>(1), a.lay<2>(0), a.ptr<const T>(3), a.lay<2>(2)); };
    r["gelu_" + sfx] = [](const args& a) { gelu(a.ptr<T>(1), a.lay<2>(0), a.ptr<const T>(3), a.lay<2>(2)); };
    r["copy_" + sfx] = [](const args& a) { copy(a.ptr<T>(1), a.lay<2>(0), a.ptr<const T>(3), a.lay<2>(2)); };
    r["scatter_" + sfx] = [](const args& a) { scatter(a.ptr<T>(1), a.lay<2>(0), a.ptr<const uint8_t>(3), a.lay<2>(2), a.scalar<T>(4)); };
    r["gather_" + sfx] = [](const args& a) { gather(a.ptr<T>(1), a.lay<2>(0), a.ptr<const T>(3), a.lay<2>(2), a.ptr<const int32_t>(5), a.lay<2>(4)); };
    r["gt_" + sfx] = [](const args& a) { gt(a.ptr<uint8_t>(1), a.lay<2>(0), a.ptr<const T>(3), a.lay<2>(2), a.scalar<T>(4)); };
    r["le_" + sfx] = [](const args& a) { le(a.ptr<uint8_t>(1), a.lay<2>(0), a.ptr<const T>(3), a.lay<2>(2), a.scalar<T>(4)); };
    r["roll_" + sfx] = [](const args& a) { roll(a.ptr<T>(1), a.lay<1>(0), a.ptr<const T>(3), a.lay<1>(2), a.scalar<uint32_t>(4), a.scalar<uint32_t>(5), a.scalar<uint32_t>(6)); };
}
const std::map<std::string, kernel_fn>& registry()
{
    static const std::map<std::string, kernel_fn> r = [] {
        std::map<std::string, kernel_fn> m;
        reg_typed<bf>(m, "bfloat");
        reg_typed<float>(m, "float");
        m["rope_freqs_float"] = [](const args& a) {
            rope_freqs(a.ptr<float>(1), a.lay<2>(0), a.ptr<float>(3), a.lay<2>(2), a.scalar<uint32_t>(4), a.scalar<uint32_t>(5), a.scalar<float>(6));
        };
        m["copy_int32_t"] = [](const args& a) { copy(a.ptr<uint32_t>(1), a.lay<2>(0), a.ptr<const uint32_t>(3), a.lay<2>(2)); };
        m["gather_int32_t"] = [](const args& a) { gather(a.ptr<uint32_t>(1), a.lay<2>(0), a.ptr<const uint32_t>(3), a.lay<2>(2), a.ptr<const int32_t>(5), a.lay<2>(4)); };
        m["hadamard_broadcast_bfloat_int8_t_bfloat"] = [](const args& a) { hadamard_broadcast(a.ptr<bf>(1), a.lay<2>(0), a.ptr<const int8_t>(3), a.lay<2>(2), a.ptr<const bf>(5), a.lay<1>(4)); };
        m["hadamard_broadcast_bfloat_int8_t_float"] = [](const args& a) { hadamard_broadcast(a.ptr<bf>(1), a.lay<2>(0), a.ptr<const int8_t>(3), a.lay<2>(2), a.ptr<const float>(5), a.lay<1>(4)); };
        m["hadamard_broadcast_float_int8_t_bfloat"] = [](const args& a) { hadamard_broadcast(a.ptr<float>(1), a.lay<2>(0), a.ptr<const int8_t>(3), a.lay<2>(2), a.ptr<const bf>(5), a.lay<1>(4)); };
        m["hadamard_broadcast_float_int8_t_float"] = [](const args& a) { hadamard_broadcast(a.ptr<float>(1), a.lay<2>(0), a.ptr<const int8_t>(3), a.lay<2>(2), a.ptr<const float>(5), a.lay<1>(4)); };
        // the sampler's rope tables etc. are only instantiated in float; unknown names fail at lookup like newFunction does
        m.erase("rope_freqs_bfloat");
        return m;
    }();
    return r;
}

} // namespace

struct mc_kernel {
    std::string name;
    const kernel_fn* fn = nullptr;
};
struct mc_cmdbuf {
    mc_device* dev = nullptr;
    size_t capacity = 64, size = 0;
    bool committed = false;
    args a;
    std::vector<std::pair<void (*)(void*, int), void*>> handlers;
    std::string error;
};

#define ORC_BEGIN try {
#define ORC_END                                       \
    return MC_OK;                                     \
    }                                                 \
    catch (const std::invalid_argument& e)            \
    {                                                 \
        return fail(MC_ERR_INVALID, e.what());        \
    }                                                 \
    catch (const std::bad_alloc& e)                   \
    {                                                 \
        return fail(MC_ERR_ALLOC, e.what());          \
    }                                                 \
    catch (const std::exception& e)                   \
    {                                                 \
        return fail(MC_ERR_RUNTIME, e.what());        \
    }

extern "C" {

const char* mc_last_error(void) { return g_err.c_str(); }
const char* mc_version(void) { return "orc_mc_abi (CPU oracle backend of the C ABI; test infrastructure)"; }

mc_status mc_device_count(int* count)
{
    *count = 1;
    return MC_OK;
}
mc_status mc_device_create(int ordinal, mc_device** out)
{
    ORC_BEGIN
    if (ordinal != 0) throw std::invalid_argument("device ordinal out of range");
    *out = new mc_device();
    ORC_END
}
mc_status mc_device_destroy(mc_device* dev)
{
    delete dev;
    return MC_OK;
}
mc_status mc_device_name(mc_device*, char* out, size_t cap)
{
    snprintf(out, cap, "%s", "CPU oracle (orc_ops.h)");
    return MC_OK;
}
mc_status mc_device_max_buffer(mc_device*, size_t* bytes)
{
    *bytes = size_t(1) << 40;
    return MC_OK;
}
mc_status mc_device_synchronize(mc_device*) { return MC_OK; }

static mc_buffer* new_buffer(size_t size)
{
    auto* b = new mc_buffer();
    b->size = size;
    b->ptr = std::aligned_alloc(256, ((size ? size : 1) + 255) & ~size_t(255));
    if (!b->ptr) {
        delete b;
        throw std::bad_alloc();
    }
    return b;
}
mc_status mc_alloc(mc_device*, size_t size, int, mc_buffer** out)
{
    ORC_BEGIN
    *out = new_buffer(size);
    ORC_END
}
mc_status mc_alloc_copy(mc_device*, const void* src, size_t size, int, mc_buffer** out)
{
    ORC_BEGIN
    mc_buffer* b = new_buffer(size);
    if (std::getenv("MC_ORC_TRACE")) std::fprintf(stderr, "alloc_copy src %p size %zu first %08x\n", src, size, size >= 4 ? *static_cast<const uint32_t*>(src) : 0u);
    if (size) std::memcpy(b->ptr, src, size);
    *out = b;
    ORC_END
}
mc_status mc_wrap_host(mc_device*, void* host, size_t size, mc_buffer** out)
{
    ORC_BEGIN
    auto* b = new mc_buffer();
    b->ptr = host, b->size = size, b->owned = false;
    *out = b;
    ORC_END
}
mc_status mc_buffer_retain(mc_buffer* buf)
{
    buf->refs.fetch_add(1);
    return MC_OK;
}
mc_status mc_buffer_release(mc_buffer* buf)
{
    if (buf && buf->refs.fetch_sub(1) == 1) {
        if (buf->parent) mc_buffer_release(buf->parent);
        else if (buf->owned) std::free(buf->ptr);
        delete buf;
    }
    return MC_OK;
}
mc_status mc_buffer_host_ptr(mc_buffer* buf, void** out)
{
    *out = buf->ptr;
    return MC_OK;
}
mc_status mc_buffer_dev_ptr(mc_buffer* buf, void** out)
{
    *out = buf->ptr;
    return MC_OK;
}
mc_status mc_buffer_size(mc_buffer* buf, size_t* out)
{
    *out = buf->size;
    return MC_OK;
}
mc_status mc_heap_create(mc_device*, size_t capacity, mc_heap** out)
{
    ORC_BEGIN
    auto* h = new mc_heap();
    h->arena = new_buffer(capacity);
    *out = h;
    ORC_END
}
mc_status mc_heap_alloc(mc_heap* heap, size_t size, mc_buffer** out)
{
    ORC_BEGIN
    const size_t start = (heap->used + 255) & ~size_t(255);
    if (start + size > heap->arena->size) return fail(MC_ERR_ALLOC, "heap exhausted");
    auto* b = new mc_buffer();
    b->ptr = static_cast<char*>(heap->arena->ptr) + start, b->size = size, b->owned = false, b->parent = heap->arena;
    heap->arena->refs.fetch_add(1);
    heap->used = start + size;
    *out = b;
    ORC_END
}
mc_status mc_heap_reset(mc_heap* heap)
{
    heap->used = 0;
    return MC_OK;
}
mc_status mc_heap_destroy(mc_heap* heap)
{
    if (heap) {
        mc_buffer_release(heap->arena);
        delete heap;
    }
    return MC_OK;
}

mc_status mc_kernel_lookup(mc_device*, const char* name, mc_kernel** out)
{
    const auto& r = registry();
    auto it = r.find(name);
    if (it == r.end()) return fail(MC_ERR_NOT_FOUND, std::string("kernel not found: ") + name);
    auto* k = new mc_kernel();
    k->name = name, k->fn = &it->second;
    *out = k;
    return MC_OK;
}
mc_status mc_kernel_release(mc_kernel* k)
{
    delete k;
    return MC_OK;
}
mc_status mc_kernel_name(mc_kernel* k, const char** out)
{
    *out = k->name.c_str();
    return MC_OK;
}
mc_status mc_kernel_max_threads(mc_kernel*, size_t* out)
{
    *out = 1024; // the oracle fixes Apple's maxTotalThreadsPerThreadgroup at 1024: it sets the reduction partitions
    return MC_OK;
}
mc_status mc_kernel_count(int* count)
{
    *count = int(registry().size());
    return MC_OK;
}

mc_status mc_stream_begin(mc_device* dev, size_t capacity, mc_cmdbuf** out)
{
    auto* cb = new mc_cmdbuf();
    cb->dev = dev, cb->capacity = capacity ? capacity : 64;
    *out = cb;
    return MC_OK;
}
mc_status mc_set_bytes(mc_cmdbuf* cb, uint32_t index, const void* bytes, size_t size)
{
    if (index >= uint32_t(kSlots) || size == 0 || size > sizeof(cb->a.s[0].data)) return fail(MC_ERR_INVALID, "setBytes: bad slot or size");
    slot& s = cb->a.s[index];
    s.kind = 1, s.nbytes = size;
    std::memcpy(s.data, bytes, size);
    return MC_OK;
}
mc_status mc_set_buffer(mc_cmdbuf* cb, uint32_t index, mc_buffer* buf, size_t offset)
{
    if (index >= uint32_t(kSlots) || !buf) return fail(MC_ERR_INVALID, "setBuffer: bad slot or buffer");
    slot& s = cb->a.s[index];
    s.kind = 2, s.buf = buf, s.offset = offset;
    return MC_OK;
}
mc_status mc_barrier(mc_cmdbuf*, mc_buffer*) { return MC_OK; }
mc_status mc_dispatch(mc_cmdbuf* cb, mc_kernel* k, const uint32_t grid[3], const uint32_t group[3])
{
    ORC_BEGIN
    if (cb->committed) return fail(MC_ERR_RUNTIME, "command buffer already committed");
    if (cb->size >= cb->capacity) return fail(MC_ERR_FULL, "command buffer is full");
    const uint64_t threads = uint64_t(group[0]) * group[1] * group[2];
    if (threads == 0 || threads > 1024) throw std::invalid_argument("kernel: thread group exceeds 1024 threads");
    (void)grid;
    if (std::getenv("MC_ORC_TRACE")) {
        std::fprintf(stderr, "dispatch %s grid <%u,%u,%u> group <%u,%u,%u>:", k->name.c_str(), grid[0], grid[1], grid[2], group[0], group[1], group[2]);
        for (int i = 0; i < kSlots && cb->a.s[i].kind; i++) {
            if (cb->a.s[i].kind == 1) {
                std::fprintf(stderr, " [%d] %zu bytes (", i, cb->a.s[i].nbytes);
                for (size_t w = 0; w < cb->a.s[i].nbytes / 4 && w < 9; w++) std::fprintf(stderr, "%u ", reinterpret_cast<const uint32_t*>(cb->a.s[i].data)[w]);
                std::fprintf(stderr, ")");
            }
            else std::fprintf(stderr, " [%d] buf %p+%zu", i, cb->a.s[i].buf->ptr, cb->a.s[i].offset);
        }
        std::fprintf(stderr, "\n");
    }
    (*k->fn)(cb->a);
    if (std::getenv("MC_ORC_TRACE")) {
        for (int i = 0; i < kSlots && cb->a.s[i].kind; i++)
            if (cb->a.s[i].kind == 2) std::fprintf(stderr, "   after: [%d] first words %08x %08x (%g)\n", i, cb->a.ptr<uint32_t>(i)[0], cb->a.ptr<uint32_t>(i)[1], cb->a.ptr<float>(i)[0]);
    }
    cb->dev->launches.fetch_add(1);
    cb->size++;
    for (auto& s : cb->a.s) s = slot();
    ORC_END
}
mc_status mc_on_completed(mc_cmdbuf* cb, void (*fn)(void*, int), void* user)
{
    cb->handlers.emplace_back(fn, user);
    return MC_OK;
}
mc_status mc_commit(mc_cmdbuf* cb)
{
    if (cb->committed) return fail(MC_ERR_RUNTIME, "command buffer already committed");
    cb->committed = true;
    for (auto& h : cb->handlers) h.first(h.second, 0);
    return MC_OK;
}
mc_status mc_wait(mc_cmdbuf* cb, char* err, size_t cap)
{
    if (err && cap) err[0] = 0;
    if (!cb->committed) return mc_commit(cb);
    return MC_OK;
}
mc_status mc_cmdbuf_size(mc_cmdbuf* cb, size_t* n)
{
    *n = cb->size;
    return MC_OK;
}
mc_status mc_cmdbuf_release(mc_cmdbuf* cb)
{
    delete cb;
    return MC_OK;
}
mc_status mc_launch_count(mc_device* dev, uint64_t* launches)
{
    *launches = dev->launches.load();
    return MC_OK;
}

} // extern "C"
