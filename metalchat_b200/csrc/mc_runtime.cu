// metalchat_b200/csrc/mc_runtime.cu — device, memory, kernel lookup and command-buffer
// entry points of the C ABI (include/mc_cuda.h).  This file is the CUDA counterpart of the
// reference's five Metal-bound translation units: src/metal.cc, src/accelerator.cc,
// src/allocator.cc, src/kernel.cc and src/kernel_thread.cc.  A Metal command queue with a
// chain of command buffers becomes ONE in-order CUDA stream; "commit" records an event,
// "wait" synchronises it and surfaces asynchronous errors (src/kernel_thread.cc:134-199).
#include "mc_common.cuh"

#include <mutex>

namespace mc {

static thread_local std::string g_last_error;
void set_last_error(const std::string& m) { g_last_error = m; }
mc_status fail(mc_status code, const std::string& m)
{
    g_last_error = m;
    return code;
}

static void use(mc_device* dev)
{
    MC_REQUIRE(dev != nullptr, "null device");
    MC_CUDA_CHECK(cudaSetDevice(dev->ordinal));
}

} // namespace mc

using namespace mc;

extern "C" {

const char* mc_last_error(void) { return g_last_error.c_str(); }
const char* mc_version(void) { return "metalchat_b200 0.1 (sm_100a)"; }

// ---- device -------------------------------------------------------------------------------
mc_status mc_device_count(int* count)
{
    MC_API_BEGIN
    MC_REQUIRE(count, "null out pointer");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        cudaGetLastError();
    }
    MC_API_END
}

mc_status mc_device_create(int ordinal, mc_device** out)
{
    MC_API_BEGIN
    MC_REQUIRE(out, "null out pointer");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        // the reference throws runtime_error when no Metal device/library is found (test/test_accelerator.cc:15-21)
        throw error(MC_ERR_RUNTIME, "cuda: no CUDA device found (libmc_cuda has no CPU fallback)");
    }
    MC_REQUIRE(ordinal >= 0 && ordinal < n, "device ordinal out of range");
    auto* dev = new mc_device();
    dev->ordinal = ordinal;
    MC_CUDA_CHECK(cudaSetDevice(ordinal));
    MC_CUDA_CHECK(cudaGetDeviceProperties(&dev->prop, ordinal));
    if (dev->prop.major != 10) {
        std::string nm = dev->prop.name;
        delete dev;
        throw error(MC_ERR_RUNTIME, "cuda: device '" + nm + "' is not sm_100; this library only ships sm_100a code");
    }
    {
        const cudaError_t e = cudaStreamCreateWithFlags(&dev->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete dev;
            MC_CUDA_CHECK(e);
        }
    }
    *out = dev;
    MC_API_END
}

mc_status mc_device_destroy(mc_device* dev)
{
    MC_API_BEGIN
    if (dev) {
        cudaSetDevice(dev->ordinal);
        if (dev->stream) {
            cudaStreamSynchronize(dev->stream);
            cudaStreamDestroy(dev->stream);
        }
        if (dev->bad_ids_host) cudaFreeHost(dev->bad_ids_host);
        delete dev;
    }
    MC_API_END
}

mc_status mc_device_name(mc_device* dev, char* out, size_t cap)
{
    MC_API_BEGIN
    MC_REQUIRE(dev && out && cap > 0, "bad arguments");
    snprintf(out, cap, "%s", dev->prop.name);
    MC_API_END
}

mc_status mc_device_max_buffer(mc_device* dev, size_t* bytes)
{
    MC_API_BEGIN
    MC_REQUIRE(dev && bytes, "bad arguments");
    *bytes = dev->prop.totalGlobalMem;
    MC_API_END
}

mc_status mc_device_sm_count(mc_device* dev, int* count)
{
    MC_API_BEGIN
    MC_REQUIRE(dev && count, "bad arguments");
    *count = dev->prop.multiProcessorCount;
    MC_API_END
}

mc_status mc_device_synchronize(mc_device* dev)
{
    MC_API_BEGIN
    use(dev);
    MC_CUDA_CHECK(cudaStreamSynchronize(dev->stream));
    MC_API_END
}

mc_status mc_device_stream(mc_device* dev, void** cuda_stream)
{
    MC_API_BEGIN
    MC_REQUIRE(dev && cuda_stream, "bad arguments");
    *cuda_stream = static_cast<void*>(dev->stream);
    MC_API_END
}

// ---- memory ---------------------------------------------------------------------------------
static mc_buffer* new_buffer(mc_device* dev, size_t size, int flags)
{
    use(dev);
    auto* b = new mc_buffer();
    b->dev = dev;
    b->size = size;
    b->kind = flags;
    const size_t n = size ? size : 1;
    cudaError_t e;
    switch (flags) {
    case MC_MEM_DEVICE: e = cudaMalloc(&b->dptr, n); break;
    case MC_MEM_SHARED:
        e = cudaMallocManaged(&b->dptr, n, cudaMemAttachGlobal);
        b->hptr = b->dptr;
        break;
    case MC_MEM_PINNED:
        e = cudaHostAlloc(&b->hptr, n, cudaHostAllocMapped | cudaHostAllocPortable);
        if (e == cudaSuccess) e = cudaHostGetDevicePointer(&b->dptr, b->hptr, 0);
        break;
    default: delete b; throw error(MC_ERR_INVALID, "unknown memory flags");
    }
    if (e != cudaSuccess) {
        delete b;
        cudaGetLastError();
        // alloc_error in the reference (src/allocator.cc:100-107)
        throw error(MC_ERR_ALLOC, std::string("cuda: allocation of ") + std::to_string(size) + " bytes failed: " + cudaGetErrorString(e));
    }
    return b;
}

mc_status mc_alloc(mc_device* dev, size_t size, int flags, mc_buffer** out)
{
    MC_API_BEGIN
    MC_REQUIRE(out, "null out pointer");
    *out = new_buffer(dev, size, flags);
    MC_API_END
}

mc_status mc_alloc_copy(mc_device* dev, const void* src, size_t size, int flags, mc_buffer** out)
{
    MC_API_BEGIN
    MC_REQUIRE(out && (src || size == 0), "bad arguments");
    mc_buffer* b = new_buffer(dev, size, flags);
    if (size) {
        cudaError_t e = cudaMemcpyAsync(b->dptr, src, size, cudaMemcpyDefault, dev->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(dev->stream);
        if (e != cudaSuccess) {
            mc_buffer_release(b);
            MC_CUDA_CHECK(e);
        }
    }
    *out = b;
    MC_API_END
}

mc_status mc_wrap_host(mc_device* dev, void* host, size_t size, mc_buffer** out)
{
    MC_API_BEGIN
    MC_REQUIRE(out && host && size, "bad arguments");
    use(dev);
    auto* b = new mc_buffer();
    b->dev = dev;
    b->size = size;
    b->kind = 100;
    b->hptr = host;
    cudaError_t e = cudaHostRegister(host, size, cudaHostRegisterMapped | cudaHostRegisterPortable);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer(&b->dptr, host, 0);
    if (e != cudaSuccess) {
        delete b;
        cudaGetLastError();
        throw error(MC_ERR_ALLOC, std::string("cuda: cannot register host memory: ") + cudaGetErrorString(e));
    }
    *out = b;
    MC_API_END
}

mc_status mc_buffer_retain(mc_buffer* buf)
{
    MC_API_BEGIN
    MC_REQUIRE(buf, "null buffer");
    buf->refs.fetch_add(1);
    MC_API_END
}

mc_status mc_buffer_release(mc_buffer* buf)
{
    MC_API_BEGIN
    if (buf && buf->refs.fetch_sub(1) == 1) {
        if (buf->dev) cudaSetDevice(buf->dev->ordinal);
        switch (buf->kind) {
        case MC_MEM_DEVICE:
        case MC_MEM_SHARED: cudaFree(buf->dptr); break;
        case MC_MEM_PINNED: cudaFreeHost(buf->hptr); break;
        case 100: cudaHostUnregister(buf->hptr); break;
        case 101:
            if (buf->parent) mc_buffer_release(buf->parent);
            break;
        }
        delete buf;
    }
    MC_API_END
}

mc_status mc_buffer_host_ptr(mc_buffer* buf, void** out)
{
    MC_API_BEGIN
    MC_REQUIRE(buf && out, "bad arguments");
    if (!buf->hptr) throw error(MC_ERR_INVALID, "buffer is device-resident (MC_MEM_DEVICE) and not host dereferenceable");
    *out = buf->hptr;
    MC_API_END
}

mc_status mc_buffer_dev_ptr(mc_buffer* buf, void** out)
{
    MC_API_BEGIN
    MC_REQUIRE(buf && out, "bad arguments");
    *out = buf->dptr;
    MC_API_END
}

mc_status mc_buffer_size(mc_buffer* buf, size_t* out)
{
    MC_API_BEGIN
    MC_REQUIRE(buf && out, "bad arguments");
    *out = buf->size;
    MC_API_END
}

mc_status mc_memcpy_h2d(mc_device* dev, mc_buffer* dst, size_t dst_off, const void* src, size_t size)
{
    MC_API_BEGIN
    use(dev);
    MC_REQUIRE(dst && (src || !size), "bad arguments");
    MC_REQUIRE(dst_off + size <= dst->size, "copy exceeds the destination buffer");
    MC_CUDA_CHECK(cudaMemcpyAsync(static_cast<char*>(dst->dptr) + dst_off, src, size, cudaMemcpyDefault, dev->stream));
    MC_CUDA_CHECK(cudaStreamSynchronize(dev->stream));
    MC_API_END
}

mc_status mc_memcpy_d2h(mc_device* dev, void* dst, mc_buffer* src, size_t src_off, size_t size)
{
    MC_API_BEGIN
    use(dev);
    MC_REQUIRE(src && (dst || !size), "bad arguments");
    MC_REQUIRE(src_off + size <= src->size, "copy exceeds the source buffer");
    MC_CUDA_CHECK(cudaMemcpyAsync(dst, static_cast<char*>(src->dptr) + src_off, size, cudaMemcpyDefault, dev->stream));
    MC_CUDA_CHECK(cudaStreamSynchronize(dev->stream));
    MC_API_END
}

mc_status mc_memset(mc_device* dev, mc_buffer* dst, size_t dst_off, int value, size_t size)
{
    MC_API_BEGIN
    use(dev);
    MC_REQUIRE(dst && dst_off + size <= dst->size, "memset exceeds the buffer");
    MC_CUDA_CHECK(cudaMemsetAsync(static_cast<char*>(dst->dptr) + dst_off, value, size, dev->stream));
    MC_API_END
}

mc_status mc_heap_create(mc_device* dev, size_t capacity, mc_heap** out)
{
    MC_API_BEGIN
    MC_REQUIRE(out, "null out pointer");
    auto* h = new mc_heap();
    h->dev = dev;
    try {
        h->arena = new_buffer(dev, capacity, MC_MEM_DEVICE);
    } catch (...) {
        delete h;
        throw;
    }
    *out = h;
    MC_API_END
}

mc_status mc_heap_alloc(mc_heap* heap, size_t size, mc_buffer** out)
{
    MC_API_BEGIN
    MC_REQUIRE(heap && out, "bad arguments");
    const size_t start = (heap->used + 255) & ~size_t(255);
    if (start + size > heap->arena->size) {
        // hardware_heap_allocator throws alloc_error when the heap is exhausted (src/allocator.cc:100-107)
        throw error(MC_ERR_ALLOC, "heap exhausted: " + std::to_string(size) + " bytes requested, " + std::to_string(heap->arena->size - std::min(start, heap->arena->size)) + " free");
    }
    auto* b = new mc_buffer();
    b->dev = heap->dev;
    b->dptr = static_cast<char*>(heap->arena->dptr) + start;
    b->size = size;
    b->kind = 101;
    b->parent = heap->arena;
    heap->arena->refs.fetch_add(1);
    heap->used = start + size;
    *out = b;
    MC_API_END
}

mc_status mc_heap_reset(mc_heap* heap)
{
    MC_API_BEGIN
    MC_REQUIRE(heap, "null heap");
    // slices hold a reference on the arena: rewinding under live slices would hand the same bytes out twice
    if (heap->arena->refs.load() != 1) throw error(MC_ERR_INVALID, "heap reset while slices of it are still alive");
    heap->used = 0;
    MC_API_END
}

mc_status mc_heap_destroy(mc_heap* heap)
{
    MC_API_BEGIN
    if (heap) {
        mc_buffer_release(heap->arena);
        delete heap;
    }
    MC_API_END
}

// ---- kernels ------------------------------------------------------------------------------------
mc_status mc_kernel_lookup(mc_device* dev, const char* name, mc_kernel** out)
{
    MC_API_BEGIN
    MC_REQUIRE(dev && name && out, "bad arguments");
    for (const auto& e : kernel_registry()) {
        if (strcmp(e.name, name) == 0) {
            auto* k = new mc_kernel();
            k->entry = &e;
            k->dev = dev;
            *out = k;
            return MC_OK;
        }
    }
    // hardware_accelerator::load throws when newFunction fails (src/accelerator.cc:121-127)
    throw error(MC_ERR_NOT_FOUND, std::string("cuda: kernel not found: ") + name);
    MC_API_END
}

mc_status mc_kernel_release(mc_kernel* k)
{
    MC_API_BEGIN
    delete k;
    MC_API_END
}

mc_status mc_kernel_name(mc_kernel* k, const char** out)
{
    MC_API_BEGIN
    MC_REQUIRE(k && out, "bad arguments");
    *out = k->entry->name;
    MC_API_END
}

mc_status mc_kernel_max_threads(mc_kernel* k, size_t* out)
{
    MC_API_BEGIN
    MC_REQUIRE(k && out, "bad arguments");
    *out = 1024; // maxTotalThreadsPerThreadgroup on Apple GPUs and maxThreadsPerBlock on sm_100
    MC_API_END
}

mc_status mc_kernel_count(int* count)
{
    MC_API_BEGIN
    MC_REQUIRE(count, "null out pointer");
    *count = int(kernel_registry().size());
    MC_API_END
}

mc_status mc_kernel_name_at(int index, const char** out)
{
    MC_API_BEGIN
    MC_REQUIRE(out && index >= 0 && size_t(index) < kernel_registry().size(), "index out of range");
    *out = kernel_registry()[size_t(index)].name;
    MC_API_END
}

// ---- command buffers -------------------------------------------------------------------------------
mc_status mc_stream_begin(mc_device* dev, size_t capacity, mc_cmdbuf** out)
{
    MC_API_BEGIN
    use(dev);
    MC_REQUIRE(out, "null out pointer");
    auto* cb = new mc_cmdbuf();
    cb->dev = dev;
    cb->capacity = capacity ? capacity : 64; // kernel_thread default capacity (accelerator.h:84-92)
    cudaError_t e = cudaEventCreateWithFlags(&cb->done, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        delete cb;
        MC_CUDA_CHECK(e);
    }
    *out = cb;
    MC_API_END
}

static arg_slot& slot_at(mc_cmdbuf* cb, uint32_t index)
{
    MC_REQUIRE(cb, "null command buffer");
    MC_REQUIRE(index < uint32_t(kMaxSlots), "argument index out of range");
    return cb->args.slot[index];
}

mc_status mc_set_bytes(mc_cmdbuf* cb, uint32_t index, const void* bytes, size_t size)
{
    MC_API_BEGIN
    arg_slot& s = slot_at(cb, index);
    MC_REQUIRE(bytes && size > 0 && size <= sizeof(s.data), "setBytes: unsupported size");
    s.kind = arg_slot::bytes;
    s.nbytes = uint8_t(size);
    memcpy(s.data, bytes, size);
    MC_API_END
}

mc_status mc_set_buffer(mc_cmdbuf* cb, uint32_t index, mc_buffer* buf, size_t offset)
{
    MC_API_BEGIN
    arg_slot& s = slot_at(cb, index);
    MC_REQUIRE(buf, "setBuffer: null buffer");
    MC_REQUIRE(offset <= buf->size, "setBuffer: offset beyond the buffer");
    s.kind = arg_slot::buffer;
    s.buf = buf;
    s.offset = offset;
    MC_API_END
}

mc_status mc_barrier(mc_cmdbuf* cb, mc_buffer* buf)
{
    MC_API_BEGIN
    MC_REQUIRE(cb && buf, "bad arguments");
    // The reference needs memoryBarrier(resource) because its encoder is concurrent
    // (src/kernel_thread.cc:91-96).  Kernels on one CUDA stream already execute in order.
    MC_API_END
}

mc_status mc_dispatch(mc_cmdbuf* cb, mc_kernel* k, const uint32_t grid[3], const uint32_t group[3])
{
    MC_API_BEGIN
    MC_REQUIRE(cb && k && grid && group, "bad arguments");
    if (cb->committed) throw error(MC_ERR_RUNTIME, "command buffer already committed");
    if (cb->size >= cb->capacity) throw error(MC_ERR_FULL, "command buffer is full");
    // kernel_task validation (kernel.h:126-140)
    const uint64_t threads = uint64_t(group[0]) * group[1] * group[2];
    if (threads > 1024) {
        throw error(MC_ERR_INVALID, "kernel: <" + std::to_string(group[0]) + "," + std::to_string(group[1]) + "," + std::to_string(group[2]) + "> exceeds maximum number of threads per group 1024");
    }
    if (threads == 0) throw error(MC_ERR_INVALID, "kernel: empty thread group");
    for (int d = 0; d < 3; d++) {
        if (grid[d] < group[d]) throw error(MC_ERR_INVALID, "kernel: there are less threads in grid than in group");
    }
    use(cb->dev);
    cb->args.dev = cb->dev;
    k->entry->launch(cb->args, cb->dev->stream);
    MC_CUDA_CHECK(cudaGetLastError());
    cb->dev->launches.fetch_add(1);
    cb->size++;
    for (auto& s : cb->args.slot) s = arg_slot();
    MC_API_END
}

mc_status mc_on_completed(mc_cmdbuf* cb, void (*fn)(void*, int), void* user)
{
    MC_API_BEGIN
    MC_REQUIRE(cb && fn, "bad arguments");
    cb->handlers.emplace_back(fn, user);
    MC_API_END
}

struct host_call {
    void (*fn)(void*, int);
    void* user;
    mc_device* dev;
};
// cudaLaunchHostFunc callbacks do not run once the stream is in an error state, so the handlers of a failed command buffer
// are invoked from mc_wait instead, with the failure status (src/kernel_thread.cc:134-144 passes the buffer's error)
static void CUDART_CB run_host_call(void* p)
{
    auto* c = static_cast<host_call*>(p);
    int bad = 0;
    if (c->dev->bad_ids_host) bad = *reinterpret_cast<volatile int*>(c->dev->bad_ids_host);
    c->fn(c->user, bad ? int(MC_ERR_INVALID) : 0);
    delete c;
}

mc_status mc_commit(mc_cmdbuf* cb)
{
    MC_API_BEGIN
    MC_REQUIRE(cb, "null command buffer");
    if (cb->committed) throw error(MC_ERR_RUNTIME, "command buffer already committed");
    use(cb->dev);
    cb->committed = true;
    // completion handlers run on a runtime-owned thread, like Metal's (src/kernel_thread.cc:134-144)
    for (auto& h : cb->handlers) {
        auto* c = new host_call{h.first, h.second, cb->dev};
        MC_CUDA_CHECK(cudaLaunchHostFunc(cb->dev->stream, run_host_call, c));
    }
    MC_CUDA_CHECK(cudaEventRecord(cb->done, cb->dev->stream));
    MC_API_END
}

mc_status mc_wait(mc_cmdbuf* cb, char* err, size_t cap)
{
    if (err && cap) err[0] = 0;
    if (!cb) return fail(MC_ERR_INVALID, "null command buffer");
    try {
        use(cb->dev);
        if (!cb->committed) {
            mc_status s = mc_commit(cb);
            if (s != MC_OK) return s;
        }
        MC_CUDA_CHECK(cudaEventSynchronize(cb->done));
        MC_CUDA_CHECK(cudaGetLastError());
        // out-of-range ids seen by an embedding kernel of this device (the Metal kernel would read out of bounds): surfaced here,
        // like a command-buffer error at waitUntilCompleted
        if (cb->dev->bad_ids_host && *reinterpret_cast<volatile int*>(cb->dev->bad_ids_host)) {
            *reinterpret_cast<volatile int*>(cb->dev->bad_ids_host) = 0;
            throw error(MC_ERR_INVALID, "embedding: token id out of range");
        }
        return MC_OK;
    } catch (const error& e) {
        if (err && cap) snprintf(err, cap, "%s", e.what());
        return fail(e.code, e.what());
    }
}

mc_status mc_cmdbuf_size(mc_cmdbuf* cb, size_t* n)
{
    MC_API_BEGIN
    MC_REQUIRE(cb && n, "bad arguments");
    *n = cb->size;
    MC_API_END
}

mc_status mc_cmdbuf_release(mc_cmdbuf* cb)
{
    MC_API_BEGIN
    if (cb) {
        if (cb->done) cudaEventDestroy(cb->done);
        delete cb;
    }
    MC_API_END
}

mc_status mc_launch_count(mc_device* dev, uint64_t* launches)
{
    MC_API_BEGIN
    MC_REQUIRE(dev && launches, "bad arguments");
    *launches = dev->launches.load();
    MC_API_END
}

} // extern "C"
