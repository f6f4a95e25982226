// metalchat_b200/facade/mc_metal.cc — replaces src/metal.cc of the reference: metal::data / metal::size and the handle
// factories, over the C ABI of include/mc_cuda.h.
#include <format>

#include <metalchat/allocator.h>

#include "mc_metal_impl.h"


namespace metalchat {
namespace metal {


void
check(mc_status status, const char* what)
{
    if (status == MC_OK) {
        return;
    }
    std::string message = mc_last_error();
    if (what != nullptr) {
        message = std::string(what) + ": " + message;
    }
    switch (status) {
    case MC_ERR_INVALID:
    case MC_ERR_NOT_FOUND:
        throw std::invalid_argument(message);
    case MC_ERR_ALLOC:
        throw alloc_error(message);
    default:
        throw std::runtime_error(message);
    }
}


// metal.h:17-22.  The reference returns MTL::Buffer::contents(): a pointer the HOST may dereference (unified memory).
// B200 has no unified memory; façade buffers are CUDA managed allocations (MC_MEM_SHARED), valid on the host once the
// command buffer that wrote them has completed -- which is when the reference's own code reads them (future_tensor::get).
void*
data(const shared_buffer& buffer)
{
    return buffer->host;
}


std::size_t
size(const shared_buffer& buffer)
{
    return buffer->bytes;
}


void
buffer_deleter::operator()(buffer* b)
{
    for (auto& deleter : deleters) {
        deleter(b);
    }
    if (b->owns_handle && b->handle != nullptr) {
        mc_buffer_release(b->handle);
    }
    b->handle = nullptr;
    delete b;
}


static buffer*
wrap(mc_buffer* handle)
{
    void* host = nullptr;
    std::size_t bytes = 0;
    check(mc_buffer_host_ptr(handle, &host), "metal: buffer is not host visible");
    check(mc_buffer_size(handle, &bytes));
    return new buffer(handle, /*owns=*/true, host, 0, bytes);
}


shared_buffer
make_buffer(mc_buffer* handle)
{
    return shared_buffer(wrap(handle), buffer_deleter());
}


shared_buffer
make_buffer(mc_buffer* handle, buffer::deleter_type deleter)
{
    return shared_buffer(wrap(handle), buffer_deleter(std::move(deleter)));
}


shared_buffer
make_slice(const shared_buffer& arena, std::size_t offset, std::size_t size, buffer::deleter_type deleter)
{
    auto host = static_cast<std::uint8_t*>(arena->host) + offset;
    auto slice = new buffer(arena->handle, /*owns=*/false, host, arena->offset + offset, size);
    return shared_buffer(slice, buffer_deleter(std::move(deleter)));
}


device::~device()
{
    if (handle != nullptr) {
        mc_device_destroy(handle);
    }
}


shared_device
make_device()
{
    // MTL::CreateSystemDefaultDevice (src/metal.cc:51-55): the first device of the process
    mc_device* handle = nullptr;
    check(mc_device_create(0, &handle));
    return std::make_shared<device>(handle);
}


kernel::~kernel()
{
    if (handle != nullptr) {
        mc_kernel_release(handle);
    }
}


shared_library
make_library(const std::filesystem::path& p, shared_device)
{
    // newLibrary(url) fails for a file that is not there, and the reference's users (and test/test_accelerator.cc:15-21)
    // rely on that error; the kernels themselves live in the backend library, so any existing file is accepted.
    std::error_code ec;
    if (!std::filesystem::exists(p, ec)) {
        throw std::runtime_error("metal: library not found");
    }
    return std::make_shared<library>(library{p});
}


} // namespace metal
} // namespace metalchat
