// metalchat_b200/facade/mc_accelerator.cc — replaces src/accelerator.cc of the reference: hardware_accelerator over one
// mc_device (one GPU + one in-order stream) and the by-name kernel registry of libmc_cuda.so.
#include <format>

#include <metalchat/accelerator.h>
#include <metalchat/kernel.h>

#include "mc_metal_impl.h"


namespace metalchat {


// accelerator.h:29-43: the reference hashes kernel names with rapidhash; any good 64-bit byte hash does (FNV-1a here)
std::size_t
_StringHash::operator()(const void* s, std::size_t len) const noexcept
{
    auto bytes = static_cast<const unsigned char*>(s);
    std::uint64_t h = 0xcbf29ce484222325ull;
    for (std::size_t i = 0; i < len; i++) {
        h = (h ^ bytes[i]) * 0x100000001b3ull;
    }
    return std::size_t(h);
}


// accelerator.h:84-92.  `path` is where the reference finds metalchat.metallib; the CUDA kernels are part of the backend
// library, so the file only has to exist (src/accelerator.cc:26-34 + test/test_accelerator.cc:15-21 keep their behaviour).
hardware_accelerator::hardware_accelerator(
    const std::filesystem::path& path, std::size_t thread_capacity
)
: _M_device(metal::make_device()),
  _M_library(metal::make_library(path, _M_device)),
  _M_kernels(),
  _M_thread(std::make_shared<recursive_kernel_thread>(_M_device, thread_capacity))
{}


// accelerator.h:101: the reference looks the shader library up in its framework bundle (src/accelerator.cc:37-71);
// here there is nothing to look up.
hardware_accelerator::hardware_accelerator(std::size_t thread_capacity)
: _M_device(metal::make_device()),
  _M_library(std::make_shared<metal::library>()),
  _M_kernels(),
  _M_thread(std::make_shared<recursive_kernel_thread>(_M_device, thread_capacity))
{}


std::size_t
hardware_accelerator::max_buffer_size() const
{
    std::size_t bytes = 0;
    metal::check(mc_device_max_buffer(_M_device->handle, &bytes));
    return bytes;
}


std::shared_ptr<kernel_thread>
hardware_accelerator::get_this_thread()
{
    return _M_thread->get_this_thread();
}


metal::shared_device
hardware_accelerator::get_metal_device()
{
    return _M_device;
}


hardware_accelerator::allocator_type
hardware_accelerator::get_allocator() const
{
    return _M_thread->get_allocator();
}


void
hardware_accelerator::set_allocator(hardware_accelerator::allocator_type alloc)
{
    _M_thread->set_allocator(alloc);
}


std::string
hardware_accelerator::name() const
{
    char buf[256] = {0};
    metal::check(mc_device_name(_M_device->handle, buf, sizeof(buf)));
    return std::string(buf);
}


// src/accelerator.cc:117-158: look the function up by its mangled host name ("softmax_bfloat", "cumsum_2_float", ...),
// cache the kernel object.  There is no pipeline to build: mc_kernel_lookup answers from the registry of compiled kernels.
const basic_kernel&
hardware_accelerator::load(const std::string& name)
{
    if (auto it = _M_kernels.find(name); it != _M_kernels.end()) {
        return it->second;
    }

    mc_kernel* handle = nullptr;
    if (mc_kernel_lookup(_M_device->handle, name.c_str(), &handle) != MC_OK) {
        throw std::invalid_argument(
            std::format("hardware_accelerator: function {} not found in a shader library", name)
        );
    }

    auto kernel_ptr = std::make_shared<metal::kernel>(handle, name);
    _M_kernels.insert_or_assign(name, basic_kernel(kernel_ptr, *this));
    return _M_kernels.at(name);
}


const basic_kernel&
hardware_accelerator::load(const std::string& name, const std::string& type)
{
    return load(std::format("{}_{}", name, type));
}


} // namespace metalchat
