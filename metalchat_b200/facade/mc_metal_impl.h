// metalchat_b200/facade/mc_metal_impl.h — what the opaque metal::{buffer,device,kernel,library} handles of the reference
// (include/metalchat/metal.h:14-34) are on B200: thin owners of the C-ABI handles of include/mc_cuda.h.
//
// The reference defines these structs in src/metal_impl.h:19-112 over metal-cpp objects; every public header only sees
// std::shared_ptr to the forward declarations, which is what makes the backend replaceable.  This directory replaces the
// reference's five Metal-bound translation units (src/{metal,accelerator,allocator,kernel,kernel_thread}.cc) and nothing
// else: the reference's headers and its portable sources (src/{layer,container,tensor}.cc) compile unmodified on top.
#pragma once

#include <filesystem>
#include <functional>
#include <list>
#include <memory>
#include <stdexcept>
#include <string>

#include <metalchat/metal.h>

#include "../../include/mc_cuda.h"

namespace metalchat {
namespace metal {

/// Turns an mc_status into the exception type the reference throws in the same situation.
void
check(mc_status status, const char* what = nullptr);


struct buffer {
    using deleter_type = std::function<void(buffer* p)>;

    mc_buffer* handle;   // owned (released by buffer_deleter); a heap slice shares the arena's handle
    bool owns_handle;
    void* host;          // host-dereferenceable address of byte 0 of this buffer (MTL::Buffer::contents())
    std::size_t offset;  // byte offset of this buffer inside `handle` (heap slices)
    std::size_t bytes;

    buffer(mc_buffer* h, bool owns, void* host_ptr, std::size_t off, std::size_t n)
    : handle(h),
      owns_handle(owns),
      host(host_ptr),
      offset(off),
      bytes(n)
    {}
};


/// Chain of callbacks run before the handle is released (residency / heap bookkeeping), as in src/metal_impl.h:30-58.
struct buffer_deleter {
    std::list<buffer::deleter_type> deleters;

    buffer_deleter() = default;

    buffer_deleter(buffer::deleter_type deleter)
    : deleters({std::move(deleter)})
    {}

    void
    invoke_before_destroy(buffer::deleter_type&& deleter)
    {
        deleters.push_back(std::move(deleter));
    }

    void
    operator()(buffer* b);
};


shared_buffer
make_buffer(mc_buffer* handle);

shared_buffer
make_buffer(mc_buffer* handle, buffer::deleter_type deleter);

/// A slice [offset, offset + size) of an arena buffer; `deleter` gives the bytes back to the arena.
shared_buffer
make_slice(const shared_buffer& arena, std::size_t offset, std::size_t size, buffer::deleter_type deleter);


struct device {
    mc_device* handle;

    explicit device(mc_device* h)
    : handle(h)
    {}

    device(const device&) = delete;

    ~device();
};

shared_device
make_device();


struct kernel {
    mc_kernel* handle;
    std::string name;

    kernel(mc_kernel* h, std::string n)
    : handle(h),
      name(std::move(n))
    {}

    kernel(const kernel&) = delete;

    ~kernel();
};


/// The kernels are compiled into the backend library; a "shader library" only remembers where it was asked to come from.
struct library {
    std::filesystem::path path;
};

shared_library
make_library(const std::filesystem::path& p, shared_device device);


} // namespace metal
} // namespace metalchat
