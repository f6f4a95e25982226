// metalchat_b200/facade/mc_allocator.cc — replaces src/allocator.cc of the reference: the four device-bound allocator
// cores behind hardware_memory_allocator / hardware_heap_allocator / nocopy_allocator / hardware_resident_allocator
// (allocator.h:268-626).  The adapters around them (pooling, paginated, aliasing, rebind, polymorphic) are header-only
// reference code and stay untouched.
//
// Storage mode: every reference buffer is MTL::ResourceStorageModeShared -- the host reads and writes tensor memory
// directly (src/metal.cc:21-25, tensor/basic.h full(), functional/primitive.h triu()).  B200 has no unified memory, so a
// façade buffer is a CUDA managed allocation (MC_MEM_SHARED): same pointer on both sides, pages migrate on demand.  The
// fused engine (mc_llama_*) does not go through these allocators; it keeps weights and caches in plain device memory.
#include <algorithm>
#include <format>
#include <map>
#include <memory>
#include <mutex>

#include <metalchat/allocator.h>

#include "mc_metal_impl.h"


namespace metalchat {


// ---- hardware_heap_allocator: one fixed arena, first-fit sub-allocation (src/allocator.cc:18-112) ----------------------
struct _HardwareHeapAllocator::_Memory {
    metal::shared_buffer arena;
    std::map<std::size_t, std::size_t> free_blocks; // offset -> size, coalesced
    std::mutex mu;
    std::size_t size = 0; // live allocations

    static constexpr std::size_t alignment = 256;

    std::size_t
    largest_free() const
    {
        std::size_t best = 0;
        for (const auto& [offset, bytes] : free_blocks) {
            best = std::max(best, bytes);
        }
        return best;
    }

    void
    give_back(std::size_t offset, std::size_t bytes)
    {
        auto [it, inserted] = free_blocks.emplace(offset, bytes);
        auto next = std::next(it);
        if (next != free_blocks.end() && it->first + it->second == next->first) {
            it->second += next->second;
            free_blocks.erase(next);
        }
        if (it != free_blocks.begin()) {
            auto prev = std::prev(it);
            if (prev->first + prev->second == it->first) {
                prev->second += it->second;
                free_blocks.erase(it);
            }
        }
    }
};


struct _HardwareHeapAllocator::_Deleter {
    std::shared_ptr<_Memory> memory;
    std::size_t offset;
    std::size_t bytes;

    void
    operator()(metal::buffer*)
    {
        auto& mem = *memory;
        const std::scoped_lock lock(mem.mu);
        mem.give_back(offset, bytes);
        if (mem.size > 0) {
            mem.size--;
        }
    }
};


_HardwareHeapAllocator::_HardwareHeapAllocator(metal::shared_device device, std::size_t capacity)
: _M_mem(std::make_shared<_Memory>())
{
    mc_buffer* handle = nullptr;
    if (mc_alloc(device->handle, capacity, MC_MEM_SHARED, &handle) != MC_OK) {
        throw std::runtime_error("hardware_heap_allocator: failed creating a new heap");
    }
    _M_mem->arena = metal::make_buffer(handle);
    _M_mem->free_blocks.emplace(0, capacity);
}


_HardwareHeapAllocator::container_pointer
_HardwareHeapAllocator::allocate(std::size_t size)
{
    auto& mem = *_M_mem;
    const std::scoped_lock lock(mem.mu);

    const auto mask = _Memory::alignment - 1;
    const auto alloc_size = (std::max<std::size_t>(size, 1) + mask) & ~mask;

    for (auto it = mem.free_blocks.begin(); it != mem.free_blocks.end(); ++it) {
        if (it->second < alloc_size) {
            continue;
        }
        const auto offset = it->first;
        const auto rest = it->second - alloc_size;
        mem.free_blocks.erase(it);
        if (rest > 0) {
            mem.free_blocks.emplace(offset + alloc_size, rest);
        }
        mem.size++;
        auto buffer_ptr = metal::make_slice(mem.arena, offset, size, _Deleter{_M_mem, offset, alloc_size});
        return std::make_shared<container_type>(buffer_ptr);
    }

    throw alloc_error(std::format(
        "hardware_heap_allocator: failed to allocate buffer of size={}, "
        "heap remaining capacity={}",
        size, mem.largest_free()
    ));
}


// ---- hardware_memory_allocator: one allocation per buffer (src/allocator.cc:115-145) -------------------------------------
struct _HardwareMemoryAllocator::_Memory {
    metal::shared_device device;
};


_HardwareMemoryAllocator::_HardwareMemoryAllocator(metal::shared_device device)
: _M_mem(std::make_shared<_Memory>())
{
    _M_mem->device = device;
}


_HardwareMemoryAllocator::container_pointer
_HardwareMemoryAllocator::allocate(std::size_t size)
{
    mc_buffer* handle = nullptr;
    metal::check(mc_alloc(_M_mem->device->handle, size, MC_MEM_SHARED, &handle));
    return std::make_shared<container_type>(metal::make_buffer(handle));
}


_HardwareMemoryAllocator::container_pointer
_HardwareMemoryAllocator::allocate(const void* ptr, std::size_t size)
{
    mc_buffer* handle = nullptr;
    metal::check(mc_alloc_copy(_M_mem->device->handle, ptr, size, MC_MEM_SHARED, &handle));
    return std::make_shared<container_type>(metal::make_buffer(handle));
}


// ---- nocopy_allocator: wrap memory the caller owns (src/allocator.cc:148-173) ----------------------------------------------
// Metal maps the pages into the GPU's address space; CUDA pins and maps them (cudaHostRegister): the GPU reads them over
// PCIe.  Good for activations and staging, not for weights that are streamed every token -- the safetensors loader of the
// engine copies those into device memory instead (metalchat_b200/safetensors.py).
struct _HardwareNocopyAllocator::_Memory {
    metal::shared_device device;
};


_HardwareNocopyAllocator::_HardwareNocopyAllocator(metal::shared_device device)
: _M_mem(std::make_shared<_Memory>())
{
    _M_mem->device = device;
}


_HardwareNocopyAllocator::container_pointer
_HardwareNocopyAllocator::allocate(const void* ptr, std::size_t size)
{
    mc_buffer* handle = nullptr;
    if (mc_wrap_host(_M_mem->device->handle, const_cast<void*>(ptr), size, &handle) != MC_OK) {
        throw alloc_error(std::format(
            "hardware_nocopy_allocator: failed to allocate no-copy buffer of size {}", size
        ));
    }
    return std::make_shared<container_type>(metal::make_buffer(handle));
}


// ---- hardware_resident_allocator: keep a set of buffers resident (src/allocator.cc:176-275) --------------------------------
// A Metal residency set wires its allocations; CUDA allocations are always resident, so only the set's bookkeeping is kept:
// the capacity limit, the "no allocation after commit" rule and the release of the set with its last buffer.
struct _HardwareResidentAllocator::_Memory {
    bool committed = false;
    std::mutex mu;
    std::size_t size = 0;
    std::size_t capacity = 0;
};


struct _HardwareResidentAllocator::_Deleter {
    std::shared_ptr<_Memory> memory;

    void
    operator()(metal::buffer*)
    {
        auto& mem = *memory;
        const std::scoped_lock lock(mem.mu);
        if (mem.size > 0) {
            mem.size--;
        }
    }
};


_HardwareResidentAllocator::_HardwareResidentAllocator(metal::shared_device, std::size_t capacity)
: _M_mem(std::make_shared<_Memory>())
{
    _M_mem->capacity = capacity;
}


_HardwareResidentAllocator::~_HardwareResidentAllocator() { detach(); }


void
_HardwareResidentAllocator::detach()
{
    auto& mem = *_M_mem;
    const std::scoped_lock lock(mem.mu);

    if (mem.size > 0 && !mem.committed) {
        mem.committed = true;
        mem.capacity = mem.size; // a committed set takes no more allocations
    }
}


_HardwareResidentAllocator::container_pointer
_HardwareResidentAllocator::allocate(_HardwareResidentAllocator::container_pointer&& container)
{
    auto& mem = *_M_mem;
    const std::scoped_lock lock(mem.mu);

    if (mem.size >= mem.capacity) {
        throw alloc_error("hardware_resident_allocator: capacity exceeded");
    }

    auto buffer_ptr = container->storage();
    mem.size++;

    if (auto deleter_ptr = std::get_deleter<metal::buffer_deleter>(buffer_ptr)) {
        deleter_ptr->invoke_before_destroy(_Deleter{_M_mem});
    }
    return std::make_shared<container_type>(buffer_ptr);
}


} // namespace metalchat
