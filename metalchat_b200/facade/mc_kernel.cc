// metalchat_b200/facade/mc_kernel.cc — replaces src/kernel.cc of the reference: basic_kernel and the 2-D grid heuristic.
#include <metalchat/kernel.h>

#include "mc_metal_impl.h"


namespace metalchat {


// kernel.h:43-44 (src/kernel.cc:14-38).  The wrappers ask for a Metal dispatchThreads shape; the CUDA backend validates it
// (kernel.h:126-140) and then picks its own launch, so only the contract matters: group <= max_threads, grid >= group per
// dimension, grid covers [dim_size, num_rows].  Small tensors become one group, rows of up to max_threads elements one
// group per row, longer rows whole groups of max_threads.
std::tuple<dim3, dim3>
make_kernel_grid_2d(std::size_t num_rows, std::size_t dim_size, std::size_t max_threads)
{
    if (dim_size * num_rows <= max_threads) {
        return {dim3(dim_size, num_rows), dim3(dim_size, num_rows)};
    }
    if (dim_size <= max_threads) {
        return {dim3(dim_size, num_rows), dim3(dim_size)};
    }
    const auto groups = ceil_div(dim_size, max_threads);
    return {dim3(max_threads * groups, num_rows), dim3(max_threads)};
}


basic_kernel::basic_kernel(metal::shared_kernel kernel, const hardware_accelerator& accelerator)
: _M_name(kernel->name),
  _M_kernel(kernel),
  _M_accelerator(accelerator)
{}


std::string
basic_kernel::name() const
{
    return _M_name;
}


hardware_accelerator&
basic_kernel::get_accelerator()
{
    return _M_accelerator;
}


hardware_accelerator::allocator_type
basic_kernel::get_allocator() const
{
    return _M_accelerator.get_allocator();
}


const metal::shared_kernel
basic_kernel::get_metal_kernel() const
{
    return _M_kernel;
}


std::size_t
basic_kernel::max_threads_per_threadgroup()
{
    // maxTotalThreadsPerThreadgroup of the pipeline (src/kernel.cc:75-79): 1024, which also fixes the reduction partition of
    // rmsnorm / softmax / sum / sort / cumsum (block_size = ceil(D / 1024))
    std::size_t n = 0;
    metal::check(mc_kernel_max_threads(_M_kernel->handle, &n));
    return n;
}


} // namespace metalchat
