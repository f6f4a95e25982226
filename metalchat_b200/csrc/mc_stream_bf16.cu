// metalchat_b200/csrc/mc_stream_bf16.cu — instantiations of the streaming persistent decode kernel (mc_stream_kernel.cuh) for
// bf16 models: head_dim 64 / 128 x one or kStSplits CTAs per attention head.
#include "mc_stream_kernel.cuh"

namespace mc {

stream_kernel_fn stream_kernel_bf16(uint32_t head_dim, bool single)
{
    if (single) return head_dim == 64 ? decode_stream_kernel<false, 64, false, 1> : decode_stream_kernel<false, 128, false, 1>;
    return head_dim == 64 ? decode_stream_kernel<false, 64, false, kStSplits> : decode_stream_kernel<false, 128, false, kStSplits>;
}

} // namespace mc
