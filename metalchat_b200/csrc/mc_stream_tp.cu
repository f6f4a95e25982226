// metalchat_b200/csrc/mc_stream_tp.cu — instantiations of the streaming persistent decode kernel (mc_stream_kernel.cuh) for
// tensor-parallel shards of bf16 models (all-reduce fused into the wo / w2 epilogues): head_dim 64 / 128 x one or kStSplits CTAs per attention head.
#include "mc_stream_kernel.cuh"

namespace mc {

stream_kernel_fn stream_kernel_tp(uint32_t head_dim, bool single)
{
    if (single) return head_dim == 64 ? decode_stream_kernel<false, 64, true, 1> : decode_stream_kernel<false, 128, true, 1>;
    return head_dim == 64 ? decode_stream_kernel<false, 64, true, kStSplits> : decode_stream_kernel<false, 128, true, kStSplits>;
}

} // namespace mc
