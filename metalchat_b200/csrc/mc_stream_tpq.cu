// metalchat_b200/csrc/mc_stream_tpq.cu — instantiations of the streaming persistent decode kernel (mc_stream_kernel.cuh) for
// tensor-parallel shards of quantised (QLoRA layout) models: head_dim 64 / 128 x one or kStSplits CTAs per attention head.
#include "mc_stream_kernel.cuh"

namespace mc {

stream_kernel_fn stream_kernel_tp_quant(uint32_t head_dim, bool single)
{
    if (single) return head_dim == 64 ? decode_stream_kernel<true, 64, true, 1> : decode_stream_kernel<true, 128, true, 1>;
    return head_dim == 64 ? decode_stream_kernel<true, 64, true, kStSplits> : decode_stream_kernel<true, 128, true, kStSplits>;
}

} // namespace mc
