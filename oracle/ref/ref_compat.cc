// oracle/ref/ref_compat.cc — TEST INFRASTRUCTURE ONLY.
//
// nn::basic_linear<T, Container>::operator()(input_type) is declared virtual but NOT pure and is defined nowhere in the reference
// (nn/linear.h:27-28; no definition under src/).  Apple clang never emits a reference to it (the base-class vtable stores of the
// inlined constructors are elided), GNU ld sees the vtable and wants the symbol.  The interface function is never called --
// every concrete layer overrides it -- so it is defined here to say so.
#include <stdexcept>

#include <metalchat/dtype.h>
#include <metalchat/nn/linear.h>

namespace metalchat {
namespace nn {

template <typename T, contiguous_container Container>
typename basic_linear<T, Container>::result_type
basic_linear<T, Container>::operator()(input_type)
{
    throw std::logic_error("nn::basic_linear: the interface has no implementation of operator()");
}

template class basic_linear<float, hardware_memory_container<float>>;
template class basic_linear<bf16, hardware_memory_container<bf16>>;
template class basic_linear<float, random_memory_container<float>>;
template class basic_linear<bf16, random_memory_container<bf16>>;

} // namespace nn
} // namespace metalchat
