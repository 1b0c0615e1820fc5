// The two instantiations of the dense-network kernel (see nn_variants.hpp).
#include "nn_variants.hpp"
#include "nn_kernel.cuh"

namespace heyoka_b200::detail
{

nn_fn nn_kernel_step()
{
    return dev::k_nn<false>;
}

nn_fn nn_kernel_prop()
{
    return dev::k_nn<true>;
}

} // namespace heyoka_b200::detail
