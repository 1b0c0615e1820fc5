// Entry points of the dense-network kernel dev::k_nn<PROP> (nn_kernel.cuh), instantiated in nn_inst.cu.
#ifndef HEYOKA_B200_CSRC_NN_VARIANTS_HPP
#define HEYOKA_B200_CSRC_NN_VARIANTS_HPP

#include "device_program.cuh"

namespace heyoka_b200::dev
{
struct run_args;    // kernels.cuh
struct nn_dev_plan; // nn_kernel.cuh
} // namespace heyoka_b200::dev

namespace heyoka_b200::detail
{

using nn_fn = void (*)(dev::program, dev::nn_dev_plan, dev::batch, dev::run_args);

nn_fn nn_kernel_step();
nn_fn nn_kernel_prop();

} // namespace heyoka_b200::detail

#endif
