// Small non-template kernels (included by batch.cu only).
#ifndef HEYOKA_B200_CSRC_SMALL_KERNELS_CUH
#define HEYOKA_B200_CSRC_SMALL_KERNELS_CUH

#include "kernels.cuh"

namespace heyoka_b200::dev
{

__global__ void k_fill_outcome(long long *out, std::uint32_t n, long long value)
{
    const std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        out[i] = value;
    }
}

// Dense output (src/taylor_01.cpp:1015-1185): Horner, or compensated summation in high-accuracy mode.
__global__ void k_d_output(program P, std::uint32_t n, const double *tc, const double *tau, double *out)
{
    const std::uint32_t lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= n) {
        return;
    }
    const double h = tau[lane];
    const std::size_t nn = n;
    for (std::uint32_t i = 0; i < P.n_eq; ++i) {
        const double *c = tc + static_cast<std::size_t>(i) * (P.order + 1u) * n + lane;
        out[static_cast<std::size_t>(i) * n + lane]
            = eval_poly(P, [c, nn](std::uint32_t o) { return c[static_cast<std::size_t>(o) * nn]; }, h);
    }
}

} // namespace heyoka_b200::dev

#endif
