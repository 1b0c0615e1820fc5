// Descriptors shared by the host planner (nb_plan.hpp) and the device code (nb_kernel.cuh) of the N-body kernel.
#ifndef HEYOKA_B200_CSRC_NB_DESC_HPP
#define HEYOKA_B200_CSRC_NB_DESC_HPP

#include <cstdint>

namespace heyoka_b200::detail
{

// One pair interaction (64 bytes, read once per kernel by the thread that owns the pair).
struct nb_pair_desc {
    std::uint16_t pa[3], pb[3]; // position slots: d_k = pos[pa[k]] - pos[pb[k]]
    std::uint16_t om[3];        // output slots of m_k
    std::uint16_t fkind;        // f = q (0), c1 q (1), -q (2)
    std::uint32_t pad0;
    double c1;
    std::uint32_t u_d[3], u_r2, u_q, u_m[3]; // u variable indices (diagnostics / tests)
};
static_assert(sizeof(nb_pair_desc) == 64u);

// One item of a summation level (64 bytes).
//   kind 0: intermediate sum -> output slot `out`
//   kind 1: acceleration of the "velocity" state variable sv1 = out & 0xffff; (out >> 16) = 1 + its "position"
//           child (0 = none); pos = 1 + position slot of the child (0 = not read by any pair)
//   kind 2: idem with a constant right-hand side consts[cidx] (no terms)
// terms[i] = output slot | (1 + index of the multiplier in consts, 0 = none) << 16.
struct nb_sum_desc {
    std::uint32_t n_terms, kind;
    std::uint32_t terms[8];
    std::uint32_t out, pos;
    std::uint32_t cidx, pad[3];
};
static_assert(sizeof(nb_sum_desc) == 64u);

} // namespace heyoka_b200::detail

#endif
