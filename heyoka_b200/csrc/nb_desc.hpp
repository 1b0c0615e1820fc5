// Descriptors shared by the host planner (nb_plan.hpp) and the device code (nb_kernel.cuh) of the N-body kernel.
#ifndef HEYOKA_B200_CSRC_NB_DESC_HPP
#define HEYOKA_B200_CSRC_NB_DESC_HPP

#include <cstdint>

namespace heyoka_b200::detail
{

// Copies of the three step-size norms of a lane that the threads of the summation phase spread their shared-memory
// atomics over (64-bit maxima on ONE address per lane and norm serialise: 5 % of the 6-body kernel's stall samples); the
// owner thread takes the maximum of the copies. lt = lanes per team.
#if defined(__CUDACC__)
__host__ __device__
#endif
constexpr std::uint32_t nb_norm_copies(std::uint32_t lt)
{
    return lt <= 2u ? 8u : (lt <= 8u ? 2u : 1u);
}

// One pair interaction (64 bytes, read once per kernel by the thread that owns the pair).
// d_k = pos[pa[k]] - pos[pb[k]], r2 = sum_sq(d), q = pow(r2, alpha), f = c1 q, m_k = d_k f -> output slot om[k],
// and, if flags bit 0 is set, n_k = c2[k] m_k -> output slot on[k] (0xffff: that n_k does not exist).
struct nb_pair_desc {
    std::uint16_t pa[3], pb[3];
    std::uint16_t om[3], on[3];
    std::uint32_t flags; // bit 0: some n_k exists
    std::uint32_t pad0;
    double c1; // 1 / -1 when f = q / -q
    double c2[3];
};
static_assert(sizeof(nb_pair_desc) == 64u);

// One item of a summation level (64 bytes; host emulation and planner tests; the device runs nb_role records).
//   kind 0: intermediate sum -> output slot `out`
//   kind 1: acceleration of the "velocity" state variable sv1 = out & 0xffff; (out >> 16) = 1 + its "position"
//           child (0 = none); pos = 1 + position slot of the child (0 = not read by any pair)
//   kind 2: idem with a constant right-hand side consts[cidx] (no terms)
// terms[i] = output slot.
struct nb_sum_desc {
    std::uint32_t n_terms, kind;
    std::uint32_t terms[8];
    std::uint32_t out, pos;
    std::uint32_t cidx, pad[3];
};
static_assert(sizeof(nb_sum_desc) == 64u);

// What ONE thread does in ONE round of the summation phase, pre-decoded by the host for a given team shape (LT
// lanes per team, NL lanes per thread): 32 bytes, one 2 x 16-byte read per round and order pair.
//   head:  bits 0-3 number of terms, bits 4-5 kind + 1 (0 = nothing to do), bit 6 has a position child,
//          bit 7 the child has a position slot
//   t[i]:  16-byte units relative to the team's output array: (slot * LT + first lane of the thread)
//   dst:   kind 0: output slot, else position slot of the child, same units (0xffff: none)
//   sv:    sv1 | sv2 << 16
//   cidx:  constant right-hand side (kind 2)
struct nb_role {
    std::uint32_t head;
    std::uint16_t t[8];
    std::uint32_t dst;
    std::uint32_t sv;
    std::uint32_t cidx;
};
static_assert(sizeof(nb_role) == 32u);

} // namespace heyoka_b200::detail

#endif
