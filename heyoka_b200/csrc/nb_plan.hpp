// Host-side planning for the dedicated N-body kernel (nb_kernel.cuh).
//
// A program qualifies when it is made of nothing but the decomposition of model::nbody-like right-hand sides
// (src/model/nbody.cpp:97-153):
//   * gravitational pair interactions  d_k = x_k^a - x_k^b (k = 0..2), r2 = sum_sq(d_0, d_1, d_2), q = pow(r2, alpha),
//     f = c1 q | -q | q, m_k = d_k f, optionally n_k = c2 m_k (src/detail/sub.cpp, src/detail/sum_sq.cpp,
//     src/math/pow.cpp, src/math/prod.cpp);
//   * sums (<= 8 terms, possibly nested, src/math/sum.cpp) of the m_k / n_k;
//   * state variables in second-order form: "velocities" whose derivative is one of the above (or a number) and
//     "positions" whose derivative is a velocity; the pair interactions read positions only.
// For such a system x^[n+2] only depends on the accelerations of order <= n, so the accelerations of the orders
// n and n + 1 can be computed TOGETHER from x^[<= n+1]: the kernel walks the orders two at a time, which halves
// the synchronisation points and lets every operand loaded in a convolution feed two accumulators. Every u
// variable still runs its own recurrence with the reference's summation order (results are bit-identical to the
// one-order-at-a-time evaluation of src/taylor_02.cpp:1147-1185).
#ifndef HEYOKA_B200_CSRC_NB_PLAN_HPP
#define HEYOKA_B200_CSRC_NB_PLAN_HPP

#include <cstdint>
#include <string>
#include <vector>

#include "nb_desc.hpp"
#include "program.hpp"

namespace heyoka_b200::detail
{

struct nb_plan {
    bool ok = false;
    std::string why; // why the program does not qualify
    std::uint32_t n_pos = 0, n_out = 0;
    double alpha = 0.;
    std::uint32_t pow_algo = 0;
    std::vector<nb_pair_desc> pairs;
    std::vector<nb_sum_desc> sums;           // level by level
    std::vector<std::uint32_t> level_offsets; // n_levels + 1 offsets into sums
    std::vector<double> consts;              // multipliers and constant right-hand sides
    std::vector<std::uint32_t> pos_sv;       // state variable held by each position slot
    std::vector<std::uint32_t> pair_uvars;   // per pair: u variable indices of d_0..2, r2, q, m_0..2 (diagnostics / tests)
    // Table fac[n][j] = n alpha - j (alpha + 1) of the pow recurrence (src/math/pow.cpp:618-963),
    // (order + 1) rows of fac_stride doubles.
    std::vector<double> fac;
    std::uint32_t fac_stride = 0;
};

nb_plan make_nb_plan(const hy_program &);

// The per-thread role table of the summation phase for a team of `tt` threads owning `lt` lanes, `nl` lanes per thread:
// rounds x tt records, level after level (every level's items are dealt out to the threads round-robin; a level ends
// with a synchronisation). round_level_end: bit r is set if round r is the last one of its level.
struct nb_roles {
    std::vector<nb_role> table;
    std::uint32_t n_rounds = 0, round_level_end = 0;
};
nb_roles make_nb_roles(const nb_plan &, std::uint32_t tt, std::uint32_t lt, std::uint32_t nl);

} // namespace heyoka_b200::detail

#endif
