// heyoka_b200 — Taylor decomposition of an ODE system (host side).
//
// Restates the reference's decomposition pipeline (bluescarni/heyoka @ 9c91f71):
//   src/taylor_01.cpp:847-1008   taylor_decompose_sys()
//   src/taylor_01.cpp:315-443    taylor_decompose_cse()
//   src/taylor_01.cpp:454-645    taylor_sort_dc()   (Kahn, breadth first)
//   src/taylor_01.cpp:788-803    taylor_decompose_replace_numbers()
//   src/taylor_01.cpp:806-840    pow_to_explog()
//   src/expression_decompose.cpp:45-209, src/func.cpp:392-420
//   src/expression_basic.cpp:1177-1233, include/heyoka/detail/udf_split.hpp:50-98
//   src/math/sum.cpp:387-544 (sum_to_sum_sq, sum_to_sub), src/math/prod.cpp:753-908 (prod_to_div)
//   src/math/sin.cpp:115-133, src/math/cos.cpp:116-134, src/math/tanh.cpp:77-92 (hidden deps)
#ifndef HEYOKA_B200_TAYLOR_DECOMPOSE_HPP
#define HEYOKA_B200_TAYLOR_DECOMPOSE_HPP

#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include <heyoka_b200/expression.hpp>

namespace heyoka_b200
{

// [0, n_eq): state variables; [n_eq, size - n_eq): elementary u variables (function of earlier
// u variables / numbers / params) with their hidden dependencies; [size - n_eq, size): the
// definitions of the state variables' first derivatives (a u variable, a number or a param).
using taylor_dc_t = std::vector<std::pair<expression, std::vector<std::uint32_t>>>;

std::pair<taylor_dc_t, std::vector<std::uint32_t>>
taylor_decompose_sys(const std::vector<std::pair<expression, expression>> &sys,
                     const std::vector<expression> &sv_funcs = {});

// "u_12" -> 12 (src/detail/string_conv.hpp uname_to_index()).
std::uint32_t uname_to_index(const std::string &);

// Validation of an ODE system (src/detail/validate_ode_sys.cpp): unique variable LHS, RHS only in
// terms of the LHS variables. Throws std::invalid_argument.
void validate_ode_sys(const std::vector<std::pair<expression, expression>> &sys,
                      const std::vector<expression> &ev_funcs = {});

std::string dc_to_string(const taylor_dc_t &);

} // namespace heyoka_b200

#endif
