// Host-side program object: the lowered Taylor decomposition (see include/heyoka_b200.h, section B).
#ifndef HEYOKA_B200_CSRC_PROGRAM_HPP
#define HEYOKA_B200_CSRC_PROGRAM_HPP

#include <cstdint>
#include <string>
#include <vector>

#include <heyoka_b200.h>
#include <heyoka_b200/taylor_decompose.hpp>

struct hy_program {
    std::uint32_t n_eq = 0, n_uvars = 0, n_pars = 0, order = 0;
    bool high_accuracy = false;
    std::vector<hy_op> ops;
    std::vector<std::uint32_t> args;
    std::vector<double> consts;
    std::vector<std::uint32_t> sv_defs;
    // Event equations: the u variable holding each of them (terminal events first), sv_funcs_dc of
    // src/taylor_01.cpp:847-1008.
    std::vector<std::uint32_t> ev_defs;
    // Reference-shaped decomposition, kept for diagnostics (empty if built from raw arrays).
    heyoka_b200::taylor_dc_t dc;

    hy_program_desc desc() const;
};

namespace heyoka_b200::detail
{

// Lower a decomposition to the opcode program. Throws std::invalid_argument /
// not_implemented (as std::runtime_error tagged "not implemented") on unsupported input.
hy_program lower_decomposition(const taylor_dc_t &dc, std::uint32_t n_eq, std::uint32_t n_pars, std::uint32_t order,
                               bool high_accuracy);

// Structural validation of a raw program (indices in range, acyclic, known opcodes).
void validate_program(const hy_program &);

std::uint32_t taylor_order_from_tol(double tol);

struct program_costs {
    double b_min, b_tape, flops;
};
program_costs compute_costs(const hy_program &);

struct not_implemented_error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

} // namespace heyoka_b200::detail

#endif
