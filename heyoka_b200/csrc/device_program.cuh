// Device-side view of a lowered program + the batch arrays (plain structs passed by value to kernels).
#ifndef HEYOKA_B200_CSRC_DEVICE_PROGRAM_CUH
#define HEYOKA_B200_CSRC_DEVICE_PROGRAM_CUH

#include <cstdint>

#include <heyoka_b200.h>

namespace heyoka_b200::dev
{

struct program {
    std::uint32_t n_eq, n_uvars, n_pars, order, n_ops;
    int high_accuracy;
    double rhofac;      // exp(-7/10 / (p - 1)) / e^2, src/taylor_00.cpp:84-94
    double inv_p;       // 1 / p
    double inv_pm1;     // 1 / (p - 1)
    const uint4 *ops;   // hy_op reinterpreted as uint4 {opcode, a, b, c}
    const std::uint32_t *args;
    const double *consts;
    const std::uint32_t *sv_defs;
};

// Resident arrays of a batch, all batch-innermost (include/heyoka_b200.h).
struct batch {
    std::uint32_t n; // number of lanes
    double *state, *t_hi, *t_lo, *last_h, *tc;
    const double *pars;
    long long *step_outcome;
    // propagate results
    long long *prop_outcome;
    double *prop_min_h, *prop_max_h;
    unsigned long long *prop_n_steps;
    unsigned long long *prop_iters; // iterations of the lock-step loop each lane was active in (last_h semantics)
};

// Global flags written by the propagate kernel.
struct run_flags {
    unsigned int any_nf;               // some lane produced a non-finite state / time
    unsigned int any_limit;            // some lane hit the iteration limit
    unsigned long long min_nf_iter;    // smallest 1-based iteration index at which a lane went non-finite
    unsigned long long max_iter;       // largest number of iterations any lane took (the reference's loop length)
};

} // namespace heyoka_b200::dev

#endif
