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

// ---- propagate_grid() (src/taylor_adaptive_batch.cpp:1545-2055): per-lane bookkeeping and dense-output
// sampling between the lock-step steps. One thread per lane. ----
struct grid_state {
    const double *grid; // [n_pts][batch]
    std::uint32_t n_pts;
    double *out;               // [n_pts][n_eq][batch], NaN-filled
    const double *max_delta_t; // positive per-lane limits or nullptr (+inf)
    std::uint32_t *cur_idx;    // first grid point not yet written, per lane
    double *rem_hi, *rem_lo;   // remaining time to the last grid point (double-length)
    unsigned char *t_dir;      // 1: forward
    double *dt_limit;          // signed limit of the next step
    unsigned *flags;           // [0] grid points left in some lane, [1] non-finite state, [2] overflow of rem
};

// After the initial propagate_until(grid[0]): remaining times, directions, counters (:1728-1760).
__global__ void k_grid_init(batch D, grid_state G, double *min_h, double *max_h, unsigned long long *ts_count)
{
    const std::uint32_t lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= D.n) {
        return;
    }
    const dfl t{D.t_hi[lane], D.t_lo[lane]};
    const dfl rem = dfl_sub(dfl{G.grid[static_cast<std::size_t>(G.n_pts - 1u) * D.n + lane], 0.}, t);
    if (!(isfinite(rem.hi) && isfinite(rem.lo))) {
        atomicOr(G.flags + 2, 1u);
    }
    G.rem_hi[lane] = rem.hi;
    G.rem_lo[lane] = rem.lo;
    G.t_dir[lane] = dfl_ge0(rem) ? 1 : 0;
    G.cur_idx[lane] = 1u;
    min_h[lane] = CUDART_INF;
    max_h[lane] = 0.;
    ts_count[lane] = 0ull;
}

// After a step: counters, min/max |h|, remaining time, outcome (:1915-1971).
__global__ void k_grid_book(batch D, grid_state G, long long *outcome, double *min_h, double *max_h,
                            unsigned long long *ts_count)
{
    const std::uint32_t lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= D.n) {
        return;
    }
    const long long oc = D.step_outcome[lane];
    const double h = D.last_h[lane];
    if (oc == HY_OUTCOME_ERR_NF_STATE) {
        atomicOr(G.flags + 1, 1u);
    } else {
        ts_count[lane] += (h != 0.) ? 1ull : 0ull;
        if (oc == HY_OUTCOME_SUCCESS) {
            const double ah = fabs(h);
            min_h[lane] = fmin(min_h[lane], ah);
            max_h[lane] = fmax(max_h[lane], ah);
        }
        if (h == G.rem_hi[lane]) {
            G.rem_hi[lane] = 0.;
            G.rem_lo[lane] = 0.;
        } else {
            const dfl rem = dfl_sub(dfl{G.grid[static_cast<std::size_t>(G.n_pts - 1u) * D.n + lane], 0.},
                                    dfl{D.t_hi[lane], D.t_lo[lane]});
            G.rem_hi[lane] = rem.hi;
            G.rem_lo[lane] = rem.lo;
        }
    }
    outcome[lane] = oc;
}

// Dense output at every grid point covered by the last step (:1811-1886), then the limit of the next step
// (:1899-1912). Does nothing if a non-finite state was detected in this iteration.
__global__ void k_grid_sample(program P, batch D, grid_state G)
{
    const std::uint32_t lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= D.n || G.flags[1] != 0u) {
        return;
    }
    const std::size_t n = D.n;
    const dfl t{D.t_hi[lane], D.t_lo[lane]};
    const dfl cmp = dfl_sub(t, dfl{D.last_h[lane], 0.}); // start of the last step
    const dfl t0 = dfl_lt(cmp, t) ? cmp : t, t1 = dfl_lt(t, cmp) ? cmp : t;
    const dfl rem{G.rem_hi[lane], G.rem_lo[lane]};
    const bool rem0 = rem.hi == 0. && rem.lo == 0.;
    std::uint32_t idx = G.cur_idx[lane];
    while (idx < G.n_pts) {
        const dfl g{G.grid[static_cast<std::size_t>(idx) * n + lane], 0.};
        const bool avail = (!dfl_lt(g, t0) && !dfl_lt(t1, g)) || rem0;
        if (!avail) {
            break;
        }
        const double tau = dfl_sub(g, cmp).hi;
        for (std::uint32_t i = 0; i < P.n_eq; ++i) {
            const double *c = D.tc + static_cast<std::size_t>(i) * (P.order + 1u) * n + lane;
            G.out[(static_cast<std::size_t>(idx) * P.n_eq + i) * n + lane]
                = eval_poly(P, [c, n](std::uint32_t o) { return c[static_cast<std::size_t>(o) * n]; }, tau);
        }
        ++idx;
    }
    G.cur_idx[lane] = idx;
    if (idx < G.n_pts) {
        atomicOr(G.flags, 1u);
    }
    const double mdt = G.max_delta_t != nullptr ? G.max_delta_t[lane] : CUDART_INF;
    G.dt_limit[lane] = step_limit(G.t_dir[lane] != 0, rem, mdt);
}

__global__ void k_fill_double(double *out, std::size_t n, double value)
{
    const std::size_t i = static_cast<std::size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) {
        out[i] = value;
    }
}

} // namespace heyoka_b200::dev

#endif
