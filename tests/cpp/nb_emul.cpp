// TEST INFRASTRUCTURE: host emulation of the N-body kernel's jet (heyoka_b200/csrc/nb_core.hpp) on plain arrays.
//
// nb_core.hpp is the arithmetic of k_nb (nb_kernel.cuh) written once for host and device; the device only supplies
// the storage policy (shared memory + tensor memory). This file supplies a storage policy on std::vector, runs the
// phases in the kernel's order (all pair threads, then the summation levels, one "thread" after the other: the
// threads of a team only communicate across synchronisation points) and returns every coefficient, so that the
// two-orders-at-a-time index arithmetic can be checked against the oracle WITHOUT a GPU (tests/test_nb_plan.py).
// It is never linked into the product library. Build: see tests/test_nb_plan.py (g++ -ffp-contract=off).
#include <cstdint>
#include <cstring>
#include <vector>

#include <heyoka_b200.h>

#include "nb_core.hpp"
#include "nb_plan.hpp"

namespace hb = heyoka_b200;
using hb::nb::d2;

namespace
{

struct emul {
    const hb::detail::nb_plan &pl;
    std::uint32_t p, npp, n_eq;
    std::vector<d2> pos, out;
    std::vector<d2> rows; // [pair][5 rows: d0 d1 d2 r2 q][npp]
    std::vector<double> coef; // [sv][p + 1]
    std::vector<double> mk;   // [pair][3][2 * npp]: the products' coefficients (diagnostics)
    const double *state;
};

struct pair_mem {
    emul &E;
    std::uint32_t pi;
    const hb::detail::nb_pair_desc &d;
    d2 *row(int r)
    {
        return E.rows.data() + (static_cast<std::size_t>(pi) * 5u + r) * E.npp;
    }
    d2 pos_a(int k)
    {
        return E.pos[d.pa[k]];
    }
    d2 pos_b(int k)
    {
        return E.pos[d.pb[k]];
    }
    void st_d(std::uint32_t m, const d2 (&D)[3])
    {
        for (int k = 0; k < 3; ++k) {
            row(k)[m] = D[k];
        }
    }
    void st_r2(std::uint32_t m, const d2 &r)
    {
        row(3)[m] = r;
    }
    void st_q(std::uint32_t m, const d2 &q)
    {
        row(4)[m] = q;
    }
    void ld_ss(std::uint32_t ai, std::uint32_t li, d2 (&A)[3], d2 (&Lo)[3])
    {
        for (int k = 0; k < 3; ++k) {
            A[k] = row(k)[ai];
            Lo[k] = row(k)[li];
        }
    }
    void ld_a(std::uint32_t ai, d2 (&A)[3])
    {
        for (int k = 0; k < 3; ++k) {
            A[k] = row(k)[ai];
        }
    }
    void ld_main(std::uint32_t qi, std::uint32_t li, d2 &Q, d2 &Rlo, d2 (&Dlo)[3])
    {
        Q = row(4)[qi];
        Rlo = row(3)[li];
        for (int k = 0; k < 3; ++k) {
            Dlo[k] = row(k)[li];
        }
    }
    d2 fac(std::uint32_t n, std::uint32_t j)
    {
        const double *f = E.pl.fac.data() + static_cast<std::size_t>(n) * E.pl.fac_stride + j;
        return d2{f[0], f[1]};
    }
    double fac1(std::uint32_t n, std::uint32_t j)
    {
        return E.pl.fac[static_cast<std::size_t>(n) * E.pl.fac_stride + j];
    }
    std::uint32_t cur_m = 0;
    void out(int k, const d2 &v)
    {
        E.out[d.om[k]] = v;
        double *dst = E.mk.data() + (static_cast<std::size_t>(pi) * 3u + k) * 2u * E.npp + 2u * cur_m;
        dst[0] = v.x;
        dst[1] = v.y;
    }
};

struct sum_mem {
    emul &E;
    d2 out_ld(std::uint32_t slot, int)
    {
        return E.out[slot];
    }
    void out_st(std::uint32_t slot, int, const d2 &v)
    {
        E.out[slot] = v;
    }
    void pos_st(std::uint32_t slot, int, const d2 &v)
    {
        E.pos[slot] = v;
    }
    double cst(std::uint32_t i)
    {
        return E.pl.consts[i];
    }
    double rcp(std::uint32_t n)
    {
        return 1. / static_cast<double>(n);
    }
    void coef(std::uint32_t sv, std::uint32_t order, int, double v)
    {
        E.coef[static_cast<std::size_t>(sv) * (E.p + 1u) + order] = v;
    }
    double state(std::uint32_t sv, int)
    {
        return E.state[sv];
    }
};

} // namespace

extern "C" {

// Plan summary: out[0] = ok, n_pairs, n_pos, n_out, n_sums, n_levels. Returns the length of `why`.
int nb_emul_plan(const hy_program *p, std::uint32_t *out, char *why, std::size_t why_len)
{
    const auto pl = hb::detail::make_nb_plan(*p);
    out[0] = pl.ok ? 1u : 0u;
    out[1] = static_cast<std::uint32_t>(pl.pairs.size());
    out[2] = pl.n_pos;
    out[3] = pl.n_out;
    out[4] = static_cast<std::uint32_t>(pl.sums.size());
    out[5] = pl.ok ? static_cast<std::uint32_t>(pl.level_offsets.size()) - 1u : 0u;
    if (why != nullptr && why_len != 0u) {
        std::strncpy(why, pl.why.c_str(), why_len - 1u);
        why[why_len - 1u] = '\0';
    }
    return static_cast<int>(pl.why.size());
}

// Jet of ONE lane. state[n_eq] -> coef[n_eq][order + 1] (state variables), and, per pair interaction i:
// u_idx[i * 8 + {0..2: d_k, 3: r2, 4: q, 5..7: m_k}] = u variable index, u_rows[(i * 8 + r) * n_ord + o] = coefficient
// of order o (n_ord = 2 * ceil(order / 2)). Returns 0, or -1 if the program does not qualify.
int nb_emul_jet(const hy_program *p, const double *state, double *coef, std::uint32_t *u_idx, double *u_rows)
{
    const auto pl = hb::detail::make_nb_plan(*p);
    if (!pl.ok) {
        return -1;
    }
    emul E{pl, p->order, (p->order + 1u) / 2u, p->n_eq, {}, {}, {}, {}, {}, state};
    const std::uint32_t n_pairs = static_cast<std::uint32_t>(pl.pairs.size());
    E.pos.assign(pl.n_pos, d2{0., 0.});
    E.out.assign(pl.n_out, d2{0., 0.});
    E.rows.assign(static_cast<std::size_t>(n_pairs) * 5u * E.npp, d2{0., 0.});
    E.coef.assign(static_cast<std::size_t>(p->n_eq) * (p->order + 1u), 0.);
    E.mk.assign(static_cast<std::size_t>(n_pairs) * 3u * 2u * E.npp, 0.);
    sum_mem SM{E};
    const auto words = [&](std::size_t i) { return reinterpret_cast<const std::uint32_t *>(&pl.sums[i]); };
    for (std::size_t i = 0; i < pl.sums.size(); ++i) {
        hb::nb::sum_init<1>(SM, words(i));
    }
    const hb::nb::pair_consts base{0., pl.alpha, pl.pow_algo};
    for (std::uint32_t m = 0; m < E.npp; ++m) {
        for (std::uint32_t pi = 0; pi < n_pairs; ++pi) {
            pair_mem PM{E, pi, pl.pairs[pi]};
            PM.cur_m = m;
            auto C = base;
            C.c1 = pl.pairs[pi].c1;
            hb::nb::pair_block(PM, C, m);
        }
        for (std::size_t lv = 0; lv + 1u < pl.level_offsets.size(); ++lv) {
            for (std::size_t i = pl.level_offsets[lv]; i < pl.level_offsets[lv + 1u]; ++i) {
                hb::nb::sum_block<1>(SM, words(i), m, p->order);
            }
        }
    }
    std::memcpy(coef, E.coef.data(), E.coef.size() * sizeof(double));
    const std::uint32_t n_ord = 2u * E.npp;
    for (std::uint32_t pi = 0; pi < n_pairs; ++pi) {
        const auto &d = pl.pairs[pi];
        const std::uint32_t us[8] = {d.u_d[0], d.u_d[1], d.u_d[2], d.u_r2, d.u_q, d.u_m[0], d.u_m[1], d.u_m[2]};
        for (std::uint32_t r = 0; r < 8u; ++r) {
            u_idx[pi * 8u + r] = us[r];
            double *dst = u_rows + (static_cast<std::size_t>(pi) * 8u + r) * n_ord;
            if (r < 5u) {
                const d2 *src = E.rows.data() + (static_cast<std::size_t>(pi) * 5u + r) * E.npp;
                for (std::uint32_t m = 0; m < E.npp; ++m) {
                    dst[2u * m] = src[m].x;
                    dst[2u * m + 1u] = src[m].y;
                }
            } else {
                std::memcpy(dst, E.mk.data() + (static_cast<std::size_t>(pi) * 3u + (r - 5u)) * n_ord,
                            n_ord * sizeof(double));
            }
        }
    }
    return 0;
}

} // extern "C"
