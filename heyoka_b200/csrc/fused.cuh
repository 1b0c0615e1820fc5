// Superinstructions of the cooperative kernels: several elementary recurrences run back to back by ONE work
// item, found by pattern matching in make_smem_plan(). Every u variable keeps its own tape row and is computed
// by exactly the same arithmetic, in the same order, as by diff_op(): the fusion removes interpreter overhead
// (one dispatch and no synchronisation for a dozen ops), it does not change a single rounding.
#ifndef HEYOKA_B200_CSRC_FUSED_CUH
#define HEYOKA_B200_CSRC_FUSED_CUH

#include <cstdint>

#include "recurrences.cuh"

namespace heyoka_b200::dev
{

constexpr std::uint32_t FOP_FIRST = 0x100u;

// The gravitational pair interaction of model::nbody (src/model/nbody.cpp:97-153) at order n:
//   d_k = x_k^j - x_k^i (src/detail/sub.cpp)            r2 = sum_sq(d_0, d_1, d_2) (src/detail/sum_sq.cpp)
//   q = pow(r2, alpha) (src/math/pow.cpp)               f = c1 q | -q | q (src/math/prod.cpp)
//   m_k = d_k f (src/math/prod.cpp, var * var)          n_k = c2_k m_k (optional)
// aux: [a_k, b_k, d_k] x 3, r2, q, alpha (constant index), order-0 pow algorithm, f, c1 (constant index),
//      [m_k, operand order, n_k, c2_k (constant index)] x 3   (all rows as packed row references).
template <int N, typename Tape>
__device__ __forceinline__ void fused_nbody_pair(const program &P, const Tape &t, const std::uint32_t *aux,
                                                 std::uint32_t fkind, bool have_n, std::uint32_t n)
{
    using V = vd<N>;
    using Row = typename Tape::row_t;
    constexpr int S = static_cast<int>(Row::stride);

    // ---- d_k^[n] = b_k^[n] ... as SUB_VV: row(a).at(n) - row(b).at(n) ----
    Row D[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const V v = t.row(aux[3 * k]).at(n) - t.row(aux[3 * k + 1]).at(n);
        D[k] = t.row(aux[3 * k + 2]);
        D[k].set(n, v);
    }

    // ---- r2^[n]: SUM_SQ over the three differences ----
    const Row R2 = t.row(aux[9]);
    {
        V v[3];
        const bool odd = (n & 1u) != 0u;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (odd) {
                v[k] = conv_plain<N>(D[k], D[k], n, 0u, (n - 1u) / 2u);
            } else {
                const V ak2 = D[k].at(n / 2u);
                const V sq = ak2 * ak2;
                if (n > 0u) {
                    const V acc = conv_plain<N>(D[k], D[k], n, 0u, (n - 2u) / 2u);
                    v[k] = (acc + acc) + sq;
                } else {
                    v[k] = sq;
                }
            }
        }
        const V r = (v[0] + v[1]) + v[2]; // pairwise_sum of three terms
        R2.set(n, odd ? r + r : r);
    }

    // ---- q^[n] = pow(r2, alpha) ----
    const Row Q = t.row(aux[10]);
    V q;
    {
        const V alpha = splat<N>(t.cst(aux[11]));
        if (n == 0u) {
            q = pow_eval(aux[12], R2.at(0u), alpha);
        } else {
            const double nd = static_cast<double>(n);
            const V ap1 = alpha + 1.;
            const V n_alpha = nd * alpha;
            V acc = splat<N>(0.);
            const double *pb = R2.hptr(n), *pa = Q.hptr(0u);
#pragma unroll 4
            for (std::uint32_t j = 0; j < n; ++j) {
                const V fac = n_alpha - static_cast<double>(j) * ap1;
                acc = vfma(fac, Row::load(pb) * Row::load(pa), acc);
                pb -= S;
                pa += S;
            }
            q = acc / (nd * R2.at(0u));
        }
        Q.set(n, q);
    }

    // ---- f^[n] ----
    Row F = Q;
    if (fkind != 0u) {
        F = t.row(aux[13]);
        F.set(n, fkind == 1u ? t.cst(aux[14]) * q : -q);
    }

    // ---- m_k^[n] = sum_j A^[n-j] B^[j] with (A, B) = (d_k, f) or (f, d_k); the f loads are shared ----
    const bool f_first = aux[16] != 0u;
    V acc[3] = {splat<N>(0.), splat<N>(0.), splat<N>(0.)};
    if (!f_first) {
        const double *pf = F.hptr(0u);
        const double *pd0 = D[0].hptr(n), *pd1 = D[1].hptr(n), *pd2 = D[2].hptr(n);
#pragma unroll 4
        for (std::uint32_t j = 0; j <= n; ++j) {
            const V fj = Row::load(pf);
            acc[0] = vfma(Row::load(pd0), fj, acc[0]);
            acc[1] = vfma(Row::load(pd1), fj, acc[1]);
            acc[2] = vfma(Row::load(pd2), fj, acc[2]);
            pf += S;
            pd0 -= S;
            pd1 -= S;
            pd2 -= S;
        }
    } else {
        const double *pf = F.hptr(n);
        const double *pd0 = D[0].hptr(0u), *pd1 = D[1].hptr(0u), *pd2 = D[2].hptr(0u);
#pragma unroll 4
        for (std::uint32_t j = 0; j <= n; ++j) {
            const V fj = Row::load(pf);
            acc[0] = vfma(fj, Row::load(pd0), acc[0]);
            acc[1] = vfma(fj, Row::load(pd1), acc[1]);
            acc[2] = vfma(fj, Row::load(pd2), acc[2]);
            pf -= S;
            pd0 += S;
            pd1 += S;
            pd2 += S;
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        t.row(aux[15 + 4 * k]).set(n, acc[k]);
        if (have_n) {
            t.row(aux[17 + 4 * k]).set(n, t.cst(aux[18 + 4 * k]) * acc[k]);
        }
    }
}

} // namespace heyoka_b200::dev

#endif
