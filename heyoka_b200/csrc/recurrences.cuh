// Device recurrences: one normalised Taylor derivative ("coefficient") of one u variable at one order,
// for the lane owned by the calling thread. Each opcode is the hand-written counterpart of one
// taylor_c_diff_func_* of the reference (compact-mode, i.e. running-accumulator, summation order):
//
//   sum / sub                       src/math/sum.cpp:250-371, src/detail/sub.cpp:180-398
//   k*v, -v, v*v                    src/math/prod.cpp:443-705
//   div                             src/detail/div.cpp:189-431
//   square / sqrt / pow             src/math/pow.cpp:618-963; order-0 evaluation :292-355, :136-152
//   sum_sq                          src/detail/sum_sq.cpp:250-468
//   sin / cos / tanh / exp / log    src/math/sin.cpp:241-372, cos.cpp:241-372, tanh.cpp:183-318,
//                                   exp.cpp:150-285, log.cpp:164-305
//   time / constant-only functions  src/math/time.cpp:82-104, include/heyoka/detail/taylor_common.hpp:88-157
//
// Floating-point contract: this file is compiled with -fmad=false, so the ONLY fused operations are
// the explicit fma() calls below: every `acc + a*b` of the reference's accumulation loops is
// fma(a, b, acc) (a contraction LLVM is allowed to make in the reference, src/llvm_state.cpp:842-845);
// everything else rounds after each operation. tests/ compare against the oracle's sequential+FMA mode.
//
// The Tape policy gives access to the lane's private column of the derivative tape:
//   double ld(slot), void st(slot, v), with slot = u * (order + 1) + o.
#ifndef HEYOKA_B200_CSRC_RECURRENCES_CUH
#define HEYOKA_B200_CSRC_RECURRENCES_CUH

#include <cstdint>

#include "device_program.cuh"

namespace heyoka_b200::dev
{

struct lane_ctx {
    std::uint32_t lane;  // global lane index (clamped to a valid lane)
    std::uint32_t batch; // number of lanes = stride of the batch-innermost arrays
    const double *pars;
    double time;         // t_hi of the lane at the beginning of the step
};

__device__ __forceinline__ double load_par(const lane_ctx &c, std::uint32_t idx)
{
    return __ldg(c.pars + static_cast<std::size_t>(idx) * c.batch + c.lane);
}

__device__ __forceinline__ double numpar_val(const program &P, const lane_ctx &c, std::uint32_t ref)
{
    return HY_REF_KIND(ref) == HY_REF_NUM ? __ldg(P.consts + HY_REF_IDX(ref)) : load_par(c, HY_REF_IDX(ref));
}

// pairwise_reduce() of up to 8 values (src/detail/llvm_helpers_algo.cpp:271-308), registers only.
__device__ __forceinline__ double pairwise8(double (&v)[8], std::uint32_t n)
{
    double w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        w[i] = (2 * i + 1 < static_cast<int>(n)) ? v[2 * i] + v[2 * i + 1] : v[2 * i];
    }
    const std::uint32_t m = (n + 1u) / 2u;
    double x[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        x[i] = (2 * i + 1 < static_cast<int>(m)) ? w[2 * i] + w[2 * i + 1] : w[2 * i];
    }
    const std::uint32_t m2 = (m + 1u) / 2u;
    return (m2 > 1u) ? x[0] + x[1] : x[0];
}

// sum_{j=j0..j1} A^[n-j] B^[j]; sa_n = slot of A at order n, sb_0 = slot of B at order 0.
template <typename Tape>
__device__ __forceinline__ double conv_plain(const Tape &t, std::uint32_t sa_n, std::uint32_t sb_0, std::uint32_t j0,
                                             std::uint32_t j1)
{
    double acc = 0.;
    if (j1 + 1u > j0) {
#pragma unroll 4
        for (std::uint32_t j = j0; j <= j1; ++j) {
            acc = fma(t.ld(sa_n - j), t.ld(sb_0 + j), acc);
        }
    }
    return acc;
}

// sum_{j=j0..j1} j * (A^[n-j] B^[j]).
template <typename Tape>
__device__ __forceinline__ double conv_jw(const Tape &t, std::uint32_t sa_n, std::uint32_t sb_0, std::uint32_t j0,
                                          std::uint32_t j1)
{
    double acc = 0.;
    if (j1 + 1u > j0) {
#pragma unroll 4
        for (std::uint32_t j = j0; j <= j1; ++j) {
            acc = fma(static_cast<double>(j), t.ld(sa_n - j) * t.ld(sb_0 + j), acc);
        }
    }
    return acc;
}

// Exponentiation by squaring with the reference's association order (src/math/pow.cpp:136-152).
__device__ inline double pow_ebs(double base, std::uint32_t e)
{
    double mult[6];
    int nm = 0;
    double b = base;
    while (e > 1u) {
        if (e & 1u) {
            mult[nm++] = b;
            e = (e - 1u) / 2u;
        } else {
            e /= 2u;
        }
        b = b * b;
    }
    double r = (e == 0u) ? 1. : b;
    for (int i = nm - 1; i >= 0; --i) {
        r = mult[i] * r;
    }
    return r;
}

__device__ inline double pow_eval(std::uint32_t algo, double x, double expo)
{
    const std::uint32_t type = algo >> 8, n = algo & 0xffu;
    switch (type) {
        case HY_POW_POS_SMALL_INT:
            return pow_ebs(x, n);
        case HY_POW_NEG_SMALL_INT:
            return 1. / pow_ebs(x, n);
        case HY_POW_POS_SMALL_HALF:
            return pow_ebs(::sqrt(x), n);
        case HY_POW_NEG_SMALL_HALF:
            return 1. / pow_ebs(::sqrt(x), n);
        default:
            return ::pow(x, expo);
    }
}

__device__ inline std::uint32_t pow_algo_of(double e)
{
    if (isfinite(e) && e == trunc(e)) {
        if (e >= 0 && e <= 16) {
            return (HY_POW_POS_SMALL_INT << 8) | static_cast<std::uint32_t>(e);
        }
        if (e < 0 && -e <= 16) {
            return (HY_POW_NEG_SMALL_INT << 8) | static_cast<std::uint32_t>(-e);
        }
    } else if (isfinite(e)) {
        const double y = 2 * e;
        if (y == trunc(y)) {
            if (y >= 0 && y <= 16) {
                return (HY_POW_POS_SMALL_HALF << 8) | static_cast<std::uint32_t>(y);
            }
            if (y < 0 && -y <= 16) {
                return (HY_POW_NEG_SMALL_HALF << 8) | static_cast<std::uint32_t>(-y);
            }
        }
    }
    return HY_POW_GENERAL << 8;
}

// Functions whose arguments are all numbers/params: evaluated at order 0 only.
__device__ inline double cfunc_eval(const program &P, const lane_ctx &c, const uint4 &op)
{
    double v[8];
    const std::uint32_t n = op.z;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        v[k] = (k < static_cast<int>(n)) ? numpar_val(P, c, __ldg(P.args + op.y + k)) : 0.;
    }
    switch (op.x) {
        case HY_CF_IDENTITY:
            return v[0];
        case HY_CF_SUM:
            return pairwise8(v, n);
        case HY_CF_PROD:
            return v[0] * v[1];
        case HY_CF_SUB:
            return v[0] - v[1];
        case HY_CF_DIV:
            return v[0] / v[1];
        case HY_CF_POW: {
            const std::uint32_t eref = __ldg(P.args + op.y + 1u);
            const std::uint32_t algo = HY_REF_KIND(eref) == HY_REF_NUM ? pow_algo_of(v[1]) : (HY_POW_GENERAL << 8);
            return pow_eval(algo, v[0], v[1]);
        }
        case HY_CF_SUM_SQ:
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                v[k] = v[k] * v[k];
            }
            return pairwise8(v, n);
        case HY_CF_SIN:
            return ::sin(v[0]);
        case HY_CF_COS:
            return ::cos(v[0]);
        case HY_CF_TANH:
            return ::tanh(v[0]);
        case HY_CF_EXP:
            return ::exp(v[0]);
        case HY_CF_LOG:
            return ::log(v[0]);
    }
    return 0.;
}

// The order-n coefficient of u variable `u_idx` defined by `op`. pp1 = order + 1 (slot stride).
template <typename Tape>
__device__ __forceinline__ double diff_op(const program &P, const lane_ctx &c, const Tape &t, const uint4 &op,
                                          std::uint32_t u_idx, std::uint32_t n)
{
    const std::uint32_t pp1 = P.order + 1u;
    const std::uint32_t a = op.y, b = op.z, dep = op.w;

    switch (op.x) {
        case HY_OP_SUM: {
            // a^[n] = pairwise sum of the terms' order-n coefficients; numbers/params only at n = 0.
            double v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                v[k] = 0.;
                if (k < static_cast<int>(b)) {
                    const std::uint32_t ref = __ldg(P.args + a + k);
                    if (HY_REF_KIND(ref) == HY_REF_VAR) {
                        v[k] = t.ld(HY_REF_IDX(ref) * pp1 + n);
                    } else if (n == 0u) {
                        v[k] = numpar_val(P, c, ref);
                    }
                }
            }
            return pairwise8(v, b);
        }
        case HY_OP_SUM_SQ: {
            // Per term the square recurrence, then a pairwise sum over the terms.
            double v[8];
            const bool odd = (n & 1u) != 0u;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                v[k] = 0.;
                if (k < static_cast<int>(b)) {
                    const std::uint32_t ref = __ldg(P.args + a + k);
                    if (HY_REF_KIND(ref) == HY_REF_VAR) {
                        const std::uint32_t s0 = HY_REF_IDX(ref) * pp1;
                        if (odd) {
                            v[k] = conv_plain(t, s0 + n, s0, 0u, (n - 1u) / 2u);
                        } else {
                            const double ak2 = t.ld(s0 + n / 2u);
                            const double sq = ak2 * ak2;
                            if (n > 0u) {
                                const double acc = conv_plain(t, s0 + n, s0, 0u, (n - 2u) / 2u);
                                v[k] = (acc + acc) + sq;
                            } else {
                                v[k] = sq;
                            }
                        }
                    } else if (n == 0u) {
                        const double val = numpar_val(P, c, ref);
                        v[k] = val * val;
                    }
                }
            }
            const double r = pairwise8(v, b);
            return odd ? r + r : r;
        }
        case HY_OP_SUB_VV:
            return t.ld(a * pp1 + n) - t.ld(b * pp1 + n);
        case HY_OP_SUB_VN: {
            const double v = t.ld(a * pp1 + n);
            return n == 0u ? v - __ldg(P.consts + b) : v;
        }
        case HY_OP_SUB_VP: {
            const double v = t.ld(a * pp1 + n);
            return n == 0u ? v - load_par(c, b) : v;
        }
        case HY_OP_SUB_NV: {
            const double v = t.ld(b * pp1 + n);
            return n == 0u ? __ldg(P.consts + a) - v : -v;
        }
        case HY_OP_SUB_PV: {
            const double v = t.ld(b * pp1 + n);
            return n == 0u ? load_par(c, a) - v : -v;
        }
        case HY_OP_NEG:
            return -t.ld(a * pp1 + n);
        case HY_OP_MUL_NV:
            return __ldg(P.consts + a) * t.ld(b * pp1 + n);
        case HY_OP_MUL_PV:
            return load_par(c, a) * t.ld(b * pp1 + n);
        case HY_OP_MUL_VV:
            // sum_{j=0..n} b^[n-j] c^[j]
            return conv_plain(t, a * pp1 + n, b * pp1, 0u, n);
        case HY_OP_DIV_VV:
        case HY_OP_DIV_NV:
        case HY_OP_DIV_PV: {
            // (b^[n] - sum_{j=1..n} a^[n-j] c^[j]) / c^[0], a = this u variable; numerator = -sum if b is constant.
            const double c0 = t.ld(b * pp1);
            if (n == 0u) {
                const double num = op.x == HY_OP_DIV_VV ? t.ld(a * pp1)
                                                        : (op.x == HY_OP_DIV_NV ? __ldg(P.consts + a) : load_par(c, a));
                return num / c0;
            }
            const double acc = conv_plain(t, u_idx * pp1 + n, b * pp1, 1u, n);
            if (op.x == HY_OP_DIV_VV) {
                return (t.ld(a * pp1 + n) - acc) / c0;
            }
            return (-acc) / c0;
        }
        case HY_OP_DIV_VN:
            return t.ld(a * pp1 + n) / __ldg(P.consts + b);
        case HY_OP_DIV_VP:
            return t.ld(a * pp1 + n) / load_par(c, b);
        case HY_OP_SQUARE: {
            const std::uint32_t s0 = a * pp1;
            if (n == 0u) {
                const double b0 = t.ld(s0);
                return b0 * b0;
            }
            if (n & 1u) {
                const double r = conv_plain(t, s0 + n, s0, 0u, (n - 1u) / 2u);
                return r + r;
            }
            const double ak2 = t.ld(s0 + n / 2u);
            const double sq = ak2 * ak2;
            const double r = conv_plain(t, s0 + n, s0, 0u, (n - 2u) / 2u);
            return (r + r) + sq;
        }
        case HY_OP_SQRT: {
            // (b^[n] - 2 sum_{j=1..} a^[n-j] a^[j] - [n even] (a^[n/2])^2) / (2 a^[0]), a = this u variable.
            if (n == 0u) {
                return ::sqrt(t.ld(a * pp1));
            }
            const std::uint32_t s0 = u_idx * pp1;
            double div = t.ld(s0);
            div = div + div;
            double fac = t.ld(a * pp1 + n);
            const bool even = (n & 1u) == 0u;
            const std::uint32_t upper = (n - (even ? 2u : 1u)) / 2u;
            double acc = conv_plain(t, s0 + n, s0, 1u, upper);
            acc = acc + acc;
            if (even) {
                const double tmp = t.ld(s0 + n / 2u);
                fac = fac - tmp * tmp;
            }
            fac = fac - acc;
            return fac / div;
        }
        case HY_OP_POW_VN:
        case HY_OP_POW_VP: {
            // (1 / (n b0)) sum_{j=0..n-1} [n alpha - j (alpha + 1)] b^[n-j] a^[j], a = this u variable.
            const double alpha = op.x == HY_OP_POW_VN ? __ldg(P.consts + b) : load_par(c, b);
            const std::uint32_t sb = a * pp1;
            if (n == 0u) {
                return pow_eval(op.x == HY_OP_POW_VN ? dep : (HY_POW_GENERAL << 8), t.ld(sb), alpha);
            }
            const std::uint32_t sa0 = u_idx * pp1;
            const double nd = static_cast<double>(n), ap1 = alpha + 1.;
            const double n_alpha = nd * alpha;
            double acc = 0.;
#pragma unroll 4
            for (std::uint32_t j = 0; j < n; ++j) {
                const double fac = n_alpha - static_cast<double>(j) * ap1;
                acc = fma(fac, t.ld(sb + n - j) * t.ld(sa0 + j), acc);
            }
            return acc / (nd * t.ld(sb));
        }
        case HY_OP_SIN:
            // (1/n) sum_{j=1..n} j c^[n-j] b^[j], c = ::cos(b) (hidden dependency).
            if (n == 0u) {
                return ::sin(t.ld(a * pp1));
            }
            return conv_jw(t, dep * pp1 + n, a * pp1, 1u, n) / static_cast<double>(n);
        case HY_OP_COS:
            // sum / (-n), with s = ::sin(b) as hidden dependency.
            if (n == 0u) {
                return ::cos(t.ld(a * pp1));
            }
            return conv_jw(t, dep * pp1 + n, a * pp1, 1u, n) / (-static_cast<double>(n));
        case HY_OP_TANH:
            // b^[n] - (1/n) sum_{j=1..n} j c^[n-j] b^[j], c = ::tanh(b)^2 (hidden dependency).
            if (n == 0u) {
                return ::tanh(t.ld(a * pp1));
            }
            return t.ld(a * pp1 + n) - conv_jw(t, dep * pp1 + n, a * pp1, 1u, n) / static_cast<double>(n);
        case HY_OP_EXP:
            // (1/n) sum_{j=1..n} j a^[n-j] b^[j], a = this u variable.
            if (n == 0u) {
                return ::exp(t.ld(a * pp1));
            }
            return conv_jw(t, u_idx * pp1 + n, a * pp1, 1u, n) / static_cast<double>(n);
        case HY_OP_LOG: {
            // (n b^[n] - sum_{j=1..n-1} j b^[n-j] a^[j]) / (n b^[0]), a = this u variable.
            if (n == 0u) {
                return ::log(t.ld(a * pp1));
            }
            const double nd = static_cast<double>(n);
            const double nb0 = nd * t.ld(a * pp1);
            double ret = nd * t.ld(a * pp1 + n);
            if (n > 1u) {
                ret = ret - conv_jw(t, a * pp1 + n, u_idx * pp1, 1u, n - 1u);
            }
            return ret / nb0;
        }
        case HY_OP_TIME:
            return n == 0u ? c.time : (n == 1u ? 1. : 0.);
        case HY_OP_CFUNC:
            return n == 0u ? cfunc_eval(P, c, op) : 0.;
    }
    return 0.;
}

// Order-n (n >= 1) coefficient of state variable sv: (u_rhs)^[n-1] / n, a true division
// (src/taylor_02.cpp:245-287); constant right-hand sides only contribute at n == 1.
template <typename Tape>
__device__ __forceinline__ double sv_diff(const program &P, const lane_ctx &c, const Tape &t, std::uint32_t sv,
                                          std::uint32_t n)
{
    const std::uint32_t ref = __ldg(P.sv_defs + sv);
    if (HY_REF_KIND(ref) == HY_REF_VAR) {
        return t.ld(HY_REF_IDX(ref) * (P.order + 1u) + n - 1u) / static_cast<double>(n);
    }
    return n == 1u ? numpar_val(P, c, ref) : 0.;
}

// The whole jet of the lane: orders 0..p-1 of every u variable, order p of the state variables
// (evaluation order of src/taylor_02.cpp:1147-1185: per order, state variables first, then the others).
template <typename Tape>
__device__ __forceinline__ void compute_jet(const program &P, const lane_ctx &c, const Tape &t, const double *state)
{
    const std::uint32_t pp1 = P.order + 1u;

    for (std::uint32_t i = 0; i < P.n_eq; ++i) {
        t.st(i * pp1, state[static_cast<std::size_t>(i) * c.batch + c.lane]);
    }
    for (std::uint32_t n = 0; n < P.order; ++n) {
        if (n > 0u) {
            for (std::uint32_t i = 0; i < P.n_eq; ++i) {
                t.st(i * pp1 + n, sv_diff(P, c, t, i, n));
            }
        }
        for (std::uint32_t k = 0; k < P.n_ops; ++k) {
            const uint4 op = __ldg(P.ops + k);
            t.st((P.n_eq + k) * pp1 + n, diff_op(P, c, t, op, P.n_eq + k, n));
        }
    }
    for (std::uint32_t i = 0; i < P.n_eq; ++i) {
        t.st(i * pp1 + P.order, sv_diff(P, c, t, i, P.order));
    }
}

// std::max / std::min semantics of the reference's llvm_max/llvm_min (src/detail/llvm_helpers_cmp.cpp:313-329).
__device__ __forceinline__ double std_max(double a, double b)
{
    return (a < b) ? b : a;
}
__device__ __forceinline__ double std_min(double a, double b)
{
    return (b < a) ? b : a;
}

// taylor_determine_h() (src/taylor_00.cpp:102-273): Jorba-Zou step size from the infinity norms of the
// state and of the two highest-order coefficients, clamped to |max_delta_t|, signed like max_delta_t.
template <typename Tape>
__device__ __forceinline__ double determine_h(const program &P, const Tape &t, double max_delta_t)
{
    const std::uint32_t pp1 = P.order + 1u, p = P.order;
    double m0 = fabs(t.ld(0)), mp = fabs(t.ld(p)), mp1 = fabs(t.ld(p - 1u));
    for (std::uint32_t i = 1; i < P.n_eq; ++i) {
        m0 = std_max(m0, fabs(t.ld(i * pp1)));
        mp = std_max(mp, fabs(t.ld(i * pp1 + p)));
        mp1 = std_max(mp1, fabs(t.ld(i * pp1 + p - 1u)));
    }
    const double num_rho = (m0 <= 1.) ? 1. : m0;
    const double rho_o = ::pow(num_rho / mp, P.inv_p);
    const double rho_om1 = ::pow(num_rho / mp1, P.inv_pm1);
    const double rho_m = std_min(rho_o, rho_om1);
    double h = rho_m * P.rhofac;
    h = std_min(h, fabs(max_delta_t));
    return (max_delta_t < 0.) ? -h : h;
}

// State update: Horner (src/taylor_00.cpp:279-351) or compensated summation of the monomials
// (src/taylor_00.cpp:355-460) when high_accuracy; optionally the tc copy (src/taylor_00.cpp:467-584).
template <typename Tape>
__device__ __forceinline__ void update_state(const program &P, const lane_ctx &c, const Tape &t, double h, double *state,
                                             double *tc, bool write)
{
    const std::uint32_t pp1 = P.order + 1u, p = P.order;
    for (std::uint32_t i = 0; i < P.n_eq; ++i) {
        const std::uint32_t s0 = i * pp1;
        double res;
        if (!P.high_accuracy) {
            res = t.ld(s0 + p);
            for (std::uint32_t o = 1; o <= p; ++o) {
                res = fma(res, h, t.ld(s0 + p - o));
            }
        } else {
            res = t.ld(s0);
            double comp = 0., cur_h = h;
            for (std::uint32_t o = 1; o <= p; ++o) {
                const double tmp = __dmul_rn(t.ld(s0 + o), cur_h);
                const double y = __dsub_rn(tmp, comp);
                const double tt = __dadd_rn(res, y);
                comp = __dsub_rn(__dsub_rn(tt, res), y);
                res = tt;
                cur_h = __dmul_rn(cur_h, h);
            }
        }
        if (write) {
            state[static_cast<std::size_t>(i) * c.batch + c.lane] = res;
            if (tc != nullptr) {
                for (std::uint32_t o = 0; o <= p; ++o) {
                    tc[(static_cast<std::size_t>(i) * pp1 + o) * c.batch + c.lane] = t.ld(s0 + o);
                }
            }
        }
    }
}

// Double-length time arithmetic (include/heyoka/detail/dfloat.hpp:104-169).
struct dfl {
    double hi, lo;
};

__device__ __forceinline__ dfl eft_knuth(double a, double b)
{
    const double x = __dadd_rn(a, b);
    const double z = __dsub_rn(x, a);
    const double y = __dadd_rn(__dsub_rn(a, __dsub_rn(x, z)), __dsub_rn(b, z));
    return {x, y};
}
__device__ __forceinline__ dfl eft_dekker(double a, double b)
{
    const double x = __dadd_rn(a, b);
    const double y = __dadd_rn(__dsub_rn(a, x), b);
    return {x, y};
}
__device__ __forceinline__ dfl dfl_add(dfl a, dfl b)
{
    const dfl h = eft_knuth(a.hi, b.hi);
    const dfl l = eft_knuth(a.lo, b.lo);
    dfl uv = eft_dekker(h.hi, __dadd_rn(h.lo, l.hi));
    uv = eft_dekker(uv.hi, __dadd_rn(uv.lo, l.lo));
    return uv;
}
__device__ __forceinline__ dfl dfl_sub(dfl a, dfl b)
{
    return dfl_add(a, dfl{-b.hi, -b.lo});
}
__device__ __forceinline__ bool dfl_lt(dfl x, dfl y)
{
    return (x.hi < y.hi) || (x.hi == y.hi && x.lo < y.lo);
}
__device__ __forceinline__ bool dfl_ge0(dfl x)
{
    return (x.hi > 0.) || (x.hi == 0. && x.lo >= 0.);
}

} // namespace heyoka_b200::dev

#endif
