// k_nb: the dedicated sm_100a kernel for N-body-shaped programs (nb_plan.hpp): model::nbody of the outer Solar System
// (6 bodies, 15 pair interactions), the two-body step benchmark, model::nbody with 32 bodies (496 pair interactions).
//
// Same persistent structure as k_coop (kernels.cuh): a team (a warp, or the whole CTA when one lane has hundreds of pair
// interactions) owns LT lanes, claims chunks of LT lanes from an atomic counter and runs a chunk's whole
// propagate_until() loop; the state update and the per-lane bookkeeping are the functions of kernels.cuh. What
// differs is the jet:
//   * the orders are walked two at a time (nb_core.hpp), two synchronisations per PAIR of orders;
//   * a thread is bound to one (pair interaction, lane) for the whole kernel: its operands' shared-memory addresses
//     and constants live in registers, nothing is decoded per order;
//   * its private history rows are stored as (even order, odd order) pairs: d_0, d_1 in shared memory, interleaved by
//     thread ([order pair][row][thread], one 16-byte access per thread, conflict-free), r^2, d_2 and r^alpha in tensor
//     memory (12 columns per order pair: one tcgen05.ld.x8 + one .x4 per loop iteration); or all five in shared
//     memory (TMEM = false);
//   * shared memory otherwise only holds what threads exchange: the positions of the current order pair and the
//     outputs of the pair interactions / partial sums ([role][pair][lane] so that a warp writes consecutive slots);
//   * what a thread does in the summation phase is a pre-decoded 32-byte record (nb_role) per round;
//   * the three infinity norms of the step-size estimate are gathered while the coefficients are produced
//     (shared-memory atomic max on the bit patterns of |x|): no second pass over the coefficients for h.
// All shared-memory accesses use 32-bit shared-window addresses (ld.shared / st.shared): no generic addressing, no
// 64-bit pointer arithmetic in the hot loops.
// Replaces, for these programs: the JIT'd step function (src/taylor_00.cpp:712-865) and the propagate loop
// (src/taylor_adaptive_batch.cpp:1136-1534), like k_coop.
#ifndef HEYOKA_B200_CSRC_NB_KERNEL_CUH
#define HEYOKA_B200_CSRC_NB_KERNEL_CUH

#include <cstdint>

#include <cuda_runtime.h>

#include "kernels.cuh"
#include "nb_core.hpp"
#include "nb_desc.hpp"
#include "tmem.cuh"

namespace heyoka_b200::dev
{

// Systems with ONE pair interaction run one thread per lane (k_nb1, nb1_kernel.cuh). Slot s = 3 * side + k: side 0 /
// 1 = the body whose positions are the pair's pa / pb, k = coordinate. The accelerations of a side's velocities v_sv
// (whose position children are x_sv) are the pair outputs m_k (kind 0), n_k (kind 1) or the number 0 (kind 2).
// sv0_slot / sv0_is_x: where state variable 0 sits (its NaNs are the ones the step-size norms let through).
struct nb1_tab {
    std::uint32_t v_sv[6], x_sv[6], kind[2];
    std::uint32_t sv0_slot, sv0_is_x;
};

// Device-side view of an nb_plan (arrays in global memory) + the shared-memory layout chosen by the host.
struct nb_dev_plan {
    const detail::nb_pair_desc *pairs;
    const uint4 *roles; // n_rounds x TT records of 2 x uint4
    const double *consts, *fac;
    std::uint32_t n_pairs, n_pos, n_out, n_consts, npp, fac_stride;
    std::uint32_t n_rounds, round_level_end; // rounds of the summation phase; bit r: round r ends a level
    double alpha;
    std::uint32_t pow_algo;
    std::uint32_t roles_in_smem;  // 1: the role table is copied to shared memory
    std::uint32_t shared_doubles; // CTA-shared tables: fac | rcp | consts | roles
    std::uint32_t team_doubles;   // per team: positions | outputs | private rows | norms | scalars
    std::uint32_t n_slots_equiv;  // team region (without the scalars) expressed in coop_smem<LT> slots
    nb1_tab l1;                   // k_nb1 only
};

namespace nbk
{

using nb::d2;

__device__ __forceinline__ std::uint32_t saddr(const void *p)
{
    return static_cast<std::uint32_t>(__cvta_generic_to_shared(p));
}
// Pins a loop-invariant value in a register: without it the compiler re-derives shared-window addresses and kernel
// parameters (a dozen uniform-datapath instructions each) inside the per-order-pair code instead of keeping them.
__device__ __forceinline__ void keep(std::uint32_t &x)
{
    asm volatile("" : "+r"(x));
}
__device__ __forceinline__ d2 lds2(std::uint32_t a)
{
    d2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds1(std::uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds4u(std::uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts2(std::uint32_t a, const d2 &v)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void red_max_u64(std::uint32_t a, unsigned long long v)
{
    asm volatile("red.shared.max.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}
__device__ __forceinline__ d2 from_words(std::uint32_t a, std::uint32_t b, std::uint32_t c, std::uint32_t d)
{
    return d2{__hiloint2double(static_cast<int>(b), static_cast<int>(a)),
              __hiloint2double(static_cast<int>(d), static_cast<int>(c))};
}
__device__ __forceinline__ tm::words<4> to_words(const d2 &v)
{
    tm::words<4> w;
    w.w[0] = static_cast<std::uint32_t>(__double2loint(v.x));
    w.w[1] = static_cast<std::uint32_t>(__double2hiint(v.x));
    w.w[2] = static_cast<std::uint32_t>(__double2loint(v.y));
    w.w[3] = static_cast<std::uint32_t>(__double2hiint(v.y));
    return w;
}

// Storage policy of pair_block() (nb_core.hpp). TT = threads per team. All members are shared-window addresses.
// Private rows in shared memory: element (order pair op, row r) of this thread at drow + (op * NSR + r) * TT * 16,
// rows d_0, d_1 (+ d_2, r^2, r^alpha when TMEM is false). Tensor memory: columns [op * 12, op * 12 + 12) of the
// thread's TMEM lane = r^2 pair, d_2 pair, r^alpha pair.
template <int TT, bool TMEM>
struct pair_mem {
    static constexpr int NSR = TMEM ? 2 : 5;
    static constexpr std::uint32_t OPB = static_cast<std::uint32_t>(NSR) * TT * 16u; // bytes per order pair
    static constexpr std::uint32_t RB = TT * 16u;                                    // bytes per row
    std::uint32_t pa[3], pb[3]; // the six positions this pair reads (this thread's lane)
    std::uint32_t om, kstride;  // output m_0 of this (pair, lane); m_k / n_k are k / (3 + k) strides further
    std::uint32_t drow;         // this thread's slice of the private rows
    std::uint32_t fac_, fac_stride_b;
    std::uint32_t tmc; // TMEM address of this thread's column 0
    std::uint32_t flags; // bit 0: active (owns a pair), bits 1-3: n_k exists

    __device__ __forceinline__ d2 pos_a(int k) const
    {
        return lds2(pa[k]);
    }
    __device__ __forceinline__ d2 pos_b(int k) const
    {
        return lds2(pb[k]);
    }
    __device__ __forceinline__ void st_d(std::uint32_t m, const d2 (&D)[3]) const
    {
        const std::uint32_t p = drow + m * OPB;
        sts2(p, D[0]);
        sts2(p + RB, D[1]);
        if constexpr (TMEM) {
            tm::st(tmc + m * 12u + 4u, to_words(D[2]));
        } else {
            sts2(p + 2u * RB, D[2]);
        }
    }
    __device__ __forceinline__ void st_r2(std::uint32_t m, const d2 &r) const
    {
        if constexpr (TMEM) {
            tm::st(tmc + m * 12u, to_words(r));
        } else {
            sts2(drow + m * OPB + 3u * RB, r);
        }
    }
    __device__ __forceinline__ void st_q(std::uint32_t m, const d2 &q) const
    {
        if constexpr (TMEM) {
            tm::st(tmc + m * 12u + 8u, to_words(q));
        } else {
            sts2(drow + m * OPB + 4u * RB, q);
        }
    }
    __device__ __forceinline__ void ld_ss(std::uint32_t ai, std::uint32_t li, d2 (&A)[3], d2 (&Lo)[3]) const
    {
        const std::uint32_t pa_ = drow + ai * OPB, pl = drow + li * OPB;
        if constexpr (TMEM) {
            tm::words<4> wa, wl;
            tm::ld(tmc + ai * 12u + 4u, wa);
            tm::ld(tmc + li * 12u + 4u, wl);
            A[0] = lds2(pa_);
            A[1] = lds2(pa_ + RB);
            Lo[0] = lds2(pl);
            Lo[1] = lds2(pl + RB);
            tm::wait_ld(wa);
            tm::wait_ld(wl);
            A[2] = from_words(wa.w[0], wa.w[1], wa.w[2], wa.w[3]);
            Lo[2] = from_words(wl.w[0], wl.w[1], wl.w[2], wl.w[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                A[k] = lds2(pa_ + k * RB);
                Lo[k] = lds2(pl + k * RB);
            }
        }
    }
    __device__ __forceinline__ void ld_a(std::uint32_t ai, d2 (&A)[3]) const
    {
        const std::uint32_t pa_ = drow + ai * OPB;
        if constexpr (TMEM) {
            tm::words<4> wa;
            tm::ld(tmc + ai * 12u + 4u, wa);
            A[0] = lds2(pa_);
            A[1] = lds2(pa_ + RB);
            tm::wait_ld(wa);
            A[2] = from_words(wa.w[0], wa.w[1], wa.w[2], wa.w[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                A[k] = lds2(pa_ + k * RB);
            }
        }
    }
    __device__ __forceinline__ void ld_main(std::uint32_t qi, std::uint32_t li, d2 &Q, d2 &Rlo, d2 (&Dlo)[3]) const
    {
        const std::uint32_t pl = drow + li * OPB;
        if constexpr (TMEM) {
            tm::words<8> wl;
            tm::words<4> wq;
            tm::ld(tmc + li * 12u, wl);
            tm::ld(tmc + qi * 12u + 8u, wq);
            Dlo[0] = lds2(pl);
            Dlo[1] = lds2(pl + RB);
            tm::wait_ld(wl);
            tm::wait_ld(wq);
            Rlo = from_words(wl.w[0], wl.w[1], wl.w[2], wl.w[3]);
            Dlo[2] = from_words(wl.w[4], wl.w[5], wl.w[6], wl.w[7]);
            Q = from_words(wq.w[0], wq.w[1], wq.w[2], wq.w[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                Dlo[k] = lds2(pl + k * RB);
            }
            Rlo = lds2(pl + 3u * RB);
            Q = lds2(drow + qi * OPB + 4u * RB);
        }
    }
    __device__ __forceinline__ d2 fac(std::uint32_t n, std::uint32_t j) const
    {
        return lds2(fac_ + n * fac_stride_b + j * 8u);
    }
    __device__ __forceinline__ double fac1(std::uint32_t n, std::uint32_t j) const
    {
        return lds1(fac_ + n * fac_stride_b + j * 8u);
    }
    __device__ __forceinline__ void out(int k, const d2 &v) const
    {
        if ((flags & 1u) != 0u) {
            sts2(om + static_cast<std::uint32_t>(k) * kstride, v);
        }
    }
    __device__ __forceinline__ void out_n(int k, const d2 &v) const
    {
        if ((flags & (2u << k)) != 0u) {
            sts2(om + static_cast<std::uint32_t>(3 + k) * kstride, v);
        }
    }
};

// Storage policy of role_block() / role_init(): the thread's NL lanes start at lane l0 of the team's LT lanes (the
// records' units already include l0).
template <int NL>
struct role_mem {
    std::uint32_t pos_b, out_b;   // team bases
    std::uint32_t consts, rcp_;   // CTA tables
    std::uint32_t norms;          // team norms: [3][LT] u64 (|x^[0]|, |x^[p]|, |x^[p-1]|), + lane l0 of this thread
    std::uint32_t lt8;            // copies * LT * 8: stride between the three norms
    const double *state0;         // D.state + first global lane of this thread (clamped)
    std::size_t n_batch;
    double *cbase;                // coefficient store + lane offset of lane 0 of this thread
    std::size_t stride_sv, stride_o;
    std::uint32_t p;
    bool pub, track;
    bool lane_ok[NL];
    std::uint32_t ldelta; // offset (in doubles) between the thread's lanes in state / public store (0 if clamped)

    __device__ __forceinline__ d2 out_u(std::uint32_t unit, int l) const
    {
        return lds2(out_b + (unit + l) * 16u);
    }
    __device__ __forceinline__ void out_st_u(std::uint32_t unit, int l, const d2 &v) const
    {
        sts2(out_b + (unit + l) * 16u, v);
    }
    __device__ __forceinline__ void pos_st_u(std::uint32_t unit, int l, const d2 &v) const
    {
        sts2(pos_b + (unit + l) * 16u, v);
    }
    __device__ __forceinline__ double cst(std::uint32_t i) const
    {
        return lds1(consts + i * 8u);
    }
    __device__ __forceinline__ double rcp(std::uint32_t n) const
    {
        return lds1(rcp_ + n * 8u);
    }
    // Norms of the step-size estimate: NaN-skipping maximum of |v| (bit patterns of non-negative doubles order like
    // unsigned integers). which: 0 = order 0, 1 = order p, 2 = order p - 1.
    __device__ __forceinline__ void norm(std::uint32_t which, int l, double v) const
    {
        if (v == v) {
            red_max_u64(norms + which * lt8 + l * 8u,
                        static_cast<unsigned long long>(__double_as_longlong(v)) & 0x7fffffffffffffffull);
        }
    }
    __device__ __forceinline__ void track_order(std::uint32_t order, const double (&a)[NL]) const
    {
        if (order == 0u || order == p || order + 1u == p) {
            const std::uint32_t which = order == 0u ? 0u : (order == p ? 1u : 2u);
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                norm(which, l, a[l]);
            }
        }
    }
    // One order of state variable sv for the thread's lanes at element index idx of the coefficient store.
    template <typename I>
    __device__ __forceinline__ void store_lanes(I idx, const double (&a)[NL]) const
    {
        if constexpr (NL == 2) {
            if (!pub) {
                // Private store: the two lanes are adjacent and 16-byte aligned.
                *reinterpret_cast<double2 *>(cbase + idx) = make_double2(a[0], a[1]);
                return;
            }
        }
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            if (lane_ok[l]) {
                cbase[idx + (pub ? l * ldelta : l)] = a[l];
            }
        }
    }
    __device__ __forceinline__ void coef_pair(std::uint32_t sv, std::uint32_t order, const double (&a)[NL],
                                              const double (&b)[NL]) const
    {
        if (pub) {
            const std::size_t idx = sv * stride_sv + order * stride_o;
            if (order <= p) {
                store_lanes(idx, a);
            }
            if (order + 1u <= p) {
                store_lanes(idx + stride_o, b);
            }
        } else {
            // (The private store has fewer than 2^32 elements.)
            const std::uint32_t so = static_cast<std::uint32_t>(stride_o);
            const std::uint32_t idx = sv * static_cast<std::uint32_t>(stride_sv) + order * so;
            if (order <= p) {
                store_lanes(idx, a);
            }
            if (order + 1u <= p) {
                store_lanes(idx + so, b);
            }
        }
        if (track) {
            track_order(order, a);
            track_order(order + 1u, b);
        }
    }
    __device__ __forceinline__ void coef_one(std::uint32_t sv, std::uint32_t order, const double (&a)[NL]) const
    {
        store_lanes(sv * stride_sv + order * stride_o, a);
        if (track) {
            track_order(order, a);
        }
    }
    __device__ __forceinline__ double state(std::uint32_t sv, int l) const
    {
        return state0[static_cast<std::size_t>(sv) * n_batch + l * ldelta];
    }
};

} // namespace nbk

// Step size of one lane from the norms gathered during the jet (norms[0], norms[lt], norms[2 lt]: bit patterns of the
// NaN-skipping maxima of |x^[0]|, |x^[p]|, |x^[p-1]|; reset to 0 here) and the coefficients of the first state
// variable: the sequential reference loop m = (m < |x|) ? |x| : m, started from |x_0|, yields NaN iff x_0 is NaN and
// ignores every other NaN (src/taylor_00.cpp:102-273). Once per lane and step: out of line.
static __device__ __noinline__ double nb_step_size(const program &P, unsigned long long *norms, std::uint32_t lt,
                                                   const double *c, std::size_t off_p, std::size_t off_pm1,
                                                   double max_delta_t)
{
    // norms[(which * copies + r) * lt]: the maximum over the copies, which are reset.
    const std::uint32_t copies = detail::nb_norm_copies(lt);
    unsigned long long n3[3];
    for (std::uint32_t w = 0; w < 3u; ++w) {
        unsigned long long v = 0ull;
        for (std::uint32_t r = 0; r < copies; ++r) {
            unsigned long long *q = norms + (w * copies + r) * lt;
            v = *q > v ? *q : v;
            *q = 0ull;
        }
        n3[w] = v;
    }
    const double m0 = __longlong_as_double(static_cast<long long>(n3[0]));
    const double mp = __longlong_as_double(static_cast<long long>(n3[1]));
    const double mp1 = __longlong_as_double(static_cast<long long>(n3[2]));
    const double f0 = fabs(c[0]), fp = fabs(c[off_p]), fp1 = fabs(c[off_pm1]);
    return h_from_norms(P, isnan(f0) ? f0 : m0, isnan(fp) ? fp : mp, isnan(fp1) ? fp1 : mp1, max_delta_t);
}

// LT: lanes per team; CTA: a team is the whole CTA (else a warp); TMEM: r^2, d_2, r^alpha rows in tensor memory.
template <int LT, bool CTA, bool TMEM, bool PROP, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_nb(program P, nb_dev_plan NP, batch D, run_args R)
{
    using T = team<CTA>;
    constexpr int TT = CTA ? MAXT : 32; // threads per team (CTA teams are launched with exactly MAXT threads)
    constexpr int NL = LT >= 2 ? 2 : 1; // lanes per thread in the summation phase
    constexpr std::uint32_t GS = LT / NL;
    extern __shared__ __align__(16) double smem_raw[];

    // ---- CTA-shared tables: fac | rcp | consts | roles ----
    const std::uint32_t p = P.order;
    double *fac_s = smem_raw;
    const std::uint32_t n_fac = (p + 1u) * NP.fac_stride;
    double *rcp_s = fac_s + n_fac;
    const std::uint32_t n_rcp = (p + 5u) & ~1u;
    double *consts_s = rcp_s + n_rcp;
    const std::uint32_t n_cst = (NP.n_consts + 1u) & ~1u;
    uint4 *roles_s = reinterpret_cast<uint4 *>(consts_s + n_cst);
    for (std::uint32_t i = threadIdx.x; i < n_fac; i += blockDim.x) {
        fac_s[i] = __ldg(NP.fac + i);
    }
    for (std::uint32_t i = threadIdx.x; i < n_rcp; i += blockDim.x) {
        rcp_s[i] = i == 0u ? 0. : 1. / static_cast<double>(i);
    }
    for (std::uint32_t i = threadIdx.x; i < NP.n_consts; i += blockDim.x) {
        consts_s[i] = __ldg(NP.consts + i);
    }
    if (NP.roles_in_smem != 0u) {
        for (std::uint32_t i = threadIdx.x; i < NP.n_rounds * TT * 2u; i += blockDim.x) {
            roles_s[i] = __ldg(NP.roles + i);
        }
    }
    __shared__ std::uint32_t tm_base_smem;
    if constexpr (TMEM) {
        if ((threadIdx.x >> 5) == 0u) {
            tm::alloc_all(&tm_base_smem);
        }
        tm::fence_before_sync();
    }
    __syncthreads();

    const std::uint32_t tid = T::tid();
    double *region = smem_raw + NP.shared_doubles
                     + (CTA ? 0u : static_cast<std::size_t>(threadIdx.x >> 5) * NP.team_doubles);
    const coop_smem<LT> S(region, NP.n_slots_equiv);
    std::uint32_t pos_b = nbk::saddr(region);
    std::uint32_t out_b = pos_b + NP.n_pos * LT * 16u;
    const std::uint32_t drow_b = out_b + NP.n_out * LT * 16u;
    const std::uint32_t norms_b = drow_b + NP.npp * nbk::pair_mem<TT, TMEM>::OPB;
    nbk::keep(pos_b);
    nbk::keep(out_b);
    unsigned long long *norms_p
        = reinterpret_cast<unsigned long long *>(region + (static_cast<std::size_t>(NP.n_pos) + NP.n_out) * LT * 2u
                                                 + static_cast<std::size_t>(NP.npp) * nbk::pair_mem<TT, TMEM>::OPB / 8u);

    // ---- this thread's pair interaction ----
    nbk::pair_mem<TT, TMEM> PM;
    nb::pair_consts PC;
    {
        const std::uint32_t n_pt = NP.n_pairs * LT;
        const bool active = tid < n_pt;
        // Idle threads shadow pair 0 / lane 0 on their own private rows (the tensor-memory accesses are warp-wide).
        const std::uint32_t pi = active ? tid / LT : 0u, l = active ? tid % LT : 0u;
        const uint4 *dp = reinterpret_cast<const uint4 *>(NP.pairs + pi);
        const uint4 w0 = __ldg(dp), w1 = __ldg(dp + 1), w2 = __ldg(dp + 2), w3 = __ldg(dp + 3);
        // u16 fields: pa[3] pb[3] om[3] on[3] = words w0.x .. w1.y; flags = w1.z; c1 = w2.xy; c2[3] = w2.zw, w3.xy, w3.zw
        const std::uint32_t h[6] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y};
        const auto u16 = [&](int i) { return (h[i >> 1] >> ((i & 1) * 16)) & 0xffffu; };
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            PM.pa[k] = pos_b + (u16(k) * LT + l) * 16u;
            PM.pb[k] = pos_b + (u16(3 + k) * LT + l) * 16u;
        }
        PM.om = out_b + (u16(6) * LT + l) * 16u; // om[k] = k * n_pairs + pair, on[k] = (3 + k) * n_pairs + pair
        PM.kstride = NP.n_pairs * LT * 16u;
        PM.flags = (active ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (active && u16(9 + k) != 0xffffu) {
                PM.flags |= 2u << k;
            }
        }
        PC.c1 = __hiloint2double(static_cast<int>(w2.y), static_cast<int>(w2.x));
        PC.c2[0] = __hiloint2double(static_cast<int>(w2.w), static_cast<int>(w2.z));
        PC.c2[1] = __hiloint2double(static_cast<int>(w3.y), static_cast<int>(w3.x));
        PC.c2[2] = __hiloint2double(static_cast<int>(w3.w), static_cast<int>(w3.z));
        PC.alpha = NP.alpha;
        PC.pow_algo = NP.pow_algo;
        PC.have_n = (w1.z & 1u) != 0u;
        PM.drow = drow_b + tid * 16u;
        PM.fac_ = nbk::saddr(fac_s);
        PM.fac_stride_b = NP.fac_stride * 8u;
        PM.tmc = 0u;
        if constexpr (TMEM) {
            tm::fence_after_sync();
            // Warp w owns the columns [(w / 4) * cols, ...) of the 32 TMEM lanes of its quadrant w % 4.
            const std::uint32_t w = threadIdx.x >> 5;
            PM.tmc = tm_base_smem + (((w & 3u) * 32u) << 16) + (w >> 2) * (NP.npp * 12u);
        }
    }
    // ---- this thread's lanes in the summation phase ----
    const std::uint32_t l0 = (tid % GS) * NL;
    nbk::role_mem<NL> RM;
    RM.pos_b = pos_b;
    RM.out_b = out_b;
    RM.consts = nbk::saddr(consts_s);
    RM.rcp_ = nbk::saddr(rcp_s);
    constexpr std::uint32_t NC = detail::nb_norm_copies(LT);
    RM.norms = norms_b + ((tid % NC) * LT + l0) * 8u;
    RM.lt8 = NC * LT * 8u;
    RM.n_batch = D.n;
    RM.p = p;
    const std::size_t team_global = T::index();
    const coef_view cv{R.coef_base + team_global * R.coef_warp_stride, static_cast<std::size_t>(R.coef_stride_sv),
                       static_cast<std::size_t>(R.coef_stride_o), R.coef_pub != 0, !PROP && R.skip != nullptr};
    RM.stride_sv = cv.stride_sv;
    RM.stride_o = cv.stride_o;
    RM.pub = cv.pub;
    std::uint32_t roles_sa = nbk::saddr(roles_s) + tid * 32u;
    nbk::keep(roles_sa);
    nbk::keep(RM.consts);
    nbk::keep(RM.rcp_);
    nbk::keep(PM.fac_);
    const std::uint32_t n_rounds = NP.n_rounds, level_end = NP.round_level_end, roles_smem = NP.roles_in_smem;
    const uint4 *roles_g = NP.roles + tid * 2u;

    const std::uint32_t n_chunks = (D.n + LT - 1u) / LT;
    const bool owner = tid < LT;
    const std::uint32_t n_blocks = NP.npp;

    const auto load_role = [&](std::uint32_t rd, std::uint32_t (&w)[8]) {
        uint4 a, b;
        if (roles_smem != 0u) {
            a = nbk::lds4u(roles_sa + rd * (TT * 32u));
            b = nbk::lds4u(roles_sa + rd * (TT * 32u) + 16u);
        } else {
            a = __ldg(roles_g + static_cast<std::size_t>(rd) * (TT * 2u));
            b = __ldg(roles_g + static_cast<std::size_t>(rd) * (TT * 2u) + 1);
        }
        w[0] = a.x, w[1] = a.y, w[2] = a.z, w[3] = a.w, w[4] = b.x, w[5] = b.y, w[6] = b.z, w[7] = b.w;
    };

    const auto jet = [&](std::uint32_t lane0) {
        // The thread's lanes: global indices (clamped), offsets into the coefficient store.
        {
            const std::uint32_t la = lane0 + l0, lb = la + (NL - 1);
            const std::uint32_t ga = la < D.n ? la : D.n - 1u, gb = lb < D.n ? lb : D.n - 1u;
            // (A step with a skip mask leaves the lanes that are not running untouched, tc included.)
            RM.lane_ok[0] = la < D.n && !(cv.mask_idle && S.running[l0] == 0);
            if constexpr (NL == 2) {
                RM.lane_ok[1] = lb < D.n && !(cv.mask_idle && S.running[l0 + 1] == 0);
            }
            RM.ldelta = gb - ga;
            RM.state0 = D.state + ga;
            RM.cbase = cv.base + cv.lane_off(ga, l0);
        }
        RM.track = true;
        for (std::uint32_t rd = 0; rd < n_rounds; ++rd) {
            std::uint32_t w[8];
            load_role(rd, w);
            nb::role_init<NL>(RM, w);
        }
        T::sync();
        for (std::uint32_t m = 0; m < n_blocks; ++m) {
            if (TMEM || (PM.flags & 1u) != 0u) {
                nb::pair_block(PM, PC, m);
            }
            if constexpr (TMEM) {
                tm::wait_st();
            }
            T::sync();
            RM.track = m + 2u >= n_blocks;
            for (std::uint32_t rd = 0; rd < n_rounds; ++rd) {
                std::uint32_t w[8];
                load_role(rd, w);
                nb::role_block<NL>(RM, w, m, p);
                if (((level_end >> rd) & 1u) != 0u) {
                    T::sync();
                }
            }
        }
    };

    // Step size from the norms gathered during the jet (owner threads: one lane each); resets the norms.
    const auto step_size = [&](std::uint32_t lane, double max_delta_t) {
        const double *c = cv.base + cv.lane_off(lane, tid);
        return nb_step_size(P, norms_p + tid, LT, c, p * cv.stride_o, (p - 1u) * cv.stride_o, max_delta_t);
    };
    if (owner) {
        for (std::uint32_t i = 0; i < 3u * NC; ++i) {
            norms_p[i * LT + tid] = 0ull;
        }
    }
    // The per-lane bookkeeping of propagate_until() is parked in shared memory while the jet runs (it would otherwise
    // hold ~26 registers of every thread across the hot loops).
    static_assert(sizeof(lane_prop) <= 128u && alignof(lane_prop) <= 8u);
    lane_prop *const park = reinterpret_cast<lane_prop *>(norms_p + 3u * NC * LT) + (owner ? tid : 0u);

    for (std::uint32_t chunk = T::claim(R.counter); chunk < n_chunks; chunk = T::claim(R.counter)) {
        const std::uint32_t lane0 = chunk * LT;
        const std::uint32_t lane_raw = lane0 + tid;
        bool valid = owner && lane_raw < D.n;
        const std::uint32_t lane = (owner && lane_raw < D.n) ? lane_raw : D.n - 1u;

        if constexpr (!PROP) {
            if (owner) {
                const bool skipped = R.skip != nullptr && R.skip[lane] != 0u;
                valid = valid && !skipped;
                S.time[tid] = D.t_hi[lane];
                S.running[tid] = skipped ? 0 : 1;
            }
            T::sync();
            jet(lane0);
            double h = 0., mdt = 0.;
            if (owner) {
                mdt = R.max_delta_t != nullptr ? R.max_delta_t[lane] : R.default_max_delta_t;
                h = step_size(lane, mdt);
                S.h[tid] = h;
            }
            T::sync();
            unsigned nf_mask = 0u;
            coop_update_state<LT, CTA>(P, D, S, cv, lane0, nf_mask);
            nf_mask = T::template reduce_or<LT>(nf_mask);
            if (valid) {
                const dfl nt = dfl_add(dfl{D.t_hi[lane], D.t_lo[lane]}, dfl{h, 0.});
                D.t_hi[lane] = nt.hi;
                D.t_lo[lane] = nt.lo;
                D.last_h[lane] = h;
                const bool nf = !(isfinite(nt.hi) && isfinite(nt.lo)) || ((nf_mask >> tid) & 1u) != 0u;
                D.step_outcome[lane]
                    = nf ? HY_OUTCOME_ERR_NF_STATE : (h == mdt ? HY_OUTCOME_TIME_LIMIT : HY_OUTCOME_SUCCESS);
            }
        } else {
            bool running = false;
            if (owner) {
                lane_prop lp;
                lp.init(D, R, lane);
                *park = lp;
                running = lp.running;
            }
            while (T::any(running)) {
                if (owner) {
                    S.time[tid] = park->t.hi;
                    S.running[tid] = running ? 1 : 0;
                }
                T::sync();
                jet(lane0);
                double h = 0., cur_max = 0.;
                if (owner) {
                    cur_max = park->cur_max();
                    h = step_size(lane, cur_max);
                    S.h[tid] = h;
                }
                T::sync();
                unsigned nf_mask = 0u;
                coop_update_state<LT, CTA>(P, D, S, cv, lane0, nf_mask);
                nf_mask = T::template reduce_or<LT>(nf_mask);
                if (running) {
                    lane_prop lp = *park;
                    lp.advance(h, cur_max, ((nf_mask >> tid) & 1u) != 0u, R, valid);
                    *park = lp;
                    running = lp.running;
                }
            }
            if (valid) {
                park->store(D, lane);
                park->report_iters(R);
            }
        }
        T::sync();
    }
    if constexpr (TMEM) {
        tm::fence_before_sync();
        __syncthreads();
        if ((threadIdx.x >> 5) == 0u) {
            tm::dealloc_all(tm_base_smem);
        }
    }
}

} // namespace heyoka_b200::dev

#endif
