// k_nb: the dedicated sm_100a kernel for N-body-shaped programs (nb_plan.hpp): model::nbody of the outer Solar System
// (6 bodies, 15 pair interactions), the two-body step benchmark, model::nbody with 32 bodies (496 pair interactions).
//
// Same persistent structure as k_coop (kernels.cuh): a team (a warp, or the whole CTA when one lane has hundreds of pair
// interactions) owns LT lanes, claims chunks of LT lanes from an atomic counter and runs a chunk's whole
// propagate_until() loop; the step-size estimate, the state update and the per-lane bookkeeping are the functions of
// kernels.cuh. What differs is the jet:
//   * the orders are walked two at a time (nb_core.hpp), two synchronisations per PAIR of orders;
//   * a thread is bound to one (pair interaction, lane) for the whole kernel: its operands' addresses and constants
//     live in registers, nothing is decoded per order;
//   * its private history rows are stored as (even order, odd order) pairs: d_0, d_1 in shared memory, interleaved by
//     thread ([order pair][row][thread], one 16-byte access per thread, conflict-free), r^2, d_2 and r^alpha in tensor
//     memory (12 columns per order pair: one tcgen05.ld.x8 + one .x4 per loop iteration); or all five in shared
//     memory (TMEM = false);
//   * shared memory otherwise only holds what threads exchange: the positions of the current order pair and the
//     outputs of the pair interactions / partial sums.
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

// Device-side view of an nb_plan (arrays in global memory) + the shared-memory layout chosen by the host.
struct nb_dev_plan {
    const detail::nb_pair_desc *pairs;
    const std::uint32_t *sums; // 16 words per item
    const double *consts, *fac;
    std::uint32_t n_pairs, n_pos, n_out, n_levels, n_sums, n_consts, npp, fac_stride;
    std::uint32_t level_offsets[8]; // n_levels + 1 offsets into sums
    double alpha;
    std::uint32_t pow_algo;
    std::uint32_t sums_in_smem;   // 1: the sum descriptors are copied to shared memory
    std::uint32_t shared_doubles; // CTA-shared tables: fac | rcp | consts | sums
    std::uint32_t team_doubles;   // per team: positions | outputs | private rows | scalars
    std::uint32_t n_slots_equiv;  // team region expressed in coop_smem<LT> slots
};

namespace nbk
{

using nb::d2;

__device__ __forceinline__ d2 lds2(const char *p)
{
    const double2 v = *reinterpret_cast<const double2 *>(p);
    return d2{v.x, v.y};
}
__device__ __forceinline__ void sts2(char *p, const d2 &v)
{
    *reinterpret_cast<double2 *>(p) = make_double2(v.x, v.y);
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

// Storage policy of pair_block() (nb_core.hpp). TT = threads per team.
// Shared-memory private rows: element (order pair op, row r) of this thread at drow + (op * NSR + r) * TT * 16 bytes,
// rows d_0, d_1 (+ d_2, r^2, r^alpha when TMEM is false). Tensor memory: columns [op * 12, op * 12 + 12) of the
// thread's TMEM lane = r^2 pair, d_2 pair, r^alpha pair.
template <int TT, bool TMEM>
struct pair_mem {
    static constexpr int NSR = TMEM ? 2 : 5;
    static constexpr std::uint32_t OPB = static_cast<std::uint32_t>(NSR) * TT * 16u; // bytes per order pair
    static constexpr std::uint32_t RB = TT * 16u;                                    // bytes per row
    const char *pos; // team positions + this thread's lane
    char *outp;      // team outputs + this thread's lane
    char *drow;      // this thread's slice of the private rows
    const double *fac_;
    std::uint32_t fac_stride;
    std::uint32_t pa[3], pb[3], om[3]; // byte offsets
    std::uint32_t tmc;                 // TMEM address of this thread's column 0
    bool active;

    __device__ __forceinline__ d2 pos_a(int k) const
    {
        return lds2(pos + pa[k]);
    }
    __device__ __forceinline__ d2 pos_b(int k) const
    {
        return lds2(pos + pb[k]);
    }
    __device__ __forceinline__ void st_d(std::uint32_t m, const d2 (&D)[3]) const
    {
        char *p = drow + m * OPB;
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
        const char *pa_ = drow + ai * OPB, *pl = drow + li * OPB;
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
        const char *pa_ = drow + ai * OPB;
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
        const char *pl = drow + li * OPB;
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
        const double2 v = *reinterpret_cast<const double2 *>(fac_ + n * fac_stride + j);
        return d2{v.x, v.y};
    }
    __device__ __forceinline__ double fac1(std::uint32_t n, std::uint32_t j) const
    {
        return fac_[n * fac_stride + j];
    }
    __device__ __forceinline__ void out(int k, const d2 &v) const
    {
        if (active) {
            sts2(outp + om[k], v);
        }
    }
};

// Storage policy of sum_block() / sum_init(): NL lanes starting at lane l0 of the team's LT lanes.
template <int LT, int NL>
struct sum_mem {
    char *pos, *out; // team bases
    const double *consts, *rcp_;
    const batch *D;
    coef_view cv;
    std::uint32_t l0;
    std::uint32_t glane[NL];
    std::size_t loff[NL];
    bool lane_ok[NL];

    __device__ __forceinline__ d2 out_ld(std::uint32_t slot, int l) const
    {
        return lds2(out + (slot * LT + l0 + l) * 16u);
    }
    __device__ __forceinline__ void out_st(std::uint32_t slot, int l, const d2 &v) const
    {
        sts2(out + (slot * LT + l0 + l) * 16u, v);
    }
    __device__ __forceinline__ void pos_st(std::uint32_t slot, int l, const d2 &v) const
    {
        sts2(pos + (slot * LT + l0 + l) * 16u, v);
    }
    __device__ __forceinline__ double cst(std::uint32_t i) const
    {
        return consts[i];
    }
    __device__ __forceinline__ double rcp(std::uint32_t n) const
    {
        return rcp_[n];
    }
    __device__ __forceinline__ void coef(std::uint32_t sv, std::uint32_t order, int l, double v) const
    {
        if (lane_ok[l]) {
            cv.base[sv * cv.stride_sv + order * cv.stride_o + loff[l]] = v;
        }
    }
    __device__ __forceinline__ double state(std::uint32_t sv, int l) const
    {
        return D->state[static_cast<std::size_t>(sv) * D->n + glane[l]];
    }
};

} // namespace nbk

// LT: lanes per team; CTA: a team is the whole CTA (else a warp); TMEM: r^2, d_2, r^alpha rows in tensor memory.
template <int LT, bool CTA, bool TMEM, bool PROP, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_nb(program P, nb_dev_plan NP, batch D, run_args R)
{
    using T = team<CTA>;
    constexpr int TT = CTA ? MAXT : 32; // threads per team (CTA teams are launched with exactly MAXT threads)
    constexpr int NL = LT >= 2 ? 2 : 1; // lanes per thread in the summation levels
    constexpr std::uint32_t GS = LT / NL;
    extern __shared__ __align__(16) double smem_raw[];

    // ---- CTA-shared tables: fac | rcp | consts | sums ----
    const std::uint32_t p = P.order;
    double *fac_s = smem_raw;
    const std::uint32_t n_fac = (p + 1u) * NP.fac_stride;
    double *rcp_s = fac_s + n_fac;
    const std::uint32_t n_rcp = (p + 5u) & ~1u;
    double *consts_s = rcp_s + n_rcp;
    const std::uint32_t n_cst = (NP.n_consts + 1u) & ~1u;
    std::uint32_t *sums_s = reinterpret_cast<std::uint32_t *>(consts_s + n_cst);
    for (std::uint32_t i = threadIdx.x; i < n_fac; i += blockDim.x) {
        fac_s[i] = __ldg(NP.fac + i);
    }
    for (std::uint32_t i = threadIdx.x; i < n_rcp; i += blockDim.x) {
        rcp_s[i] = i == 0u ? 0. : 1. / static_cast<double>(i);
    }
    for (std::uint32_t i = threadIdx.x; i < NP.n_consts; i += blockDim.x) {
        consts_s[i] = __ldg(NP.consts + i);
    }
    const std::uint32_t *sums = NP.sums;
    if (NP.sums_in_smem != 0u) {
        for (std::uint32_t i = threadIdx.x; i < NP.n_sums * 16u; i += blockDim.x) {
            sums_s[i] = __ldg(NP.sums + i);
        }
        sums = sums_s;
    }
    __shared__ std::uint32_t tm_base_smem;
    if constexpr (TMEM) {
        if ((threadIdx.x >> 5) == 0u) {
            tm::alloc_all(&tm_base_smem);
        }
        tm::fence_before_sync();
    }
    __syncthreads();

    const std::uint32_t tid = T::tid(), nthr = T::size();
    const std::size_t team_global = T::index();
    double *region = smem_raw + NP.shared_doubles
                     + (CTA ? 0u : static_cast<std::size_t>(threadIdx.x >> 5) * NP.team_doubles);
    const coop_smem<LT> S(region, NP.n_slots_equiv);
    char *pos_b = reinterpret_cast<char *>(region);
    char *out_b = pos_b + static_cast<std::size_t>(NP.n_pos) * LT * 16u;
    char *drow_b = out_b + static_cast<std::size_t>(NP.n_out) * LT * 16u;

    // ---- this thread's pair interaction ----
    nbk::pair_mem<TT, TMEM> PM;
    nb::pair_consts PC;
    {
        const std::uint32_t n_pt = NP.n_pairs * LT;
        PM.active = tid < n_pt;
        // Idle threads shadow pair 0 / lane 0 on their own private rows (the tensor-memory accesses are warp-wide).
        const std::uint32_t pi = PM.active ? tid / LT : 0u, l = PM.active ? tid % LT : 0u;
        const uint4 *dp = reinterpret_cast<const uint4 *>(NP.pairs + pi);
        const uint4 w0 = __ldg(dp), w1 = __ldg(dp + 1);
        const std::uint32_t h[6] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y};
        const auto u16 = [&](int i) { return (h[i >> 1] >> ((i & 1) * 16)) & 0xffffu; };
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            PM.pa[k] = u16(k) * LT * 16u;
            PM.pb[k] = u16(3 + k) * LT * 16u;
            PM.om[k] = u16(6 + k) * LT * 16u;
        }
        PC.c1 = __hiloint2double(static_cast<int>(w1.w), static_cast<int>(w1.z));
        PC.alpha = NP.alpha;
        PC.pow_algo = NP.pow_algo;
        PM.pos = pos_b + l * 16u;
        PM.outp = out_b + l * 16u;
        PM.drow = drow_b + tid * 16u;
        PM.fac_ = fac_s;
        PM.fac_stride = NP.fac_stride;
        PM.tmc = 0u;
        if constexpr (TMEM) {
            tm::fence_after_sync();
            // Warp w owns the columns [(w / 4) * cols, ...) of the 32 TMEM lanes of its quadrant w % 4.
            const std::uint32_t w = threadIdx.x >> 5;
            PM.tmc = tm_base_smem + (((w & 3u) * 32u) << 16) + (w >> 2) * (NP.npp * 12u);
        }
    }
    // ---- this thread's lanes in the summation levels ----
    nbk::sum_mem<LT, NL> SM;
    SM.pos = pos_b;
    SM.out = out_b;
    SM.consts = consts_s;
    SM.rcp_ = rcp_s;
    SM.D = &D;
    SM.l0 = (tid % GS) * NL;

    const std::uint32_t n_chunks = (D.n + LT - 1u) / LT;
    const bool owner = tid < LT;
    const coef_view cv{R.coef_base + team_global * R.coef_warp_stride, static_cast<std::size_t>(R.coef_stride_sv),
                       static_cast<std::size_t>(R.coef_stride_o), R.coef_pub != 0};
    SM.cv = cv;
    const std::uint32_t n_blocks = (p + 1u) / 2u;

    const auto jet = [&](std::uint32_t lane0) {
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            const std::uint32_t l = lane0 + SM.l0 + i;
            SM.lane_ok[i] = l < D.n;
            SM.glane[i] = SM.lane_ok[i] ? l : D.n - 1u;
            SM.loff[i] = cv.lane_off(SM.glane[i], SM.l0 + i);
        }
        for (std::uint32_t it = tid; it < NP.n_sums * GS; it += nthr) {
            nb::sum_init<NL>(SM, sums + (it / GS) * 16u);
        }
        T::sync();
        for (std::uint32_t m = 0; m < n_blocks; ++m) {
            if (TMEM || PM.active) {
                nb::pair_block(PM, PC, m);
            }
            if constexpr (TMEM) {
                tm::wait_st();
            }
            T::sync();
            for (std::uint32_t lv = 0; lv < NP.n_levels; ++lv) {
                const std::uint32_t b = NP.level_offsets[lv], e = NP.level_offsets[lv + 1u];
                for (std::uint32_t it = tid; it < (e - b) * GS; it += nthr) {
                    nb::sum_block<NL>(SM, sums + (b + it / GS) * 16u, m, p);
                }
                T::sync();
            }
        }
    };

    for (std::uint32_t chunk = T::claim(R.counter); chunk < n_chunks; chunk = T::claim(R.counter)) {
        const std::uint32_t lane0 = chunk * LT;
        const std::uint32_t lane_raw = lane0 + tid;
        const bool valid = owner && lane_raw < D.n;
        const std::uint32_t lane = valid ? lane_raw : D.n - 1u;

        if constexpr (!PROP) {
            double mdt = 0.;
            dfl t0{0., 0.};
            if (owner) {
                mdt = R.max_delta_t != nullptr ? R.max_delta_t[lane] : R.default_max_delta_t;
                t0 = dfl{D.t_hi[lane], D.t_lo[lane]};
                S.time[tid] = t0.hi;
                S.running[tid] = 1;
            }
            T::sync();
            jet(lane0);
            const double h = (!CTA || threadIdx.x < 32u) ? coop_determine_h<LT>(P, D, cv, lane0, mdt) : 0.;
            if (owner) {
                S.h[tid] = h;
            }
            T::sync();
            unsigned nf_mask = 0u;
            coop_update_state<LT, CTA>(P, D, S, cv, lane0, nf_mask);
            nf_mask = T::template reduce_or<LT>(nf_mask);
            if (valid) {
                const dfl nt = dfl_add(t0, dfl{h, 0.});
                D.t_hi[lane] = nt.hi;
                D.t_lo[lane] = nt.lo;
                D.last_h[lane] = h;
                const bool nf = !(isfinite(nt.hi) && isfinite(nt.lo)) || ((nf_mask >> tid) & 1u) != 0u;
                D.step_outcome[lane]
                    = nf ? HY_OUTCOME_ERR_NF_STATE : (h == mdt ? HY_OUTCOME_TIME_LIMIT : HY_OUTCOME_SUCCESS);
            }
        } else {
            lane_prop lp;
            lp.running = false;
            if (owner) {
                lp.init(D, R, lane);
            }
            while (T::any(owner && lp.running)) {
                double cur_max = 0.;
                if (owner) {
                    cur_max = lp.cur_max();
                    S.time[tid] = lp.t.hi;
                    S.running[tid] = lp.running ? 1 : 0;
                }
                T::sync();
                jet(lane0);
                const double h = (!CTA || threadIdx.x < 32u) ? coop_determine_h<LT>(P, D, cv, lane0, cur_max) : 0.;
                if (owner) {
                    S.h[tid] = h;
                }
                T::sync();
                unsigned nf_mask = 0u;
                coop_update_state<LT, CTA>(P, D, S, cv, lane0, nf_mask);
                nf_mask = T::template reduce_or<LT>(nf_mask);
                if (owner && lp.running) {
                    lp.advance(h, cur_max, ((nf_mask >> tid) & 1u) != 0u, R, valid);
                }
            }
            if (valid) {
                lp.store(D, lane);
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
