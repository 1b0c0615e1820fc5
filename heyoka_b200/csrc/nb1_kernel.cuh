// k_nb1: the N-body kernel for systems with ONE pair interaction (the two-body step benchmark,
// benchmark/two_body_step_batch.cpp: model::nbody(2, masses = {1, 0})): one thread per lane, nothing exchanged between
// threads.
//
// k_nb (nb_kernel.cuh) with 32 lanes per warp already gives every thread one (pair interaction, lane), but its
// summation phase is written for sums whose terms come from OTHER threads: positions and pair outputs go through shared
// memory, what a thread adds up is a pre-decoded record per round, the norms of the step-size estimate are shared-memory
// atomics, and two warp synchronisations separate the phases of every order pair: 80 % of the instructions of a
// two-body step. With one pair interaction per lane every "sum" is a single pair output (or a number: the accelerations
// of a body that only massless bodies pull on), so the thread that owns the lane keeps everything in registers:
//   * the six positions of the current order pair, the outputs m_k / n_k of the pair interaction (nb_core.hpp's
//     pair_block() with a register policy), v^[n+1] = a^[n] / (n + 1), x^[n+2] = v^[n+1] / (n + 2) in straight-line code;
//   * the three infinity norms of the step-size estimate (NaN-skipping maxima, like nb_step_size());
//   * its history rows d_0, d_1 in shared memory and r^2, d_2, r^alpha in tensor memory, exactly as in k_nb;
//   * the state variables' coefficients go to the private per-warp store (or to the public tc array), 256-byte rows.
// No __syncwarp() inside a step. Same arithmetic, same order of operations as k_nb / k_coop: bit-identical results
// (tests/test_gpu_parity.py runs both on the same inputs).
// Replaces, for these programs: the JIT'd step function (src/taylor_00.cpp:712-865) and the propagate loop
// (src/taylor_adaptive_batch.cpp:1136-1534).
#ifndef HEYOKA_B200_CSRC_NB1_KERNEL_CUH
#define HEYOKA_B200_CSRC_NB1_KERNEL_CUH

#include <cstdint>

#include <cuda_runtime.h>

#include "nb_kernel.cuh"

namespace heyoka_b200::dev
{

namespace nbk
{

// pair_block()'s storage policy for a thread that owns its lane: positions in, pair outputs out are registers; the
// private history rows are those of pair_mem<32, TMEM>.
template <bool TMEM>
struct pair_mem1 : pair_mem<32, TMEM> {
    d2 xa[3], xb[3];   // (x^[n], x^[n+1]) of the two bodies
    d2 om_[3], on_[3]; // (m_k^[n], m_k^[n+1]), (n_k^[n], n_k^[n+1])

    __device__ __forceinline__ d2 pos_a(int k) const
    {
        return xa[k];
    }
    __device__ __forceinline__ d2 pos_b(int k) const
    {
        return xb[k];
    }
    __device__ __forceinline__ void out(int k, const d2 &v)
    {
        om_[k] = v;
    }
    __device__ __forceinline__ void out_n(int k, const d2 &v)
    {
        on_[k] = v;
    }
};

} // namespace nbk

// NaN-skipping running maximum of |v| (fmax() returns its other argument for a NaN).
__device__ __forceinline__ void nb1_track(double &m, double v)
{
    m = fmax(m, fabs(v));
}

template <bool TMEM, bool PROP, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_nb1(program P, nb_dev_plan NP, batch D, run_args R)
{
    using nb::d2;
    extern __shared__ __align__(16) double smem_raw[];

    // ---- CTA-shared tables: fac | rcp ----
    const std::uint32_t p = P.order;
    double *fac_s = smem_raw;
    const std::uint32_t n_fac = (p + 1u) * NP.fac_stride;
    double *rcp_s = fac_s + n_fac;
    const std::uint32_t n_rcp = (p + 5u) & ~1u;
    for (std::uint32_t i = threadIdx.x; i < n_fac; i += blockDim.x) {
        fac_s[i] = __ldg(NP.fac + i);
    }
    for (std::uint32_t i = threadIdx.x; i < n_rcp; i += blockDim.x) {
        rcp_s[i] = i == 0u ? 0. : 1. / static_cast<double>(i);
    }
    __shared__ std::uint32_t tm_base_smem;
    if constexpr (TMEM) {
        if ((threadIdx.x >> 5) == 0u) {
            tm::alloc_all(&tm_base_smem);
        }
        tm::fence_before_sync();
    }
    __syncthreads();

    const std::uint32_t tid = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    double *region = smem_raw + NP.shared_doubles + static_cast<std::size_t>(warp) * NP.team_doubles;
    using PM_t = nbk::pair_mem1<TMEM>;
    PM_t PM;
    nb::pair_consts PC;
    {
        const uint4 *dp = reinterpret_cast<const uint4 *>(NP.pairs);
        const uint4 w1 = __ldg(dp + 1), w2 = __ldg(dp + 2), w3 = __ldg(dp + 3);
        PC.c1 = __hiloint2double(static_cast<int>(w2.y), static_cast<int>(w2.x));
        PC.c2[0] = __hiloint2double(static_cast<int>(w2.w), static_cast<int>(w2.z));
        PC.c2[1] = __hiloint2double(static_cast<int>(w3.y), static_cast<int>(w3.x));
        PC.c2[2] = __hiloint2double(static_cast<int>(w3.w), static_cast<int>(w3.z));
        PC.alpha = NP.alpha;
        PC.pow_algo = NP.pow_algo;
        PC.have_n = (w1.z & 1u) != 0u;
        PM.drow = nbk::saddr(region) + tid * 16u;
        PM.fac_ = nbk::saddr(fac_s);
        PM.fac_stride_b = NP.fac_stride * 8u;
        PM.flags = 0u;
        PM.tmc = 0u;
        if constexpr (TMEM) {
            tm::fence_after_sync();
            PM.tmc = tm_base_smem + (((warp & 3u) * 32u) << 16) + (warp >> 2) * (NP.npp * 12u);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            PM.on_[k] = d2{0., 0.};
        }
    }
    std::uint32_t rcp_a = nbk::saddr(rcp_s);
    nbk::keep(rcp_a);
    nbk::keep(PM.fac_);
    nbk::keep(PM.drow);
    lane_prop *const park
        = reinterpret_cast<lane_prop *>(region + static_cast<std::size_t>(NP.npp) * PM_t::OPB / 8u) + tid;
    static_assert(sizeof(lane_prop) <= 128u && alignof(lane_prop) <= 8u);

    const std::size_t team_global = (static_cast<std::size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const bool pub = R.coef_pub != 0;
    const bool mask_idle = !PROP && R.skip != nullptr;
    double *const cstore = R.coef_base + team_global * R.coef_warp_stride;
    const std::size_t stride_sv = static_cast<std::size_t>(R.coef_stride_sv), stride_o = static_cast<std::size_t>(R.coef_stride_o);
    const std::uint32_t n_chunks = (D.n + 31u) / 32u;
    const std::uint32_t n_blocks = NP.npp;
    const nb1_tab &TB = NP.l1;

    // The jet of this thread's lane (glane: clamped global lane; ok: the lane may write to the public store).
    // Returns the step-size norms through m0 / mp / mp1.
    const auto jet = [&](std::uint32_t glane, bool ok, double &m0, double &mp, double &mp1) {
        double *const cb = cstore + (pub ? static_cast<std::size_t>(glane) : static_cast<std::size_t>(tid));
        const auto st = [&](std::uint32_t sv, std::uint32_t order, double v) {
            if (pub) {
                if (ok) {
                    cb[sv * stride_sv + order * stride_o] = v;
                }
            } else {
                cb[sv * static_cast<std::uint32_t>(stride_sv) + order * static_cast<std::uint32_t>(stride_o)] = v;
            }
        };
        m0 = 0., mp = 0., mp1 = 0.;
        // nb_step_size()'s bookkeeping: order 0 -> m0, order p -> mp, order p - 1 -> mp1 (orders beyond p: dropped).
        const auto trk = [&](std::uint32_t order, double v) {
            if (order == 0u) {
                nb1_track(m0, v);
            } else if (order == p) {
                nb1_track(mp, v);
            } else if (order + 1u == p) {
                nb1_track(mp1, v);
            }
        };
        // Order 0 (and order 1 of the positions).
#pragma unroll
        for (int s = 0; s < 6; ++s) {
            const double x0 = D.state[static_cast<std::size_t>(TB.x_sv[s]) * D.n + glane];
            const double v0 = D.state[static_cast<std::size_t>(TB.v_sv[s]) * D.n + glane];
            st(TB.v_sv[s], 0u, v0);
            st(TB.x_sv[s], 0u, x0);
            st(TB.x_sv[s], 1u, v0);
            trk(0u, v0);
            trk(0u, x0);
            trk(1u, v0);
            (s < 3 ? PM.xa[s % 3] : PM.xb[s % 3]) = d2{x0, v0};
        }
        for (std::uint32_t m = 0; m < n_blocks; ++m) {
            __syncwarp(); // (the tensor-memory accesses of pair_block() are warp-wide: converged)
            nb::pair_block(PM, PC, m);
            if constexpr (TMEM) {
                tm::wait_st();
            }
            const std::uint32_t n = 2u * m;
            const double n1 = static_cast<double>(n + 1u), n2 = static_cast<double>(n + 2u),
                         n3 = static_cast<double>(n + 3u);
            const double r1 = nbk::lds1(rcp_a + (n + 1u) * 8u), r2 = nbk::lds1(rcp_a + (n + 2u) * 8u),
                         r3 = nbk::lds1(rcp_a + (n + 3u) * 8u);
            const bool track = m + 2u >= n_blocks;
#pragma unroll
            for (int s = 0; s < 6; ++s) {
                const int k = s % 3;
                const std::uint32_t kind = TB.kind[s];
                d2 a;
                if (kind == 2u) {
                    a = d2{n == 0u ? TB.cval[s] : 0., 0.};
                } else {
                    a = kind == 1u ? PM.on_[k] : PM.om_[k];
                }
                double va, vb, xa, xb;
                if (kind == 2u && n > 0u) {
                    // A constant right-hand side: every coefficient beyond the first order is an exact zero.
                    va = vb = xa = xb = 0.;
                } else if (n + 3u <= 64u && nb::div_si_in_range2(a.x) && nb::div_si_in_range2(a.y)) {
                    va = nb::div_si_fast(a.x, n1, r1); // v^[n+1]
                    vb = nb::div_si_fast(a.y, n2, r2); // v^[n+2]
                    xa = nb::div_si_fast(va, n2, r2);  // x^[n+2]
                    xb = nb::div_si_fast(vb, n3, r3);  // x^[n+3]
                } else {
                    va = a.x == 0. ? a.x : nb::div_cold(a.x, n1);
                    vb = a.y == 0. ? a.y : nb::div_cold(a.y, n2);
                    xa = va == 0. ? va : nb::div_cold(va, n2);
                    xb = vb == 0. ? vb : nb::div_cold(vb, n3);
                }
                const std::uint32_t vs = TB.v_sv[s], xs = TB.x_sv[s];
                st(vs, n + 1u, va);
                if (n + 2u <= p) {
                    st(vs, n + 2u, vb);
                    st(xs, n + 2u, xa);
                }
                if (n + 3u <= p) {
                    st(xs, n + 3u, xb);
                }
                if (track) {
                    trk(n + 1u, va);
                    trk(n + 2u, vb);
                    trk(n + 2u, xa);
                    trk(n + 3u, xb);
                }
                (s < 3 ? PM.xa[k] : PM.xb[k]) = d2{xa, xb};
            }
        }
    };

    // Step size from the norms and the coefficients of the first state variable (nb_step_size()'s semantics).
    const auto step_size = [&](std::uint32_t glane, double m0, double mp, double mp1, double max_delta_t) {
        const double *c = cstore + (pub ? static_cast<std::size_t>(glane) : static_cast<std::size_t>(tid));
        const double f0 = fabs(c[0]), fp = fabs(c[p * stride_o]), fp1 = fabs(c[(p - 1u) * stride_o]);
        return h_from_norms(P, isnan(f0) ? f0 : m0, isnan(fp) ? fp : mp, isnan(fp1) ? fp1 : mp1, max_delta_t);
    };
    // State update of the lane; returns true if a non-finite value was produced.
    const auto update = [&](std::uint32_t glane, bool write, double h) {
        constexpr int K = 4;
        const double *c0 = cstore + (pub ? static_cast<std::size_t>(glane) : static_cast<std::size_t>(tid));
        bool nf = false;
        for (std::uint32_t sv = 0; sv < P.n_eq; sv += K) {
            const double *c[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                c[k] = c0 + (sv + k < P.n_eq ? sv + k : sv) * stride_sv;
            }
            double res[K];
            eval_poly_k<K>(P, c, stride_o, h, res);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (write && sv + k < P.n_eq) {
                    D.state[static_cast<std::size_t>(sv + k) * D.n + glane] = res[k];
                    nf = nf || !isfinite(res[k]);
                }
            }
        }
        return nf;
    };

    for (std::uint32_t chunk = claim_chunk_warp(R.counter); chunk < n_chunks; chunk = claim_chunk_warp(R.counter)) {
        const std::uint32_t lane_raw = chunk * 32u + tid;
        bool valid = lane_raw < D.n;
        const std::uint32_t lane = valid ? lane_raw : D.n - 1u;

        if constexpr (!PROP) {
            const bool skipped = R.skip != nullptr && R.skip[lane] != 0u;
            valid = valid && !skipped;
            double m0, mp, mp1;
            jet(lane, lane_raw < D.n && !(mask_idle && skipped), m0, mp, mp1);
            const double mdt = R.max_delta_t != nullptr ? R.max_delta_t[lane] : R.default_max_delta_t;
            const double h = step_size(lane, m0, mp, mp1, mdt);
            const bool state_nf = update(lane, valid, h);
            if (valid) {
                const dfl nt = dfl_add(dfl{D.t_hi[lane], D.t_lo[lane]}, dfl{h, 0.});
                D.t_hi[lane] = nt.hi;
                D.t_lo[lane] = nt.lo;
                D.last_h[lane] = h;
                const bool nf = !(isfinite(nt.hi) && isfinite(nt.lo)) || state_nf;
                D.step_outcome[lane]
                    = nf ? HY_OUTCOME_ERR_NF_STATE : (h == mdt ? HY_OUTCOME_TIME_LIMIT : HY_OUTCOME_SUCCESS);
            }
        } else {
            bool running;
            {
                lane_prop lp;
                lp.init(D, R, lane);
                *park = lp;
                running = lp.running;
            }
            while (__any_sync(0xffffffffu, running)) {
                double m0, mp, mp1;
                jet(lane, lane_raw < D.n, m0, mp, mp1);
                const double cur_max = park->cur_max();
                const double h = step_size(lane, m0, mp, mp1, cur_max);
                const bool state_nf = update(lane, valid && running, h);
                if (running) {
                    lane_prop lp = *park;
                    lp.advance(h, cur_max, state_nf, R, valid);
                    *park = lp;
                    running = lp.running;
                }
            }
            if (valid) {
                park->store(D, lane);
                park->report_iters(R);
            }
        }
        __syncwarp();
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
