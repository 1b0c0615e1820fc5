// C ABI, device part (include/heyoka_b200.h section C): device-resident batch state and the
// persistent step / propagate kernels for sm_100a.
//
// Execution model ("G1"): one thread owns one lane (one ODE system of the batch); a warp
// therefore owns 32 consecutive lanes, and every global access of the warp — state, parameters, the
// derivative tape — is one coalesced 256-byte row. Warps are persistent: each one repeatedly claims a
// chunk of 32 lanes from an atomic counter, runs those lanes to completion (one step, or the whole
// propagate_until() loop), and moves on. All lanes of a warp execute the same opcode program, so the
// interpreter's control flow is warp-uniform. The tape of a warp (n_uvars * (order + 1) rows of 32
// doubles) lives in a per-warp slab of a scratch buffer in HBM and is re-used for every chunk.
//
// Replaces: the JIT'd step function (src/taylor_00.cpp:712-865), step_impl()
// (src/taylor_adaptive_batch.cpp:632-727), propagate_until_impl() (:1136-1534), d_out_f
// (src/taylor_01.cpp:1015-1185).
#include <heyoka_b200.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <limits>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <math_constants.h>

#include "capi_common.hpp"
#include "device_program.cuh"
#include "program.hpp"
#include "recurrences.cuh"

namespace hy = heyoka_b200;
using hy::detail::cuda_error;
using hy::detail::translate_exception;

#define HY_CUDA_CHECK(expr)                                                                                            \
    do {                                                                                                               \
        const cudaError_t err_ = (expr);                                                                               \
        if (err_ != cudaSuccess) {                                                                                     \
            throw cuda_error(std::string("CUDA error in " #expr ": ") + cudaGetErrorString(err_));                     \
        }                                                                                                              \
    } while (0)

namespace heyoka_b200::dev
{

// A warp's private view of its tape slab: row `slot` holds the 32 lanes' values of one coefficient.
struct warp_tape {
    double *base; // slab + lane-in-warp
    __device__ __forceinline__ double ld(std::uint32_t slot) const
    {
        return base[static_cast<std::size_t>(slot) * 32u];
    }
    __device__ __forceinline__ void st(std::uint32_t slot, double v) const
    {
        base[static_cast<std::size_t>(slot) * 32u] = v;
    }
};

__device__ __forceinline__ bool lane_state_nonfinite(const program &P, const batch &D, std::uint32_t lane)
{
    bool nf = false;
    for (std::uint32_t i = 0; i < P.n_eq; ++i) {
        nf = nf || !isfinite(D.state[static_cast<std::size_t>(i) * D.n + lane]);
    }
    return nf;
}

__device__ __forceinline__ std::uint32_t claim_chunk(unsigned int *counter)
{
    unsigned int c = 0;
    if ((threadIdx.x & 31u) == 0u) {
        c = atomicAdd(counter, 1u);
    }
    return __shfl_sync(0xffffffffu, c, 0);
}

// ------------------------------------------------------------------------------------------------
// One step for every lane: step()/step_backward()/step(max_delta_ts) + the bookkeeping of step_impl().
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_step(program P, batch D, double *scratch, std::size_t slab_doubles, unsigned int *counter,
           const double *max_delta_t, double default_max_delta_t, int write_tc)
{
    const std::uint32_t lane_in_warp = threadIdx.x & 31u;
    const std::size_t warp_global = (static_cast<std::size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const warp_tape tape{scratch + warp_global * slab_doubles + lane_in_warp};
    const std::uint32_t n_chunks = (D.n + 31u) / 32u;

    for (std::uint32_t chunk = claim_chunk(counter); chunk < n_chunks; chunk = claim_chunk(counter)) {
        const std::uint32_t lane_raw = chunk * 32u + lane_in_warp;
        const bool valid = lane_raw < D.n;
        const std::uint32_t lane = valid ? lane_raw : D.n - 1u;

        const double mdt = max_delta_t != nullptr ? max_delta_t[lane] : default_max_delta_t;
        const dfl t0{D.t_hi[lane], D.t_lo[lane]};
        const lane_ctx c{lane, D.n, D.pars, t0.hi};

        compute_jet(P, c, tape, D.state);
        const double h = determine_h(P, tape, mdt);
        update_state(P, c, tape, h, D.state, write_tc ? D.tc : nullptr, valid);

        if (valid) {
            const dfl nt = dfl_add(t0, dfl{h, 0.});
            D.t_hi[lane] = nt.hi;
            D.t_lo[lane] = nt.lo;
            D.last_h[lane] = h;
            const bool nf = !(isfinite(nt.hi) && isfinite(nt.lo)) || lane_state_nonfinite(P, D, lane);
            D.step_outcome[lane]
                = nf ? HY_OUTCOME_ERR_NF_STATE : (h == mdt ? HY_OUTCOME_TIME_LIMIT : HY_OUTCOME_SUCCESS);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// propagate_until(): every lane loops to its own final time (src/taylor_adaptive_batch.cpp:1372-1527,
// per-lane part). iter_cap == 0: unlimited. replay != 0: the cap reproduces a global early exit, lanes
// that hit it keep the outcome of their last step instead of step_limit.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_propagate(program P, batch D, double *scratch, std::size_t slab_doubles, unsigned int *counter,
                const double *tf_hi, const double *tf_lo, const double *max_delta_t, unsigned long long iter_cap,
                int replay, int write_tc, run_flags *flags)
{
    const std::uint32_t lane_in_warp = threadIdx.x & 31u;
    const std::size_t warp_global = (static_cast<std::size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const warp_tape tape{scratch + warp_global * slab_doubles + lane_in_warp};
    const std::uint32_t n_chunks = (D.n + 31u) / 32u;

    for (std::uint32_t chunk = claim_chunk(counter); chunk < n_chunks; chunk = claim_chunk(counter)) {
        const std::uint32_t lane_raw = chunk * 32u + lane_in_warp;
        const bool valid = lane_raw < D.n;
        const std::uint32_t lane = valid ? lane_raw : D.n - 1u;

        const dfl tf{tf_hi[lane], tf_lo != nullptr ? tf_lo[lane] : 0.};
        const double mdt = max_delta_t != nullptr ? max_delta_t[lane] : CUDART_INF;
        dfl t{D.t_hi[lane], D.t_lo[lane]};
        dfl rem = dfl_sub(tf, t);
        // Integration direction, fixed at the start (src/taylor_adaptive_batch.cpp:1265-1273).
        const bool dir = dfl_ge0(rem);

        unsigned long long ts_count = 0, iter = 0;
        double min_h = CUDART_INF, max_h = 0., last_h = 0.;
        long long outcome = HY_OUTCOME_TIME_LIMIT;
        bool running = true;

        while (__any_sync(0xffffffffu, running)) {
            // Time limit of this step (src/taylor_adaptive_batch.cpp:1378-1387). A lane that is not
            // running takes a zero-length step: the jet is computed (keeps the warp converged) but
            // nothing is written.
            const dfl lim = dir ? (dfl_lt(rem, dfl{mdt, 0.}) ? rem : dfl{mdt, 0.})
                                : (dfl_lt(rem, dfl{-mdt, 0.}) ? dfl{-mdt, 0.} : rem);
            const double cur_max = running ? lim.hi : 0.;

            const lane_ctx c{lane, D.n, D.pars, t.hi};
            compute_jet(P, c, tape, D.state);
            const double h = determine_h(P, tape, cur_max);
            const bool wr = running && valid;
            update_state(P, c, tape, h, D.state, write_tc ? D.tc : nullptr, wr);

            if (running) {
                t = dfl_add(t, dfl{h, 0.});
                last_h = h;
                ++iter;
                const bool nf = !(isfinite(t.hi) && isfinite(t.lo)) || lane_state_nonfinite(P, D, lane);
                if (nf) {
                    outcome = HY_OUTCOME_ERR_NF_STATE;
                    running = false;
                    if (valid) {
                        atomicOr(&flags->any_nf, 1u);
                        atomicMin(&flags->min_nf_iter, iter);
                    }
                } else {
                    const bool time_limit = (h == cur_max);
                    outcome = time_limit ? HY_OUTCOME_TIME_LIMIT : HY_OUTCOME_SUCCESS;
                    ts_count += (h != 0.) ? 1u : 0u;
                    if (!time_limit) {
                        const double ah = fabs(h);
                        min_h = fmin(min_h, ah);
                        max_h = fmax(max_h, ah);
                    }
                    if (h == rem.hi) {
                        // Final time reached (the outcome is necessarily time_limit).
                        rem = dfl{0., 0.};
                        running = false;
                    } else {
                        rem = dfl_sub(tf, t);
                        if (iter == iter_cap) {
                            running = false;
                            if (!replay) {
                                outcome = HY_OUTCOME_STEP_LIMIT;
                                if (valid) {
                                    atomicOr(&flags->any_limit, 1u);
                                }
                            }
                        }
                    }
                }
            }
        }

        if (valid) {
            D.t_hi[lane] = t.hi;
            D.t_lo[lane] = t.lo;
            D.last_h[lane] = last_h;
            D.prop_outcome[lane] = outcome;
            D.prop_min_h[lane] = min_h;
            D.prop_max_h[lane] = max_h;
            D.prop_n_steps[lane] = ts_count;
        }
    }
}

__global__ void k_fill_outcome(long long *out, std::uint32_t n, long long value)
{
    const std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        out[i] = value;
    }
}

// Dense output (src/taylor_01.cpp:1015-1185): Horner, or compensated summation in high-accuracy mode.
__global__ void k_d_output(std::uint32_t n_eq, std::uint32_t order, int high_accuracy, std::uint32_t n, const double *tc,
                           const double *tau, double *out)
{
    const std::uint32_t lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= n) {
        return;
    }
    const double h = tau[lane];
    for (std::uint32_t i = 0; i < n_eq; ++i) {
        const double *cf = tc + static_cast<std::size_t>(i) * (order + 1u) * n + lane;
        double res;
        if (!high_accuracy) {
            res = cf[static_cast<std::size_t>(order) * n];
            for (std::uint32_t o = 1; o <= order; ++o) {
                res = fma(res, h, cf[static_cast<std::size_t>(order - o) * n]);
            }
        } else {
            res = cf[0];
            double comp = 0., cur_h = h;
            for (std::uint32_t o = 1; o <= order; ++o) {
                const double tmp = __dmul_rn(cf[static_cast<std::size_t>(o) * n], cur_h);
                const double y = __dsub_rn(tmp, comp);
                const double tt = __dadd_rn(res, y);
                comp = __dsub_rn(__dsub_rn(tt, res), y);
                res = tt;
                cur_h = __dmul_rn(cur_h, h);
            }
        }
        out[static_cast<std::size_t>(i) * n + lane] = res;
    }
}

} // namespace heyoka_b200::dev

namespace dev = heyoka_b200::dev;

// ------------------------------------------------------------------------------------------------
// Host object.
// ------------------------------------------------------------------------------------------------
struct hy_batch {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::uint32_t n = 0;
    std::uint32_t n_eq = 0, n_pars = 0, order = 0;
    bool high_accuracy = false;

    // Device copies of the program arrays.
    uint4 *d_ops = nullptr;
    std::uint32_t *d_args = nullptr, *d_sv_defs = nullptr;
    double *d_consts = nullptr;
    dev::program prog{};

    // Resident arrays.
    double *d_state = nullptr, *d_pars = nullptr, *d_t_hi = nullptr, *d_t_lo = nullptr, *d_last_h = nullptr,
           *d_tc = nullptr, *d_d_out = nullptr;
    long long *d_step_outcome = nullptr, *d_prop_outcome = nullptr;
    double *d_prop_min_h = nullptr, *d_prop_max_h = nullptr;
    unsigned long long *d_prop_n_steps = nullptr;

    // Scratch.
    double *d_scratch = nullptr; // per-warp tape slabs
    std::size_t slab_doubles = 0;
    double *d_tmp = nullptr;      // 3 * n doubles: staged per-lane inputs (t_final hi/lo, max_delta_t)
    double *d_snapshot = nullptr; // state + time snapshot for the global-exit replay
    unsigned int *d_counter = nullptr;
    dev::run_flags *d_flags = nullptr;

    // Launch geometry.
    std::uint32_t block_threads = 256, blocks_per_sm = 0, n_sms = 0, grid = 0;
    std::uint64_t n_launches = 0;

    ~hy_batch();
    void free_all() noexcept;
    void alloc_scratch();
    dev::batch view() const;
    template <typename T>
    T *dalloc(std::size_t count);
};

template <typename T>
T *hy_batch::dalloc(std::size_t count)
{
    void *p = nullptr;
    HY_CUDA_CHECK(cudaMalloc(&p, std::max<std::size_t>(count, 1u) * sizeof(T)));
    return static_cast<T *>(p);
}

void hy_batch::free_all() noexcept
{
    for (void *p : {static_cast<void *>(d_ops), static_cast<void *>(d_args), static_cast<void *>(d_sv_defs),
                    static_cast<void *>(d_consts), static_cast<void *>(d_state), static_cast<void *>(d_pars),
                    static_cast<void *>(d_t_hi), static_cast<void *>(d_t_lo), static_cast<void *>(d_last_h),
                    static_cast<void *>(d_tc), static_cast<void *>(d_d_out), static_cast<void *>(d_step_outcome),
                    static_cast<void *>(d_prop_outcome), static_cast<void *>(d_prop_min_h),
                    static_cast<void *>(d_prop_max_h), static_cast<void *>(d_prop_n_steps),
                    static_cast<void *>(d_scratch), static_cast<void *>(d_tmp), static_cast<void *>(d_snapshot),
                    static_cast<void *>(d_counter), static_cast<void *>(d_flags)}) {
        if (p != nullptr) {
            cudaFree(p);
        }
    }
}

hy_batch::~hy_batch()
{
    int cur = 0;
    if (cudaGetDevice(&cur) == cudaSuccess) {
        cudaSetDevice(device);
        free_all();
        cudaSetDevice(cur);
    }
}

dev::batch hy_batch::view() const
{
    dev::batch b{};
    b.n = n;
    b.state = d_state;
    b.t_hi = d_t_hi;
    b.t_lo = d_t_lo;
    b.last_h = d_last_h;
    b.tc = d_tc;
    b.pars = d_pars;
    b.step_outcome = d_step_outcome;
    b.prop_outcome = d_prop_outcome;
    b.prop_min_h = d_prop_min_h;
    b.prop_max_h = d_prop_max_h;
    b.prop_n_steps = d_prop_n_steps;
    return b;
}

void hy_batch::alloc_scratch()
{
    // One slab per resident warp; never more warps than chunks of 32 lanes.
    const std::uint32_t warps_per_block = block_threads / 32u;
    const std::uint32_t n_chunks = (n + 31u) / 32u;
    std::uint32_t blocks = n_sms * blocks_per_sm;
    const std::uint32_t needed_blocks = (n_chunks + warps_per_block - 1u) / warps_per_block;
    blocks = std::max(1u, std::min(blocks, needed_blocks));
    grid = blocks;

    if (d_scratch != nullptr) {
        HY_CUDA_CHECK(cudaFree(d_scratch));
        d_scratch = nullptr;
    }
    const std::size_t n_warps = static_cast<std::size_t>(blocks) * warps_per_block;
    d_scratch = dalloc<double>(n_warps * slab_doubles);
}

namespace
{

struct device_guard {
    int prev = 0;
    explicit device_guard(int dev)
    {
        HY_CUDA_CHECK(cudaGetDevice(&prev));
        if (prev != dev) {
            HY_CUDA_CHECK(cudaSetDevice(dev));
        }
    }
    ~device_guard()
    {
        cudaSetDevice(prev);
    }
};

void launch_reset(hy_batch *b)
{
    HY_CUDA_CHECK(cudaMemsetAsync(b->d_counter, 0, sizeof(unsigned int), b->stream));
}

// Stage a host (or device) array of n doubles into slot `slot` of d_tmp; returns the device pointer.
const double *stage(hy_batch *b, const double *src, int on_device, std::uint32_t slot)
{
    if (src == nullptr) {
        return nullptr;
    }
    if (on_device) {
        return src;
    }
    double *dst = b->d_tmp + static_cast<std::size_t>(slot) * b->n;
    HY_CUDA_CHECK(cudaMemcpyAsync(dst, src, sizeof(double) * b->n, cudaMemcpyHostToDevice, b->stream));
    return dst;
}

void run_propagate(hy_batch *b, const double *d_tf_hi, const double *d_tf_lo, const double *d_mdt,
                   unsigned long long iter_cap, int replay, int write_tc)
{
    launch_reset(b);
    dev::k_propagate<<<b->grid, b->block_threads, 0, b->stream>>>(b->prog, b->view(), b->d_scratch, b->slab_doubles,
                                                                  b->d_counter, d_tf_hi, d_tf_lo, d_mdt, iter_cap,
                                                                  replay, write_tc, b->d_flags);
    HY_CUDA_CHECK(cudaGetLastError());
    ++b->n_launches;
}

int propagate_impl(hy_batch *b, const double *d_tf_hi, const double *d_tf_lo, const double *d_mdt, uint64_t max_steps,
                   int write_tc, int *any_flag)
{
    const std::size_t state_doubles = static_cast<std::size_t>(b->n_eq) * b->n;

    // Snapshot (state, t_hi, t_lo) so that a global early exit can be replayed exactly.
    HY_CUDA_CHECK(cudaMemcpyAsync(b->d_snapshot, b->d_state, sizeof(double) * state_doubles, cudaMemcpyDeviceToDevice,
                                  b->stream));
    HY_CUDA_CHECK(cudaMemcpyAsync(b->d_snapshot + state_doubles, b->d_t_hi, sizeof(double) * b->n,
                                  cudaMemcpyDeviceToDevice, b->stream));
    HY_CUDA_CHECK(cudaMemcpyAsync(b->d_snapshot + state_doubles + b->n, b->d_t_lo, sizeof(double) * b->n,
                                  cudaMemcpyDeviceToDevice, b->stream));

    const dev::run_flags init{0u, 0u, ~0ull};
    HY_CUDA_CHECK(cudaMemcpyAsync(b->d_flags, &init, sizeof(init), cudaMemcpyHostToDevice, b->stream));

    run_propagate(b, d_tf_hi, d_tf_lo, d_mdt, max_steps, 0, write_tc);

    dev::run_flags fl{};
    HY_CUDA_CHECK(cudaMemcpyAsync(&fl, b->d_flags, sizeof(fl), cudaMemcpyDeviceToHost, b->stream));
    HY_CUDA_CHECK(cudaStreamSynchronize(b->stream));

    if (fl.any_nf != 0u) {
        // The reference stops EVERY lane at the first iteration in which any lane goes non-finite
        // (src/taylor_adaptive_batch.cpp:1462-1467). Lanes are independent, so re-running from the
        // snapshot with the iteration count capped at that index reproduces it exactly (the index is
        // never beyond max_steps, because the first run was capped there).
        const unsigned long long cap = fl.min_nf_iter;
        HY_CUDA_CHECK(cudaMemcpyAsync(b->d_state, b->d_snapshot, sizeof(double) * state_doubles,
                                      cudaMemcpyDeviceToDevice, b->stream));
        HY_CUDA_CHECK(cudaMemcpyAsync(b->d_t_hi, b->d_snapshot + state_doubles, sizeof(double) * b->n,
                                      cudaMemcpyDeviceToDevice, b->stream));
        HY_CUDA_CHECK(cudaMemcpyAsync(b->d_t_lo, b->d_snapshot + state_doubles + b->n, sizeof(double) * b->n,
                                      cudaMemcpyDeviceToDevice, b->stream));
        HY_CUDA_CHECK(cudaMemcpyAsync(b->d_flags, &init, sizeof(init), cudaMemcpyHostToDevice, b->stream));
        run_propagate(b, d_tf_hi, d_tf_lo, d_mdt, cap, 1, write_tc);
        HY_CUDA_CHECK(cudaMemcpyAsync(&fl, b->d_flags, sizeof(fl), cudaMemcpyDeviceToHost, b->stream));
        HY_CUDA_CHECK(cudaStreamSynchronize(b->stream));
    }

    if (fl.any_nf == 0u && fl.any_limit != 0u) {
        // Iteration limit: every lane reports step_limit (src/taylor_adaptive_batch.cpp:1516-1526).
        dev::k_fill_outcome<<<(b->n + 255u) / 256u, 256, 0, b->stream>>>(b->d_prop_outcome, b->n,
                                                                         HY_OUTCOME_STEP_LIMIT);
        HY_CUDA_CHECK(cudaGetLastError());
        ++b->n_launches;
    }

    if (any_flag != nullptr) {
        *any_flag = (fl.any_nf != 0u ? 1 : 0) | (fl.any_limit != 0u ? 2 : 0);
    }
    return HY_OK;
}

} // namespace

extern "C" {

int hy_batch_create(const hy_program *p, uint32_t batch, int device, hy_batch **out)
{
    hy_batch *b = nullptr;
    try {
        if (p == nullptr || out == nullptr) {
            throw std::invalid_argument("Null pointer passed to hy_batch_create()");
        }
        if (batch == 0u) {
            throw std::invalid_argument("The batch size in an adaptive Taylor integrator cannot be zero");
        }

        int n_dev = 0;
        if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
            throw cuda_error("No usable CUDA device: heyoka_b200 has no CPU fallback");
        }
        if (device < 0) {
            HY_CUDA_CHECK(cudaGetDevice(&device));
        }
        if (device >= n_dev) {
            throw std::invalid_argument("Invalid CUDA device index " + std::to_string(device));
        }

        // Overflow checks on the buffer sizes, like src/taylor_adaptive_batch.cpp:256-264,375-378.
        const std::uint64_t tc_size = static_cast<std::uint64_t>(p->n_eq) * (p->order + 1u) * batch;
        if (tc_size > (std::numeric_limits<std::uint64_t>::max() >> 4)) {
            throw std::overflow_error("Overflow detected while computing the size of the Taylor coefficients buffer");
        }

        b = new hy_batch;
        b->device = device;
        device_guard guard(device);

        b->n = batch;
        b->n_eq = p->n_eq;
        b->n_pars = p->n_pars;
        b->order = p->order;
        b->high_accuracy = p->high_accuracy;

        cudaDeviceProp prop{};
        HY_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        b->n_sms = static_cast<std::uint32_t>(prop.multiProcessorCount);

        // Program arrays.
        static_assert(sizeof(hy_op) == sizeof(uint4), "hy_op must be 16 bytes");
        b->d_ops = b->dalloc<uint4>(p->ops.size());
        b->d_args = b->dalloc<std::uint32_t>(p->args.size());
        b->d_consts = b->dalloc<double>(p->consts.size());
        b->d_sv_defs = b->dalloc<std::uint32_t>(p->sv_defs.size());
        HY_CUDA_CHECK(cudaMemcpy(b->d_ops, p->ops.data(), p->ops.size() * sizeof(hy_op), cudaMemcpyHostToDevice));
        HY_CUDA_CHECK(cudaMemcpy(b->d_args, p->args.data(), p->args.size() * sizeof(std::uint32_t),
                                 cudaMemcpyHostToDevice));
        HY_CUDA_CHECK(cudaMemcpy(b->d_consts, p->consts.data(), p->consts.size() * sizeof(double),
                                 cudaMemcpyHostToDevice));
        HY_CUDA_CHECK(cudaMemcpy(b->d_sv_defs, p->sv_defs.data(), p->sv_defs.size() * sizeof(std::uint32_t),
                                 cudaMemcpyHostToDevice));

        auto &P = b->prog;
        P.n_eq = p->n_eq;
        P.n_uvars = p->n_uvars;
        P.n_pars = p->n_pars;
        P.order = p->order;
        P.n_ops = p->n_uvars - p->n_eq;
        P.high_accuracy = p->high_accuracy ? 1 : 0;
        // taylor_determine_h_rhofac(), src/taylor_00.cpp:84-94 (host libm, like the reference's number arithmetic).
        P.rhofac = std::exp((-7. / 10.) / static_cast<double>(p->order - 1u)) / (std::exp(1.) * std::exp(1.));
        P.inv_p = 1. / static_cast<double>(p->order);
        P.inv_pm1 = 1. / static_cast<double>(p->order - 1u);
        P.ops = b->d_ops;
        P.args = b->d_args;
        P.consts = b->d_consts;
        P.sv_defs = b->d_sv_defs;

        // Resident arrays.
        const std::size_t n = batch;
        b->d_state = b->dalloc<double>(n * p->n_eq);
        b->d_pars = b->dalloc<double>(n * p->n_pars);
        b->d_t_hi = b->dalloc<double>(n);
        b->d_t_lo = b->dalloc<double>(n);
        b->d_last_h = b->dalloc<double>(n);
        b->d_tc = b->dalloc<double>(tc_size);
        b->d_d_out = b->dalloc<double>(n * p->n_eq);
        b->d_step_outcome = b->dalloc<long long>(n);
        b->d_prop_outcome = b->dalloc<long long>(n);
        b->d_prop_min_h = b->dalloc<double>(n);
        b->d_prop_max_h = b->dalloc<double>(n);
        b->d_prop_n_steps = b->dalloc<unsigned long long>(n);
        b->d_tmp = b->dalloc<double>(3u * n);
        b->d_snapshot = b->dalloc<double>(n * (p->n_eq + 2u));
        b->d_counter = b->dalloc<unsigned int>(1);
        b->d_flags = b->dalloc<dev::run_flags>(1);

        HY_CUDA_CHECK(cudaMemset(b->d_state, 0, sizeof(double) * n * p->n_eq));
        HY_CUDA_CHECK(cudaMemset(b->d_pars, 0, sizeof(double) * std::max<std::size_t>(n * p->n_pars, 1u)));
        HY_CUDA_CHECK(cudaMemset(b->d_t_hi, 0, sizeof(double) * n));
        HY_CUDA_CHECK(cudaMemset(b->d_t_lo, 0, sizeof(double) * n));
        HY_CUDA_CHECK(cudaMemset(b->d_last_h, 0, sizeof(double) * n));
        HY_CUDA_CHECK(cudaMemset(b->d_tc, 0, sizeof(double) * tc_size));

        // Launch geometry: as many resident 256-thread blocks per SM as the kernel allows.
        b->slab_doubles = static_cast<std::size_t>(p->n_uvars) * (p->order + 1u) * 32u;
        int occ = 0;
        HY_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dev::k_propagate,
                                                                    static_cast<int>(b->block_threads), 0));
        b->blocks_per_sm = static_cast<std::uint32_t>(std::max(occ, 1));
        b->alloc_scratch();

        *out = b;
        return HY_OK;
    } catch (...) {
        delete b;
        return translate_exception();
    }
}

void hy_batch_destroy(hy_batch *b)
{
    delete b;
}

int hy_batch_set_stream(hy_batch *b, void *cuda_stream)
{
    if (b == nullptr) {
        hy::detail::set_last_error("Null batch");
        return HY_ERR_INVALID_ARG;
    }
    b->stream = static_cast<cudaStream_t>(cuda_stream);
    return HY_OK;
}

int hy_batch_sync(hy_batch *b)
{
    try {
        device_guard guard(b->device);
        HY_CUDA_CHECK(cudaStreamSynchronize(b->stream));
        return HY_OK;
    } catch (...) {
        return translate_exception();
    }
}

int hy_batch_set_launch_config(hy_batch *b, uint32_t block_threads, uint32_t blocks_per_sm)
{
    try {
        device_guard guard(b->device);
        if (block_threads != 0u) {
            if (block_threads % 32u != 0u || block_threads > 256u) {
                throw std::invalid_argument("block_threads must be a multiple of 32 not larger than 256");
            }
            b->block_threads = block_threads;
        }
        if (blocks_per_sm != 0u) {
            b->blocks_per_sm = blocks_per_sm;
        }
        HY_CUDA_CHECK(cudaStreamSynchronize(b->stream));
        b->alloc_scratch();
        return HY_OK;
    } catch (...) {
        return translate_exception();
    }
}

int hy_batch_upload(hy_batch *b, const double *state, const double *pars, const double *t_hi, const double *t_lo)
{
    try {
        device_guard guard(b->device);
        const std::size_t n = b->n;
        if (state != nullptr) {
            HY_CUDA_CHECK(cudaMemcpyAsync(b->d_state, state, sizeof(double) * n * b->n_eq, cudaMemcpyHostToDevice,
                                          b->stream));
        }
        if (pars != nullptr && b->n_pars > 0u) {
            HY_CUDA_CHECK(cudaMemcpyAsync(b->d_pars, pars, sizeof(double) * n * b->n_pars, cudaMemcpyHostToDevice,
                                          b->stream));
        }
        if (t_hi != nullptr) {
            HY_CUDA_CHECK(cudaMemcpyAsync(b->d_t_hi, t_hi, sizeof(double) * n, cudaMemcpyHostToDevice, b->stream));
        }
        if (t_lo != nullptr) {
            HY_CUDA_CHECK(cudaMemcpyAsync(b->d_t_lo, t_lo, sizeof(double) * n, cudaMemcpyHostToDevice, b->stream));
        }
        HY_CUDA_CHECK(cudaStreamSynchronize(b->stream));
        return HY_OK;
    } catch (...) {
        return translate_exception();
    }
}

int hy_batch_download(hy_batch *b, double *state, double *t_hi, double *t_lo, double *last_h)
{
    try {
        device_guard guard(b->device);
        const std::size_t n = b->n;
        if (state != nullptr) {
            HY_CUDA_CHECK(cudaMemcpyAsync(state, b->d_state, sizeof(double) * n * b->n_eq, cudaMemcpyDeviceToHost,
                                          b->stream));
        }
        if (t_hi != nullptr) {
            HY_CUDA_CHECK(cudaMemcpyAsync(t_hi, b->d_t_hi, sizeof(double) * n, cudaMemcpyDeviceToHost, b->stream));
        }
        if (t_lo != nullptr) {
            HY_CUDA_CHECK(cudaMemcpyAsync(t_lo, b->d_t_lo, sizeof(double) * n, cudaMemcpyDeviceToHost, b->stream));
        }
        if (last_h != nullptr) {
            HY_CUDA_CHECK(cudaMemcpyAsync(last_h, b->d_last_h, sizeof(double) * n, cudaMemcpyDeviceToHost, b->stream));
        }
        HY_CUDA_CHECK(cudaStreamSynchronize(b->stream));
        return HY_OK;
    } catch (...) {
        return translate_exception();
    }
}

int hy_batch_download_step_res(hy_batch *b, int64_t *outcome, double *h)
{
    try {
        device_guard guard(b->device);
        const std::size_t n = b->n;
        if (outcome != nullptr) {
            HY_CUDA_CHECK(cudaMemcpyAsync(outcome, b->d_step_outcome, sizeof(int64_t) * n, cudaMemcpyDeviceToHost,
                                          b->stream));
        }
        if (h != nullptr) {
            HY_CUDA_CHECK(cudaMemcpyAsync(h, b->d_last_h, sizeof(double) * n, cudaMemcpyDeviceToHost, b->stream));
        }
        HY_CUDA_CHECK(cudaStreamSynchronize(b->stream));
        return HY_OK;
    } catch (...) {
        return translate_exception();
    }
}

int hy_batch_download_prop_res(hy_batch *b, int64_t *outcome, double *min_h, double *max_h, uint64_t *n_steps)
{
    try {
        device_guard guard(b->device);
        const std::size_t n = b->n;
        if (outcome != nullptr) {
            HY_CUDA_CHECK(cudaMemcpyAsync(outcome, b->d_prop_outcome, sizeof(int64_t) * n, cudaMemcpyDeviceToHost,
                                          b->stream));
        }
        if (min_h != nullptr) {
            HY_CUDA_CHECK(cudaMemcpyAsync(min_h, b->d_prop_min_h, sizeof(double) * n, cudaMemcpyDeviceToHost,
                                          b->stream));
        }
        if (max_h != nullptr) {
            HY_CUDA_CHECK(cudaMemcpyAsync(max_h, b->d_prop_max_h, sizeof(double) * n, cudaMemcpyDeviceToHost,
                                          b->stream));
        }
        if (n_steps != nullptr) {
            HY_CUDA_CHECK(cudaMemcpyAsync(n_steps, b->d_prop_n_steps, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost,
                                          b->stream));
        }
        HY_CUDA_CHECK(cudaStreamSynchronize(b->stream));
        return HY_OK;
    } catch (...) {
        return translate_exception();
    }
}

int hy_batch_download_tc(hy_batch *b, double *tc)
{
    try {
        device_guard guard(b->device);
        const std::size_t sz = static_cast<std::size_t>(b->n_eq) * (b->order + 1u) * b->n;
        HY_CUDA_CHECK(cudaMemcpyAsync(tc, b->d_tc, sizeof(double) * sz, cudaMemcpyDeviceToHost, b->stream));
        HY_CUDA_CHECK(cudaStreamSynchronize(b->stream));
        return HY_OK;
    } catch (...) {
        return translate_exception();
    }
}

int hy_batch_get_ptrs(hy_batch *b, hy_batch_ptrs *out)
{
    if (b == nullptr || out == nullptr) {
        hy::detail::set_last_error("Null pointer passed to hy_batch_get_ptrs()");
        return HY_ERR_INVALID_ARG;
    }
    out->state = b->d_state;
    out->pars = b->d_pars;
    out->t_hi = b->d_t_hi;
    out->t_lo = b->d_t_lo;
    out->last_h = b->d_last_h;
    out->tc = b->d_tc;
    out->d_out = b->d_d_out;
    out->step_outcome = reinterpret_cast<int64_t *>(b->d_step_outcome);
    out->prop_outcome = reinterpret_cast<int64_t *>(b->d_prop_outcome);
    out->prop_min_h = b->d_prop_min_h;
    out->prop_max_h = b->d_prop_max_h;
    out->prop_n_steps = reinterpret_cast<uint64_t *>(b->d_prop_n_steps);
    return HY_OK;
}

int hy_batch_step(hy_batch *b, const double *max_delta_t, int on_device, int backward, int write_tc)
{
    try {
        device_guard guard(b->device);
        if (max_delta_t != nullptr && !on_device) {
            // step(max_delta_ts): NaN limits are rejected (src/taylor_adaptive_batch.cpp:1060-1075).
            for (std::uint32_t i = 0; i < b->n; ++i) {
                if (std::isnan(max_delta_t[i])) {
                    throw std::invalid_argument("Cannot use a nan max_delta_t in the step() function of an adaptive "
                                                "Taylor integrator in batch mode");
                }
            }
        }
        const double *d_mdt = stage(b, max_delta_t, on_device, 0);
        const double def = backward ? -std::numeric_limits<double>::infinity() : std::numeric_limits<double>::infinity();
        launch_reset(b);
        dev::k_step<<<b->grid, b->block_threads, 0, b->stream>>>(b->prog, b->view(), b->d_scratch, b->slab_doubles,
                                                                 b->d_counter, d_mdt, def, write_tc);
        HY_CUDA_CHECK(cudaGetLastError());
        ++b->n_launches;
        return HY_OK;
    } catch (...) {
        return translate_exception();
    }
}

int hy_batch_propagate_until(hy_batch *b, const double *t_final_hi, const double *t_final_lo, const double *max_delta_t,
                             uint64_t max_steps, int write_tc)
{
    try {
        device_guard guard(b->device);
        if (t_final_hi == nullptr) {
            throw std::invalid_argument("Null final times passed to hy_batch_propagate_until()");
        }
        // Argument checks of propagate_until_impl(), src/taylor_adaptive_batch.cpp:1212-1241.
        for (std::uint32_t i = 0; i < b->n; ++i) {
            if (!std::isfinite(t_final_hi[i]) || (t_final_lo != nullptr && !std::isfinite(t_final_lo[i]))) {
                throw std::invalid_argument("A non-finite time was passed to the propagate_until() function of an "
                                            "adaptive Taylor integrator in batch mode");
            }
            if (max_delta_t != nullptr) {
                if (std::isnan(max_delta_t[i])) {
                    throw std::invalid_argument("A nan max_delta_t was passed to the propagate_until() function of an "
                                                "adaptive Taylor integrator in batch mode");
                }
                if (max_delta_t[i] <= 0) {
                    throw std::invalid_argument("A non-positive max_delta_t was passed to the propagate_until() "
                                                "function of an adaptive Taylor integrator in batch mode");
                }
            }
        }
        const double *d_hi = stage(b, t_final_hi, 0, 0);
        const double *d_lo = stage(b, t_final_lo, 0, 1);
        const double *d_mdt = stage(b, max_delta_t, 0, 2);
        return propagate_impl(b, d_hi, d_lo, d_mdt, max_steps, write_tc, nullptr);
    } catch (...) {
        return translate_exception();
    }
}

int hy_batch_propagate_until_dev(hy_batch *b, const double *d_t_final_hi, const double *d_t_final_lo,
                                 const double *d_max_delta_t, uint64_t max_steps, int write_tc, int *any_nf_or_limit)
{
    try {
        device_guard guard(b->device);
        if (d_t_final_hi == nullptr) {
            throw std::invalid_argument("Null final times passed to hy_batch_propagate_until_dev()");
        }
        return propagate_impl(b, d_t_final_hi, d_t_final_lo, d_max_delta_t, max_steps, write_tc, any_nf_or_limit);
    } catch (...) {
        return translate_exception();
    }
}

int hy_batch_d_output(hy_batch *b, const double *tau, double *out)
{
    try {
        device_guard guard(b->device);
        const double *d_tau = stage(b, tau, 0, 0);
        dev::k_d_output<<<(b->n + 127u) / 128u, 128, 0, b->stream>>>(b->n_eq, b->order, b->high_accuracy ? 1 : 0, b->n,
                                                                     b->d_tc, d_tau, b->d_d_out);
        HY_CUDA_CHECK(cudaGetLastError());
        ++b->n_launches;
        if (out != nullptr) {
            HY_CUDA_CHECK(cudaMemcpyAsync(out, b->d_d_out, sizeof(double) * b->n * b->n_eq, cudaMemcpyDeviceToHost,
                                          b->stream));
        }
        HY_CUDA_CHECK(cudaStreamSynchronize(b->stream));
        return HY_OK;
    } catch (...) {
        return translate_exception();
    }
}

int hy_batch_launch_count(const hy_batch *b, uint64_t *n_launches)
{
    if (b == nullptr || n_launches == nullptr) {
        hy::detail::set_last_error("Null pointer passed to hy_batch_launch_count()");
        return HY_ERR_INVALID_ARG;
    }
    *n_launches = b->n_launches;
    return HY_OK;
}

} // extern "C"
