// taylor_adaptive_batch<double>: host-side mirror of the reference class on top of the C ABI.
// Reference: src/taylor_adaptive_batch.cpp (ctor :78-427, step :1039-1078, propagate :1081-1534,
// dense output :2251-2327, getters), include/heyoka/detail/dfloat.hpp.
#include <heyoka_b200/taylor.hpp>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <exception>
#include <limits>
#include <string>
#include <thread>

#include "program.hpp"

namespace heyoka_b200
{

namespace
{

// Double-length arithmetic (include/heyoka/detail/dfloat.hpp:104-169). volatile: no contraction / reassociation.
struct dfl {
    double hi, lo;
};
inline dfl eft_knuth(double a, double b)
{
    volatile double x = a + b;
    volatile double z = x - a;
    volatile double y = (a - (x - z)) + (b - z);
    return {x, y};
}
inline dfl eft_dekker(double a, double b)
{
    volatile double x = a + b;
    volatile double y = (a - x) + b;
    return {x, y};
}
inline dfl dfl_add(dfl a, dfl b)
{
    const dfl h = eft_knuth(a.hi, b.hi), l = eft_knuth(a.lo, b.lo);
    dfl uv = eft_dekker(h.hi, h.lo + l.hi);
    uv = eft_dekker(uv.hi, uv.lo + l.lo);
    return uv;
}
inline dfl dfl_sub(dfl a, dfl b)
{
    return dfl_add(a, dfl{-b.hi, -b.lo});
}
inline bool dfl_lt(dfl x, dfl y)
{
    return (x.hi < y.hi) || (x.hi == y.hi && x.lo < y.lo);
}
inline bool dfl_ge0(dfl x)
{
    return (x.hi > 0.) || (x.hi == 0. && x.lo >= 0.);
}

std::string fp_to_string(double x)
{
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%.17g", x);
    return buf;
}

[[noreturn]] void throw_from_status(int st)
{
    const std::string msg = hy_last_error();
    switch (st) {
        case HY_ERR_NOT_IMPLEMENTED:
            throw not_implemented_error(msg);
        case HY_ERR_OVERFLOW:
            throw std::overflow_error(msg);
        case HY_ERR_CUDA:
            throw std::runtime_error(msg);
        default:
            throw std::invalid_argument(msg);
    }
}

inline void check(int st)
{
    if (st != HY_OK) {
        throw_from_status(st);
    }
}

} // namespace

struct taylor_adaptive_batch<double>::impl {
    std::vector<std::pair<expression, expression>> sys;
    taylor_dc_t dc;
    // The lowered program is shared between copies, like the JIT-compiled code of the reference
    // (src/detail/i_data.cpp:335-352); every copy owns its device buffers.
    std::shared_ptr<hy_program> prog;
    hy_batch *batch = nullptr;
    std::uint32_t batch_size = 0, dim = 0, order = 0, n_pars = 0;
    double tol = 0;
    bool high_accuracy = false, compact_mode = false;
    int device = -1;
    std::vector<int> devices; // non-empty: sharded over these GPUs (hy_batch_create_multi())
    int tape_mode = 0;
    std::uint32_t k_lpw = 0, k_lpt = 0, k_threads = 0, k_bpsm = 0;

    std::vector<double> state, pars, time_hi, time_lo, tc, last_h, d_out;
    std::vector<std::tuple<taylor_outcome, double>> step_res;
    std::vector<std::tuple<taylor_outcome, double, double, std::size_t>> prop_res;
    // scratch
    std::vector<std::int64_t> oc;
    std::vector<double> tmp_a, tmp_b;
    std::vector<std::uint64_t> tmp_n;
    bool tc_valid = false; // the device holds the Taylor coefficients mirrored in `tc`

    impl() = default;
    impl(const impl &o)
        : sys(o.sys), dc(o.dc), prog(o.prog), batch_size(o.batch_size), dim(o.dim), order(o.order), n_pars(o.n_pars),
          tol(o.tol), high_accuracy(o.high_accuracy), compact_mode(o.compact_mode), device(o.device), devices(o.devices),
          tape_mode(o.tape_mode), k_lpw(o.k_lpw), k_lpt(o.k_lpt), k_threads(o.k_threads), k_bpsm(o.k_bpsm),
          state(o.state), pars(o.pars), time_hi(o.time_hi), time_lo(o.time_lo), tc(o.tc), last_h(o.last_h),
          d_out(o.d_out), step_res(o.step_res), prop_res(o.prop_res), oc(o.oc), tmp_a(o.tmp_a), tmp_b(o.tmp_b),
          tmp_n(o.tmp_n), tc_valid(o.tc_valid)
    {
        if (prog) {
            make_batch();
            if (o.batch != nullptr && tc_valid) {
                // The copy can serve update_d_output() right away (src/detail/i_data.cpp:335-352 copies m_tc).
                check(hy_batch_upload_tc(batch, tc.data()));
            }
        }
    }
    ~impl()
    {
        hy_batch_destroy(batch);
    }
    void make_batch()
    {
        if (devices.empty()) {
            check(hy_batch_create(prog.get(), batch_size, device, &batch));
        } else {
            check(hy_batch_create_multi(prog.get(), batch_size, devices.data(), static_cast<std::uint32_t>(devices.size()),
                                        &batch));
        }
        if (tape_mode != 0 || k_lpw != 0 || k_lpt != 0 || k_threads != 0 || k_bpsm != 0) {
            check(hy_batch_set_kernel(batch, tape_mode, k_lpw, k_lpt, k_threads, k_bpsm));
        }
    }
    void push()
    {
        check(hy_batch_upload(batch, state.data(), n_pars ? pars.data() : nullptr, time_hi.data(), time_lo.data()));
    }
    void pull(bool wtc)
    {
        check(hy_batch_download(batch, state.data(), time_hi.data(), time_lo.data(), last_h.data()));
        if (wtc) {
            check(hy_batch_download_tc(batch, tc.data()));
            tc_valid = true;
        }
    }
    void pull_step_res()
    {
        check(hy_batch_download_step_res(batch, oc.data(), tmp_a.data()));
        for (std::uint32_t i = 0; i < batch_size; ++i) {
            step_res[i] = std::tuple{static_cast<taylor_outcome>(oc[i]), tmp_a[i]};
        }
    }
};

taylor_adaptive_batch<double>::taylor_adaptive_batch() : m_impl(std::make_unique<impl>()) {}

taylor_adaptive_batch<double>::taylor_adaptive_batch(const taylor_adaptive_batch &o)
    : m_impl(std::make_unique<impl>(*o.m_impl))
{
}

taylor_adaptive_batch<double>::taylor_adaptive_batch(taylor_adaptive_batch &&) noexcept = default;

taylor_adaptive_batch<double> &taylor_adaptive_batch<double>::operator=(const taylor_adaptive_batch &o)
{
    if (this != &o) {
        *this = taylor_adaptive_batch(o);
    }
    return *this;
}

taylor_adaptive_batch<double> &taylor_adaptive_batch<double>::operator=(taylor_adaptive_batch &&) noexcept = default;

taylor_adaptive_batch<double>::~taylor_adaptive_batch() = default;

// finalise_ctor_impl(), src/taylor_adaptive_batch.cpp:78-427 (validation order and messages).
void taylor_adaptive_batch<double>::finalise_ctor(std::vector<std::pair<expression, expression>> sys,
                                                  std::vector<double> state, std::uint32_t batch_size, ctor_opts o)
{
    auto &m = *m_impl;

    if (o.with_events) {
        throw not_implemented_error("Event detection is not supported by the B200 batch integrator");
    }
    validate_ode_sys(sys);

    m.batch_size = batch_size;
    m.high_accuracy = o.high_accuracy;
    m.compact_mode = o.compact_mode;
    m.device = o.device;
    m.devices = o.devices;

    if (batch_size == 0u) {
        throw std::invalid_argument("The batch size in an adaptive Taylor integrator cannot be zero");
    }
    if (state.size() % batch_size != 0u) {
        throw std::invalid_argument("Invalid size detected in the initialization of an adaptive Taylor integrator: "
                                    "the state vector has a size of "
                                    + std::to_string(state.size()) + ", which is not a multiple of the batch size ("
                                    + std::to_string(batch_size) + ")");
    }
    if (state.empty()) {
        state.resize(sys.size() * batch_size);
    }
    if (state.size() / batch_size != sys.size()) {
        throw std::invalid_argument("Inconsistent sizes detected in the initialization of an adaptive Taylor "
                                    "integrator: the state vector has a dimension of "
                                    + std::to_string(state.size() / batch_size) + " and a batch size of "
                                    + std::to_string(batch_size) + ", while the number of equations is "
                                    + std::to_string(sys.size()));
    }
    m.state = std::move(state);

    if (o.time_is_scalar) {
        m.time_hi.assign(batch_size, o.time_scalar);
    } else {
        m.time_hi = std::move(o.time);
    }
    if (m.time_hi.size() != batch_size) {
        throw std::invalid_argument("Invalid size detected in the initialization of an adaptive Taylor integrator: "
                                    "the time vector has a size of "
                                    + std::to_string(m.time_hi.size()) + ", which is not equal to the batch size ("
                                    + std::to_string(batch_size) + ")");
    }
    m.time_lo.assign(batch_size, 0.);

    if (o.tol && (!std::isfinite(*o.tol) || *o.tol < 0)) {
        throw std::invalid_argument("The tolerance in an adaptive Taylor integrator must be finite and positive, "
                                    "but it is "
                                    + fp_to_string(*o.tol) + " instead");
    }
    m.tol = (o.tol && *o.tol != 0) ? *o.tol : std::numeric_limits<double>::epsilon();
    m.dim = static_cast<std::uint32_t>(sys.size());

    std::vector<expression> all_rhs;
    for (const auto &p : sys) {
        all_rhs.push_back(p.second);
    }
    m.n_pars = get_param_size(all_rhs);
    const auto pars_req = static_cast<std::size_t>(m.n_pars) * batch_size;
    if (o.pars.empty()) {
        o.pars.resize(pars_req);
    } else if (o.pars.size() != pars_req) {
        throw std::invalid_argument("Invalid number of parameter values passed to the constructor of an adaptive "
                                    "Taylor integrator in batch mode: "
                                    + std::to_string(o.pars.size())
                                    + " parameter value(s) were passed, but the ODE system contains "
                                    + std::to_string(m.n_pars) + " parameter(s) (in batches of "
                                    + std::to_string(batch_size) + ")");
    }
    m.pars = std::move(o.pars);

    m.order = detail::taylor_order_from_tol(m.tol);

    // Decompose, lower, create the device-resident batch (replaces taylor_add_adaptive_step() + JIT).
    m.dc = taylor_decompose_sys(sys, {}).first;
    try {
        m.prog = std::make_shared<hy_program>(
            detail::lower_decomposition(m.dc, m.dim, m.n_pars, m.order, m.high_accuracy));
    } catch (const detail::not_implemented_error &e) {
        throw not_implemented_error(e.what());
    }
    m.sys = std::move(sys);
    m.make_batch();

    // Buffers (src/taylor_adaptive_batch.cpp:361-402).
    const auto n = static_cast<std::size_t>(batch_size);
    m.tc.assign(static_cast<std::size_t>(m.dim) * (m.order + 1u) * n, 0.);
    m.last_h.assign(n, 0.);
    m.d_out.assign(static_cast<std::size_t>(m.dim) * n, 0.);
    m.step_res.assign(n, std::tuple{taylor_outcome::success, 0.});
    m.prop_res.assign(n, std::tuple{taylor_outcome::success, 0., 0., std::size_t(0)});
    m.oc.assign(n, 0);
    m.tmp_a.assign(n, 0.);
    m.tmp_b.assign(n, 0.);
    m.tmp_n.assign(n, 0);

    // Non-finite initial conditions are rejected (src/taylor_adaptive_batch.cpp:404-423).
    for (const auto x : m.state) {
        if (!std::isfinite(x)) {
            throw std::invalid_argument(
                "A non-finite value was detected in the initial state of an adaptive Taylor integrator");
        }
    }
    for (const auto x : m.time_hi) {
        if (!std::isfinite(x)) {
            throw std::invalid_argument(
                "A non-finite initial time was detected in the initialisation of an adaptive Taylor integrator");
        }
    }
}

const taylor_dc_t &taylor_adaptive_batch<double>::get_decomposition() const
{
    return m_impl->dc;
}
std::uint32_t taylor_adaptive_batch<double>::get_batch_size() const
{
    return m_impl->batch_size;
}
std::uint32_t taylor_adaptive_batch<double>::get_order() const
{
    return m_impl->order;
}
double taylor_adaptive_batch<double>::get_tol() const
{
    return m_impl->tol;
}
bool taylor_adaptive_batch<double>::get_high_accuracy() const
{
    return m_impl->high_accuracy;
}
bool taylor_adaptive_batch<double>::get_compact_mode() const
{
    return m_impl->compact_mode;
}
std::uint32_t taylor_adaptive_batch<double>::get_dim() const
{
    return m_impl->dim;
}
const std::vector<std::pair<expression, expression>> &taylor_adaptive_batch<double>::get_sys() const noexcept
{
    return m_impl->sys;
}
const std::vector<double> &taylor_adaptive_batch<double>::get_time() const
{
    return m_impl->time_hi;
}
const double *taylor_adaptive_batch<double>::get_time_data() const
{
    return m_impl->time_hi.data();
}

// set_time() / set_dtime(): src/taylor_adaptive_batch.cpp:2120-2232.
void taylor_adaptive_batch<double>::set_time(const std::vector<double> &t)
{
    auto &m = *m_impl;
    if (t.size() != m.batch_size) {
        throw std::invalid_argument("Invalid number of new times specified in a Taylor integrator in batch mode: the "
                                    "batch size is "
                                    + std::to_string(m.batch_size) + ", but the number of specified times is "
                                    + std::to_string(t.size()));
    }
    m.time_hi = t;
    std::fill(m.time_lo.begin(), m.time_lo.end(), 0.);
}
void taylor_adaptive_batch<double>::set_time(double t)
{
    std::fill(m_impl->time_hi.begin(), m_impl->time_hi.end(), t);
    std::fill(m_impl->time_lo.begin(), m_impl->time_lo.end(), 0.);
}
std::pair<const std::vector<double> &, const std::vector<double> &> taylor_adaptive_batch<double>::get_dtime() const
{
    return {m_impl->time_hi, m_impl->time_lo};
}
void taylor_adaptive_batch<double>::set_dtime(const std::vector<double> &hi, const std::vector<double> &lo)
{
    auto &m = *m_impl;
    if (hi.size() != m.batch_size || lo.size() != m.batch_size) {
        throw std::invalid_argument("Invalid number of new times specified in a Taylor integrator in batch mode: the "
                                    "batch size is "
                                    + std::to_string(m.batch_size) + ", but the number of specified times is ("
                                    + std::to_string(hi.size()) + ", " + std::to_string(lo.size()) + ")");
    }
    for (std::uint32_t i = 0; i < m.batch_size; ++i) {
        if (std::isfinite(hi[i]) && std::isfinite(lo[i]) && std::abs(hi[i]) < std::abs(lo[i])) {
            throw std::invalid_argument("The first component of a double-length time must not be smaller in "
                                        "magnitude than the second");
        }
    }
    for (std::uint32_t i = 0; i < m.batch_size; ++i) {
        const auto r = eft_dekker(hi[i], lo[i]); // normalise
        m.time_hi[i] = r.hi;
        m.time_lo[i] = r.lo;
    }
}
void taylor_adaptive_batch<double>::set_dtime(double hi, double lo)
{
    set_dtime(std::vector<double>(m_impl->batch_size, hi), std::vector<double>(m_impl->batch_size, lo));
}

const std::vector<double> &taylor_adaptive_batch<double>::get_state() const
{
    return m_impl->state;
}
const double *taylor_adaptive_batch<double>::get_state_data() const
{
    return m_impl->state.data();
}
double *taylor_adaptive_batch<double>::get_state_data()
{
    return m_impl->state.data();
}
const std::vector<double> &taylor_adaptive_batch<double>::get_pars() const
{
    return m_impl->pars;
}
const double *taylor_adaptive_batch<double>::get_pars_data() const
{
    return m_impl->pars.data();
}
double *taylor_adaptive_batch<double>::get_pars_data()
{
    return m_impl->pars.data();
}
const std::vector<double> &taylor_adaptive_batch<double>::get_tc() const
{
    return m_impl->tc;
}
const std::vector<double> &taylor_adaptive_batch<double>::get_last_h() const
{
    return m_impl->last_h;
}
const std::vector<double> &taylor_adaptive_batch<double>::get_d_output() const
{
    return m_impl->d_out;
}
const std::vector<std::tuple<taylor_outcome, double>> &taylor_adaptive_batch<double>::get_step_res() const
{
    return m_impl->step_res;
}
const std::vector<std::tuple<taylor_outcome, double, double, std::size_t>> &
taylor_adaptive_batch<double>::get_propagate_res() const
{
    return m_impl->prop_res;
}
hy_batch *taylor_adaptive_batch<double>::get_device_batch()
{
    return m_impl->batch;
}
int taylor_adaptive_batch<double>::get_device() const
{
    return m_impl->device;
}
// Re-creates the device-resident batch on another GPU (or sharded over several); the host mirrors are the master
// copy of state / parameters / time, the Taylor coefficients of the last step are restored if they were valid.
void taylor_adaptive_batch<double>::set_devices(const std::vector<int> &devices)
{
    auto &m = *m_impl;
    if (devices == m.devices && (!devices.empty() || m.batch != nullptr)) {
        return;
    }
    hy_batch_destroy(m.batch);
    m.batch = nullptr;
    m.devices = devices;
    m.make_batch();
    if (m.tc_valid) {
        check(hy_batch_upload_tc(m.batch, m.tc.data()));
    }
}
void taylor_adaptive_batch<double>::set_device(int device)
{
    auto &m = *m_impl;
    if (m.devices.empty() && (device == m.device || device < 0)) {
        return;
    }
    hy_batch_destroy(m.batch);
    m.batch = nullptr;
    m.devices.clear();
    m.device = device;
    m.make_batch();
    if (m.tc_valid) {
        check(hy_batch_upload_tc(m.batch, m.tc.data()));
    }
}

namespace detail
{

int ensemble_device_count()
{
    return hy_device_count();
}

void ensemble_for_each(std::size_t n_iter, const std::function<void(std::size_t, int)> &fn)
{
    const int n_dev = std::max(1, ensemble_device_count());
    const std::size_t n_workers = std::min<std::size_t>(static_cast<std::size_t>(n_dev), n_iter);
    if (n_workers <= 1u) {
        for (std::size_t i = 0; i < n_iter; ++i) {
            fn(i, -1);
        }
        return;
    }
    std::vector<std::exception_ptr> errs(n_workers);
    std::vector<std::thread> thr;
    for (std::size_t w = 0; w < n_workers; ++w) {
        thr.emplace_back([&, w] {
            try {
                for (std::size_t i = w; i < n_iter; i += n_workers) {
                    fn(i, static_cast<int>(w));
                }
            } catch (...) {
                errs[w] = std::current_exception();
            }
        });
    }
    for (auto &t : thr) {
        t.join();
    }
    for (const auto &e : errs) {
        if (e) {
            std::rethrow_exception(e);
        }
    }
}

} // namespace detail

void taylor_adaptive_batch<double>::set_kernel(int tape_mode, std::uint32_t lpw, std::uint32_t lpt,
                                               std::uint32_t threads, std::uint32_t bpsm)
{
    auto &m = *m_impl;
    check(hy_batch_set_kernel(m.batch, tape_mode, lpw, lpt, threads, bpsm));
    m.tape_mode = tape_mode;
    m.k_lpw = lpw;
    m.k_lpt = lpt;
    m.k_threads = threads;
    m.k_bpsm = bpsm;
}

// ---- stepping (src/taylor_adaptive_batch.cpp:1039-1078) ----
void taylor_adaptive_batch<double>::step_impl(const std::vector<double> *max_delta_ts, bool backward, bool wtc)
{
    auto &m = *m_impl;
    m.push();
    check(hy_batch_step(m.batch, max_delta_ts != nullptr ? max_delta_ts->data() : nullptr, 0, backward ? 1 : 0,
                        wtc ? 1 : 0));
    m.pull(wtc);
    m.pull_step_res();
}

void taylor_adaptive_batch<double>::step(bool wtc)
{
    step_impl(nullptr, false, wtc);
}

void taylor_adaptive_batch<double>::step_backward(bool wtc)
{
    step_impl(nullptr, true, wtc);
}

void taylor_adaptive_batch<double>::step(const std::vector<double> &max_delta_ts, bool wtc)
{
    const auto &m = *m_impl;
    if (max_delta_ts.size() != m.batch_size) {
        throw std::invalid_argument("Invalid number of max timesteps specified in a Taylor integrator in batch mode: "
                                    "the batch size is "
                                    + std::to_string(m.batch_size) + ", but the number of specified timesteps is "
                                    + std::to_string(max_delta_ts.size()));
    }
    for (const auto x : max_delta_ts) {
        if (std::isnan(x)) {
            throw std::invalid_argument(
                "Cannot use a nan max_delta_t in the step() function of an adaptive Taylor integrator in batch mode");
        }
    }
    step_impl(&max_delta_ts, false, wtc);
}

void taylor_adaptive_batch<double>::check_max_delta_t_size(std::size_t n) const
{
    if (n != m_impl->batch_size) {
        throw std::invalid_argument("Invalid number of max timesteps specified in a Taylor integrator in batch mode: "
                                    "the batch size is "
                                    + std::to_string(m_impl->batch_size)
                                    + ", but the number of specified timesteps is " + std::to_string(n));
    }
}

// ---- propagation (src/taylor_adaptive_batch.cpp:1081-1534) ----
std::tuple<std::optional<continuous_output_batch<double>>, step_callback_batch<double>>
taylor_adaptive_batch<double>::propagate_for_vec(const std::vector<double> &delta_ts, prop_opts o)
{
    auto &m = *m_impl;
    if (delta_ts.size() != m.batch_size) {
        throw std::invalid_argument("Invalid number of time intervals specified in a Taylor integrator in batch "
                                    "mode: the batch size is "
                                    + std::to_string(m.batch_size)
                                    + ", but the number of specified time intervals is "
                                    + std::to_string(delta_ts.size()));
    }
    std::vector<double> hi(m.batch_size), lo(m.batch_size);
    for (std::uint32_t i = 0; i < m.batch_size; ++i) {
        const auto r = dfl_add(dfl{m.time_hi[i], m.time_lo[i]}, dfl{delta_ts[i], 0.});
        hi[i] = r.hi;
        lo[i] = r.lo;
    }
    return propagate_until_impl(hi, lo, std::move(o));
}

std::tuple<std::optional<continuous_output_batch<double>>, step_callback_batch<double>>
taylor_adaptive_batch<double>::propagate_until_vec(const std::vector<double> &ts, prop_opts o)
{
    auto &m = *m_impl;
    if (ts.size() != m.batch_size) {
        throw std::invalid_argument("Invalid number of time limits specified in a Taylor integrator in batch mode: "
                                    "the batch size is "
                                    + std::to_string(m.batch_size) + ", but the number of specified time limits is "
                                    + std::to_string(ts.size()));
    }
    return propagate_until_impl(ts, std::vector<double>(m.batch_size, 0.), std::move(o));
}

// ---- continuous_output_batch<double> ----
continuous_output_batch<double>::continuous_output_batch(hy_cout *h, std::uint32_t batch_size, std::uint32_t dim)
    : m_h(h, [](hy_cout *p) { hy_cout_destroy(p); }), m_batch_size(batch_size), m_dim(dim),
      m_output(static_cast<std::size_t>(batch_size) * dim)
{
}

void continuous_output_batch<double>::check_valid() const
{
    if (!m_h) {
        throw std::invalid_argument("Cannot use a default-constructed continuous_output_batch object");
    }
}

const std::vector<double> &continuous_output_batch<double>::operator()(const std::vector<double> &tm)
{
    check_valid();
    if (tm.size() != m_batch_size) {
        throw std::invalid_argument("An invalid time vector was passed to the call operator of continuous_output_batch: "
                                    "the vector size is "
                                    + std::to_string(tm.size()) + ", but a size of " + std::to_string(m_batch_size)
                                    + " was expected instead");
    }
    check(hy_cout_eval(m_h.get(), tm.data(), m_output.data()));
    return m_output;
}

const std::vector<double> &continuous_output_batch<double>::operator()(double tm)
{
    check_valid();
    return (*this)(std::vector<double>(m_batch_size, tm));
}

std::pair<std::vector<double>, std::vector<double>> continuous_output_batch<double>::get_bounds() const
{
    check_valid();
    std::vector<double> lb(m_batch_size), ub(m_batch_size);
    check(hy_cout_get_bounds(m_h.get(), lb.data(), ub.data()));
    return {std::move(lb), std::move(ub)};
}

std::size_t continuous_output_batch<double>::get_n_steps() const
{
    check_valid();
    return static_cast<std::size_t>(hy_cout_n_steps(m_h.get()));
}

// propagate_grid(): src/taylor_adaptive_batch.cpp:1545-2055. The size checks that need the reference's wording are
// done here, the grid checks and the integration by hy_batch_propagate_grid().
std::tuple<step_callback_batch<double>, std::vector<double>>
taylor_adaptive_batch<double>::propagate_grid_impl(const std::vector<double> &grid, prop_opts o)
{
    auto &m = *m_impl;
    const auto n = m.batch_size;
    if (o.cb) {
        throw not_implemented_error("Callbacks are not supported by propagate_grid() in the B200 batch integrator");
    }
    if (grid.empty()) {
        throw std::invalid_argument(
            "Cannot invoke propagate_grid() in an adaptive Taylor integrator in batch mode if the time grid is empty");
    }
    if (grid.size() % n != 0u) {
        throw std::invalid_argument(
            "Invalid grid size detected in propagate_grid() for an adaptive Taylor integrator in batch mode: "
            "the grid has a size of "
            + std::to_string(grid.size()) + ", which is not a multiple of the batch size (" + std::to_string(n) + ")");
    }
    if (!o.max_delta_t.empty() && o.max_delta_t.size() != n) {
        throw std::invalid_argument("Invalid number of max timesteps specified in a Taylor integrator in batch mode: "
                                    "the batch size is "
                                    + std::to_string(n) + ", but the number of specified timesteps is "
                                    + std::to_string(o.max_delta_t.size()));
    }
    std::vector<double> retval(grid.size() * m.dim);
    m.push();
    check(hy_batch_propagate_grid(m.batch, grid.data(), grid.size() / n,
                                  o.max_delta_t.empty() ? nullptr : o.max_delta_t.data(), o.max_steps, retval.data()));
    m.pull(true);
    check(hy_batch_download_prop_res(m.batch, m.oc.data(), m.tmp_a.data(), m.tmp_b.data(), m.tmp_n.data()));
    for (std::uint32_t i = 0; i < n; ++i) {
        m.prop_res[i] = std::tuple{static_cast<taylor_outcome>(m.oc[i]), m.tmp_a[i], m.tmp_b[i],
                                   static_cast<std::size_t>(m.tmp_n[i])};
    }
    return {std::move(o.cb), std::move(retval)};
}

std::tuple<std::optional<continuous_output_batch<double>>, step_callback_batch<double>>
taylor_adaptive_batch<double>::propagate_until_impl(const std::vector<double> &hi, const std::vector<double> &lo,
                                                    prop_opts o)
{
    auto &m = *m_impl;
    const auto n = m.batch_size;

    if (o.c_output && o.cb) {
        throw not_implemented_error("Continuous output together with a callback is not supported by the B200 batch "
                                    "integrator");
    }

    // Validation, src/taylor_adaptive_batch.cpp:1212-1273.
    for (std::uint32_t i = 0; i < n; ++i) {
        if (!std::isfinite(m.time_hi[i]) || !std::isfinite(m.time_lo[i])) {
            throw std::invalid_argument("Cannot invoke the propagate_until() function of an adaptive Taylor "
                                        "integrator in batch mode if one of the current times is not finite");
        }
    }
    for (std::uint32_t i = 0; i < n; ++i) {
        if (!std::isfinite(hi[i]) || !std::isfinite(lo[i])) {
            throw std::invalid_argument("A non-finite time was passed to the propagate_until() function of an "
                                        "adaptive Taylor integrator in batch mode");
        }
    }
    for (const auto dt : o.max_delta_t) {
        if (std::isnan(dt)) {
            throw std::invalid_argument("A nan max_delta_t was passed to the propagate_until() function of an "
                                        "adaptive Taylor integrator in batch mode");
        }
        if (dt <= 0) {
            throw std::invalid_argument("A non-positive max_delta_t was passed to the propagate_until() function of "
                                        "an adaptive Taylor integrator in batch mode");
        }
    }
    std::vector<dfl> rem(n);
    for (std::uint32_t i = 0; i < n; ++i) {
        rem[i] = dfl_sub(dfl{hi[i], lo[i]}, dfl{m.time_hi[i], m.time_lo[i]});
        if (!std::isfinite(rem[i].hi) || !std::isfinite(rem[i].lo)) {
            throw std::invalid_argument("The final time passed to the propagate_until() function of an adaptive "
                                        "Taylor integrator in batch mode results in an overflow condition");
        }
    }
    const double *mdt = o.max_delta_t.empty() ? nullptr : o.max_delta_t.data();

    if (o.c_output) {
        // The reference's lock-step loop with the recording of the Taylor coefficients, on the device
        // (hy_batch_propagate_until_cout()).
        m.push();
        hy_cout *co = nullptr;
        check(hy_batch_propagate_until_cout(m.batch, hi.data(), lo.data(), mdt, o.max_steps, &co));
        std::optional<continuous_output_batch<double>> ret;
        if (co != nullptr) {
            ret.emplace(co, n, m.dim);
        }
        m.pull(true);
        check(hy_batch_download_prop_res(m.batch, m.oc.data(), m.tmp_a.data(), m.tmp_b.data(), m.tmp_n.data()));
        for (std::uint32_t i = 0; i < n; ++i) {
            m.prop_res[i] = std::tuple{static_cast<taylor_outcome>(m.oc[i]), m.tmp_a[i], m.tmp_b[i],
                                       static_cast<std::size_t>(m.tmp_n[i])};
        }
        return {std::move(ret), std::move(o.cb)};
    }

    if (!o.cb) {
        // Fast path: the whole loop runs on the device.
        m.push();
        check(hy_batch_propagate_until(m.batch, hi.data(), lo.data(), mdt, o.max_steps, o.write_tc ? 1 : 0));
        m.pull(o.write_tc);
        check(hy_batch_download_prop_res(m.batch, m.oc.data(), m.tmp_a.data(), m.tmp_b.data(), m.tmp_n.data()));
        for (std::uint32_t i = 0; i < n; ++i) {
            m.prop_res[i] = std::tuple{static_cast<taylor_outcome>(m.oc[i]), m.tmp_a[i], m.tmp_b[i],
                                       static_cast<std::size_t>(m.tmp_n[i])};
        }
        return {std::nullopt, std::move(o.cb)};
    }

    // Callback path: the reference's lock-step loop (src/taylor_adaptive_batch.cpp:1372-1527) on the host, one
    // device step per iteration (a host callback per step forces a synchronisation anyway).
    constexpr auto cb_time_errmsg
        = "The invocation of the callback passed to propagate_until() resulted in the alteration of the "
          "time coordinate of the integrator - this is not supported";
    std::vector<char> t_dir(n);
    std::vector<std::size_t> ts_count(n, 0);
    std::vector<double> min_h(n, std::numeric_limits<double>::infinity()), max_h(n, 0.), cur_max(n);
    for (std::uint32_t i = 0; i < n; ++i) {
        t_dir[i] = dfl_ge0(rem[i]) ? 1 : 0;
    }
    std::size_t iter_counter = 0;
    while (true) {
        for (std::uint32_t i = 0; i < n; ++i) {
            const double md = mdt != nullptr ? mdt[i] : std::numeric_limits<double>::infinity();
            const dfl lim = t_dir[i] ? (dfl_lt(rem[i], dfl{md, 0.}) ? rem[i] : dfl{md, 0.})
                                     : (dfl_lt(rem[i], dfl{-md, 0.}) ? dfl{-md, 0.} : rem[i]);
            cur_max[i] = lim.hi;
        }
        step_impl(&cur_max, false, o.write_tc);

        std::uint32_t n_done = 0;
        bool nfs = false;
        for (std::uint32_t i = 0; i < n; ++i) {
            const auto [oc, h] = m.step_res[i];
            if (oc == taylor_outcome::err_nf_state) {
                nfs = true;
            } else {
                ts_count[i] += static_cast<std::size_t>(h != 0);
                if (oc == taylor_outcome::success) {
                    const auto ah = std::abs(h);
                    min_h[i] = std::min(min_h[i], ah);
                    max_h[i] = std::max(max_h[i], ah);
                }
                const bool cur_done = (h == rem[i].hi);
                n_done += cur_done ? 1u : 0u;
                if (cur_done) {
                    rem[i] = dfl{0., 0.};
                } else {
                    rem[i] = dfl_sub(dfl{hi[i], lo[i]}, dfl{m.time_hi[i], m.time_lo[i]});
                }
            }
            m.prop_res[i] = std::tuple{oc, min_h[i], max_h[i], ts_count[i]};
        }
        if (nfs) {
            return {std::nullopt, std::move(o.cb)};
        }
        ++iter_counter;
        {
            const auto thi = m.time_hi, tlo = m.time_lo;
            const bool ret_cb = o.cb(*this);
            if (m.time_hi != thi || m.time_lo != tlo) {
                throw std::runtime_error(cb_time_errmsg);
            }
            if (!ret_cb) {
                for (auto &r : m.prop_res) {
                    std::get<0>(r) = taylor_outcome::cb_stop;
                }
                return {std::nullopt, std::move(o.cb)};
            }
        }
        if (n_done == n) {
            return {std::nullopt, std::move(o.cb)};
        }
        if (iter_counter == o.max_steps) {
            for (auto &r : m.prop_res) {
                std::get<0>(r) = taylor_outcome::step_limit;
            }
            return {std::nullopt, std::move(o.cb)};
        }
    }
}

// ---- dense output (src/taylor_adaptive_batch.cpp:2251-2327) ----
const std::vector<double> &taylor_adaptive_batch<double>::update_d_output(const std::vector<double> &t, bool rel_time)
{
    auto &m = *m_impl;
    if (t.size() != m.batch_size) {
        throw std::invalid_argument("Invalid number of time coordinates specified for the dense output in a Taylor "
                                    "integrator in batch mode: the batch size is "
                                    + std::to_string(m.batch_size) + ", but the number of time coordinates is "
                                    + std::to_string(t.size()));
    }
    std::vector<double> tau(m.batch_size);
    for (std::uint32_t i = 0; i < m.batch_size; ++i) {
        if (rel_time) {
            // Relative to the CURRENT time; the kernel expands about the start of the last step (:2276-2280).
            tau[i] = m.last_h[i] + t[i];
        } else {
            // tau = t - (time - last_h) in double-length arithmetic (:2276-2286).
            const auto t0 = dfl_sub(dfl{m.time_hi[i], m.time_lo[i]}, dfl{m.last_h[i], 0.});
            tau[i] = dfl_sub(dfl{t[i], 0.}, t0).hi;
        }
    }
    check(hy_batch_d_output(m.batch, tau.data(), m.d_out.data()));
    return m.d_out;
}

const std::vector<double> &taylor_adaptive_batch<double>::update_d_output(double t, bool rel_time)
{
    return update_d_output(std::vector<double>(m_impl->batch_size, t), rel_time);
}

} // namespace heyoka_b200
