// C++ API tests for heyoka_b200::taylor_adaptive_batch<double>, written after the reference's own
// test/taylor_adaptive_batch.cpp ("batch consistency" :105-144, "propagate for_until" :528-731, ctor errors
// :955-1016) and test/ensemble_propagate.cpp:355-467. Run by tests/test_cpp_api.py.
//
//   test_batch_api cpu   -> argument validation only (no CUDA device needed)
//   test_batch_api gpu   -> everything
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include <heyoka_b200/heyoka_b200.hpp>

using namespace heyoka_b200;

static int n_fail = 0;
#define REQUIRE(cond)                                                                                                  \
    do {                                                                                                               \
        if (!(cond)) {                                                                                                 \
            std::printf("REQUIRE failed at %s:%d: %s\n", __FILE__, __LINE__, #cond);                                   \
            ++n_fail;                                                                                                  \
        }                                                                                                              \
    } while (0)
#define REQUIRE_THROWS_MSG(expr, exc, msg)                                                                             \
    do {                                                                                                               \
        bool ok_ = false;                                                                                              \
        try {                                                                                                          \
            expr;                                                                                                      \
        } catch (const exc &e_) {                                                                                      \
            ok_ = std::string(e_.what()).find(msg) != std::string::npos;                                               \
            if (!ok_) {                                                                                                \
                std::printf("wrong message at %s:%d: %s\n", __FILE__, __LINE__, e_.what());                            \
            }                                                                                                          \
        } catch (...) {                                                                                                \
        }                                                                                                              \
        if (!ok_) {                                                                                                    \
            std::printf("REQUIRE_THROWS failed at %s:%d: %s\n", __FILE__, __LINE__, #expr);                            \
            ++n_fail;                                                                                                  \
        }                                                                                                              \
    } while (0)

static bool approx(double a, double b, double tol_eps)
{
    const double eps = std::numeric_limits<double>::epsilon();
    return std::abs(a - b) <= eps * tol_eps * std::max(std::abs(a), std::abs(b));
}

static void test_ctor_errors()
{
    auto [x, v] = make_vars("x", "v");
    using ta_t = taylor_adaptive_batch<double>;
    const std::vector<std::pair<expression, expression>> sys{prime(x) = v, prime(v) = -9.8 * sin(x)};

    REQUIRE_THROWS_MSG((ta_t{sys, {0.05, 0.06, 0.025, 0.026}, 0u}), std::invalid_argument,
                       "The batch size in an adaptive Taylor integrator cannot be zero");
    REQUIRE_THROWS_MSG((ta_t{sys, {0.05, 0.06, 0.025}, 2u}), std::invalid_argument,
                       "which is not a multiple of the batch size (2)");
    REQUIRE_THROWS_MSG((ta_t{sys, {0.05, 0.06}, 2u}), std::invalid_argument,
                       "the state vector has a dimension of 1 and a batch size of 2, while the number of equations is 2");
    REQUIRE_THROWS_MSG((ta_t{sys, {0.05, 0.06, 0.025, 0.026}, 2u, kw::time = std::vector<double>{0.}}),
                       std::invalid_argument, "the time vector has a size of 1, which is not equal to the batch size (2)");
    REQUIRE_THROWS_MSG((ta_t{sys, {0.05, 0.06, 0.025, 0.026}, 2u, kw::tol = -1.}), std::invalid_argument,
                       "The tolerance in an adaptive Taylor integrator must be finite and positive");
    REQUIRE_THROWS_MSG((ta_t{{prime(x) = v, prime(v) = -par[0] * sin(x)}, {0.05, 0.06, 0.025, 0.026}, 2u,
                             kw::pars = std::vector<double>{1.}}),
                       std::invalid_argument,
                       "1 parameter value(s) were passed, but the ODE system contains 1 parameter(s) (in batches of 2)");
    REQUIRE_THROWS_MSG((ta_t{{prime(x) = v, prime(x) = x}, {0.05, 0.06, 0.025, 0.026}, 2u}), std::invalid_argument,
                       "appears twice");
    {
        auto [y] = make_vars("y");
        REQUIRE_THROWS_MSG((ta_t{sys, {0.05, 0.06, 0.025, 0.026}, 2u, kw::t_events = {t_event_batch<double>(x + y)}}),
                           std::invalid_argument,
                           "an event function contains the variable 'y', which is not a state variable");
    }
}

static void test_batch_consistency()
{
    // test/taylor_adaptive_batch.cpp:105-144: forced damped pendulum, batch 4 vs 4 scalar-like runs (here: the
    // same lanes run as four batches of size 1), 1000 eps.
    auto [x, v] = make_vars("x", "v");
    const std::vector<std::pair<expression, expression>> sys{prime(x) = v,
                                                             prime(v) = cos(heyoka_b200::time) - .1 * v - sin(x)};
    const std::vector<double> st{0.05, 0.06, 0.07, 0.08, 0.025, 0.026, 0.027, 0.028};
    for (const bool ha : {false, true}) {
        taylor_adaptive_batch<double> ta{sys, st, 4u, kw::high_accuracy = ha};
        std::vector<taylor_adaptive_batch<double>> scal;
        for (unsigned i = 0; i < 4u; ++i) {
            scal.emplace_back(sys, std::vector<double>{st[i], st[4u + i]}, 1u, kw::high_accuracy = ha);
        }
        for (int step = 0; step < 200; ++step) {
            ta.step();
            for (unsigned i = 0; i < 4u; ++i) {
                scal[i].step();
                REQUIRE(std::get<0>(ta.get_step_res()[i]) == taylor_outcome::success);
                REQUIRE(approx(ta.get_state()[i], scal[i].get_state()[0], 1000.));
                REQUIRE(approx(ta.get_state()[4u + i], scal[i].get_state()[1], 1000.));
                REQUIRE(approx(ta.get_time()[i], scal[i].get_time()[0], 1000.));
            }
        }
    }
}

static void test_propagate_for_until()
{
    // test/taylor_adaptive_batch.cpp:528-731.
    auto [x, v] = make_vars("x", "v");
    for (const bool cm : {true, false}) {
        auto ta = taylor_adaptive_batch<double>{
            {prime(x) = v, prime(v) = -9.8 * sin(x)}, {0.05, 0.06, 0.025, 0.026}, 2u, kw::compact_mode = cm};
        auto ta_copy = ta;

        REQUIRE_THROWS_MSG(ta.propagate_until({0., std::numeric_limits<double>::infinity()}), std::invalid_argument,
                           "A non-finite time was passed to the propagate_until() function of an adaptive "
                           "Taylor integrator in batch mode");
        REQUIRE_THROWS_MSG(ta.propagate_until({10., 11.}, kw::max_delta_t = std::vector<double>{1}),
                           std::invalid_argument,
                           "Invalid number of max timesteps specified in a Taylor integrator in batch mode: the batch "
                           "size is 2, but the number of specified timesteps is 1");
        REQUIRE_THROWS_MSG(
            ta.propagate_until({10., 11.}, kw::max_delta_t = {1., std::numeric_limits<double>::quiet_NaN()}),
            std::invalid_argument, "A nan max_delta_t was passed to the propagate_until() function");
        REQUIRE_THROWS_MSG(ta.propagate_until({10., 11.}, kw::max_delta_t = {1., -1.}), std::invalid_argument,
                           "A non-positive max_delta_t was passed to the propagate_until() function");
        ta.set_time({0., std::numeric_limits<double>::lowest()});
        REQUIRE_THROWS_MSG(ta.propagate_until({10., std::numeric_limits<double>::max()}), std::invalid_argument,
                           "results in an overflow condition");
        ta.set_time({0., 0.});

        unsigned long counter0 = 0, counter1 = 0;
        auto cb = [&counter0, &counter1](taylor_adaptive_batch<double> &t) {
            if (t.get_last_h()[0] != 0) {
                ++counter0;
            }
            if (t.get_last_h()[1] != 0) {
                ++counter1;
            }
            return true;
        };

        // Callback path (host lock-step loop) with tiny max_delta_t: scaled down from the reference's 1e-4 / 5e-5
        // to keep the per-step host round trips affordable; the fused device path below runs the original.
        // (powers of two, so that the step counts are exact whatever the rounding of the last step.)
        ta.propagate_until({1., 1.25}, kw::max_delta_t = {0.0078125, 0.00390625}, kw::callback = cb);
        REQUIRE((ta.get_time() == std::vector<double>{1., 1.25}));
        REQUIRE(counter0 == 128ul);
        REQUIRE(counter1 == 320ul);
        for (const auto &r : ta.get_propagate_res()) {
            REQUIRE(std::get<0>(r) == taylor_outcome::time_limit);
        }

        // Same limits through the device loop (no callback): exact step counts 100000 / 220000 (:586-598).
        auto ta2 = ta_copy;
        ta2.propagate_until({10., 11.}, kw::max_delta_t = {1e-4, 5e-5});
        ta_copy.propagate_until({10., 11.});
        REQUIRE((ta2.get_time() == std::vector<double>{10., 11.}));
        REQUIRE(std::get<3>(ta2.get_propagate_res()[0]) == 100000u);
        REQUIRE(std::get<3>(ta2.get_propagate_res()[1]) == 220000u);
        REQUIRE((ta_copy.get_time() == std::vector<double>{10., 11.}));
        for (unsigned i = 0; i < 4u; ++i) {
            REQUIRE(approx(ta2.get_state()[i], ta_copy.get_state()[i], 1000.));
        }

        // Scalar vs vector final time (:606-612).
        auto c2 = ta_copy, c3 = ta_copy;
        c2.propagate_until(20.);
        c3.propagate_until({20., 20.});
        REQUIRE(c2.get_state() == c3.get_state());
        c2.propagate_for(5., kw::max_delta_t = 1e-2);
        c3.propagate_for({5., 5.}, kw::max_delta_t = {1e-2, 1e-2});
        REQUIRE(c2.get_state() == c3.get_state());
        REQUIRE(c2.get_propagate_res() == c3.get_propagate_res());

        // Callback interruption: all outcomes become cb_stop (:853-).
        int n_calls = 0;
        c2.propagate_for(10., kw::callback = [&n_calls](taylor_adaptive_batch<double> &) { return ++n_calls < 3; });
        REQUIRE(n_calls == 3);
        for (const auto &r : c2.get_propagate_res()) {
            REQUIRE(std::get<0>(r) == taylor_outcome::cb_stop);
        }
        // max_steps: step_limit for everybody.
        c3.propagate_for(10., kw::max_steps = 2u);
        for (const auto &r : c3.get_propagate_res()) {
            REQUIRE(std::get<0>(r) == taylor_outcome::step_limit);
            REQUIRE(std::get<3>(r) == 2u);
        }
    }
}

static void test_ensemble()
{
    // test/ensemble_propagate.cpp:355-467: harmonic oscillator, members must equal sequential runs bit for bit.
    auto [x, v] = make_vars("x", "v");
    taylor_adaptive_batch<double> ta{{prime(x) = v, prime(v) = -x}, {0., 0., 1., 1.}, 2u};
    const std::size_t n_iter = 16;
    const auto gen = [](taylor_adaptive_batch<double> tac, std::size_t i) {
        tac.get_state_data()[0] += i / 100.;
        tac.get_state_data()[1] += i / 100.;
        tac.get_state_data()[2] += i / 100.;
        tac.get_state_data()[3] += i / 100.;
        return tac;
    };
    auto res = ensemble_propagate_until_batch(ta, 20., n_iter, gen);
    REQUIRE(res.size() == n_iter);
    for (std::size_t i = 0; i < n_iter; ++i) {
        auto seq = gen(ta, i);
        seq.propagate_until(20.);
        REQUIRE(std::get<0>(res[i]).get_state() == seq.get_state());
        REQUIRE(std::get<0>(res[i]).get_time() == seq.get_time());
        REQUIRE(std::get<0>(res[i]).get_propagate_res() == seq.get_propagate_res());
    }
}

// Sharding over devices (kw::devices, set_device()): the lanes of one integrator split over several device-resident
// batches, bit-identical to the unsharded integrator; copies and moves between devices keep state, time and tc.
// A star and five planets on circular, coplanar orbits of radii 5, 9.5, 19, 30, 39 (years as time unit), lane i rotated
// and stretched a little: [x, y, z, vx, vy, vz] per body, batch innermost (the layout of model::nbody).
static std::vector<double> planetary_ic(std::uint32_t B, double G, double m0)
{
    const double radii[6] = {0., 5., 9.5, 19., 30., 39.};
    std::vector<double> st(36u * B, 0.);
    for (std::uint32_t l = 0; l < B; ++l) {
        for (std::uint32_t b = 1; b < 6u; ++b) {
            const double r = radii[b] * (1. + 1e-3 * l), ph = 0.9 * b + 0.05 * l, vc = std::sqrt(G * m0 / r);
            st[(b * 6u + 0u) * B + l] = r * std::cos(ph);
            st[(b * 6u + 1u) * B + l] = r * std::sin(ph);
            st[(b * 6u + 3u) * B + l] = -vc * std::sin(ph);
            st[(b * 6u + 4u) * B + l] = vc * std::cos(ph);
        }
    }
    return st;
}

static void test_sharded_and_placement()
{
    const std::vector<double> masses{1.00000597682, 1 / 1047.355, 1 / 3501.6, 1 / 22869., 1 / 19314., 7.4074074e-09};
    const double G = 0.01720209895 * 0.01720209895 * 365 * 365;
    auto sys = model::nbody(6, kw::masses = masses, kw::Gconst = G);
    const std::uint32_t B = 7u;
    const auto st = planetary_ic(B, G, masses[0]);
    taylor_adaptive_batch<double> one{sys, st, B, kw::high_accuracy = true};
    taylor_adaptive_batch<double> many{sys, st, B, kw::high_accuracy = true, kw::devices = std::vector<int>{0, 0, 0}};
    REQUIRE(hy_batch_n_shards(many.get_device_batch()) == 3u);
    one.step(true);
    many.step(true);
    REQUIRE(one.get_state() == many.get_state());
    REQUIRE(one.get_tc() == many.get_tc());
    REQUIRE(one.get_last_h() == many.get_last_h());
    std::vector<double> tf(B);
    for (std::uint32_t i = 0; i < B; ++i) {
        tf[i] = 3. + 0.7 * i;
    }
    one.propagate_until(tf);
    many.propagate_until(tf);
    for (const auto x : many.get_state()) {
        REQUIRE(std::isfinite(x));
    }
    REQUIRE(one.get_state() == many.get_state());
    REQUIRE(one.get_time() == many.get_time());
    REQUIRE(one.get_last_h() == many.get_last_h());
    REQUIRE(one.get_propagate_res() == many.get_propagate_res());
    // A copy of a sharded integrator is sharded the same way and serves dense output right away (tc is copied).
    many.step(true);
    one.step(true);
    REQUIRE(one.get_state() == many.get_state());
    REQUIRE(one.get_tc() == many.get_tc());
    REQUIRE(one.get_last_h() == many.get_last_h());
    auto cp = many;
    REQUIRE(hy_batch_n_shards(cp.get_device_batch()) == 3u);
    REQUIRE(cp.update_d_output(0., true) == many.update_d_output(0., true));
    REQUIRE(one.update_d_output(0., true) == many.update_d_output(0., true));
    for (std::size_t i = 0, n_shown = 0; i < st.size(); ++i) {
        const bool ok = approx(cp.get_d_output()[i], many.get_state()[i], 100.)
                        || std::abs(cp.get_d_output()[i] - many.get_state()[i]) < 1e-13; // (z components: ~0)
        REQUIRE(ok);
        if (!ok && n_shown++ < 4u) {
            std::fprintf(stderr, "d_output %zu: %.17g vs state %.17g (last_h %.17g)\n", i, cp.get_d_output()[i],
                         many.get_state()[i], many.get_last_h()[i % B]);
        }
    }
    // Back to one device.
    cp.set_device(0);
    REQUIRE(hy_batch_n_shards(cp.get_device_batch()) == 0u);
    cp.step();
    many.step();
    REQUIRE(cp.get_state() == many.get_state());
}

// test/ensemble_propagate.cpp ("batch grid"): every member of a grid ensemble equals its own sequential run.
static void test_ensemble_grid()
{
    auto [x, v] = make_vars("x", "v");
    taylor_adaptive_batch<double> ta{{prime(x) = v, prime(v) = -x}, {0., 0., 1., 1.}, 2u};
    const auto gen = [](taylor_adaptive_batch<double> tint, std::size_t i) {
        tint.get_state_data()[0] = 0.05 * static_cast<double>(i);
        tint.get_state_data()[1] = 0.05 * static_cast<double>(i) + 0.01;
        return tint;
    };
    const std::vector<double> grid{0., 1., 2.5, 4., 7.25};
    auto res = ensemble_propagate_grid_batch(ta, grid, 5u, gen);
    REQUIRE(res.size() == 5u);
    for (std::size_t i = 0; i < res.size(); ++i) {
        auto single = gen(ta, i);
        std::vector<double> g2;
        for (const auto t : grid) {
            g2.push_back(t);
            g2.push_back(t);
        }
        auto [cb, out] = single.propagate_grid(g2);
        REQUIRE(out == std::get<2>(res[i]));
        REQUIRE(single.get_state() == std::get<0>(res[i]).get_state());
        REQUIRE(std::get<2>(res[i]).size() == grid.size() * 2u * 2u);
    }
}

static void test_models_and_dense_output()
{
    const std::vector<double> masses{1.00000597682, 1 / 1047.355, 1 / 3501.6, 1 / 22869., 1 / 19314., 7.4074074e-09};
    const double G = 0.01720209895 * 0.01720209895 * 365 * 365;
    auto sys = model::nbody(6, kw::masses = masses, kw::Gconst = G);
    REQUIRE(sys.size() == 36u);
    std::vector<double> st(36u * 3u);
    for (std::size_t i = 0; i < st.size(); ++i) {
        st[i] = 0.1 + 0.37 * static_cast<double>((i * 7919u) % 101u) / 101.;
    }
    taylor_adaptive_batch<double> ta{sys, st, 3u, kw::high_accuracy = true, kw::tol = 1e-12};
    REQUIRE(ta.get_order() == 15u);
    REQUIRE(ta.get_decomposition().size() == 36u + 198u + 36u);
    const auto st0 = ta.get_state();
    ta.step(true);
    const auto &h = ta.get_last_h();
    // Dense output relative to the CURRENT time (src/taylor_adaptive_batch.cpp:2276-2280): 0 reproduces the state,
    // -last_h the state before the step.
    auto d1 = ta.update_d_output(0., true);
    for (std::size_t i = 0; i < st.size(); ++i) {
        REQUIRE(approx(d1[i], ta.get_state()[i], 100.));
    }
    std::vector<double> mh(h.size());
    for (std::size_t i = 0; i < h.size(); ++i) {
        mh[i] = -h[i];
    }
    auto d0 = ta.update_d_output(mh, true);
    for (std::size_t i = 0; i < st.size(); ++i) {
        REQUIRE(d0[i] == st0[i]);
    }
    // tc[var][0][lane] is the state before the step (src/taylor_00.cpp:574-580 layout)
    for (unsigned var = 0; var < 36u; ++var) {
        for (unsigned l = 0; l < 3u; ++l) {
            REQUIRE(ta.get_tc()[(var * 16u + 0u) * 3u + l] == st0[var * 3u + l]);
        }
    }
    auto psys = model::pendulum(kw::gconst = 9.8, kw::length = 1.);
    REQUIRE(psys.size() == 2u);
}

// test/taylor_adaptive_batch.cpp:162-323 ("propagate grid"): error messages, trivial grids, the harmonic
// oscillator on a dense grid forward and backward against the closed form (10000 eps).
static void test_propagate_grid()
{
    auto [x, v] = make_vars("x", "v");
    const double inf = std::numeric_limits<double>::infinity();
    taylor_adaptive_batch<double> ta{{prime(x) = v, prime(v) = -9.8 * sin(x)},
                                     {0.05, 0.025, 0.051, 0.0251, 0.052, 0.0252, 0.053, 0.0253},
                                     4u};
    REQUIRE_THROWS_MSG(ta.propagate_grid({}), std::invalid_argument,
                       "Cannot invoke propagate_grid() in an adaptive Taylor integrator in batch mode "
                       "if the time grid is empty");
    REQUIRE_THROWS_MSG(ta.propagate_grid({1., 2., 3., 4., 5.}), std::invalid_argument,
                       "the grid has a size of 5, which is not a multiple of the batch size (4)");
    REQUIRE_THROWS_MSG(ta.propagate_grid({0., 0., 1., 4.}), std::invalid_argument,
                       "batch index 2 has a value of 1, while the current time coordinate is 0");
    ta.set_time({0., 0., inf, 0.});
    REQUIRE_THROWS_MSG(ta.propagate_grid({0., 0., 0., 0.}), std::invalid_argument, "the current time is not finite");
    ta.set_time({0., 0., 0., 0.});
    REQUIRE_THROWS_MSG(ta.propagate_grid({0., 0., inf, 0.}), std::invalid_argument, "A non-finite time value");
    REQUIRE_THROWS_MSG(ta.propagate_grid({0., 0., 0., 0., 1., 1., -1., 1.}), std::invalid_argument,
                       "A non-monotonic time grid");
    REQUIRE_THROWS_MSG(ta.propagate_grid({0., 0., 0., 0., 1., 1., 1., 1., 2., 2., 1., 2.}), std::invalid_argument,
                       "A non-monotonic time grid");
    {
        auto [cb, ret] = ta.propagate_grid({0., 0., 0., 0.});
        REQUIRE(!cb);
        REQUIRE((ret == std::vector<double>{0.05, 0.025, 0.051, 0.0251, 0.052, 0.0252, 0.053, 0.0253}));
        for (const auto &[oc, min_h, max_h, nsteps] : ta.get_propagate_res()) {
            REQUIRE(oc == taylor_outcome::time_limit);
            REQUIRE(min_h == inf);
            REQUIRE(max_h == 0);
            REQUIRE(nsteps == 0u);
        }
    }
    for (const double sign : {1., -1.}) {
        taylor_adaptive_batch<double> osc{{prime(x) = v, prime(v) = -x}, {0., 0., 0., 0., 1., 1.1, 1.2, 1.3}, 4u};
        std::vector<double> grid;
        for (auto i = 0u; i < 1000u; ++i) {
            for (auto j = 0; j < 4; ++j) {
                grid.push_back(sign * (i / 100.));
                if (i != 0u) {
                    grid.back() += sign * (j / 10.);
                }
            }
        }
        auto [cb, ret] = osc.propagate_grid(grid);
        REQUIRE(!cb);
        REQUIRE(ret.size() == 8000u);
        for (auto i = 0u; i < 4u; ++i) {
            REQUIRE(std::get<0>(osc.get_propagate_res()[i]) == taylor_outcome::time_limit);
            REQUIRE(osc.get_time()[i] == grid[3996u + i]);
        }
        bool ok = true;
        for (auto i = 0u; i < 1000u; ++i) {
            for (auto j = 0u; j < 4u; ++j) {
                const double sv = (1 + j / 10.) * std::sin(grid[i * 4u + j]), cv = (1 + j / 10.) * std::cos(grid[i * 4u + j]);
                // approximately(): relative to the computed value, absolute below the tolerance.
                const double tol = std::numeric_limits<double>::epsilon() * 10000.;
                const auto near = [tol](double c, double e) {
                    return std::abs(c) < tol ? std::abs(c - e) <= tol : std::abs((c - e) / c) <= tol;
                };
                ok = ok && near(ret[8u * i + j], sv) && near(ret[8u * i + j + 4u], cv);
            }
        }
        REQUIRE(ok);
    }
    // A step callback (src/taylor_adaptive_batch.cpp:2004-2039): invoked after every step of the grid loop, the results
    // are those of the loop without it; returning false stops the propagation with cb_stop in every batch element and
    // leaves the grid points not reached as NaN; altering the time is an error.
    {
        const std::vector<double> ic{0., 0., 0., 0., 1., 1.1, 1.2, 1.3};
        std::vector<double> grid;
        for (auto i = 0u; i < 200u; ++i) {
            for (auto j = 0; j < 4; ++j) {
                grid.push_back(i / 20. + (i != 0u ? j / 10. : 0.));
            }
        }
        taylor_adaptive_batch<double> ref{{prime(x) = v, prime(v) = -x}, ic, 4u};
        auto [cb0, out0] = ref.propagate_grid(grid);
        REQUIRE(!cb0);
        taylor_adaptive_batch<double> ta{{prime(x) = v, prime(v) = -x}, ic, 4u};
        std::size_t calls = 0;
        auto [cb1, out1] = ta.propagate_grid(grid, kw::callback = [&calls](taylor_adaptive_batch<double> &) {
            ++calls;
            return true;
        });
        REQUIRE(static_cast<bool>(cb1));
        REQUIRE(calls > 0u);
        REQUIRE(out1 == out0);
        REQUIRE(ta.get_state() == ref.get_state());
        REQUIRE(ta.get_time() == ref.get_time());
        for (auto i = 0u; i < 4u; ++i) {
            REQUIRE(ta.get_propagate_res()[i] == ref.get_propagate_res()[i]);
        }
        taylor_adaptive_batch<double> tb{{prime(x) = v, prime(v) = -x}, ic, 4u};
        std::size_t n_cb = 0;
        auto [cb2, out2] = tb.propagate_grid(grid, kw::callback = [&n_cb](taylor_adaptive_batch<double> &) { return ++n_cb < 3u; });
        REQUIRE(n_cb == 3u);
        for (auto i = 0u; i < 4u; ++i) {
            REQUIRE(std::get<0>(tb.get_propagate_res()[i]) == taylor_outcome::cb_stop);
            REQUIRE(std::get<3>(tb.get_propagate_res()[i]) == 3u);
        }
        REQUIRE(std::isnan(out2.back()));
        REQUIRE(out2[0] == 0. && out2[4] == 1.);
        taylor_adaptive_batch<double> tc{{prime(x) = v, prime(v) = -x}, ic, 4u};
        REQUIRE_THROWS_MSG(tc.propagate_grid(grid, kw::callback =
                                                       [](taylor_adaptive_batch<double> &t) {
                                                           t.set_time(-1.);
                                                           return true;
                                                       }),
                           std::runtime_error, "resulted in the alteration of the time coordinate");
    }
}

// test/c_output.cpp:289-420 ("batch"): default-constructed object, continuous output against a grid propagation.
static void test_continuous_output()
{
    continuous_output_batch<double> co0;
    REQUIRE(co0.get_output().empty());
    REQUIRE(co0.get_batch_size() == 0u);
    REQUIRE_THROWS_MSG(co0(std::vector<double>{0., 0.}), std::invalid_argument,
                       "Cannot use a default-constructed continuous_output_batch object");
    REQUIRE_THROWS_MSG(co0.get_bounds(), std::invalid_argument,
                       "Cannot use a default-constructed continuous_output_batch object");
    REQUIRE_THROWS_MSG(co0.get_n_steps(), std::invalid_argument,
                       "Cannot use a default-constructed continuous_output_batch object");

    auto [x, v] = make_vars("x", "v");
    const unsigned batch_size = 4;
    std::vector<double> ic, final_tm, init_tm(batch_size, 0.);
    for (auto i = 0u; i < batch_size; ++i) {
        ic.push_back(i / 100.);
    }
    for (auto i = 0u; i < batch_size; ++i) {
        ic.push_back(1 + i / 100.);
        final_tm.push_back(10. + i / 100.);
    }
    const auto n_points = 10u;
    std::vector<double> grid(batch_size * n_points);
    for (auto i = 0u; i < batch_size; ++i) {
        grid[i] = 0;
        for (auto j = 1u; j + 1u < n_points; ++j) {
            grid[j * batch_size + i] = final_tm[i] * (j - 0.37 + 0.05 * i) / (n_points - 1.);
        }
        grid[grid.size() - batch_size + i] = final_tm[i];
    }
    taylor_adaptive_batch<double> ta{{prime(x) = v, prime(v) = -x}, ic, batch_size};
    auto [d_out, cb] = ta.propagate_until(final_tm, kw::c_output = true);
    REQUIRE(!cb);
    REQUIRE(d_out.has_value());
    REQUIRE(d_out->get_output().size() == 2u * batch_size);
    REQUIRE(d_out->get_n_steps() >= 5u);
    const auto bounds = d_out->get_bounds();
    REQUIRE(bounds.first == init_tm);
    REQUIRE(bounds.second == final_tm);
    REQUIRE_THROWS_MSG((*d_out)(std::vector<double>{0.}), std::invalid_argument,
                       "the vector size is 1, but a size of 4 was expected instead");

    std::copy(ic.begin(), ic.end(), ta.get_state_data());
    ta.set_time(init_tm);
    auto [cb_grid, grid_out] = ta.propagate_grid(grid);
    REQUIRE(!cb_grid);
    bool ok = true;
    const double tol = std::numeric_limits<double>::epsilon() * 100.;
    const auto near = [tol](double c, double e) {
        return std::abs(c) < tol ? std::abs(c - e) <= tol : std::abs((c - e) / c) <= tol;
    };
    std::vector<double> loc_time(batch_size);
    for (auto i = 0u; i < n_points; ++i) {
        for (auto j = 0u; j < batch_size; ++j) {
            loc_time[j] = grid[i * batch_size + j];
        }
        (*d_out)(loc_time);
        for (auto j = 0u; j < batch_size; ++j) {
            ok = ok && near(d_out->get_output()[j], grid_out[2u * i * batch_size + j])
                 && near(d_out->get_output()[batch_size + j], grid_out[2u * i * batch_size + batch_size + j]);
        }
        // The scalar overload.
        (*d_out)(loc_time[0]);
        ok = ok && near(d_out->get_output()[0], grid_out[2u * i * batch_size]);
    }
    REQUIRE(ok);
    // Continuous output TOGETHER with a step callback (src/taylor_adaptive_batch.cpp:1476-1500): the callback runs after
    // every recorded iteration and sees the integrator's current state; the recording is the one of the run without it;
    // returning false stops with cb_stop in every batch element and keeps what was recorded; the callback may alter the
    // state (here: not) but not the time.
    {
        auto [x2, v2] = make_vars("x", "v");
        const std::vector<double> ic2{0., 0.01, 0.02, 0.03, 1., 1.01, 1.02, 1.03}, tf2{10., 10.01, 10.02, 10.03};
        taylor_adaptive_batch<double> ref{{prime(x2) = v2, prime(v2) = -x2}, ic2, 4u};
        auto [co_ref, cb_ref] = ref.propagate_until(tf2, kw::c_output = true);
        REQUIRE(co_ref.has_value() && !cb_ref);
        taylor_adaptive_batch<double> ta2{{prime(x2) = v2, prime(v2) = -x2}, ic2, 4u};
        std::size_t calls = 0;
        double last_seen_t = -1.;
        auto [co_cb, cb_out] = ta2.propagate_until(
            tf2, kw::c_output = true, kw::callback = [&](taylor_adaptive_batch<double> &t) {
                ++calls;
                last_seen_t = t.get_time()[0];
                return true;
            });
        REQUIRE(co_cb.has_value() && static_cast<bool>(cb_out));
        REQUIRE(calls == co_ref->get_n_steps() && co_cb->get_n_steps() == co_ref->get_n_steps());
        REQUIRE(last_seen_t == 10.);
        REQUIRE(ta2.get_state() == ref.get_state() && ta2.get_time() == ref.get_time());
        for (const double tq : {0.5, 3.3, 9.99}) {
            REQUIRE((*co_cb)(tq) == (*co_ref)(tq));
        }
        taylor_adaptive_batch<double> ta3{{prime(x2) = v2, prime(v2) = -x2}, ic2, 4u};
        std::size_t n3 = 0;
        auto [co3, cb3] = ta3.propagate_until(tf2, kw::c_output = true,
                                              kw::callback = [&n3](taylor_adaptive_batch<double> &) { return ++n3 < 4u; });
        REQUIRE(n3 == 4u && co3.has_value() && co3->get_n_steps() == 4u);
        for (auto i = 0u; i < 4u; ++i) {
            REQUIRE(std::get<0>(ta3.get_propagate_res()[i]) == taylor_outcome::cb_stop);
        }
        REQUIRE(co3->get_bounds().second == ta3.get_time());
        taylor_adaptive_batch<double> ta4{{prime(x2) = v2, prime(v2) = -x2}, ic2, 4u};
        REQUIRE_THROWS_MSG(ta4.propagate_until(tf2, kw::c_output = true,
                                               kw::callback =
                                                   [](taylor_adaptive_batch<double> &t) {
                                                       t.set_time(-1.);
                                                       return true;
                                                   }),
                           std::runtime_error, "resulted in the alteration of the time coordinate");
        taylor_adaptive_batch<double> ta5{{prime(x2) = v2, prime(v2) = -x2}, ic2, 4u};
        REQUIRE_THROWS_MSG(ta5.propagate_until(tf2, kw::c_output = true,
                                               kw::callback = [](taylor_adaptive_batch<double> &) -> bool {
                                                   throw std::domain_error("from the callback");
                                               }),
                           std::domain_error, "from the callback");
    }
    // Continuous output of an integrator WITH events (the host lock-step loop records every iteration on the device,
    // hy_cout_rec_*): x = sin t, the non-terminal event v = 0 fires at pi / 2 + k pi; a stopping terminal event ends the
    // recording at the event time.
    {
        auto [x3, v3] = make_vars("x", "v");
        const std::vector<double> ic3{0., 0., 0., 0., 1., 1., 1., 1.};
        std::size_t n_ev_calls = 0;
        nt_event_batch<double> nte(v3, [&n_ev_calls](taylor_adaptive_batch<double> &, double, int, std::uint32_t) {
            ++n_ev_calls;
        });
        taylor_adaptive_batch<double> te_ta{{prime(x3) = v3, prime(v3) = -x3}, ic3, 4u, kw::nt_events = {nte}};
        auto [co_ev, cb_ev] = te_ta.propagate_until(10., kw::c_output = true);
        REQUIRE(co_ev.has_value());
        REQUIRE(n_ev_calls == 12u); // three zeros of cos t below 10 in each of the four lanes
        REQUIRE(co_ev->get_bounds().second == std::vector<double>(4u, 10.));
        for (const double tq : {0.3, 2., 5.5, 9.9}) {
            const auto &sv = (*co_ev)(tq);
            REQUIRE(std::abs(sv[0] - std::sin(tq)) < 1e-13);
            REQUIRE(std::abs(sv[4] - std::cos(tq)) < 1e-13);
        }
        t_event_batch<double> te(x3 + 0.5, kw::direction = event_direction::negative);
        taylor_adaptive_batch<double> te_tb{{prime(x3) = v3, prime(v3) = -x3}, ic3, 4u, kw::t_events = {te}};
        auto [co_te, cb_te] = te_tb.propagate_until(20., kw::c_output = true);
        REQUIRE(co_te.has_value());
        const double t_ev = 7. * 3.141592653589793 / 6.;
        REQUIRE(std::abs(te_tb.get_time()[0] - t_ev) < 1e-13);
        REQUIRE(co_te->get_bounds().second == te_tb.get_time());
        REQUIRE(static_cast<std::int64_t>(std::get<0>(te_tb.get_propagate_res()[0])) == -1);
    }
}

// Event detection through the drop-in class: blocks of test/batch_event_detection.cpp.
static void test_events()
{
    using t_ev_t = taylor_adaptive_batch<double>::t_event_t;
    using nt_ev_t = taylor_adaptive_batch<double>::nt_event_t;
    auto [x, v] = make_vars("x", "v");
    const double inf = std::numeric_limits<double>::infinity();

    // Event classes (src/t_event.cpp, src/nt_event.cpp; "te def ctor" :1818-1826).
    {
        t_ev_t te;
        REQUIRE(!te.get_callback());
        REQUIRE(te.get_direction() == event_direction::any);
        REQUIRE(te.get_cooldown() == -1.);
        REQUIRE_THROWS_MSG(t_ev_t(v, kw::cooldown = inf), std::invalid_argument,
                           "Cannot set a non-finite cooldown value for a terminal event");
        REQUIRE_THROWS_MSG(nt_ev_t(v, nt_ev_t::callback_t{}), std::invalid_argument,
                           "Cannot construct a non-terminal event with an empty callback");
        REQUIRE_THROWS_MSG(t_ev_t(v, kw::direction = static_cast<event_direction>(5)), std::invalid_argument,
                           "Invalid value selected for the direction of a terminal event");
    }
    // "nte linear box" / "te linear box" (:262-327): an event at the very end of a step fires in the next one.
    {
        auto counter = 0u;
        taylor_adaptive_batch<double> ta{{prime(x) = par[0]},
                                         {0., 0., 0., 0.},
                                         4u,
                                         kw::nt_events = {nt_ev_t(x - 1.,
                                                                  [&counter](auto &tint, double tm, int, std::uint32_t idx) {
                                                                      REQUIRE(approx(tm, 1 / tint.get_pars()[idx], 100.));
                                                                      ++counter;
                                                                  })},
                                         kw::pars = {1., 2., 4., 8.}};
        REQUIRE(ta.with_events());
        REQUIRE(ta.get_nt_events().size() == 1u);
        ta.step({1., 1 / 2., 1 / 4., 1 / 8.});
        REQUIRE(counter == 0u);
        for (const auto &r : ta.get_step_res()) {
            REQUIRE(std::get<0>(r) == taylor_outcome::time_limit);
        }
        ta.step({1., 1 / 2., 1 / 4., 1 / 8.});
        REQUIRE(counter == 4u);

        counter = 0u;
        taylor_adaptive_batch<double> tb{{prime(x) = par[0]},
                                         {0., 0., 0., 0.},
                                         4u,
                                         kw::t_events = {t_ev_t(x - 1., kw::callback =
                                                                            [&counter](auto &, int, std::uint32_t) {
                                                                                ++counter;
                                                                                return true;
                                                                            })},
                                         kw::pars = {1., 2., 4., 8.}};
        tb.step({1., 1 / 2., 1 / 4., 1 / 8.});
        REQUIRE(counter == 0u);
        tb.step({1., 1 / 2., 1 / 4., 1 / 8.});
        REQUIRE(counter == 4u);
        for (const auto &r : tb.get_step_res()) {
            REQUIRE(std::get<0>(r) == taylor_outcome{0});
            REQUIRE(std::abs(std::get<1>(r)) < 1e-14);
        }
    }
    // "te propagate_for" (:1440-1477).
    {
        std::vector<unsigned> counter(4u, 0u);
        t_ev_t ev(v, kw::callback = [&counter](auto &, int, std::uint32_t idx) {
            ++counter[idx];
            return true;
        });
        taylor_adaptive_batch<double> ta{
            {prime(x) = v, prime(v) = -9.8 * sin(x)}, {0, 0.01, 0.02, 0.03, .25, .26, .27, .28}, 4u, kw::t_events = {ev}};
        ta.propagate_for(100.);
        for (std::uint32_t i = 0; i < 4u; ++i) {
            REQUIRE(std::get<0>(ta.get_propagate_res()[i]) == taylor_outcome::time_limit);
            REQUIRE(ta.get_time()[i] == 100.);
            REQUIRE(counter[i] == 100u);
        }
        taylor_adaptive_batch<double> tb{{prime(x) = v, prime(v) = -9.8 * sin(x)},
                                         {0, 0.01, 0.02, 0.03, .25, .26, .27, .28},
                                         4u,
                                         kw::t_events = {t_ev_t(v)}};
        tb.propagate_for(100.);
        for (std::uint32_t i = 0; i < 4u; ++i) {
            REQUIRE(static_cast<std::int64_t>(std::get<0>(tb.get_propagate_res()[i])) == -1);
        }
        // Cooldowns can be cleared for one batch element or for all of them.
        tb.reset_cooldowns(1u);
        tb.reset_cooldowns();
        REQUIRE_THROWS_MSG(tb.reset_cooldowns(4u), std::invalid_argument,
                           "Cannot reset the cooldowns at batch index 4: the batch size for this integrator is only 4");
    }
    // "te propagate_grid" (:1479-1534).
    {
        std::vector<unsigned> counter(4u, 0u);
        t_ev_t ev(v, kw::callback = [&counter](auto &, int, std::uint32_t idx) {
            ++counter[idx];
            return true;
        });
        taylor_adaptive_batch<double> ta{
            {prime(x) = v, prime(v) = -9.8 * sin(x)}, {0, 0.01, 0.02, 0.03, .25, .26, .27, .28}, 4u, kw::t_events = {ev}};
        std::vector<double> grid;
        for (auto i = 0; i < 101; ++i) {
            grid.insert(grid.end(), 4u, static_cast<double>(i));
        }
        auto [cb, out] = ta.propagate_grid(grid);
        REQUIRE(!cb);
        REQUIRE(out.size() == 202u * 4u);
        REQUIRE(std::all_of(out.begin() + 1, out.end(), [](const auto &val) { return val != 0 && std::isfinite(val); }));
        for (std::uint32_t i = 0; i < 4u; ++i) {
            REQUIRE(counter[i] == 100u);
            REQUIRE(std::get<0>(ta.get_propagate_res()[i]) == taylor_outcome::time_limit);
        }
        taylor_adaptive_batch<double> tb{{prime(x) = v, prime(v) = -9.8 * sin(x)},
                                         {0, 0.01, 0.02, 0.03, .25, .26, .27, .28},
                                         4u,
                                         kw::t_events = {t_ev_t(v)}};
        std::tie(cb, out) = tb.propagate_grid(grid);
        REQUIRE(std::all_of(out.begin() + 8, out.end(), [](const auto &val) { return std::isnan(val); }));
        for (std::uint32_t i = 0; i < 4u; ++i) {
            REQUIRE(static_cast<std::int64_t>(std::get<0>(tb.get_propagate_res()[i])) == -1);
        }
    }
    // "te damped pendulum" (:1580-1645): the callback changes a parameter through get_pars_data().
    {
        std::vector<std::vector<double>> zero_vel_times(4u);
        t_ev_t ev(v, kw::callback = [&zero_vel_times](auto &ta, int, std::uint32_t idx) {
            ta.get_pars_data()[idx] = ta.get_pars()[idx] == 0 ? 1. : 0.;
            zero_vel_times[idx].push_back(ta.get_time()[idx]);
            return true;
        });
        taylor_adaptive_batch<double> ta{{prime(x) = v, prime(v) = -9.8 * sin(x) - par[0] * v},
                                         {0.05, 0.051, 0.052, 0.053, 0.025, 0.0251, 0.0252, 0.0253},
                                         4u,
                                         kw::t_events = {ev}};
        ta.propagate_until(100.);
        for (auto i = 0u; i < 4u; ++i) {
            REQUIRE(zero_vel_times[i].size() == 99u);
        }
        ta.step();
        for (auto i = 0u; i < 4u; ++i) {
            REQUIRE(zero_vel_times[i].size() == 100u);
        }
        // A copy carries the events (callbacks included) and detects on its own.
        auto cp = ta;
        REQUIRE(cp.with_events());
        cp.propagate_for(1.5); // (the velocity vanishes every ~1.0035 time units)
        for (auto i = 0u; i < 4u; ++i) {
            REQUIRE(zero_vel_times[i].size() == 101u);
        }
    }
    // An event function must be written in terms of the state variables (src/detail/validate_ode_sys.cpp:131-145).
    {
        auto [y] = make_vars("y");
        REQUIRE_THROWS_MSG((taylor_adaptive_batch<double>{{prime(x) = v, prime(v) = -x}, {0., 1.}, 1u,
                                                          kw::t_events = {t_ev_t(x + y)}}),
                           std::invalid_argument,
                           "Invalid system of differential equations detected: an event function contains the variable "
                           "'y', which is not a state variable");
    }
}

int main(int argc, char **argv)
{
    const bool gpu = argc > 1 && std::strcmp(argv[1], "gpu") == 0;
    test_ctor_errors();
    if (gpu) {
        test_batch_consistency();
        test_propagate_for_until();
        test_ensemble();
        test_models_and_dense_output();
        test_propagate_grid();
        test_continuous_output();
        test_ensemble_grid();
        test_sharded_and_placement();
        test_events();
    }
    if (n_fail == 0) {
        std::printf("ALL PASSED (%s)\n", gpu ? "gpu" : "cpu");
        return 0;
    }
    std::printf("%d FAILURES\n", n_fail);
    return 1;
}
