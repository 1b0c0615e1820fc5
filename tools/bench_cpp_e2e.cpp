// End-to-end throughput of the headline workload through the drop-in C++ class (the call a heyoka user makes):
// taylor_adaptive_batch<double> on the outer Solar System (benchmark/outer_ss_long_term_batch.cpp), every step = new
// initial conditions written through get_state_data() + set_time(0) + propagate_until(tf) + a read of the final state
// and of the propagation results. Prints one JSON line per synchronisation mode (host_sync::strict = the reference's
// raw-pointer contract, host_sync::lazy). Usage: bench_cpp_e2e <batch> <steps> <tfinal> [perturb]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include <heyoka_b200/heyoka_b200.hpp>

using namespace heyoka_b200;

int main(int argc, char **argv)
{
    const std::uint32_t batch = argc > 1 ? static_cast<std::uint32_t>(std::atoll(argv[1])) : 65536u;
    const int steps = argc > 2 ? std::atoi(argv[2]) : 3;
    const double tf = argc > 3 ? std::atof(argv[3]) : 20.;
    const double perturb = argc > 4 ? std::atof(argv[4]) : 1e-3;

    const std::vector<double> masses{1.00000597682, 1 / 1047.355, 1 / 3501.6, 1 / 22869., 1 / 19314., 7.4074074e-09};
    const double G = 0.01720209895 * 0.01720209895 * 365 * 365;
    // benchmark/outer_ss_long_term_batch.cpp:75-94 (AU, AU/day).
    const double ic0[36]
        = {-4.06428567034226e-3, -6.08813756435987e-3, -1.66162304225834e-6, +6.69048890636161e-6, -6.33922479583593e-6,
           -3.13202145590767e-9, +3.40546614227466e+0, +3.62978190075864e+0, +3.42386261766577e-2, -5.59797969310664e-3,
           +5.51815399480116e-3, -2.66711392865591e-6, +6.60801554403466e+0, +6.38084674585064e+0, -1.36145963724542e-1,
           -4.17354020307064e-3, +3.99723751748116e-3, +1.67206320571441e-5, +1.11636331405597e+1, +1.60373479057256e+1,
           +3.61783279369958e-1, -3.25884806151064e-3, +2.06438412905916e-3, -2.17699042180559e-5, -3.01777243405203e+1,
           +1.91155314998064e+0, -1.53887595621042e-1, -2.17471785045538e-4, -3.11361111025884e-3, +3.58344705491441e-5,
           -2.13858977531573e+1, +3.20719104739886e+1, +2.49245689556096e+0, -1.76936577252484e-3, -2.06720938381724e-3,
           +6.58091931493844e-4};
    std::mt19937_64 rng(42);
    std::uniform_real_distribution<double> u(-1., 1.);
    std::vector<double> ic(36u * static_cast<std::size_t>(batch));
    for (std::uint32_t v = 0; v < 36u; ++v) {
        const double base = ic0[v] * ((v % 6u) >= 3u ? 365. : 1.);
        for (std::uint32_t l = 0; l < batch; ++l) {
            ic[static_cast<std::size_t>(v) * batch + l] = base + std::abs(base) * u(rng) * perturb;
        }
    }
    auto sys = model::nbody(6, kw::masses = masses, kw::Gconst = G);
    taylor_adaptive_batch<double> ta{sys, ic, batch, kw::high_accuracy = true};

    for (const auto mode : {host_sync::strict, host_sync::lazy}) {
        ta.set_host_sync(mode);
        double sink = 0.;
        std::size_t lane_steps = 0;
        const auto one = [&]() {
            std::memcpy(ta.get_state_data(), ic.data(), ic.size() * sizeof(double));
            ta.set_time(0.);
            ta.propagate_until(tf);
            sink += ta.get_state()[0] + ta.get_time()[batch - 1u];
            lane_steps = 0;
            for (const auto &r : ta.get_propagate_res()) {
                lane_steps += std::get<3>(r);
            }
        };
        one(); // warm-up
        const auto t0 = std::chrono::steady_clock::now();
        for (int k = 0; k < steps; ++k) {
            one();
        }
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("{\"api\": \"taylor_adaptive_batch<double>::propagate_until\", \"host_sync\": \"%s\", \"batch\": %u, "
                    "\"steps\": %d, \"lane_steps_per_step\": %zu, \"ms_per_step\": %.3f, \"lane_steps_per_s\": %.6g, "
                    "\"h2d_bytes_per_step\": %zu, \"d2h_bytes_per_step\": %zu, \"check\": %.17g}\n",
                    mode == host_sync::strict ? "strict" : "lazy", batch, steps, lane_steps, 1e3 * secs / steps,
                    static_cast<double>(lane_steps) * steps / secs, ic.size() * 8u + 16u * batch,
                    ic.size() * 8u + (24u + 32u) * static_cast<std::size_t>(batch), sink);
    }
    return 0;
}
