// heyoka_b200 — ODE model builders named by the benchmark configurations.
//
// Reference (bluescarni/heyoka @ 9c91f71):
//   model::nbody      src/model/nbody.cpp:53-173, include/heyoka/model/nbody.hpp:34-83
//   model::pendulum   src/model/pendulum.cpp:24-29, include/heyoka/model/pendulum.hpp
//   model::ffnn       src/model/ffnn.cpp:36-142, include/heyoka/model/ffnn.hpp:34-120
#ifndef HEYOKA_B200_MODEL_HPP
#define HEYOKA_B200_MODEL_HPP

#include <cstdint>
#include <functional>
#include <utility>
#include <vector>

#include <heyoka_b200/expression.hpp>
#include <heyoka_b200/kw.hpp>

namespace heyoka_b200
{

namespace model
{

namespace detail
{

std::vector<std::pair<expression, expression>> nbody_impl(std::uint32_t, const expression &,
                                                          const std::vector<expression> &);
expression nbody_energy_impl(std::uint32_t, const expression &, const std::vector<expression> &);
std::vector<std::pair<expression, expression>> pendulum_impl(const expression &, const expression &);
expression pendulum_energy_impl(const expression &, const expression &);
std::vector<expression> ffnn_impl(const std::vector<expression> &, const std::vector<std::uint32_t> &, std::uint32_t,
                                  const std::vector<std::function<expression(const expression &)>> &,
                                  const std::vector<expression> &);

template <typename T>
std::vector<expression> to_ex_vector(const T &r)
{
    std::vector<expression> ret;
    for (const auto &x : r) {
        ret.emplace_back(x);
    }
    return ret;
}

template <typename... KwArgs>
auto nbody_common_opts(std::uint32_t n, const KwArgs &...kw_args)
{
    static_assert(kw::allowed_tags<kw::Gconst_tag, kw::masses_tag>::template all<KwArgs...>(),
                  "Invalid named argument(s) for an N-body model");
    expression G{1.};
    kw::visit(kw::Gconst, [&G](const auto &v) { G = expression{v}; }, kw_args...);
    std::vector<expression> masses_vec;
    if constexpr (kw::has_tag<kw::masses_tag, KwArgs...>()) {
        kw::visit(kw::masses, [&masses_vec](const auto &v) { masses_vec = to_ex_vector(v); }, kw_args...);
    } else {
        masses_vec.resize(n, expression{1.});
    }
    return std::pair{std::move(G), std::move(masses_vec)};
}

} // namespace detail

template <typename... KwArgs>
std::vector<std::pair<expression, expression>> nbody(std::uint32_t n, const KwArgs &...kw_args)
{
    auto [G, m] = detail::nbody_common_opts(n, kw_args...);
    return detail::nbody_impl(n, G, m);
}

template <typename... KwArgs>
expression nbody_energy(std::uint32_t n, const KwArgs &...kw_args)
{
    auto [G, m] = detail::nbody_common_opts(n, kw_args...);
    return detail::nbody_energy_impl(n, G, m);
}

template <typename... KwArgs>
std::vector<std::pair<expression, expression>> pendulum(const KwArgs &...kw_args)
{
    static_assert(kw::allowed_tags<kw::gconst_tag, kw::length_tag>::template all<KwArgs...>(),
                  "Invalid named argument(s) for a pendulum model");
    expression g{1.}, l{1.};
    kw::visit(kw::gconst, [&g](const auto &v) { g = expression{v}; }, kw_args...);
    kw::visit(kw::length, [&l](const auto &v) { l = expression{v}; }, kw_args...);
    return detail::pendulum_impl(g, l);
}

template <typename... KwArgs>
expression pendulum_energy(const KwArgs &...kw_args)
{
    expression g{1.}, l{1.};
    kw::visit(kw::gconst, [&g](const auto &v) { g = expression{v}; }, kw_args...);
    kw::visit(kw::length, [&l](const auto &v) { l = expression{v}; }, kw_args...);
    return detail::pendulum_energy_impl(g, l);
}

// model::ffnn(kw::inputs = ..., kw::nn_hidden = ..., kw::n_out = ..., kw::activations = ..., [kw::nn_wb = ...]).
// Without kw::nn_wb, weights and biases are the runtime parameters par[0..n_wb).
template <typename... KwArgs>
std::vector<expression> ffnn(const KwArgs &...kw_args)
{
    static_assert(kw::allowed_tags<kw::inputs_tag, kw::nn_hidden_tag, kw::n_out_tag, kw::activations_tag,
                                   kw::nn_wb_tag>::template all<KwArgs...>(),
                  "Invalid named argument(s) for a FFNN model");
    static_assert(kw::has_tag<kw::inputs_tag, KwArgs...>() && kw::has_tag<kw::nn_hidden_tag, KwArgs...>()
                      && kw::has_tag<kw::n_out_tag, KwArgs...>() && kw::has_tag<kw::activations_tag, KwArgs...>(),
                  "kw::inputs, kw::nn_hidden, kw::n_out and kw::activations are required");
    std::vector<expression> in, nn_wb;
    std::vector<std::uint32_t> nn_hidden;
    std::uint32_t n_out = 0;
    std::vector<std::function<expression(const expression &)>> acts;
    kw::visit(kw::inputs, [&in](const auto &v) { in = detail::to_ex_vector(v); }, kw_args...);
    kw::visit(
        kw::nn_hidden,
        [&nn_hidden](const auto &v) {
            for (const auto &x : v) {
                nn_hidden.push_back(static_cast<std::uint32_t>(x));
            }
        },
        kw_args...);
    kw::visit(kw::n_out, [&n_out](const auto &v) { n_out = static_cast<std::uint32_t>(v); }, kw_args...);
    kw::visit(
        kw::activations,
        [&acts](const auto &v) {
            for (const auto &f : v) {
                acts.emplace_back(f);
            }
        },
        kw_args...);
    if constexpr (kw::has_tag<kw::nn_wb_tag, KwArgs...>()) {
        kw::visit(kw::nn_wb, [&nn_wb](const auto &v) { nn_wb = detail::to_ex_vector(v); }, kw_args...);
    } else {
        std::vector<std::uint32_t> n_neurons{static_cast<std::uint32_t>(in.size())};
        n_neurons.insert(n_neurons.end(), nn_hidden.begin(), nn_hidden.end());
        n_neurons.push_back(n_out);
        std::uint32_t n_wb = 0;
        for (std::size_t i = 1; i < n_neurons.size(); ++i) {
            n_wb += n_neurons[i - 1u] * n_neurons[i] + n_neurons[i];
        }
        for (std::uint32_t i = 0; i < n_wb; ++i) {
            nn_wb.push_back(par[i]);
        }
    }
    return detail::ffnn_impl(in, nn_hidden, n_out, acts, nn_wb);
}

} // namespace model

} // namespace heyoka_b200

#endif
