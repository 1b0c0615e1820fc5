// Model builders: see include/heyoka_b200/model.hpp for the reference map.
#include <heyoka_b200/model.hpp>

#include <stdexcept>
#include <string>

namespace heyoka_b200::model::detail
{

namespace
{

void nbody_checks(std::uint32_t n, const std::vector<expression> &masses_vec)
{
    if (n < 2u) {
        throw std::invalid_argument("Cannot construct an N-body system with N == " + std::to_string(n)
                                    + ": at least 2 bodies are needed");
    }
    if (masses_vec.size() > n) {
        throw std::invalid_argument("In an N-body system the number of particles with mass ("
                                    + std::to_string(masses_vec.size())
                                    + ") cannot be greater than the total number of particles (" + std::to_string(n)
                                    + ")");
    }
}

} // namespace

// State variable order: x_i, y_i, z_i, vx_i, vy_i, vz_i for each body (massive bodies first).
// Pair interactions use r**-3 = pow(sum(dx**2, dy**2, dz**2), -3/2); when G and m_j are numbers
// the j->i acceleration is computed once and the i->j one obtained by a constant rescaling
// (src/model/nbody.cpp:97-153).
std::vector<std::pair<expression, expression>> nbody_impl(std::uint32_t n, const expression &Gconst,
                                                          const std::vector<expression> &masses_vec)
{
    nbody_checks(n, masses_vec);

    std::vector<expression> x_vars, y_vars, z_vars, vx_vars, vy_vars, vz_vars;
    for (std::uint32_t i = 0; i < n; ++i) {
        const auto s = std::to_string(i);
        x_vars.emplace_back(variable{"x_" + s});
        y_vars.emplace_back(variable{"y_" + s});
        z_vars.emplace_back(variable{"z_" + s});
        vx_vars.emplace_back(variable{"vx_" + s});
        vy_vars.emplace_back(variable{"vy_" + s});
        vz_vars.emplace_back(variable{"vz_" + s});
    }

    std::vector<std::pair<expression, expression>> retval;
    std::vector<std::vector<expression>> x_acc(n), y_acc(n), z_acc(n);

    const auto n_massive = static_cast<std::uint32_t>(masses_vec.size());

    for (std::uint32_t i = 0; i < n_massive; ++i) {
        retval.push_back(prime(x_vars[i]) = vx_vars[i]);
        retval.push_back(prime(y_vars[i]) = vy_vars[i]);
        retval.push_back(prime(z_vars[i]) = vz_vars[i]);

        for (std::uint32_t j = i + 1u; j < n; ++j) {
            const auto diff_x = x_vars[j] - x_vars[i];
            const auto diff_y = y_vars[j] - y_vars[i];
            const auto diff_z = z_vars[j] - z_vars[i];

            const auto r_m3 = pow(sum({pow(diff_x, 2_dbl), pow(diff_y, 2_dbl), pow(diff_z, 2_dbl)}),
                                  expression{-3. / 2});

            const auto j_massive = j < n_massive;
            const auto opt_grouping = j_massive && masses_vec[j].is_number() && masses_vec[j].num() != 0.
                                      && Gconst.is_number();

            if (opt_grouping) {
                const auto fac_j = Gconst * masses_vec[j] * r_m3;
                const auto c_ij = -masses_vec[i] / masses_vec[j];

                x_acc[i].push_back(diff_x * fac_j);
                y_acc[i].push_back(diff_y * fac_j);
                z_acc[i].push_back(diff_z * fac_j);

                x_acc[j].push_back(x_acc[i].back() * c_ij);
                y_acc[j].push_back(y_acc[i].back() * c_ij);
                z_acc[j].push_back(z_acc[i].back() * c_ij);
            } else {
                const auto G_r_m3 = Gconst * r_m3;

                const auto fac_i = -masses_vec[i] * G_r_m3;
                x_acc[j].push_back(diff_x * fac_i);
                y_acc[j].push_back(diff_y * fac_i);
                z_acc[j].push_back(diff_z * fac_i);

                if (j_massive) {
                    const auto fac_j = masses_vec[j] * G_r_m3;
                    x_acc[i].push_back(diff_x * fac_j);
                    y_acc[i].push_back(diff_y * fac_j);
                    z_acc[i].push_back(diff_z * fac_j);
                }
            }
        }

        retval.push_back(prime(vx_vars[i]) = sum(x_acc[i]));
        retval.push_back(prime(vy_vars[i]) = sum(y_acc[i]));
        retval.push_back(prime(vz_vars[i]) = sum(z_acc[i]));
    }

    for (auto i = n_massive; i < n; ++i) {
        retval.push_back(prime(x_vars[i]) = vx_vars[i]);
        retval.push_back(prime(y_vars[i]) = vy_vars[i]);
        retval.push_back(prime(z_vars[i]) = vz_vars[i]);

        retval.push_back(prime(vx_vars[i]) = sum(x_acc[i]));
        retval.push_back(prime(vy_vars[i]) = sum(y_acc[i]));
        retval.push_back(prime(vz_vars[i]) = sum(z_acc[i]));
    }

    return retval;
}

// Total energy (kinetic + potential) of the N-body system (src/model/nbody.cpp:176-260).
expression nbody_energy_impl(std::uint32_t n, const expression &Gconst, const std::vector<expression> &masses_vec)
{
    nbody_checks(n, masses_vec);
    const auto n_massive = static_cast<std::uint32_t>(masses_vec.size());

    std::vector<expression> kin, pot;
    for (std::uint32_t i = 0; i < n_massive; ++i) {
        const auto s = std::to_string(i);
        const expression vx{variable{"vx_" + s}}, vy{variable{"vy_" + s}}, vz{variable{"vz_" + s}};
        kin.push_back(masses_vec[i] * sum({pow(vx, 2_dbl), pow(vy, 2_dbl), pow(vz, 2_dbl)}));
        for (std::uint32_t j = i + 1u; j < n_massive; ++j) {
            const auto t = std::to_string(j);
            const auto dx = expression{variable{"x_" + t}} - expression{variable{"x_" + s}};
            const auto dy = expression{variable{"y_" + t}} - expression{variable{"y_" + s}};
            const auto dz = expression{variable{"z_" + t}} - expression{variable{"z_" + s}};
            pot.push_back(masses_vec[i] * masses_vec[j]
                          * pow(sum({pow(dx, 2_dbl), pow(dy, 2_dbl), pow(dz, 2_dbl)}), expression{-.5}));
        }
    }
    return 0.5_dbl * sum(kin) - Gconst * sum(pot);
}

std::vector<std::pair<expression, expression>> pendulum_impl(const expression &gconst, const expression &l)
{
    auto [x, v] = make_vars("x", "v");
    return {prime(x) = v, prime(v) = -gconst / l * sin(x)};
}

expression pendulum_energy_impl(const expression &gconst, const expression &l)
{
    auto [x, v] = make_vars("x", "v");
    return 0.5_dbl * pow(l, 2_dbl) * pow(v, 2_dbl) + gconst * l * (1_dbl - cos(x));
}

namespace
{

// One dense layer: out_i = activation(sum_j(W_ij * in_j) + b_i) (src/model/ffnn.cpp:36-68).
std::vector<expression> compute_layer(std::uint32_t layer_id, const std::vector<expression> &inputs,
                                      const std::vector<std::uint32_t> &n_neurons,
                                      const std::function<expression(const expression &)> &activation,
                                      const std::vector<expression> &nn_wb, std::uint32_t n_net_w,
                                      std::uint32_t &wcounter, std::uint32_t &bcounter)
{
    const auto n_prev = static_cast<std::uint32_t>(inputs.size());
    const auto n_cur = n_neurons[layer_id];

    std::vector<expression> retval, tmp_sum;
    retval.reserve(n_cur);

    for (std::uint32_t i = 0; i < n_cur; ++i) {
        tmp_sum.clear();
        for (std::uint32_t j = 0; j < n_prev; ++j) {
            tmp_sum.push_back(nn_wb[wcounter] * inputs[j]);
            ++wcounter;
        }
        tmp_sum.push_back(nn_wb[bcounter + n_net_w]);
        ++bcounter;
        retval.push_back(activation(sum(tmp_sum)));
    }

    return retval;
}

} // namespace

// Weights/biases layout: [W01, W12, ..., B1, B2, ...], each W row-major (src/model/ffnn.cpp:70-142).
std::vector<expression> ffnn_impl(const std::vector<expression> &in, const std::vector<std::uint32_t> &nn_hidden,
                                  std::uint32_t n_out,
                                  const std::vector<std::function<expression(const expression &)>> &activations,
                                  const std::vector<expression> &nn_wb)
{
    if (activations.empty()) {
        throw std::invalid_argument("Cannot create a FFNN with an empty list of activation functions");
    }
    if (nn_hidden.size() != activations.size() - 1u) {
        throw std::invalid_argument(
            "The number of hidden layers, as detected from the inputs, was " + std::to_string(nn_hidden.size())
            + ", while the number of activation function supplied was " + std::to_string(activations.size())
            + ". A FFNN needs exactly one more activation function than the number of hidden layers.");
    }
    if (in.empty()) {
        throw std::invalid_argument("The inputs provided to the FFNN is an empty vector.");
    }
    if (n_out == 0u) {
        throw std::invalid_argument("The number of network outputs cannot be zero.");
    }
    for (const auto item : nn_hidden) {
        if (item == 0u) {
            throw std::invalid_argument("The number of neurons for each hidden layer must be greater than zero!");
        }
    }
    for (const auto &f : activations) {
        if (!f) {
            throw std::invalid_argument("The list of activation functions cannot contain empty functions");
        }
    }

    const auto n_layers = static_cast<std::uint32_t>(nn_hidden.size()) + 2u;
    std::vector<std::uint32_t> n_neurons{static_cast<std::uint32_t>(in.size())};
    n_neurons.insert(n_neurons.end(), nn_hidden.begin(), nn_hidden.end());
    n_neurons.push_back(n_out);

    std::uint32_t n_net_wb = 0, n_net_w = 0;
    for (std::uint32_t i = 1; i < n_layers; ++i) {
        n_net_wb += n_neurons[i - 1u] * n_neurons[i];
        n_net_w += n_neurons[i - 1u] * n_neurons[i];
        n_net_wb += n_neurons[i];
    }
    if (nn_wb.size() != n_net_wb) {
        throw std::invalid_argument("The number of network parameters, detected from its structure to be "
                                    + std::to_string(n_net_wb)
                                    + ", does not match the size of the corresponding expressions: "
                                    + std::to_string(nn_wb.size()) + ".");
    }

    std::vector<expression> retval = in;
    std::uint32_t wcounter = 0, bcounter = 0;
    for (std::uint32_t i = 1; i < n_layers; ++i) {
        retval = compute_layer(i, retval, n_neurons, activations[i - 1u], nn_wb, n_net_w, wcounter, bcounter);
    }
    return retval;
}

} // namespace heyoka_b200::model::detail
