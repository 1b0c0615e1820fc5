// Model builders: see include/heyoka_b200/model.hpp for the reference map.
#include <heyoka_b200/model.hpp>

#include <algorithm>
#include <array>
#include <iterator>
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

// The N-body right-hand side. What must match the reference (src/model/nbody.cpp:52-173) is the EXPRESSION that comes
// out - operand order and association decide the decomposition - not how it is assembled. Here: a table of state
// variables indexed [body][axis], one pass over the ordered pairs (i < j, i massive) that appends each pair's
// contribution to the acceleration lists of its two bodies, then one pass over the bodies that emits the six
// equations of each. Per pair, with d = r_j - r_i and q = |d|^-3 = pow(sum(d_x^2, d_y^2, d_z^2), -3/2):
//   * "rescaled" form, when G and a non-zero m_j are numbers: a_i += d (G m_j q), a_j += (d (G m_j q)) (-m_i / m_j)
//     (the second product reuses the first one: one convolution per axis and pair instead of two);
//   * general form: a_j += d (-m_i (G q)) and, if j is massive, a_i += d (m_j (G q)).
std::vector<std::pair<expression, expression>> nbody_impl(std::uint32_t n, const expression &Gconst,
                                                          const std::vector<expression> &masses_vec)
{
    nbody_checks(n, masses_vec);
    const auto n_massive = static_cast<std::uint32_t>(masses_vec.size());
    constexpr const char *axis_name[3] = {"x", "y", "z"};

    // pos[b][a], vel[b][a]; acc[b][a] = the terms of the acceleration of body b along axis a.
    std::vector<std::array<expression, 3>> pos(n), vel(n);
    std::vector<std::array<std::vector<expression>, 3>> acc(n);
    for (std::uint32_t b = 0; b < n; ++b) {
        for (int a = 0; a < 3; ++a) {
            const auto suffix = "_" + std::to_string(b);
            pos[b][a] = expression{variable{axis_name[a] + suffix}};
            vel[b][a] = expression{variable{std::string("v") + axis_name[a] + suffix}};
        }
    }

    for (std::uint32_t i = 0; i < n_massive; ++i) {
        for (std::uint32_t j = i + 1u; j < n; ++j) {
            std::array<expression, 3> d;
            std::vector<expression> squares;
            for (int a = 0; a < 3; ++a) {
                d[a] = pos[j][a] - pos[i][a];
                squares.push_back(pow(d[a], 2_dbl));
            }
            const auto inv_r3 = pow(sum(squares), expression{-3. / 2});

            const bool j_has_mass = j < n_massive;
            const bool rescaled
                = j_has_mass && Gconst.is_number() && masses_vec[j].is_number() && masses_vec[j].num() != 0.;
            if (rescaled) {
                const auto pull_on_i = Gconst * masses_vec[j] * inv_r3;
                const auto ratio = -masses_vec[i] / masses_vec[j];
                for (int a = 0; a < 3; ++a) {
                    acc[i][a].push_back(d[a] * pull_on_i);
                }
                for (int a = 0; a < 3; ++a) {
                    acc[j][a].push_back(acc[i][a].back() * ratio);
                }
            } else {
                const auto g_inv_r3 = Gconst * inv_r3;
                const auto pull_on_j = -masses_vec[i] * g_inv_r3;
                for (int a = 0; a < 3; ++a) {
                    acc[j][a].push_back(d[a] * pull_on_j);
                }
                if (j_has_mass) {
                    const auto pull_on_i = masses_vec[j] * g_inv_r3;
                    for (int a = 0; a < 3; ++a) {
                        acc[i][a].push_back(d[a] * pull_on_i);
                    }
                }
            }
        }
    }

    std::vector<std::pair<expression, expression>> eqs;
    eqs.reserve(static_cast<std::size_t>(n) * 6u);
    for (std::uint32_t b = 0; b < n; ++b) {
        for (int a = 0; a < 3; ++a) {
            eqs.push_back(prime(pos[b][a]) = vel[b][a]);
        }
        for (int a = 0; a < 3; ++a) {
            eqs.push_back(prime(vel[b][a]) = sum(acc[b][a]));
        }
    }
    return eqs;
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

// A feed-forward network as expressions (src/model/ffnn.cpp:36-142). Layer l maps the n_{l-1} outputs of the previous
// layer to n_l neurons: out_i = activation_l(sum(W_l[i][0] in_0, ..., W_l[i][n_{l-1} - 1] in_{n_{l-1}-1}, b_l[i])).
// nn_wb holds every weight matrix, row-major and layer after layer, followed by every bias vector; the offsets of a
// layer's block are prefix sums of the layer sizes (no running counters).
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
    if (std::find(nn_hidden.begin(), nn_hidden.end(), 0u) != nn_hidden.end()) {
        throw std::invalid_argument("The number of neurons for each hidden layer must be greater than zero!");
    }
    if (std::any_of(activations.begin(), activations.end(), [](const auto &f) { return !f; })) {
        throw std::invalid_argument("The list of activation functions cannot contain empty functions");
    }

    // widths[l]: number of values entering layer l + 1 (widths[0] = inputs, widths.back() = outputs).
    std::vector<std::uint32_t> widths;
    widths.push_back(static_cast<std::uint32_t>(in.size()));
    widths.insert(widths.end(), nn_hidden.begin(), nn_hidden.end());
    widths.push_back(n_out);
    const std::size_t n_maps = widths.size() - 1u;

    // Offsets of every layer's weights and biases inside nn_wb.
    std::vector<std::size_t> w_off(n_maps), b_off(n_maps);
    std::size_t n_weights = 0, n_biases = 0;
    for (std::size_t l = 0; l < n_maps; ++l) {
        w_off[l] = n_weights;
        b_off[l] = n_biases;
        n_weights += static_cast<std::size_t>(widths[l]) * widths[l + 1u];
        n_biases += widths[l + 1u];
    }
    if (nn_wb.size() != n_weights + n_biases) {
        throw std::invalid_argument("The number of network parameters, detected from its structure to be "
                                    + std::to_string(n_weights + n_biases)
                                    + ", does not match the size of the corresponding expressions: "
                                    + std::to_string(nn_wb.size()) + ".");
    }

    std::vector<expression> values = in;
    for (std::size_t l = 0; l < n_maps; ++l) {
        const auto fan_in = widths[l], fan_out = widths[l + 1u];
        std::vector<expression> next;
        next.reserve(fan_out);
        for (std::uint32_t neuron = 0; neuron < fan_out; ++neuron) {
            const auto row = nn_wb.begin() + static_cast<std::ptrdiff_t>(w_off[l] + static_cast<std::size_t>(neuron) * fan_in);
            std::vector<expression> terms;
            terms.reserve(static_cast<std::size_t>(fan_in) + 1u);
            std::transform(row, row + fan_in, values.begin(), std::back_inserter(terms),
                           [](const expression &w, const expression &x) { return w * x; });
            terms.push_back(nn_wb[n_weights + b_off[l] + neuron]);
            next.push_back(activations[l](sum(terms)));
        }
        values = std::move(next);
    }
    return values;
}

} // namespace heyoka_b200::model::detail
