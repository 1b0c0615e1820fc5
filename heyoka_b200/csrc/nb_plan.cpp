// See nb_plan.hpp.
#include "nb_plan.hpp"

#include <algorithm>
#include <cstring>
#include <map>

namespace heyoka_b200::detail
{

namespace
{

struct fail {
    std::string why;
};

} // namespace

nb_plan make_nb_plan(const hy_program &p)
{
    nb_plan pl;
    const std::uint32_t n_eq = p.n_eq, n_uvars = p.n_uvars, n_ops = n_uvars - n_eq;
    try {
        if (n_ops == 0u || p.n_pars != 0u) {
            throw fail{"no operations / runtime parameters"};
        }
        const auto op_of = [&](std::uint32_t u) -> const hy_op & { return p.ops[u - n_eq]; };

        // ---- who reads each u variable ----
        std::vector<std::vector<std::uint32_t>> users(n_uvars); // op indices
        for (std::uint32_t i = 0; i < n_ops; ++i) {
            const auto &op = p.ops[i];
            switch (op.opcode) {
                case HY_OP_SUM:
                case HY_OP_SUM_SQ:
                    for (std::uint32_t k = 0; k < op.b; ++k) {
                        const auto ref = p.args[op.a + k];
                        if (HY_REF_KIND(ref) != HY_REF_VAR) {
                            throw fail{"sum / sum_sq with a non-variable term"};
                        }
                        users[HY_REF_IDX(ref)].push_back(i);
                    }
                    break;
                case HY_OP_SUB_VV:
                case HY_OP_MUL_VV:
                    users[op.a].push_back(i);
                    users[op.b].push_back(i);
                    break;
                case HY_OP_POW_VN:
                case HY_OP_NEG:
                    users[op.a].push_back(i);
                    break;
                case HY_OP_MUL_NV:
                    users[op.b].push_back(i);
                    break;
                default:
                    throw fail{"opcode " + std::to_string(op.opcode) + " is not part of an N-body right-hand side"};
            }
        }
        std::vector<std::vector<std::uint32_t>> sv_of_u(n_uvars); // state variables whose derivative is u
        for (std::uint32_t s = 0; s < n_eq; ++s) {
            const auto ref = p.sv_defs[s];
            if (HY_REF_KIND(ref) == HY_REF_VAR) {
                sv_of_u[HY_REF_IDX(ref)].push_back(s);
            } else if (HY_REF_KIND(ref) != HY_REF_NUM) {
                throw fail{"state variable with a parameter as right-hand side"};
            }
        }

        // ---- state variables: velocities (derivative = u variable or number) and positions (derivative = velocity) ----
        // role: 1 velocity, 2 position.
        std::vector<int> role(n_eq, 0);
        std::vector<std::uint32_t> child_of(n_eq, ~0u); // velocity -> its position
        for (std::uint32_t s = 0; s < n_eq; ++s) {
            const auto ref = p.sv_defs[s];
            if (HY_REF_KIND(ref) == HY_REF_NUM || HY_REF_IDX(ref) >= n_eq) {
                role[s] = 1;
            }
        }
        for (std::uint32_t s = 0; s < n_eq; ++s) {
            if (role[s] == 1) {
                continue;
            }
            const auto par = HY_REF_IDX(p.sv_defs[s]);
            if (role[par] != 1) {
                throw fail{"state variable whose derivative is not a velocity"};
            }
            if (child_of[par] != ~0u) {
                throw fail{"two state variables derive from the same velocity"};
            }
            child_of[par] = s;
            role[s] = 2;
        }
        for (std::uint32_t s = 0; s < n_eq; ++s) {
            if (role[s] == 1 && !users[s].empty()) {
                throw fail{"a velocity is read by the right-hand side"};
            }
        }

        // ---- pair interactions ----
        std::vector<char> taken(n_ops, 0);
        // term_of[u]: output slot of the u variables that may appear in sums / as right-hand sides. While the pair
        // interactions are being collected the slots are provisional codes: (pair * 3 + k) for m_k, n_flag | the
        // same for n_k, sum_flag | index for the intermediate sums; final numbers are assigned below.
        constexpr std::uint32_t n_flag = 1u << 29, sum_flag = 1u << 30;
        std::vector<std::uint32_t> term_of(n_uvars, ~0u);
        std::vector<std::uint32_t> pos_slot(n_eq, ~0u);
        const auto pos_of = [&](std::uint32_t s) {
            if (role[s] != 2) {
                throw fail{"a pair interaction reads a state variable that is not a position"};
            }
            if (pos_slot[s] == ~0u) {
                pos_slot[s] = static_cast<std::uint32_t>(pl.pos_sv.size());
                pl.pos_sv.push_back(s);
            }
            return pos_slot[s];
        };
        const auto add_const = [&](double c) {
            for (std::size_t i = 0; i < pl.consts.size(); ++i) {
                if (std::memcmp(&pl.consts[i], &c, sizeof(double)) == 0) {
                    return static_cast<std::uint32_t>(i);
                }
            }
            pl.consts.push_back(c);
            return static_cast<std::uint32_t>(pl.consts.size() - 1u);
        };
        bool have_alpha = false;
        std::uint32_t n_out = 0;
        for (std::uint32_t qi = 0; qi < n_ops; ++qi) {
            const auto &qop = p.ops[qi];
            if (qop.opcode != HY_OP_POW_VN) {
                continue;
            }
            if (qop.a < n_eq) {
                throw fail{"pow of a state variable"};
            }
            const auto r2u = qop.a, qu = n_eq + qi;
            const auto &sop = op_of(r2u);
            if (sop.opcode != HY_OP_SUM_SQ || sop.b != 3u || users[r2u].size() != 1u || !sv_of_u[r2u].empty()
                || !sv_of_u[qu].empty()) {
                throw fail{"pow whose argument is not a private 3-term sum of squares"};
            }
            const double alpha = p.consts[qop.b];
            if (have_alpha && (std::memcmp(&alpha, &pl.alpha, sizeof(double)) != 0 || pl.pow_algo != qop.c)) {
                throw fail{"pair interactions with different exponents"};
            }
            have_alpha = true;
            pl.alpha = alpha;
            pl.pow_algo = qop.c;
            nb_pair_desc d{};
            std::uint32_t du[3], mu[3] = {0u, 0u, 0u};
            for (std::uint32_t k = 0; k < 3u; ++k) {
                du[k] = HY_REF_IDX(p.args[sop.a + k]);
                if (du[k] < n_eq || op_of(du[k]).opcode != HY_OP_SUB_VV || op_of(du[k]).a >= n_eq
                    || op_of(du[k]).b >= n_eq || !sv_of_u[du[k]].empty()) {
                    throw fail{"sum of squares of something else than coordinate differences"};
                }
                d.pa[k] = static_cast<std::uint16_t>(pos_of(op_of(du[k]).a));
                d.pb[k] = static_cast<std::uint16_t>(pos_of(op_of(du[k]).b));
                pl.pair_uvars.push_back(du[k]);
            }
            if (du[0] == du[1] || du[1] == du[2] || du[0] == du[2]) {
                throw fail{"repeated coordinate difference"};
            }
            // f
            std::uint32_t fu = qu;
            d.c1 = 1.;
            if (users[qu].size() == 1u) {
                const auto &fop = p.ops[users[qu][0]];
                if (fop.opcode == HY_OP_MUL_NV && fop.b == qu) {
                    fu = n_eq + users[qu][0];
                    d.c1 = p.consts[fop.a];
                } else if (fop.opcode == HY_OP_NEG && fop.a == qu) {
                    fu = n_eq + users[qu][0];
                    d.c1 = -1.;
                }
            }
            if (!sv_of_u[fu].empty()) {
                throw fail{"r^alpha is a right-hand side"};
            }
            // The products m_k = d_k * f (d first: the convolution pairs d^[n-j] with f^[j]).
            if (users[fu].size() != 3u) {
                throw fail{"r^alpha is not read by exactly three products"};
            }
            for (std::uint32_t k = 0; k < 3u; ++k) {
                std::uint32_t found = ~0u;
                for (const auto x : users[du[k]]) {
                    if (x == r2u - n_eq) {
                        continue;
                    }
                    const auto &mop = p.ops[x];
                    if (found != ~0u || mop.opcode != HY_OP_MUL_VV || mop.a != du[k] || mop.b != fu) {
                        throw fail{"coordinate difference read by something else than d * f"};
                    }
                    found = x;
                }
                if (found == ~0u || users[du[k]].size() != 2u) {
                    throw fail{"coordinate difference without its product"};
                }
                mu[k] = n_eq + found;
                d.on[k] = 0xffffu;
                term_of[mu[k]] = static_cast<std::uint32_t>(pl.pairs.size()) * 3u + k;
                taken[found] = 1;
                taken[du[k] - n_eq] = 1;
            }
            pl.pair_uvars.push_back(r2u);
            pl.pair_uvars.push_back(qu);
            pl.pair_uvars.insert(pl.pair_uvars.end(), {mu[0], mu[1], mu[2]});
            taken[r2u - n_eq] = 1;
            taken[qi] = 1;
            if (fu != qu) {
                taken[fu - n_eq] = 1;
            }
            pl.pairs.push_back(d);
        }
        if (pl.pairs.empty()) {
            throw fail{"no pair interaction"};
        }
        // ---- scaled outputs n_k = c * m_k / -m_k (read by sums or right-hand sides only): computed by the pair's thread ----
        for (std::uint32_t i = 0; i < n_ops; ++i) {
            if (taken[i]) {
                continue;
            }
            const auto &op = p.ops[i];
            std::uint32_t src = ~0u;
            double c = 0.;
            if (op.opcode == HY_OP_MUL_NV) {
                src = op.b;
                c = p.consts[op.a];
            } else if (op.opcode == HY_OP_NEG) {
                src = op.a;
                c = -1.;
            } else {
                continue;
            }
            if (src < n_eq || term_of[src] == ~0u || term_of[src] >= n_flag) {
                throw fail{"scaling of something else than a pair interaction's output"};
            }
            auto &d = pl.pairs[term_of[src] / 3u];
            const auto k = term_of[src] % 3u;
            if (d.on[k] != 0xffffu) {
                throw fail{"a pair interaction's output is rescaled twice"};
            }
            d.on[k] = 0u; // (assigned below)
            d.c2[k] = c;
            d.flags |= 1u;
            term_of[n_eq + i] = n_flag | term_of[src];
            taken[i] = 1;
        }
        // Final output slots: [m_0 of every pair | m_1 | m_2 | n_0 | n_1 | n_2 | intermediate sums]: the threads of a warp
        // (consecutive pairs) write consecutive 16-byte slots.
        const std::uint32_t NPR = static_cast<std::uint32_t>(pl.pairs.size());
        bool any_n = false;
        for (std::uint32_t pi = 0; pi < NPR; ++pi) {
            for (std::uint32_t k = 0; k < 3u; ++k) {
                pl.pairs[pi].om[k] = static_cast<std::uint16_t>(k * NPR + pi);
                if (pl.pairs[pi].on[k] != 0xffffu) {
                    pl.pairs[pi].on[k] = static_cast<std::uint16_t>((3u + k) * NPR + pi);
                    any_n = true;
                }
            }
        }
        n_out = (any_n ? 6u : 3u) * NPR;
        if (n_out >= 0xffffu) {
            throw fail{"too many pair interactions"};
        }
        const auto final_slot = [&](std::uint32_t code) {
            if (code & sum_flag) {
                return code & ~sum_flag;
            }
            const bool isn = (code & n_flag) != 0u;
            const auto c = code & ~n_flag;
            return ((isn ? 3u : 0u) + c % 3u) * NPR + c / 3u;
        };
        // ---- sums ----
        std::vector<std::uint32_t> sum_level(n_uvars, 0u);
        std::vector<std::uint32_t> sum_ops;
        for (std::uint32_t i = 0; i < n_ops; ++i) {
            if (taken[i]) {
                continue;
            }
            const auto &op = p.ops[i];
            if (op.opcode != HY_OP_SUM || op.b > 8u || op.b == 0u) {
                throw fail{"operation outside of pair interactions and sums"};
            }
            std::uint32_t lvl = 0;
            for (std::uint32_t k = 0; k < op.b; ++k) {
                const auto u = HY_REF_IDX(p.args[op.a + k]);
                if (u < n_eq || (term_of[u] == ~0u && (op_of(u).opcode != HY_OP_SUM || taken[u - n_eq]))) {
                    throw fail{"sum of something else than pair outputs / sums"};
                }
                if (term_of[u] == ~0u || (term_of[u] & sum_flag)) {
                    lvl = std::max(lvl, sum_level[u] + 1u); // (ops are topologically sorted)
                }
            }
            sum_level[n_eq + i] = lvl;
            sum_ops.push_back(i);
        }
        // A sum that defines a velocity must define exactly one and not be a term of another sum; the others are
        // intermediate and get an output slot.
        std::uint32_t n_levels = 1;
        for (const auto i : sum_ops) {
            const auto u = n_eq + i;
            n_levels = std::max(n_levels, sum_level[u] + 1u);
            if (sv_of_u[u].empty()) {
                if (users[u].empty()) {
                    throw fail{"unused sum"};
                }
                term_of[u] = sum_flag | n_out++;
            } else if (sv_of_u[u].size() != 1u || !users[u].empty()) {
                throw fail{"a sum is the right-hand side of several state variables or is reused"};
            }
        }
        // Pair outputs / scaled outputs that are right-hand sides must not be shared either.
        for (std::uint32_t u = n_eq; u < n_uvars; ++u) {
            if (op_of(u).opcode != HY_OP_SUM && sv_of_u[u].size() > 1u) {
                throw fail{"an output is the right-hand side of several state variables"};
            }
        }
        // Everything that is produced must be consumed by what the kernel computes.
        for (std::uint32_t u = n_eq; u < n_uvars; ++u) {
            if (term_of[u] != ~0u && op_of(u).opcode != HY_OP_SUM) {
                for (const auto x : users[u]) {
                    const auto oc = p.ops[x].opcode;
                    if (oc != HY_OP_SUM && oc != HY_OP_MUL_NV && oc != HY_OP_NEG) {
                        throw fail{"pair output read by an unsupported operation"};
                    }
                }
            }
        }
        if (n_out >= 0xffffu || pl.pos_sv.size() >= 0xffffu || n_eq >= 0xffffu) {
            throw fail{"too large"};
        }
        // ---- emit the levels ----
        const auto final_fields = [&](nb_sum_desc &sd, std::uint32_t s1) {
            const auto ch = child_of[s1];
            sd.out = s1 | ((ch == ~0u ? 0u : ch + 1u) << 16);
            sd.pos = (ch == ~0u || pos_slot[ch] == ~0u) ? 0u : pos_slot[ch] + 1u;
        };
        std::vector<std::vector<nb_sum_desc>> levels(n_levels);
        for (const auto i : sum_ops) {
            const auto &op = p.ops[i];
            const auto u = n_eq + i;
            nb_sum_desc sd{};
            sd.n_terms = op.b;
            for (std::uint32_t k = 0; k < op.b; ++k) {
                sd.terms[k] = final_slot(term_of[HY_REF_IDX(p.args[op.a + k])]);
            }
            if (sv_of_u[u].empty()) {
                sd.kind = 0;
                sd.out = final_slot(term_of[u]);
            } else {
                sd.kind = 1;
                final_fields(sd, sv_of_u[u][0]);
            }
            levels[sum_level[u]].push_back(sd);
        }
        // Velocities whose derivative is a single (scaled) pair output or a number.
        for (std::uint32_t s = 0; s < n_eq; ++s) {
            if (role[s] != 1) {
                continue;
            }
            const auto ref = p.sv_defs[s];
            nb_sum_desc sd{};
            if (HY_REF_KIND(ref) == HY_REF_NUM) {
                sd.kind = 2;
                sd.n_terms = 0;
                sd.cidx = add_const(p.consts[HY_REF_IDX(ref)]);
            } else {
                const auto u = HY_REF_IDX(ref);
                if (op_of(u).opcode == HY_OP_SUM) {
                    continue; // emitted above
                }
                if (term_of[u] == ~0u) {
                    throw fail{"right-hand side that is neither a sum nor a pair output"};
                }
                sd.kind = 1;
                sd.n_terms = 1;
                sd.terms[0] = final_slot(term_of[u]);
            }
            final_fields(sd, s);
            levels[0].push_back(sd);
        }
        pl.level_offsets.push_back(0u);
        for (auto &lv : levels) {
            // Finals first (their state-variable propagation is the longer job).
            std::stable_sort(lv.begin(), lv.end(), [](const nb_sum_desc &x, const nb_sum_desc &y) { return x.kind > y.kind; });
            pl.sums.insert(pl.sums.end(), lv.begin(), lv.end());
            pl.level_offsets.push_back(static_cast<std::uint32_t>(pl.sums.size()));
        }
        pl.n_pos = static_cast<std::uint32_t>(pl.pos_sv.size());
        pl.n_out = n_out;
        // ---- fac table ----
        const std::uint32_t order = p.order;
        pl.fac_stride = (order + 2u) & ~1u;
        pl.fac.assign(static_cast<std::size_t>(order + 1u) * pl.fac_stride, 0.);
        const double ap1 = pl.alpha + 1.;
        for (std::uint32_t n = 0; n <= order; ++n) {
            const double n_alpha = static_cast<double>(n) * pl.alpha;
            for (std::uint32_t j = 0; j < pl.fac_stride; ++j) {
                pl.fac[static_cast<std::size_t>(n) * pl.fac_stride + j] = n_alpha - static_cast<double>(j) * ap1;
            }
        }
        pl.ok = true;
    } catch (const fail &f) {
        pl = nb_plan{};
        pl.ok = false;
        pl.why = f.why;
    }
    return pl;
}

nb_roles make_nb_roles(const nb_plan &pl, std::uint32_t tt, std::uint32_t lt, std::uint32_t nl)
{
    nb_roles r;
    const std::uint32_t gs = lt / nl; // lane groups per team
    for (std::size_t lv = 0; lv + 1u < pl.level_offsets.size(); ++lv) {
        const std::uint32_t b = pl.level_offsets[lv], e = pl.level_offsets[lv + 1u];
        const std::uint32_t n_items = (e - b) * gs;
        const std::uint32_t rounds = std::max(1u, (n_items + tt - 1u) / tt);
        for (std::uint32_t rd = 0; rd < rounds; ++rd) {
            for (std::uint32_t t = 0; t < tt; ++t) {
                nb_role ro{};
                const std::uint32_t it = rd * tt + t;
                if (it < n_items) {
                    // (tt is a multiple of gs: thread t always works on the lanes (t % gs) * nl ...)
                    const auto &sd = pl.sums[b + it / gs];
                    const std::uint32_t l0 = (it % gs) * nl;
                    const auto unit = [&](std::uint32_t slot) { return slot * lt + l0; };
                    const std::uint32_t sv1 = sd.out & 0xffffu, sv2p1 = sd.out >> 16;
                    ro.head = sd.n_terms | ((sd.kind + 1u) << 4) | (sd.kind != 0u && sv2p1 != 0u ? 1u << 6 : 0u)
                              | (sd.kind != 0u && sd.pos != 0u ? 1u << 7 : 0u);
                    for (std::uint32_t k = 0; k < sd.n_terms; ++k) {
                        ro.t[k] = static_cast<std::uint16_t>(unit(sd.terms[k]));
                    }
                    if (sd.kind == 0u) {
                        ro.dst = unit(sd.out);
                    } else {
                        ro.dst = sd.pos != 0u ? unit(sd.pos - 1u) : 0xffffu;
                        ro.sv = sv1 | ((sv2p1 != 0u ? sv2p1 - 1u : 0xffffu) << 16);
                        ro.cidx = sd.cidx;
                    }
                }
                r.table.push_back(ro);
            }
            ++r.n_rounds;
        }
        r.round_level_end |= 1u << (r.n_rounds - 1u);
    }
    return r;
}

} // namespace heyoka_b200::detail
