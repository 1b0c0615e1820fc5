// Taylor decomposition: see include/heyoka_b200/taylor_decompose.hpp for the reference map.
#include <heyoka_b200/taylor_decompose.hpp>

#include <algorithm>
#include <cassert>
#include <cmath>
#include <deque>
#include <functional>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <unordered_map>
#include <unordered_set>

namespace heyoka_b200
{

std::uint32_t uname_to_index(const std::string &s)
{
    assert(s.size() > 2u && s[0] == 'u' && s[1] == '_');
    return static_cast<std::uint32_t>(std::stoul(s.substr(2)));
}

namespace
{

std::string uname(std::size_t i)
{
    return "u_" + std::to_string(i);
}

expression uvar(std::size_t i)
{
    return expression{variable{uname(i)}};
}

using func_cache_t = std::unordered_map<const void *, expression>;

// Post-order node transformation with a cache on node identity, so that shared subexpressions
// stay shared after the transformation (src/detail/ex_traversal.cpp:35-180).
expression transform_nodes(func_cache_t &cache, const expression &e,
                           const std::function<expression(const expression &)> &leaf_tfunc,
                           const std::function<expression(const expression &)> &branch_tfunc)
{
    struct frame {
        const expression *ex;
        bool visited;
    };
    std::vector<frame> stack;
    std::vector<std::optional<expression>> out;

    stack.push_back({&e, false});

    while (!stack.empty()) {
        const auto [cur, visited] = stack.back();
        stack.pop_back();

        if (cur->is_func()) {
            const auto *id = cur->fn_id();

            if (!visited) {
                if (const auto it = cache.find(id); it != cache.end()) {
                    out.emplace_back(it->second);
                    continue;
                }
                stack.push_back({cur, true});
                for (const auto &a : cur->fn().args) {
                    stack.push_back({&a, false});
                }
                out.emplace_back();
            } else {
                const auto &f = cur->fn();
                std::vector<expression> new_args;
                new_args.reserve(f.args.size());
                for (std::size_t i = 0; i < f.args.size(); ++i) {
                    new_args.push_back(std::move(*out.back()));
                    out.pop_back();
                }
                expression copy{f.kind, std::move(new_args)};
                if (branch_tfunc) {
                    copy = branch_tfunc(copy);
                }
                cache.emplace(id, copy);
                out.back().emplace(std::move(copy));
            }
        } else {
            out.emplace_back(leaf_tfunc ? leaf_tfunc(*cur) : *cur);
        }
    }

    assert(out.size() == 1u);
    return std::move(*out.back());
}

std::vector<expression> transform_all(const std::vector<expression> &v,
                                      const std::function<expression(const expression &)> &leaf_tfunc,
                                      const std::function<expression(const expression &)> &branch_tfunc)
{
    func_cache_t cache;
    std::vector<expression> ret;
    ret.reserve(v.size());
    for (const auto &e : v) {
        ret.push_back(transform_nodes(cache, e, leaf_tfunc, branch_tfunc));
    }
    return ret;
}

bool is_kind(const expression &e, func_kind k)
{
    return e.is_func() && e.fn().kind == k;
}

// ---- pass 2: x**y -> exp(y*log(x)) when y is not a number (src/taylor_01.cpp:806-840) ----
std::vector<expression> pow_to_explog(const std::vector<expression> &v)
{
    return transform_all(v, {}, [](const expression &ex) {
        if (is_kind(ex, func_kind::pow) && !ex.fn().args[1].is_number()) {
            const auto &args = ex.fn().args;
            // NOTE: log node built directly (no constant folding of the base).
            return exp(args[1] * expression{func_kind::log, {args[0]}});
        }
        return ex;
    });
}

// ---- pass 3: a + (-1*b) -> sub(a, b) (src/math/sum.cpp:461-544) ----
std::vector<expression> sum_to_sub(const std::vector<expression> &v)
{
    return transform_all(v, {}, [](const expression &ex) {
        if (!is_kind(ex, func_kind::sum)) {
            return ex;
        }

        auto new_args = ex.fn().args;
        const auto fpart = [](const expression &arg) {
            if (is_kind(arg, func_kind::prod) && arg.fn().args.size() >= 2u && arg.fn().args[0].is_number()) {
                return !(arg.fn().args[0].num() == -1.);
            }
            return true;
        };
        const auto it = std::stable_partition(new_args.begin(), new_args.end(), fpart);

        if (it == new_args.end()) {
            return ex;
        }

        std::vector<expression> sub_args;
        for (auto jt = it; jt != new_args.end(); ++jt) {
            const auto &f = jt->fn();
            std::vector<expression> tmp(f.args.begin() + 1, f.args.end());
            sub_args.push_back(prod(std::move(tmp)));
        }

        auto st = sum(std::move(sub_args));

        if (it == new_args.begin()) {
            return prod({expression{-1.}, std::move(st)});
        }

        new_args.erase(it, new_args.end());
        auto mend = sum(std::move(new_args));

        return expression{func_kind::sub, {std::move(mend), std::move(st)}};
    });
}

// ---- passes 4 & 7: nested split of associative n-ary functions (udf_split.hpp:50-98) ----
expression udf_split(const expression &e, func_kind kind, std::uint32_t split)
{
    assert(split >= 2u);

    if (!is_kind(e, kind) || e.fn().args.size() <= split) {
        return e;
    }

    std::vector<expression> ret_seq, tmp;
    for (const auto &arg : e.fn().args) {
        tmp.push_back(arg);
        if (tmp.size() == split) {
            ret_seq.emplace_back(kind, std::move(tmp));
            tmp.clear();
        }
    }

    if (!tmp.empty()) {
        if (tmp.size() == 1u) {
            ret_seq.push_back(std::move(tmp[0]));
        } else {
            ret_seq.emplace_back(kind, std::move(tmp));
        }
    }

    return udf_split(expression{kind, std::move(ret_seq)}, kind, split);
}

std::vector<expression> split_sums(const std::vector<expression> &v)
{
    // Power of two, so that the pairwise sums inside each chunk round the same way
    // (src/expression_basic.cpp:1183-1186).
    return transform_all(v, {}, [](const expression &ex) { return udf_split(ex, func_kind::sum, 8); });
}

std::vector<expression> split_prods(const std::vector<expression> &v, std::uint32_t split)
{
    return transform_all(v, {}, [split](const expression &ex) { return udf_split(ex, func_kind::prod, split); });
}

// ---- pass 5: sum(x**2, y**2, ...) -> sum_sq(x, y, ...) (src/math/sum.cpp:387-455) ----
const expression *is_square(const expression &ex)
{
    if (!is_kind(ex, func_kind::pow)) {
        return nullptr;
    }
    const auto &args = ex.fn().args;
    if (args[1].is_number() && args[1].num() == 2.) {
        return &args[0];
    }
    return nullptr;
}

std::vector<expression> sums_to_sum_sqs(const std::vector<expression> &v)
{
    return transform_all(v, {}, [](const expression &ex) {
        if (!is_kind(ex, func_kind::sum)) {
            return ex;
        }
        std::vector<expression> new_args;
        new_args.reserve(ex.fn().args.size());
        for (const auto &arg : ex.fn().args) {
            const auto *sq = is_square(arg);
            if (sq == nullptr) {
                return ex;
            }
            new_args.push_back(*sq);
        }
        return expression{func_kind::sum_sq, std::move(new_args)};
    });
}

// ---- pass 6: prod(..., pow(z, -1)) -> div(..., z) (src/math/prod.cpp:753-908) ----
std::vector<expression> prod_to_div(const std::vector<expression> &v)
{
    return transform_all(v, {}, [](const expression &ex) {
        if (!is_kind(ex, func_kind::prod)) {
            return ex;
        }

        auto new_args = ex.fn().args;
        const auto fpart = [](const expression &e) {
            if (!is_kind(e, func_kind::pow)) {
                return true;
            }
            const auto &expo = e.fn().args[1];
            return !(expo.is_number() && expo.num() == -1.);
        };
        const auto it = std::stable_partition(new_args.begin(), new_args.end(), fpart);

        if (it == new_args.end()) {
            return ex;
        }

        std::vector<expression> div_args;
        for (auto jt = it; jt != new_args.end(); ++jt) {
            const auto &f = jt->fn();
            div_args.push_back(pow(f.args[0], expression{-f.args[1].num()}));
        }

        auto divisor = prod(std::move(div_args));
        new_args.erase(it, new_args.end());
        auto num = prod(std::move(new_args));

        return expression{func_kind::div, {std::move(num), std::move(divisor)}};
    });
}

std::vector<expression> rename_variables(const std::vector<expression> &v,
                                         const std::unordered_map<std::string, std::string> &repl)
{
    return transform_all(
        v,
        [&repl](const expression &leaf) {
            if (leaf.is_variable()) {
                if (const auto it = repl.find(leaf.var_name()); it != repl.end()) {
                    return expression{variable{it->second}};
                }
            }
            return leaf;
        },
        {});
}

// Rename on an already-decomposed definition (args are leaves): no cache needed.
expression rename_shallow(const expression &ex, const std::unordered_map<std::string, std::string> &repl)
{
    const auto ren = [&repl](const expression &leaf) {
        if (leaf.is_variable()) {
            if (const auto it = repl.find(leaf.var_name()); it != repl.end()) {
                return expression{variable{it->second}};
            }
        }
        return leaf;
    };

    if (!ex.is_func()) {
        return ren(ex);
    }
    std::vector<expression> args;
    args.reserve(ex.fn().args.size());
    for (const auto &a : ex.fn().args) {
        assert(!a.is_func());
        args.push_back(ren(a));
    }
    return expression{ex.fn().kind, std::move(args)};
}

// ---- per-function decomposition hook (src/func.cpp:392-420) ----
std::size_t func_taylor_decompose(expression fn, taylor_dc_t &dc)
{
    const auto &f = fn.fn();
    std::size_t ret = 0;

    switch (f.kind) {
        case func_kind::sin: {
            // cos first, then sin, cross-linked (src/math/sin.cpp:115-133).
            dc.emplace_back(expression{func_kind::cos, {f.args[0]}}, std::vector<std::uint32_t>{});
            dc.emplace_back(std::move(fn), std::vector<std::uint32_t>{});
            (dc.end() - 2)->second.push_back(static_cast<std::uint32_t>(dc.size() - 1u));
            (dc.end() - 1)->second.push_back(static_cast<std::uint32_t>(dc.size() - 2u));
            ret = dc.size() - 1u;
            break;
        }
        case func_kind::cos: {
            // sin first, then cos (src/math/cos.cpp:116-134).
            dc.emplace_back(expression{func_kind::sin, {f.args[0]}}, std::vector<std::uint32_t>{});
            dc.emplace_back(std::move(fn), std::vector<std::uint32_t>{});
            (dc.end() - 2)->second.push_back(static_cast<std::uint32_t>(dc.size() - 1u));
            (dc.end() - 1)->second.push_back(static_cast<std::uint32_t>(dc.size() - 2u));
            ret = dc.size() - 1u;
            break;
        }
        case func_kind::tanh: {
            // tanh, then tanh**2 as hidden dependency (src/math/tanh.cpp:77-92).
            dc.emplace_back(std::move(fn), std::vector<std::uint32_t>{});
            dc.emplace_back(pow(uvar(dc.size() - 1u), expression{2.}), std::vector<std::uint32_t>{});
            (dc.end() - 2)->second.push_back(static_cast<std::uint32_t>(dc.size() - 1u));
            ret = dc.size() - 2u;
            break;
        }
        case func_kind::sigmoid: {
            // sigmoid, then sigmoid**2 as hidden dependency (src/math/sigmoid.cpp:99-113).
            dc.emplace_back(std::move(fn), std::vector<std::uint32_t>{});
            dc.emplace_back(pow(uvar(dc.size() - 1u), expression{2.}), std::vector<std::uint32_t>{});
            (dc.end() - 2)->second.push_back(static_cast<std::uint32_t>(dc.size() - 1u));
            ret = dc.size() - 2u;
            break;
        }
        default:
            ret = dc.size();
            dc.emplace_back(std::move(fn), std::vector<std::uint32_t>{});
    }

    if (ret == 0u || ret >= dc.size()) {
        throw std::invalid_argument("Invalid value returned by the Taylor decomposition of a function");
    }
    return ret;
}

// ---- depth-first decomposition of one expression (src/expression_decompose.cpp:45-209) ----
// NOTE: arguments are pushed in order and popped from the back, hence the LAST argument of a
// function is decomposed first, exactly like in the reference.
std::optional<std::size_t> taylor_decompose(std::unordered_map<const void *, std::size_t> &func_map,
                                            const expression &e, taylor_dc_t &dc)
{
    struct frame {
        const expression *ex;
        bool visited;
    };
    std::vector<frame> stack;
    // outer optional: slot filled or not; inner optional: index into dc or leaf.
    std::vector<std::optional<std::optional<std::size_t>>> out;

    stack.push_back({&e, false});

    while (!stack.empty()) {
        const auto [cur, visited] = stack.back();
        stack.pop_back();

        if (cur->is_func()) {
            const auto *id = cur->fn_id();

            if (!visited) {
                if (const auto it = func_map.find(id); it != func_map.end()) {
                    out.emplace_back(std::optional<std::size_t>{it->second});
                    continue;
                }
                stack.push_back({cur, true});
                for (const auto &a : cur->fn().args) {
                    stack.push_back({&a, false});
                }
                out.emplace_back();
            } else {
                const auto &f = cur->fn();
                std::vector<expression> new_args;
                new_args.reserve(f.args.size());
                for (std::size_t i = 0; i < f.args.size(); ++i) {
                    const auto opt_idx = *out.back();
                    if (opt_idx) {
                        new_args.push_back(uvar(*opt_idx));
                    } else {
                        new_args.push_back(f.args[i]);
                    }
                    out.pop_back();
                }

                const auto ret = func_taylor_decompose(expression{f.kind, std::move(new_args)}, dc);
                func_map.emplace(id, ret);
                out.back().emplace(std::optional<std::size_t>{ret});
            }
        } else {
            out.emplace_back(std::optional<std::size_t>{});
        }
    }

    assert(out.size() == 1u);
    return *out.back();
}

// ---- CSE (src/taylor_01.cpp:315-443) ----
struct ex_hash {
    std::size_t operator()(const expression &e) const
    {
        return hash_value(e);
    }
};

void taylor_decompose_cse(taylor_dc_t &dc, std::vector<std::uint32_t> &sv_funcs_dc, std::size_t n_eq)
{
    taylor_dc_t new_dc;
    std::unordered_map<expression, std::size_t, ex_hash> ex_map;
    std::unordered_map<std::string, std::string> rename;

    for (std::size_t i = 0; i < n_eq; ++i) {
        new_dc.push_back(dc[i]);
        rename.emplace(uname(i), uname(i));
    }

    for (auto i = n_eq; i < dc.size() - n_eq; ++i) {
        auto new_ex = rename_shallow(dc[i].first, rename);

        if (const auto it = ex_map.find(new_ex); it == ex_map.end()) {
            new_dc.emplace_back(new_ex, dc[i].second);
            ex_map.emplace(std::move(new_ex), new_dc.size() - 1u);
            rename.emplace(uname(i), uname(new_dc.size() - 1u));
        } else {
            rename.emplace(uname(i), uname(it->second));
        }
    }

    for (auto i = dc.size() - n_eq; i < dc.size(); ++i) {
        new_dc.emplace_back(rename_shallow(dc[i].first, rename), dc[i].second);
    }

    for (auto &p : new_dc) {
        for (auto &idx : p.second) {
            idx = uname_to_index(rename.at(uname(idx)));
        }
    }
    for (auto &idx : sv_funcs_dc) {
        idx = uname_to_index(rename.at(uname(idx)));
    }

    dc = std::move(new_dc);
}

// ---- breadth-first topological re-sort (src/taylor_01.cpp:454-645) ----
// Vertex 0 is a virtual root; vertex i+1 is u_i. Edges only from the explicit variables of a
// definition (hidden deps are not edges). Kahn's algorithm with out-edges visited in increasing
// target order.
void taylor_sort_dc(taylor_dc_t &dc, std::vector<std::uint32_t> &sv_funcs_dc, std::size_t n_eq)
{
    const auto n_u = dc.size() - n_eq;
    const auto n_v = n_u + 1u;

    std::vector<std::vector<std::size_t>> out_edges(n_v);
    std::vector<std::size_t> in_deg(n_v, 0);

    for (std::size_t i = 0; i < n_eq; ++i) {
        out_edges[0].push_back(i + 1u);
        ++in_deg[i + 1u];
    }

    for (auto i = n_eq; i < n_u; ++i) {
        const auto vars = get_variables(dc[i].first);
        if (vars.empty()) {
            out_edges[0].push_back(i + 1u);
            ++in_deg[i + 1u];
        } else {
            for (const auto &var : vars) {
                const auto idx = uname_to_index(var);
                out_edges[idx + 1u].push_back(i + 1u);
                ++in_deg[i + 1u];
            }
        }
    }

    std::vector<std::size_t> v_idx;
    v_idx.reserve(dc.size() + 1u);
    std::deque<std::size_t> tmp;
    tmp.push_back(0);

    while (!tmp.empty()) {
        const auto v = tmp.front();
        tmp.pop_front();
        v_idx.push_back(v);

        auto &edges = out_edges[v];
        std::sort(edges.begin(), edges.end());
        for (const auto t : edges) {
            if (--in_deg[t] == 0u) {
                tmp.push_back(t);
            }
        }
    }

    assert(v_idx.size() == n_v);

    for (std::size_t i = 0; i + 1u < v_idx.size(); ++i) {
        v_idx[i] = v_idx[i + 1u] - 1u;
    }
    v_idx.resize(dc.size());
    for (auto i = n_u; i < dc.size(); ++i) {
        v_idx[i] = i;
    }

    std::unordered_map<std::string, std::string> remap;
    for (std::size_t i = 0; i < n_eq; ++i) {
        assert(v_idx[i] == i);
        remap.emplace(uname(i), uname(i));
    }
    for (auto i = n_eq; i < n_u; ++i) {
        remap.emplace(uname(v_idx[i]), uname(i));
    }

    taylor_dc_t new_dc;
    new_dc.reserve(dc.size());
    for (const auto idx : v_idx) {
        const auto &[ex, deps] = dc[idx];
        std::vector<std::uint32_t> new_deps;
        new_deps.reserve(deps.size());
        for (const auto d : deps) {
            new_deps.push_back(uname_to_index(remap.at(uname(d))));
        }
        new_dc.emplace_back(rename_shallow(ex, remap), std::move(new_deps));
    }

    for (auto &idx : sv_funcs_dc) {
        idx = uname_to_index(remap.at(uname(idx)));
    }

    dc = std::move(new_dc);
}

// ---- bare numbers -> num_identity (src/taylor_01.cpp:788-803) ----
void replace_numbers(taylor_dc_t &dc, std::size_t n_eq)
{
    for (auto i = n_eq; i < dc.size() - n_eq; ++i) {
        auto &[ex, deps] = dc[i];
        if (ex.is_number()) {
            ex = expression{func_kind::num_identity, {ex}};
            deps.clear();
        }
    }
}

} // namespace

void validate_ode_sys(const std::vector<std::pair<expression, expression>> &sys, const std::vector<expression> &ev_funcs)
{
    if (sys.empty()) {
        throw std::invalid_argument("Cannot integrate a system of zero equations");
    }

    std::unordered_set<std::string> lhs_vars;
    for (const auto &[lhs, rhs] : sys) {
        if (!lhs.is_variable()) {
            throw std::invalid_argument(
                "Error in the left-hand side of an ODE system: the left-hand side contains the expression '"
                + to_string(lhs) + "', which is not a variable");
        }
        if (!lhs_vars.insert(lhs.var_name()).second) {
            throw std::invalid_argument("Error in the left-hand side of an ODE system: the variable '" + lhs.var_name()
                                        + "' appears twice");
        }
    }

    for (const auto &[lhs, rhs] : sys) {
        for (const auto &var : get_variables(rhs)) {
            if (lhs_vars.find(var) == lhs_vars.end()) {
                throw std::invalid_argument("Error in the right-hand side of an ODE system: the variable '" + var
                                            + "' appears in the right-hand side but not in the left-hand side");
            }
        }
    }

    // The event functions may only contain state variables (src/detail/validate_ode_sys.cpp:131-145).
    for (const auto &ev : ev_funcs) {
        for (const auto &var : get_variables(ev)) {
            if (lhs_vars.find(var) == lhs_vars.end()) {
                throw std::invalid_argument("Invalid system of differential equations detected: an event function "
                                            "contains the variable '"
                                            + var + "', which is not a state variable");
            }
        }
    }
}

std::pair<taylor_dc_t, std::vector<std::uint32_t>>
taylor_decompose_sys(const std::vector<std::pair<expression, expression>> &sys, const std::vector<expression> &sv_funcs)
{
    const auto n_eq = sys.size();

    std::unordered_map<std::string, std::string> repl_map;
    for (std::size_t i = 0; i < n_eq; ++i) {
        repl_map.emplace(sys[i].first.var_name(), uname(i));
    }

    std::vector<expression> all_ex;
    all_ex.reserve(n_eq + sv_funcs.size());
    for (const auto &p : sys) {
        all_ex.push_back(p.second);
    }
    all_ex.insert(all_ex.end(), sv_funcs.begin(), sv_funcs.end());

    all_ex = pow_to_explog(all_ex);
    all_ex = sum_to_sub(all_ex);
    all_ex = split_sums(all_ex);
    all_ex = sums_to_sum_sqs(all_ex);
    all_ex = prod_to_div(all_ex);
    all_ex = split_prods(all_ex, 2);
    all_ex = rename_variables(all_ex, repl_map);

    taylor_dc_t u_vars_defs;
    u_vars_defs.reserve(n_eq);
    for (const auto &p : sys) {
        u_vars_defs.emplace_back(p.first, std::vector<std::uint32_t>{});
    }

    taylor_dc_t outs;
    outs.reserve(n_eq);

    std::unordered_map<const void *, std::size_t> func_map;
    for (std::size_t i = 0; i < n_eq; ++i) {
        const auto &ex = all_ex[i];
        if (const auto dres = taylor_decompose(func_map, ex, u_vars_defs)) {
            outs.emplace_back(uvar(*dres), std::vector<std::uint32_t>{});
        } else {
            outs.emplace_back(ex, std::vector<std::uint32_t>{});
        }
    }

    std::vector<std::uint32_t> sv_funcs_dc;
    for (auto i = n_eq; i < all_ex.size(); ++i) {
        const auto &sv_ex = all_ex[i];
        if (sv_ex.is_variable()) {
            sv_funcs_dc.push_back(uname_to_index(sv_ex.var_name()));
        } else if (const auto dres = taylor_decompose(func_map, sv_ex, u_vars_defs)) {
            sv_funcs_dc.push_back(static_cast<std::uint32_t>(*dres));
        } else {
            throw std::invalid_argument(
                "The extra functions in a Taylor decomposition cannot be constants or parameters");
        }
    }

    u_vars_defs.insert(u_vars_defs.end(), outs.begin(), outs.end());

    taylor_decompose_cse(u_vars_defs, sv_funcs_dc, n_eq);
    taylor_sort_dc(u_vars_defs, sv_funcs_dc, n_eq);
    // NOTE: the reference's sincos_combine_taylor() (src/detail/sincos_combine.cpp:95-99) only
    // changes how sin/cos of the same argument are *evaluated* at order 0 (one sincos call); the
    // device code always evaluates the pair with one sincos(), so there is nothing to rewrite.
    replace_numbers(u_vars_defs, n_eq);

    return {std::move(u_vars_defs), std::move(sv_funcs_dc)};
}

std::string dc_to_string(const taylor_dc_t &dc)
{
    std::ostringstream oss;
    for (std::size_t i = 0; i < dc.size(); ++i) {
        oss << "u_" << i << " = " << dc[i].first;
        if (!dc[i].second.empty()) {
            oss << "  [deps:";
            for (auto d : dc[i].second) {
                oss << ' ' << d;
            }
            oss << ']';
        }
        oss << '\n';
    }
    return oss.str();
}

} // namespace heyoka_b200
