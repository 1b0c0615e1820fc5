// Lowering of a Taylor decomposition to the flat opcode program consumed by the device kernels.
//
// The reference does the equivalent work when it groups the u variables of a segment by
// (function, argument kinds) and builds the per-call argument tables of compact mode
// (src/taylor_02.cpp:830-953, src/detail/cm_utils.cpp:79-157); the specialisation by argument kind
// mirrors the taylor_diff overload sets (e.g. src/math/prod.cpp:316-410, src/detail/div.cpp:64-160).
#include "program.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>

namespace heyoka_b200::detail
{

namespace
{

struct lowering_ctx {
    hy_program &p;

    std::uint32_t add_const(double v)
    {
        // Deduplicate bit-identical constants.
        for (std::size_t i = 0; i < p.consts.size(); ++i) {
            if (std::memcmp(&p.consts[i], &v, sizeof(double)) == 0) {
                return static_cast<std::uint32_t>(i);
            }
        }
        p.consts.push_back(v);
        return static_cast<std::uint32_t>(p.consts.size() - 1u);
    }

    std::uint32_t ref(const expression &e)
    {
        if (e.is_variable()) {
            return HY_REF(HY_REF_VAR, uname_to_index(e.var_name()));
        }
        if (e.is_number()) {
            return HY_REF(HY_REF_NUM, add_const(e.num()));
        }
        if (e.is_param()) {
            return HY_REF(HY_REF_PAR, e.par_idx());
        }
        throw std::invalid_argument("Function argument found in a decomposed expression: " + to_string(e));
    }

    std::uint32_t add_args(const std::vector<expression> &args)
    {
        const auto off = static_cast<std::uint32_t>(p.args.size());
        for (const auto &a : args) {
            p.args.push_back(ref(a));
        }
        return off;
    }
};

char kind_of(const expression &e)
{
    return e.is_variable() ? 'V' : (e.is_number() ? 'N' : 'P');
}

bool all_const(const std::vector<expression> &args)
{
    for (const auto &a : args) {
        if (a.is_variable()) {
            return false;
        }
    }
    return true;
}

// get_pow_eval_algo(), src/math/pow.cpp:292-355.
std::uint32_t pow_eval_algo(double e)
{
    constexpr double max_small = 16;
    if (std::isfinite(e) && e == std::trunc(e)) {
        if (e >= 0 && e <= max_small) {
            return (HY_POW_POS_SMALL_INT << 8) | static_cast<std::uint32_t>(e);
        }
        if (e < 0 && -e <= max_small) {
            return (HY_POW_NEG_SMALL_INT << 8) | static_cast<std::uint32_t>(-e);
        }
    } else if (std::isfinite(e)) {
        const auto y = 2 * e;
        if (y == std::trunc(y)) {
            if (y >= 0 && y <= max_small) {
                return (HY_POW_POS_SMALL_HALF << 8) | static_cast<std::uint32_t>(y);
            }
            if (y < 0 && -y <= max_small) {
                return (HY_POW_NEG_SMALL_HALF << 8) | static_cast<std::uint32_t>(-y);
            }
        }
    }
    return HY_POW_GENERAL << 8;
}

std::uint32_t cfunc_of(func_kind k)
{
    switch (k) {
        case func_kind::num_identity:
            return HY_CF_IDENTITY;
        case func_kind::sum:
            return HY_CF_SUM;
        case func_kind::prod:
            return HY_CF_PROD;
        case func_kind::sub:
            return HY_CF_SUB;
        case func_kind::div:
            return HY_CF_DIV;
        case func_kind::pow:
            return HY_CF_POW;
        case func_kind::sum_sq:
            return HY_CF_SUM_SQ;
        case func_kind::sin:
            return HY_CF_SIN;
        case func_kind::cos:
            return HY_CF_COS;
        case func_kind::tanh:
            return HY_CF_TANH;
        case func_kind::exp:
            return HY_CF_EXP;
        case func_kind::log:
            return HY_CF_LOG;
        case func_kind::sigmoid:
            return HY_CF_SIGMOID;
        case func_kind::relu:
            return HY_CF_RELU;
        case func_kind::relup:
            return HY_CF_RELUP;
        default:
            throw not_implemented_error(std::string("Constant folding of function '") + func_kind_name(k)
                                        + "' is not implemented");
    }
}

} // namespace

std::uint32_t taylor_order_from_tol(double tol)
{
    // include/heyoka/detail/taylor_common.hpp:165-191.
    auto order_f = std::ceil(-std::log(tol) / 2 + 1);
    if (!std::isfinite(order_f)) {
        throw std::invalid_argument(
            "The computation of the Taylor order in an adaptive Taylor stepper produced a non-finite value");
    }
    order_f = std::max(2., order_f);
    if (order_f > static_cast<double>(std::numeric_limits<std::uint32_t>::max())) {
        throw std::overflow_error(
            "The computation of the Taylor order in an adaptive Taylor stepper resulted in an overflow condition");
    }
    return static_cast<std::uint32_t>(order_f);
}

hy_program lower_decomposition(const taylor_dc_t &dc, std::uint32_t n_eq, std::uint32_t n_pars, std::uint32_t order,
                               bool high_accuracy)
{
    if (dc.size() < 2u * static_cast<std::size_t>(n_eq)) {
        throw std::invalid_argument("Invalid Taylor decomposition: too few entries");
    }
    if (dc.size() - n_eq > 0x3fffffffu) {
        throw std::overflow_error("The Taylor decomposition is too large");
    }

    hy_program p;
    p.n_eq = n_eq;
    p.n_uvars = static_cast<std::uint32_t>(dc.size() - n_eq);
    p.n_pars = n_pars;
    p.order = order;
    p.high_accuracy = high_accuracy;
    p.dc = dc;

    lowering_ctx ctx{p};

    for (std::uint32_t i = n_eq; i < p.n_uvars; ++i) {
        const auto &[ex, deps] = dc[i];
        if (!ex.is_func()) {
            throw std::invalid_argument("Invalid Taylor decomposition: u_" + std::to_string(i)
                                        + " is not a function");
        }
        const auto &f = ex.fn();
        const auto &a = f.args;
        hy_op op{0, 0, 0, 0};

        const auto need_args = [&](std::size_t n) {
            if (a.size() != n) {
                throw std::invalid_argument(std::string("Invalid number of arguments for '") + func_kind_name(f.kind)
                                            + "' in a Taylor decomposition");
            }
        };
        const auto need_dep = [&]() {
            if (deps.size() != 1u) {
                throw std::invalid_argument(std::string("The function '") + func_kind_name(f.kind)
                                            + "' needs exactly one hidden dependency");
            }
            return deps[0];
        };

        if (f.kind == func_kind::time) {
            op.opcode = HY_OP_TIME;
        } else if (all_const(a)) {
            // All arguments are numbers/params (include/heyoka/detail/taylor_common.hpp:88-157).
            op.opcode = HY_OP_CFUNC;
            op.a = cfunc_of(f.kind);
            op.b = ctx.add_args(a);
            op.c = static_cast<std::uint32_t>(a.size());
        } else {
            switch (f.kind) {
                case func_kind::sum:
                    op.opcode = HY_OP_SUM;
                    op.a = ctx.add_args(a);
                    op.b = static_cast<std::uint32_t>(a.size());
                    break;
                case func_kind::sum_sq:
                    op.opcode = HY_OP_SUM_SQ;
                    op.a = ctx.add_args(a);
                    op.b = static_cast<std::uint32_t>(a.size());
                    break;
                case func_kind::sub: {
                    need_args(2);
                    const std::string k{kind_of(a[0]), kind_of(a[1])};
                    op.opcode = k == "VV"   ? HY_OP_SUB_VV
                                : k == "VN" ? HY_OP_SUB_VN
                                : k == "NV" ? HY_OP_SUB_NV
                                : k == "VP" ? HY_OP_SUB_VP
                                            : HY_OP_SUB_PV;
                    op.a = HY_REF_IDX(ctx.ref(a[0]));
                    op.b = HY_REF_IDX(ctx.ref(a[1]));
                    break;
                }
                case func_kind::prod: {
                    if (a.size() != 2u) {
                        throw std::invalid_argument(
                            "The Taylor derivative of a product can be computed only for products of 2 terms, but "
                            "the current product has "
                            + std::to_string(a.size()) + " term(s) instead");
                    }
                    const std::string k{kind_of(a[0]), kind_of(a[1])};
                    if (k == "VV") {
                        op.opcode = HY_OP_MUL_VV;
                        op.a = HY_REF_IDX(ctx.ref(a[0]));
                        op.b = HY_REF_IDX(ctx.ref(a[1]));
                    } else {
                        // number/param times variable, in either order (src/math/prod.cpp:348-373).
                        const auto &c = a[0].is_variable() ? a[1] : a[0];
                        const auto &v = a[0].is_variable() ? a[0] : a[1];
                        if (c.is_number() && c.num() == -1.) {
                            op.opcode = HY_OP_NEG;
                            op.a = HY_REF_IDX(ctx.ref(v));
                        } else {
                            op.opcode = c.is_number() ? HY_OP_MUL_NV : HY_OP_MUL_PV;
                            op.a = HY_REF_IDX(ctx.ref(c));
                            op.b = HY_REF_IDX(ctx.ref(v));
                        }
                    }
                    break;
                }
                case func_kind::div: {
                    need_args(2);
                    const std::string k{kind_of(a[0]), kind_of(a[1])};
                    op.opcode = k == "VV"   ? HY_OP_DIV_VV
                                : k == "NV" ? HY_OP_DIV_NV
                                : k == "PV" ? HY_OP_DIV_PV
                                : k == "VN" ? HY_OP_DIV_VN
                                            : HY_OP_DIV_VP;
                    op.a = HY_REF_IDX(ctx.ref(a[0]));
                    op.b = HY_REF_IDX(ctx.ref(a[1]));
                    break;
                }
                case func_kind::pow: {
                    need_args(2);
                    if (!a[0].is_variable()) {
                        // number ** variable was rewritten to exp(y*log(x)) by pow_to_explog().
                        throw std::invalid_argument("An invalid argument type was encountered while trying to build "
                                                    "the Taylor derivative of a pow()");
                    }
                    op.a = HY_REF_IDX(ctx.ref(a[0]));
                    if (a[1].is_number()) {
                        const auto e = a[1].num();
                        const auto algo = pow_eval_algo(e);
                        if (algo == ((HY_POW_POS_SMALL_INT << 8) | 2u)) {
                            op.opcode = HY_OP_SQUARE;
                        } else if (algo == ((HY_POW_POS_SMALL_HALF << 8) | 1u)) {
                            op.opcode = HY_OP_SQRT;
                        } else {
                            op.opcode = HY_OP_POW_VN;
                            op.b = ctx.add_const(e);
                            op.c = algo;
                        }
                    } else if (a[1].is_param()) {
                        op.opcode = HY_OP_POW_VP;
                        op.b = a[1].par_idx();
                    } else {
                        throw std::invalid_argument("An invalid argument type was encountered while trying to build "
                                                    "the Taylor derivative of a pow()");
                    }
                    break;
                }
                case func_kind::sin:
                    need_args(1);
                    op.opcode = HY_OP_SIN;
                    op.a = HY_REF_IDX(ctx.ref(a[0]));
                    op.c = need_dep();
                    break;
                case func_kind::cos:
                    need_args(1);
                    op.opcode = HY_OP_COS;
                    op.a = HY_REF_IDX(ctx.ref(a[0]));
                    op.c = need_dep();
                    break;
                case func_kind::tanh:
                    need_args(1);
                    op.opcode = HY_OP_TANH;
                    op.a = HY_REF_IDX(ctx.ref(a[0]));
                    op.c = need_dep();
                    break;
                case func_kind::sigmoid:
                    need_args(1);
                    op.opcode = HY_OP_SIGMOID;
                    op.a = HY_REF_IDX(ctx.ref(a[0]));
                    op.c = need_dep();
                    break;
                case func_kind::relu:
                case func_kind::relup:
                    need_args(2);
                    if (!a[1].is_number()) {
                        throw std::invalid_argument("The slope of a ReLU must be a number");
                    }
                    op.opcode = f.kind == func_kind::relu ? HY_OP_RELU : HY_OP_RELUP;
                    op.a = HY_REF_IDX(ctx.ref(a[0]));
                    op.b = ctx.add_const(a[1].num());
                    break;
                case func_kind::exp:
                    need_args(1);
                    op.opcode = HY_OP_EXP;
                    op.a = HY_REF_IDX(ctx.ref(a[0]));
                    break;
                case func_kind::log:
                    need_args(1);
                    op.opcode = HY_OP_LOG;
                    op.a = HY_REF_IDX(ctx.ref(a[0]));
                    break;
                default:
                    throw not_implemented_error(std::string("Taylor derivative of function '")
                                                + func_kind_name(f.kind) + "' is not implemented");
            }
        }

        p.ops.push_back(op);
    }

    for (auto i = dc.size() - n_eq; i < dc.size(); ++i) {
        p.sv_defs.push_back(ctx.ref(dc[i].first));
    }

    validate_program(p);

    return p;
}

void validate_program(const hy_program &p)
{
    const auto fail = [](const std::string &msg) { throw std::invalid_argument("Invalid program: " + msg); };

    if (p.n_eq == 0u || p.n_uvars < p.n_eq) {
        fail("inconsistent n_eq/n_uvars");
    }
    if (p.order < 2u) {
        fail("the Taylor order must be at least 2");
    }
    if (p.ops.size() != p.n_uvars - p.n_eq) {
        fail("the number of ops must be n_uvars - n_eq");
    }
    if (p.sv_defs.size() != p.n_eq) {
        fail("the number of state variable definitions must be n_eq");
    }
    // Overflow check on the tape size, like src/taylor_02.cpp:1227-1233.
    if (static_cast<std::uint64_t>(p.n_uvars) * (p.order + 1u) > 0x7fffffffull) {
        throw std::overflow_error("An overflow condition was detected while computing the size of the Taylor tape");
    }

    const auto check_ref = [&](std::uint32_t r, std::uint32_t cur, const char *what) {
        const auto idx = HY_REF_IDX(r);
        switch (HY_REF_KIND(r)) {
            case HY_REF_VAR:
                if (idx >= cur) {
                    fail(std::string(what) + ": a u variable is used before its definition");
                }
                break;
            case HY_REF_NUM:
                if (idx >= p.consts.size()) {
                    fail(std::string(what) + ": constant index out of range");
                }
                break;
            case HY_REF_PAR:
                if (idx >= p.n_pars) {
                    fail(std::string(what) + ": parameter index out of range");
                }
                break;
            default:
                fail(std::string(what) + ": invalid reference kind");
        }
    };
    const auto var = [&](std::uint32_t idx, std::uint32_t cur) { check_ref(HY_REF(HY_REF_VAR, idx), cur, "op"); };
    const auto num = [&](std::uint32_t idx) { check_ref(HY_REF(HY_REF_NUM, idx), 0, "op"); };
    const auto par = [&](std::uint32_t idx) { check_ref(HY_REF(HY_REF_PAR, idx), 0, "op"); };

    for (std::uint32_t i = 0; i < p.ops.size(); ++i) {
        const auto cur = p.n_eq + i;
        const auto &op = p.ops[i];
        switch (op.opcode) {
            case HY_OP_SUM:
            case HY_OP_SUM_SQ:
                if (op.b == 0u || static_cast<std::uint64_t>(op.a) + op.b > p.args.size()) {
                    fail("n-ary argument table out of range");
                }
                if (op.b > 8u && op.opcode == HY_OP_SUM) {
                    fail("sums must have at most 8 terms");
                }
                for (std::uint32_t k = 0; k < op.b; ++k) {
                    check_ref(p.args[op.a + k], cur, "n-ary op");
                }
                break;
            case HY_OP_SUB_VV:
            case HY_OP_MUL_VV:
            case HY_OP_DIV_VV:
                var(op.a, cur);
                var(op.b, cur);
                break;
            case HY_OP_SUB_VN:
            case HY_OP_DIV_VN:
                var(op.a, cur);
                num(op.b);
                break;
            case HY_OP_SUB_NV:
            case HY_OP_MUL_NV:
            case HY_OP_DIV_NV:
                num(op.a);
                var(op.b, cur);
                break;
            case HY_OP_SUB_VP:
            case HY_OP_DIV_VP:
                var(op.a, cur);
                par(op.b);
                break;
            case HY_OP_SUB_PV:
            case HY_OP_MUL_PV:
            case HY_OP_DIV_PV:
                par(op.a);
                var(op.b, cur);
                break;
            case HY_OP_NEG:
            case HY_OP_SQUARE:
            case HY_OP_SQRT:
            case HY_OP_EXP:
            case HY_OP_LOG:
                var(op.a, cur);
                break;
            case HY_OP_POW_VN:
                var(op.a, cur);
                num(op.b);
                if ((op.c >> 8) > HY_POW_NEG_SMALL_HALF) {
                    fail("invalid pow evaluation algorithm");
                }
                break;
            case HY_OP_POW_VP:
                var(op.a, cur);
                par(op.b);
                break;
            case HY_OP_RELU:
            case HY_OP_RELUP:
                var(op.a, cur);
                if (op.b >= p.consts.size()) {
                    fail("constant index out of range");
                }
                break;
            case HY_OP_SIN:
            case HY_OP_COS:
            case HY_OP_TANH:
            case HY_OP_SIGMOID:
                var(op.a, cur);
                // The hidden dependency may come right after the op (sin/cos pairs, tanh -> tanh^2).
                if (op.c < p.n_eq || op.c >= p.n_uvars || op.c == cur) {
                    fail("hidden dependency out of range");
                }
                break;
            case HY_OP_TIME:
                break;
            case HY_OP_CFUNC:
                if (op.a > HY_CF_RELUP) {
                    fail("invalid constant function");
                }
                if (op.c == 0u || static_cast<std::uint64_t>(op.b) + op.c > p.args.size()) {
                    fail("constant function argument table out of range");
                }
                for (std::uint32_t k = 0; k < op.c; ++k) {
                    if (HY_REF_KIND(p.args[op.b + k]) == HY_REF_VAR) {
                        fail("constant function with a variable argument");
                    }
                    check_ref(p.args[op.b + k], cur, "constant function");
                }
                break;
            default:
                throw not_implemented_error("Unknown opcode " + std::to_string(op.opcode));
        }
    }

    // Hidden dependencies must be mutually consistent: a forward dependency (index > cur) is only
    // legal if, at every order, it can be computed from data available before it is read. sin/cos
    // read each other's lower orders only, tanh reads tanh^2 lower orders only: both fine.
    for (std::uint32_t i = 0; i < p.n_eq; ++i) {
        check_ref(p.sv_defs[i], p.n_uvars, "state variable definition");
    }
}

program_costs compute_costs(const hy_program &p)
{
    const double n_eq = p.n_eq, n_pars = p.n_pars, n_uvars = p.n_uvars, ord = p.order;
    program_costs c{};
    // SURVEY.md §8(d).
    c.b_min = 8. * (2. * n_eq + n_pars + 7.);
    c.b_tape = c.b_min + 16. * (n_uvars * ord + n_eq);

    // Flop model: one fused multiply-add = 2 flops; per op, summed over orders 0..p-1.
    double fl = 0;
    const auto conv = [&](double per_term, double extra) {
        // sum_{n=1}^{p-1} (n * per_term + extra)
        double s = 0;
        for (std::uint32_t n = 1; n < p.order; ++n) {
            s += n * per_term + extra;
        }
        return s;
    };
    for (const auto &op : p.ops) {
        switch (op.opcode) {
            case HY_OP_SUM:
                fl += (op.b - 1.) * ord;
                break;
            case HY_OP_SUM_SQ:
                fl += op.b * conv(1., 2.) + (op.b - 1.) * ord;
                break;
            case HY_OP_MUL_VV:
                fl += conv(2., 2.);
                break;
            case HY_OP_DIV_VV:
            case HY_OP_DIV_NV:
            case HY_OP_DIV_PV:
                fl += conv(2., 2.);
                break;
            case HY_OP_SQUARE:
                fl += conv(1., 2.);
                break;
            case HY_OP_SQRT:
                fl += conv(1., 4.);
                break;
            case HY_OP_POW_VN:
            case HY_OP_POW_VP:
                fl += conv(6., 3.);
                break;
            case HY_OP_SIN:
            case HY_OP_COS:
            case HY_OP_TANH:
            case HY_OP_EXP:
            case HY_OP_LOG:
            case HY_OP_SIGMOID:
                fl += conv(3., 2.);
                break;
            case HY_OP_TIME:
            case HY_OP_CFUNC:
                break;
            default:
                fl += ord;
        }
    }
    // State-variable derivatives (one division each per order), step-size estimate, state update.
    fl += n_eq * ord + 3. * n_eq + (p.high_accuracy ? 6. : 2.) * n_eq * ord;
    c.flops = fl;
    return c;
}

} // namespace heyoka_b200::detail

hy_program_desc hy_program::desc() const
{
    hy_program_desc d{};
    d.n_eq = n_eq;
    d.n_uvars = n_uvars;
    d.n_pars = n_pars;
    d.order = order;
    d.n_args = static_cast<std::uint32_t>(args.size());
    d.n_consts = static_cast<std::uint32_t>(consts.size());
    d.high_accuracy = high_accuracy ? 1 : 0;
    d.ops = ops.data();
    d.args = args.data();
    d.consts = consts.data();
    d.sv_defs = sv_defs.data();
    d.n_ev = static_cast<std::uint32_t>(ev_defs.size());
    d.ev_defs = ev_defs.empty() ? nullptr : ev_defs.data();
    return d;
}
