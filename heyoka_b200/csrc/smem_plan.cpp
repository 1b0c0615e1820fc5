// See smem_plan.hpp.
#include "smem_plan.hpp"

#include <algorithm>
#include <numeric>
#include <stdexcept>

namespace heyoka_b200::detail
{

namespace
{

// How one op uses u variables at order n: `now` = read at order n only, `hist` = read at lower orders too.
struct op_uses {
    std::vector<std::uint32_t> now, hist;
    bool self_hist = false;
};

op_uses uses_of(const hy_program &p, const hy_op &op)
{
    op_uses u;
    const auto nary = [&](bool hist) {
        for (std::uint32_t k = 0; k < op.b; ++k) {
            const auto ref = p.args[op.a + k];
            if (HY_REF_KIND(ref) == HY_REF_VAR) {
                (hist ? u.hist : u.now).push_back(HY_REF_IDX(ref));
            }
        }
    };
    switch (op.opcode) {
        case HY_OP_SUM:
            nary(false);
            break;
        case HY_OP_SUM_SQ:
            nary(true);
            break;
        case HY_OP_SUB_VV:
            u.now = {op.a, op.b};
            break;
        case HY_OP_SUB_VN:
        case HY_OP_SUB_VP:
        case HY_OP_DIV_VN:
        case HY_OP_DIV_VP:
        case HY_OP_NEG:
            u.now = {op.a};
            break;
        case HY_OP_SUB_NV:
        case HY_OP_SUB_PV:
        case HY_OP_MUL_NV:
        case HY_OP_MUL_PV:
            u.now = {op.b};
            break;
        case HY_OP_MUL_VV:
            u.hist = {op.a, op.b};
            break;
        case HY_OP_DIV_VV:
            u.now = {op.a};
            u.hist = {op.b};
            u.self_hist = true;
            break;
        case HY_OP_DIV_NV:
        case HY_OP_DIV_PV:
            u.hist = {op.b};
            u.self_hist = true;
            break;
        case HY_OP_SQUARE:
            u.hist = {op.a};
            break;
        case HY_OP_SQRT:
            u.now = {op.a};
            u.self_hist = true;
            break;
        case HY_OP_POW_VN:
        case HY_OP_POW_VP:
        case HY_OP_EXP:
        case HY_OP_LOG:
            u.hist = {op.a};
            u.self_hist = true;
            break;
        case HY_OP_SIN:
        case HY_OP_COS:
        case HY_OP_TANH:
            // op.c is the hidden dependency: read at lower orders only, but it needs its history.
            u.hist = {op.a};
            break;
        default:
            break;
    }
    return u;
}

} // namespace

smem_plan make_smem_plan(const hy_program &p)
{
    smem_plan pl;
    const auto n_eq = p.n_eq, n_uvars = p.n_uvars, order = p.order;
    const auto n_ops = n_uvars - n_eq;

    // ---- history analysis ----
    std::vector<char> hist(n_uvars, 0);
    for (std::uint32_t i = 0; i < n_ops; ++i) {
        const auto &op = p.ops[i];
        const auto u = uses_of(p, op);
        for (const auto v : u.hist) {
            hist[v] = 1;
        }
        if (u.self_hist) {
            hist[n_eq + i] = 1;
        }
        if (op.opcode == HY_OP_SIN || op.opcode == HY_OP_COS || op.opcode == HY_OP_TANH) {
            hist[op.c] = 1;
        }
    }

    // ---- slot assignment ----
    // Odd history stride: consecutive history rows then start in different shared-memory bank groups.
    const std::uint32_t hstride = (order + 1u) | 1u;
    std::vector<std::uint32_t> row(n_uvars, 0);
    std::uint32_t next = 0;
    const auto alloc = [&](std::uint32_t kind, std::uint32_t n) {
        const auto r = (next << 2) | kind;
        next += n;
        return r;
    };
    for (std::uint32_t i = 0; i < n_uvars; ++i) {
        if (hist[i]) {
            row[i] = alloc(ROW_H, hstride);
        } else if (i < n_eq) {
            row[i] = alloc(ROW_SV, 2u);
        } else {
            row[i] = alloc(ROW_T, 1u);
        }
    }
    if (next >= (1u << 29)) {
        throw std::overflow_error("The Taylor tape is too large");
    }
    pl.n_slots = next;
    pl.sv_rows.assign(row.begin(), row.begin() + n_eq);

    // ---- segments (src/taylor_02.cpp:105-207): a new one starts when an op reads, at the current order,
    // a u variable defined in the current segment. Hidden dependencies are not dependencies. ----
    std::vector<std::uint32_t> seg_begin{0};
    std::uint32_t cur_limit = n_eq;
    for (std::uint32_t i = 0; i < n_ops; ++i) {
        const auto u = uses_of(p, p.ops[i]);
        bool dep = false;
        for (const auto v : u.now) {
            dep = dep || v >= cur_limit;
        }
        for (const auto v : u.hist) {
            // convolution operands are read at order n too (e.g. b^[n] c^[0])
            dep = dep || v >= cur_limit;
        }
        if (dep) {
            seg_begin.push_back(i);
            cur_limit = n_eq + i;
        }
    }
    seg_begin.push_back(n_ops);
    pl.n_segments = static_cast<std::uint32_t>(seg_begin.size() - 1u);

    // ---- n-ary argument table and state-variable definitions with row references ----
    pl.args = p.args;
    // Which args entries belong to constant-function ops (no variables there) is irrelevant: only VAR refs change.
    for (auto &ref : pl.args) {
        if (HY_REF_KIND(ref) == HY_REF_VAR) {
            ref = HY_REF(HY_REF_VAR, row[HY_REF_IDX(ref)]);
        }
    }
    pl.sv_defs = p.sv_defs;
    for (auto &ref : pl.sv_defs) {
        if (HY_REF_KIND(ref) == HY_REF_VAR) {
            ref = HY_REF(HY_REF_VAR, row[HY_REF_IDX(ref)]);
        }
    }

    // ---- ops: segment by segment, grouped by opcode inside a segment (warp-uniform control flow) ----
    pl.seg_offsets.push_back(0);
    for (std::uint32_t s = 0; s < pl.n_segments; ++s) {
        std::vector<std::uint32_t> idx(seg_begin[s + 1u] - seg_begin[s]);
        std::iota(idx.begin(), idx.end(), seg_begin[s]);
        std::stable_sort(idx.begin(), idx.end(),
                         [&](std::uint32_t x, std::uint32_t y) { return p.ops[x].opcode < p.ops[y].opcode; });
        pl.max_seg_width = std::max<std::uint32_t>(pl.max_seg_width, static_cast<std::uint32_t>(idx.size()));
        for (const auto i : idx) {
            auto op = p.ops[i];
            const auto var = [&](std::uint32_t &f) { f = row[f]; };
            switch (op.opcode) {
                case HY_OP_SUB_VV:
                case HY_OP_MUL_VV:
                case HY_OP_DIV_VV:
                    var(op.a);
                    var(op.b);
                    break;
                case HY_OP_SUB_VN:
                case HY_OP_SUB_VP:
                case HY_OP_DIV_VN:
                case HY_OP_DIV_VP:
                case HY_OP_NEG:
                case HY_OP_SQUARE:
                case HY_OP_SQRT:
                case HY_OP_POW_VN:
                case HY_OP_POW_VP:
                case HY_OP_EXP:
                case HY_OP_LOG:
                    var(op.a);
                    break;
                case HY_OP_SUB_NV:
                case HY_OP_SUB_PV:
                case HY_OP_MUL_NV:
                case HY_OP_MUL_PV:
                case HY_OP_DIV_NV:
                case HY_OP_DIV_PV:
                    var(op.b);
                    break;
                case HY_OP_SIN:
                case HY_OP_COS:
                case HY_OP_TANH:
                    var(op.a);
                    var(op.c);
                    break;
                default:
                    // SUM / SUM_SQ / CFUNC go through the argument table, TIME has no operands.
                    break;
            }
            pl.ops.push_back(op);
            pl.dst.push_back(row[n_eq + i]);
        }
        pl.seg_offsets.push_back(static_cast<std::uint32_t>(pl.ops.size()));
    }

    return pl;
}

} // namespace heyoka_b200::detail
