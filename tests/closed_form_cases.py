"""The systems and initial conditions of the reference's closed-form jet tests (test/taylor_*.cpp), batch 3,
tol = .1 (order 3), as backend-neutral lambdas: `m` provides sin/cos/.../pow/c(onstant)/t(ime); x, y are the state
variables (batch 3; the relu blocks of the reference use batch 2). Used by tests/golden/make_closed_form_jets.py (sympy backend -> expected jets) and by the tests
(heyoka_b200 backend -> oracle / GPU jets). State layout as in the reference: [x lanes..., y lanes...]."""

CASES = [
    # (name, reference file:line of the batch-3 block, rhs(m, x, y) -> (x', y'), state, time or None)
    ("sincos_num", "test/taylor_sincos.cpp:253", lambda m, x, y: (m.sin(m.c(2)) + m.cos(m.c(3)), x + y),
     [2, -4, -1, 3, 5, -2], None),
    ("sincos_var", "test/taylor_sincos.cpp:461", lambda m, x, y: (m.sin(y), m.cos(x)), [2, -1, -5, 3, -4, 6], None),
    ("tanh_num", "test/taylor_tanh.cpp:266", lambda m, x, y: (m.tanh(m.c(2)), x + y), [2, -4, -1, 3, 5, -2], None),
    ("tanh_var", "test/taylor_tanh.cpp:462", lambda m, x, y: (m.tanh(y), m.tanh(x)), [2, -1, -5, 3, -4, 6], None),
    ("exp_num", "test/taylor_exp.cpp:231", lambda m, x, y: (m.exp(m.c(2)), x + y), [2, -2, 1, 3, -3, 0], None),
    ("exp_var", "test/taylor_exp.cpp:425", lambda m, x, y: (m.exp(y), m.exp(x)), [2, 4, 3, 3, 5, 6], None),
    ("log_num", "test/taylor_log.cpp:231", lambda m, x, y: (m.log(m.c(2)), x + y), [2, -2, 1, 3, -3, 0], None),
    ("log_var", "test/taylor_log.cpp:425", lambda m, x, y: (m.log(y), m.log(x)), [2, 4, 3, 3, 5, 6], None),
    ("sqrt_num", "test/taylor_sqrt.cpp:202", lambda m, x, y: (m.sqrt(m.c(2)), x + y), [2, -2, 1, 3, -3, 0], None),
    ("sqrt_var", "test/taylor_sqrt.cpp:396", lambda m, x, y: (m.sqrt(y), m.sqrt(x)), [2, 4, 3, 3, 5, 6], None),
    ("square_num", "test/taylor_square.cpp:218", lambda m, x, y: (m.square(m.c(2)), x + y), [2, -2, 1, 3, -3, 0],
     None),
    ("square_var", "test/taylor_square.cpp:433", lambda m, x, y: (m.square(y), m.square(x)), [2, 4, 3, 3, 5, 6], None),
    ("pow_num", "test/taylor_pow.cpp:315", lambda m, x, y: (m.pow(m.c(3), m.c(1) / m.c(3)), x + y),
     [2, -1, -4, 3, 5, 6], None),
    ("pow_var", "test/taylor_pow.cpp:574",
     lambda m, x, y: (m.pow(y, m.c(3) / m.c(2)), m.pow(x, m.c(-1) / m.c(3))), [2, 5, 1, 3, 4, 6], None),
    ("div_num", "test/taylor_div.cpp:210", lambda m, x, y: (m.c(1) / m.c(3), x + y), [2, 1, -6, 3, -4, 2], None),
    ("div_var_num", "test/taylor_div.cpp:445", lambda m, x, y: (y / m.c(2), x / m.c(-4)), [2, 1, -5, 3, -4, 2], None),
    ("div_num_var", "test/taylor_div.cpp:685", lambda m, x, y: (m.c(2) / y, m.c(-4) / x), [2, -4, 1, 3, 5, -2], None),
    ("div_var_var", "test/taylor_div.cpp:895", lambda m, x, y: (x / y, y / x), [2, -5, 1, 3, 4, -2], None),
    ("mul_num", "test/taylor_mul.cpp:194", lambda m, x, y: (m.c(2) * m.c(3), x + y), [2, -2, -1, 3, 2, 4], None),
    ("mul_var_num", "test/taylor_mul.cpp:416", lambda m, x, y: (y * m.c(2), x * m.c(-4)), [2, -1, 0, 3, 4, -5], None),
    ("mul_num_var", "test/taylor_mul.cpp:650", lambda m, x, y: (m.c(2) * y, m.c(-4) * x), [2, -1, 0, 3, 4, -5], None),
    ("mul_var_var", "test/taylor_mul.cpp:843", lambda m, x, y: (x * y, y * x), [2, 1, 3, 3, -4, 6], None),
    ("sub_num", "test/taylor_sub.cpp:210", lambda m, x, y: (m.c(2) - m.c(3), x + y), [2, -2, -1, 3, 2, 4], None),
    ("sub_var_num", "test/taylor_sub.cpp:431", lambda m, x, y: (y - m.c(2), x - m.c(-4)), [2, -1, 0, 3, 4, -5], None),
    ("sub_num_var", "test/taylor_sub.cpp:666", lambda m, x, y: (m.c(2) - y, m.c(-4) - x), [2, -1, 0, 3, 4, -5], None),
    ("sub_var_var", "test/taylor_sub.cpp:860", lambda m, x, y: (x - y, y - x), [2, 1, 3, 3, -4, 6], None),
    ("sum_sq_num", "test/taylor_sum_sq.cpp:240",
     lambda m, x, y: (m.square(m.c(2)) + m.square(m.c(3)) + m.square(m.c(1)), x + y), [2, -2, 1, 3, -3, 0], None),
    ("sum_sq_var", "test/taylor_sum_sq.cpp:435",
     lambda m, x, y: (m.square(y) + m.square(x) + m.square(m.c(1)), m.square(x) + m.square(y) + m.square(m.c(2))),
     [2, 4, 3, 3, 5, 6], None),
    ("neg_num", "test/taylor_neg.cpp:205", lambda m, x, y: (-m.c(2), x + y), [2, -2, 1, 3, -3, 0], None),
    ("neg_var", "test/taylor_neg.cpp:399", lambda m, x, y: (-y, -x), [2, 4, 3, 3, 5, 6], None),
    ("sigmoid_num", "test/taylor_sigmoid.cpp:284", lambda m, x, y: (m.sigmoid(m.c(2)), x + y), [2, -4, -1, 3, 5, -2],
     None),
    ("sigmoid_var", "test/taylor_sigmoid.cpp:480", lambda m, x, y: (m.sigmoid(y), m.sigmoid(x)),
     [2, -1, -5, 3, -4, 6], None),
    # relu / relup and their leaky variants (batch 2 in the reference).
    ("relu_var", "test/taylor_relu.cpp:112", lambda m, x, y: (m.relu(x) + m.relup(y), x + y), [2, -1, 3, 5], None),
    ("leaky_relu_var", "test/taylor_relu.cpp:223",
     lambda m, x, y: (m.relu(x, 0.01) + m.relup(y, 0.02), x + y), [2, -1, -3, 5], None),
    ("time", "test/taylor_time.cpp:197", lambda m, x, y: (m.t() + x, x + y), [2, -2, 1, 3, -3, 0], [-5, 6, -1]),
]

# Tolerance in units of epsilon: 100 (test/test_utils.hpp:51) unless the reference's block says otherwise
# (1 - tanh^2 cancels for |x| >= 4: test/taylor_tanh.cpp:489-511 uses 10000).
# sigmoid: a - a^2 cancels for |x| >= 5 the same way; the reference's expected values (test/taylor_sigmoid.cpp:508-
# 530) are formed from the computed jets and share that rounding, the exact closed forms used here do not.
EPS_MUL = {"tanh_var": 10000.0, "sigmoid_var": 1000.0}

BATCH = 3  # (2 for the relu blocks: batch_of(case))
TOL = 0.1  # -> order 3 (include/heyoka/detail/taylor_common.hpp:165-191)
ORDER = 3


class hb_backend:
    """Builds heyoka_b200 expressions."""

    def __init__(self, hb):
        self.hb = hb
        for f in ("sin", "cos", "tanh", "exp", "log", "sqrt", "square", "pow", "sigmoid", "relu", "relup"):
            setattr(self, f, getattr(hb, f))

    def c(self, v):
        return self.hb.expression(float(v))

    def t(self):
        return self.hb.time


def hb_system(hb, case):
    x, y = hb.make_vars("x", "y")
    rx, ry = case[2](hb_backend(hb), x, y)
    return [(x, rx), (y, ry)]


def batch_of(case):
    return len(case[3]) // 2
