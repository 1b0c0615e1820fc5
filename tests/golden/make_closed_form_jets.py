#!/usr/bin/env python3
"""Expected Taylor jets of the reference's closed-form jet tests (test/taylor_*.cpp, the batch-3 / order-3 blocks
listed in tests/closed_form_cases.py), computed independently of any Taylor recurrence: the normalised derivatives
x^[k] = (1/k!) d^k x / dt^k are obtained by differentiating the right-hand side symbolically along the flow (sympy)
and evaluated with 40 significant digits. The reference's tests hand-write the same closed forms
(e.g. test/taylor_sincos.cpp:482-506) and accept 100 epsilon (test/test_utils.hpp:47-80).

    python tests/golden/make_closed_form_jets.py      # writes tests/golden/closed_form_jets.json
"""
import json
import os
import sys

import sympy as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from closed_form_cases import CASES, ORDER, batch_of  # noqa: E402


class sympy_backend:
    sin, cos, tanh, exp, log, sqrt = sp.sin, sp.cos, sp.tanh, sp.exp, sp.log, sp.sqrt
    T = sp.Symbol("t")

    @staticmethod
    def square(e):
        return e ** 2

    @staticmethod
    def sigmoid(e):
        return 1 / (1 + sp.exp(-e))

    @staticmethod
    def relu(e, slope=0):
        return sp.Piecewise((e, e > 0), (sp.nsimplify(slope) * e, True))

    @staticmethod
    def relup(e, slope=0):
        return sp.Piecewise((sp.Integer(1), e > 0), (sp.nsimplify(slope), True))

    @staticmethod
    def pow(b, e):  # noqa: A003
        return b ** e

    @staticmethod
    def c(v):
        return sp.Integer(v)

    @classmethod
    def t(cls):
        return cls.T


def main():
    x, y = sp.symbols("x y")
    t = sympy_backend.T
    out = {"order": ORDER, "cases": []}
    for case in CASES:
        name, cite, rhs, state, time = case
        BATCH = batch_of(case)
        f = [sp.sympify(e) for e in rhs(sympy_backend, x, y)]
        # derivs[k][i] = d^k x_i / dt^k as an expression of (x, y, t)
        derivs = [[x, y], f]
        for _ in range(2, ORDER + 1):
            prev = derivs[-1]
            derivs.append([sp.diff(e, x) * f[0] + sp.diff(e, y) * f[1] + sp.diff(e, t) for e in prev])
        tc = [[[None] * BATCH for _ in range(ORDER + 1)] for _ in range(2)]
        for lane in range(BATCH):
            subs = {x: sp.nsimplify(state[lane]), y: sp.nsimplify(state[BATCH + lane]),
                    t: sp.Integer(time[lane]) if time else sp.Integer(0)}
            for k in range(ORDER + 1):
                for i in range(2):
                    val = sp.N(derivs[k][i].subs(subs) / sp.factorial(k), 40)
                    tc[i][k][lane] = float(val)  # (Piecewise: the branch of the initial conditions)
        out["cases"].append({"name": name, "cite": cite, "state": state, "time": time, "tc": tc})
    with open(os.path.join(HERE, "closed_form_jets.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote %d cases" % len(out["cases"]))


if __name__ == "__main__":
    main()
