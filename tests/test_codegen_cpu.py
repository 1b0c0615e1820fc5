"""Pin the code-generating CPU baseline (oracle/codegen.py: the reference's straight-line, per-order unrolled SIMD stepper
restated through gcc instead of LLVM) against the same known answers as the interpreting oracle (CPU only):
  1. doc/tut_batch_mode.rst printed output (tests/golden/tut_batch_mode.json): step sizes, states, step counts 34/38/41/44
     and 40/38/35/34, the full order-20 array of Taylor coefficients;
  2. test/taylor_adaptive_batch.cpp:586-598 (exact 100000 / 220000 step counts under max_delta_t);
  3. the interpreting oracle in the reference's default (pairwise) summation mode on the 6-body system and on random
     right-hand sides covering every opcode (tolerance: a few ulp, the generated code is compiled with contraction
     allowed, like the reference's LLVM setting src/llvm_state.cpp:842-845).
Both SIMD widths the bench reports (4 and 8 lanes) are checked."""
import os
import sys

import numpy as np
import pytest

import heyoka_b200 as hb
import oracle
from common import (decimals_equal, golden, nbody_rel_err, outer_ss_batch_state, sig_digits_equal, sys_outer_ss, sys_pendulum,
                    sys_tutorial, sys_two_body, two_body_batch_state)

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import codegen  # noqa: E402

OC = {"success": hb.taylor_outcome.success, "time_limit": hb.taylor_outcome.time_limit}
WIDTHS = [4, 8]


class installed:
    """Context manager: the generated jet of program P is the one the oracle's driver runs (for width W)."""

    def __init__(self, P, W):
        self.jet, self.W = codegen.Jet(P, W), W

    def __enter__(self):
        self.jet.install(oracle.lib)
        return self.jet

    def __exit__(self, *a):
        codegen.Jet.uninstall(oracle.lib, self.W)


@pytest.mark.parametrize("W", WIDTHS)
def test_tutorial_batch_mode_codegen(W):
    g = golden("tut_batch_mode.json")
    P = hb.Program(sys_tutorial())
    with installed(P, W):
        o = oracle.OracleIntegrator(P, [g["x0"], g["v0"]], 4, pars=[g["alpha"]], mode=oracle.PAIRWISE, width=W)
        o.step()
        assert [int(x) for x in o.step_outcome] == [OC[r["outcome"]] for r in g["first_step"]]
        assert sig_digits_equal(o.last_h, [r["h"] for r in g["first_step"]])
        assert decimals_equal(o.state, g["states"][0])
        o.step(g["clamped_step_limits"])
        assert np.all(o.last_h == np.array(g["clamped_step_limits"]))
        assert decimals_equal(o.state, g["states"][1])
        hi, lo = hb._dfloat_add(o.t_hi, o.t_lo, np.array(g["propagate_for"]["delta_ts"]), np.zeros(4))
        o.propagate_until(hi, lo)
        res = g["propagate_for"]["res"]
        assert [int(x) for x in o.n_steps] == [r["n_steps"] for r in res]          # 34, 38, 41, 44
        assert sig_digits_equal(o.min_h, [r["min_h"] for r in res])
        assert sig_digits_equal(o.max_h, [r["max_h"] for r in res])
        assert decimals_equal(o.state, g["states"][2])
        o.propagate_until(g["propagate_until"]["ts"])
        res = g["propagate_until"]["res"]
        assert [int(x) for x in o.n_steps] == [r["n_steps"] for r in res]          # 40, 38, 35, 34
        assert decimals_equal(o.state, g["states"][3])
        assert np.all(o.t_hi == np.array(g["propagate_until"]["ts"]))
        o.step(write_tc=True)
        assert sig_digits_equal(o.tc, g["tc_after_final_step"], 7)


def test_exact_step_counts_codegen():
    P = hb.Program(sys_pendulum())
    with installed(P, 8):
        o = oracle.OracleIntegrator(P, [[0.05, 0.06], [0.025, 0.026]], 2, mode=oracle.PAIRWISE, width=8)
        o.propagate_until([10., 11.], max_delta_t=[1e-4, 5e-5])
        assert np.all(o.t_hi == [10., 11.])
        assert [int(x) for x in o.n_steps] == [100000, 220000]


@pytest.mark.parametrize("W", WIDTHS)
def test_outer_ss_codegen_matches_interpreter(W):
    """6-body system (the bench workload): Taylor coefficients of a step to a few ulp of the interpreting oracle
    (default-mode summation), identical step counts and 1e-12 on the state after 50 years."""
    P = hb.Program(sys_outer_ss(), high_accuracy=True)
    st = outer_ss_batch_state(2 * W + 3)
    ref = oracle.OracleIntegrator(P, st, st.shape[1], mode=oracle.PAIRWISE)
    ref.step(write_tc=True)
    with installed(P, W):
        o = oracle.OracleIntegrator(P, st, st.shape[1], mode=oracle.PAIRWISE, width=W)
        o.step(write_tc=True)
        assert np.max(np.abs(o.last_h / ref.last_h - 1)) < 1e-13
        scale = np.max(np.abs(ref.tc), axis=0, keepdims=True)
        assert np.max(np.abs(o.tc - ref.tc) / scale) < 1e-13
        o.propagate_until(50.0, lockstep=False)
    ref.propagate_until(50.0)
    assert np.array_equal(o.n_steps, ref.n_steps)
    assert nbody_rel_err(o.state, ref.state) < 1e-12


def test_two_body_codegen():
    P = hb.Program(sys_two_body())
    st = two_body_batch_state(16)
    ref = oracle.OracleIntegrator(P, st, 16, mode=oracle.PAIRWISE)
    ref.step(write_tc=True)
    with installed(P, 8):
        o = oracle.OracleIntegrator(P, st, 16, mode=oracle.PAIRWISE, width=8)
        o.step(write_tc=True)
    assert np.max(np.abs(o.tc - ref.tc)) < 1e-13 * max(1.0, np.max(np.abs(ref.tc)))


@pytest.mark.parametrize("seed", range(8))
def test_random_expressions_codegen(seed):
    """Every opcode the generator emits (sums, sub, products, quotients, sin/cos/tanh/exp/log/sqrt/square/pow/sigmoid)
    on seeded random right-hand sides: generated Taylor coefficients against the interpreting oracle, weighted by
    their contribution to the step."""
    from test_random_expressions import build_case
    sys_, _, _, ic = build_case(seed)
    P = hb.Program(sys_, tol=1e-12)
    ref = oracle.OracleIntegrator(P, ic, 3, mode=oracle.PAIRWISE)
    ref.step(write_tc=True)
    with installed(P, 4):
        o = oracle.OracleIntegrator(P, ic, 3, mode=oracle.PAIRWISE, width=4)
        o.step(write_tc=True)
    w = np.abs(ref.last_h)[None, None, :] ** np.arange(P.order + 1)[None, :, None]
    scale = np.maximum(np.max(np.abs(ref.tc[:, 0, :]), axis=0), 1.0)[None, None, :]
    assert np.max(np.abs(o.tc - ref.tc) * w / scale) < 1e-13, seed
    assert np.max(np.abs(o.last_h / ref.last_h - 1)) < 1e-11


@pytest.mark.parametrize("W", WIDTHS)
@pytest.mark.parametrize("ha", [False, True])
@pytest.mark.parametrize("mode", [oracle.PAIRWISE, oracle.FMA])
def test_simd_tail_equals_scalar_tail(W, ha, mode):
    """The timed ports deduce the step size and update the state on SIMD vectors (oracle/taylor_oracle.c
    step_tail_v(), like the reference's JIT-compiled step function src/taylor_00.cpp:102-460); the per-lane scalar
    tail (determine_h() / update_state()) stays selectable: both give the same bits - in particular the compensated
    summation of high_accuracy is not contracted - on full and on ragged groups of lanes, with step limits of both
    signs, and over a propagation."""
    P = hb.Program(sys_outer_ss(), high_accuracy=ha)
    n = 2 * W + 3
    st = outer_ss_batch_state(n)
    lim = np.where(np.arange(n) % 3 == 0, 0.11, np.inf) * np.where(np.arange(n) % 5 == 4, -1., 1.)
    a = oracle.OracleIntegrator(P, st, n, mode=mode, width=W)
    b = oracle.OracleIntegrator(P, st, n, mode=mode | oracle.SCALAR_TAIL, width=W)
    for o in (a, b):
        o.step(lim, write_tc=True)
    assert np.array_equal(a.last_h, b.last_h) and np.array_equal(a.state, b.state) and np.array_equal(a.tc, b.tc)
    assert np.array_equal(a.step_outcome, b.step_outcome)
    assert np.count_nonzero(a.last_h == lim) >= n // 3 - 1 and np.any(a.last_h < 0)
    tf = a.t_hi + np.where(a.last_h < 0, -7.0, 7.0)
    for o in (a, b):
        o.propagate_until(tf, lockstep=False)
    assert np.array_equal(a.n_steps, b.n_steps) and np.array_equal(a.state, b.state)
    assert np.array_equal(a.t_hi, b.t_hi) and np.array_equal(a.t_lo, b.t_lo)


@pytest.mark.parametrize("W", WIDTHS)
def test_tutorial_ensemble_codegen(W):
    """doc/tut_ensemble.rst (tests/golden/tut_ensemble.json): the generated SIMD stepper (with the SIMD tail) on the ten
    initial conditions of the tutorial as lanes - full and ragged groups; member 9 after propagate_until(20) as the
    reference prints it (17 digits) to 2e-14, 124 steps. doc/tut_adaptive.rst: time and state after the first step to
    1e-14, 24 + 72 + 97 steps."""
    g = golden("tut_ensemble.json")
    P = hb.Program(sys_pendulum())
    with installed(P, W):
        o = oracle.OracleIntegrator(P, np.array(g["ics"]).T.copy(), g["n_iter"], mode=oracle.PAIRWISE, width=W)
        o.propagate_until(g["t_final"], lockstep=False)
        m = g["member"]
        assert int(o.n_steps[m]) == g["n_steps"] and np.all(o.t_hi == 20.)
        assert sig_digits_equal(o.min_h[m], g["min_h"]) and sig_digits_equal(o.max_h[m], g["max_h"])
        assert np.max(np.abs(o.state[:, m] / np.array(g["state"]) - 1)) < 2e-14
        a = golden("tut_adaptive.json")
        o = oracle.OracleIntegrator(P, [[a["x0"]] * 3, [a["v0"]] * 3], 3, mode=oracle.PAIRWISE, width=W)
        o.step()
        assert np.max(np.abs(o.t_hi / a["first_step"]["time"] - 1)) < 1e-14
        assert np.max(np.abs(o.state / np.array(a["first_step"]["state"])[:, None] - 1)) < 1e-14
        o.state[:] = np.array([[a["x0"]] * 3, [a["v0"]] * 3])
        o.t_hi[:] = 0
        for r, tf in zip(a["propagate"], (5., 20., 0.)):
            o.propagate_until(tf, lockstep=False)
            assert [int(x) for x in o.n_steps] == [r["n_steps"]] * 3
