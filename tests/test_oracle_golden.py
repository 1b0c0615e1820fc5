"""Pin the CPU oracle against the reference's own known answers (CPU only, no CUDA calls).

Fixtures:
  1. doc/tut_batch_mode.rst printed output (tests/golden/tut_batch_mode.json)
  2. README.md scalar pendulum (tests/golden/readme_pendulum.json)
  3. closed-form jets of test/taylor_*.cpp (tests/closed_form_cases.py, tests/golden/closed_form_jets.json)
  4. test/timestep_check.cpp:33-86 (step-size formula recomputed from the Taylor coefficients)
  5. test/taylor_adaptive_batch.cpp:586-598 (exact step counts under max_delta_t, exact final times)
  6. doc/tut_adaptive.rst and doc/tut_d_output.rst printed outputs (tests/golden/tut_adaptive.json, tut_d_output.json:
     state after one step to 16 digits, step counts 24 / 72 / 97, reversibility to an ulp, propagate_grid, dense and
     continuous output: 48 recorded steps and samples); the scalar integrators of those pages as batches
  7. doc/tut_ensemble.rst (the ensemble's members as the lanes of one batch: member 9 to 16 digits, 124 steps),
     doc/tut_param.rst (runtime parameters), doc/tut_nonauto.rst (time-dependent right-hand side, 25 printed values),
     doc/tut_adaptive_custom.rst (tol = 1e-9: order 12, the printed state after 0 -> 10 -> 0)
All three summation modes of the oracle must satisfy them, like the reference sweeps compact_mode/opt_level.
"""
import numpy as np
import pytest

import heyoka_b200 as hb
import oracle
from common import (approx, decimals_equal, golden, outer_ss_ic, sig_digits_equal, sys_outer_ss, sys_pendulum, sys_tutorial)

MODES = [oracle.PAIRWISE, oracle.SEQ, oracle.FMA]
OC = {"success": hb.taylor_outcome.success, "time_limit": hb.taylor_outcome.time_limit}


@pytest.mark.parametrize("mode", MODES)
def test_readme_pendulum(mode):
    g = golden("readme_pendulum.json")
    P = hb.Program(sys_pendulum())
    assert P.order == 20
    o = oracle.OracleIntegrator(P, [g["x0"], g["v0"]], 1, mode=mode)
    o.propagate_until(g["t"])
    assert o.t_hi[0] == 10.0
    assert sig_digits_equal(o.state[0, 0], g["x"])
    assert sig_digits_equal(o.state[1, 0], g["v"])


@pytest.mark.parametrize("mode", MODES)
def test_tutorial_batch_mode(mode):
    g = golden("tut_batch_mode.json")
    P = hb.Program(sys_tutorial())
    assert (P.n_eq, P.n_pars, P.order) == (2, 1, 20)
    o = oracle.OracleIntegrator(P, [g["x0"], g["v0"]], 4, pars=[g["alpha"]], mode=mode)

    # step()
    o.step()
    assert [int(x) for x in o.step_outcome] == [OC[r["outcome"]] for r in g["first_step"]]
    assert sig_digits_equal(o.last_h, [r["h"] for r in g["first_step"]])
    assert decimals_equal(o.state, g["states"][0])
    assert decimals_equal(o.t_hi, g["times"][0])

    # step(max_delta_ts)
    o.step(g["clamped_step_limits"])
    assert [int(x) for x in o.step_outcome] == [OC[r["outcome"]] for r in g["clamped_step"]]
    assert np.all(o.last_h == np.array(g["clamped_step_limits"]))
    assert decimals_equal(o.state, g["states"][1])
    assert decimals_equal(o.t_hi, g["times"][1])

    # propagate_for({10, 11, 12, 13}): final time in double-length arithmetic.
    hi, lo = hb._dfloat_add(o.t_hi, o.t_lo, np.array(g["propagate_for"]["delta_ts"]), np.zeros(4))
    o.propagate_until(hi, lo)
    res = g["propagate_for"]["res"]
    assert [int(x) for x in o.prop_outcome] == [OC[r["outcome"]] for r in res]
    assert [int(x) for x in o.n_steps] == [r["n_steps"] for r in res]          # 34, 38, 41, 44
    assert sig_digits_equal(o.min_h, [r["min_h"] for r in res])
    assert sig_digits_equal(o.max_h, [r["max_h"] for r in res])
    assert decimals_equal(o.state, g["states"][2])
    assert decimals_equal(o.t_hi, g["times"][2])

    # propagate_until({20, 21, 22, 23})
    o.propagate_until(g["propagate_until"]["ts"])
    res = g["propagate_until"]["res"]
    assert [int(x) for x in o.n_steps] == [r["n_steps"] for r in res]          # 40, 38, 35, 34
    assert sig_digits_equal(o.min_h, [r["min_h"] for r in res])
    assert sig_digits_equal(o.max_h, [r["max_h"] for r in res])
    assert decimals_equal(o.state, g["states"][3])
    assert np.all(o.t_hi == np.array(g["propagate_until"]["ts"]))

    # step(true): the full order-20 array of Taylor coefficients, [var][order][batch].
    o.step(write_tc=True)
    assert sig_digits_equal(o.tc, g["tc_after_final_step"], 7)


@pytest.mark.parametrize("mode", MODES)
def test_exact_step_counts_under_max_delta_t(mode):
    """test/taylor_adaptive_batch.cpp:586-598: max_delta_t = {1e-4, 5e-5} up to t = {10, 11} takes exactly
    100000 / 220000 non-zero steps and lands exactly on the final times."""
    P = hb.Program(sys_pendulum())
    o = oracle.OracleIntegrator(P, [[0.05, 0.06], [0.025, 0.026]], 2, mode=mode)
    o.propagate_until([10., 11.], max_delta_t=[1e-4, 5e-5])
    assert np.all(o.t_hi == [10., 11.])
    assert [int(x) for x in o.n_steps] == [100000, 220000]
    assert np.all(o.prop_outcome == hb.taylor_outcome.time_limit)
    # ... and agrees with the unclamped propagation to 1000 eps (same test, :600-603).
    o2 = oracle.OracleIntegrator(P, [[0.05, 0.06], [0.025, 0.026]], 2, mode=mode)
    o2.propagate_until([10., 11.])
    assert approx(o.state, o2.state, 1000.)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("ha", [False, True])
def test_timestep_formula(mode, ha):
    """test/timestep_check.cpp:33-86: h recomputed from the Taylor coefficients with Jorba's formula."""
    P = hb.Program(sys_outer_ss(), high_accuracy=ha)
    assert (P.n_eq, P.n_uvars, P.order) == (36, 234, 20)
    o = oracle.OracleIntegrator(P, outer_ss_ic(), 1, mode=mode)
    order = P.order
    for _ in range(10):
        o.step(write_tc=True)
        assert int(o.step_outcome[0]) == hb.taylor_outcome.success
        tc = o.tc[:, :, 0]
        max_abs_state = np.max(np.abs(tc[:, 0]))
        max_abs_o = np.max(np.abs(tc[:, order]))
        max_abs_om1 = np.max(np.abs(tc[:, order - 1]))
        num = 1. if max_abs_state <= 1 else max_abs_state
        rho_o = (num / max_abs_o) ** (1. / order)
        rho_om1 = (num / max_abs_om1) ** (1. / (order - 1))
        rho_m = min(rho_o, rho_om1)
        h = rho_m * np.exp(-7. / 10 / (order - 1)) / (np.exp(1.) * np.exp(1.))
        assert approx(o.last_h[0], h, 100.)


def test_modes_agree_outer_ss():
    """The three summation modes agree to rounding over a few hundred steps (the reference's own
    compact-vs-default comparisons use 100-1000 eps)."""
    P = hb.Program(sys_outer_ss(), high_accuracy=True)
    res = []
    for mode in MODES:
        o = oracle.OracleIntegrator(P, outer_ss_ic(), 1, mode=mode)
        o.propagate_until(100.)
        res.append((o.state.copy(), int(o.n_steps[0])))
    assert res[0][1] == res[1][1] == res[2][1]
    for s, _ in res[1:]:
        assert np.max(np.abs(s - res[0][0]) / np.maximum(np.abs(res[0][0]), 1e-3)) < 1e-12


def test_vector_width_port_matches_scalar():
    """The 8-lane timed port (oracle_*_w8) computes the same thing as the scalar oracle."""
    from common import outer_ss_batch_state
    P = hb.Program(sys_outer_ss(), high_accuracy=True)
    st = outer_ss_batch_state(13, perturb=1e-3)
    a = oracle.OracleIntegrator(P, st, 13, mode=oracle.FMA, width=1)
    b = oracle.OracleIntegrator(P, st, 13, mode=oracle.FMA, width=8)
    a.propagate_until(20.)
    b.propagate_until(20., lockstep=False, n_threads=2)
    assert np.array_equal(a.n_steps, b.n_steps)
    assert np.max(np.abs(a.state - b.state) / np.maximum(np.abs(a.state), 1e-3)) < 1e-13
    assert np.all(b.t_hi == 20.)


# ---- 3. closed-form jets of the reference's per-operation tests (test/taylor_*.cpp) ----

def approximately(cmp, value, eps_mul=100.0):
    """test/test_utils.hpp:47-80: relative to the computed value, absolute below the tolerance."""
    cmp, value = np.asarray(cmp, dtype=float), np.asarray(value, dtype=float)
    tol = np.finfo(float).eps * eps_mul
    err = np.where(np.abs(cmp) < tol, np.abs(cmp - value), np.abs((cmp - value) / np.where(cmp == 0, 1.0, cmp)))
    return bool(np.all(err <= tol))


def closed_form_cases():
    from closed_form_cases import CASES
    g = golden("closed_form_jets.json")
    assert [c["name"] for c in g["cases"]] == [c[0] for c in CASES]
    return list(zip(CASES, g["cases"]))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case,gold", closed_form_cases(), ids=lambda v: v[0] if isinstance(v, tuple) else "")
def test_closed_form_jets(case, gold, mode):
    """Every batch-3 / order-3 block of test/taylor_{sincos,tanh,exp,log,sqrt,square,pow,div,mul,sub,sum_sq,neg,
    time}.cpp (file:line in tests/closed_form_cases.py): the jet of one step against the closed forms, to the
    reference's own 100 epsilon."""
    from closed_form_cases import EPS_MUL, ORDER, TOL, batch_of, hb_system
    BATCH = batch_of(case)
    P = hb.Program(hb_system(hb, case), tol=TOL)
    assert P.order == ORDER
    o = oracle.OracleIntegrator(P, gold["state"], BATCH, time=gold["time"] if gold["time"] else 0.0, mode=mode)
    o.step(write_tc=True)
    assert approximately(o.tc, gold["tc"], EPS_MUL.get(case[0], 100.0)), (case[0], o.tc, gold["tc"])


# ---- propagate_grid: the harmonic-oscillator fixtures of test/taylor_adaptive_batch.cpp ----

def grid_fixtures():
    """(name, grid [1000, 4], tolerance in epsilon): test/taylor_adaptive_batch.cpp:269-323 (regular grids, forward
    and backward, 10000 eps) and :332-385 (random grids, 400000 / 800000 eps; mt19937 replaced by numpy's
    generator: the closed form does not depend on the draws)."""
    out = []
    for sign, nm in ((1.0, "fwd"), (-1.0, "bwd")):
        g = np.zeros((1000, 4))
        for i in range(1000):
            for j in range(4):
                g[i, j] = sign * (i / 100.0) + (sign * (j / 10.0) if i != 0 else 0.0)
        out.append(("regular-" + nm, g, 10000.0))
    rng = np.random.default_rng(20240924)
    for sign, nm, tol in ((1.0, "fwd", 400000.0), (-1.0, "bwd", 800000.0)):
        g = np.zeros((1000, 4))
        g[1:] = np.cumsum(sign * rng.uniform(0.0, 0.1, (999, 4)), axis=0)
        out.append(("random-" + nm, g, tol))
    return out


def sys_oscillator():
    x, v = hb.make_vars("x", "v")
    return [(x, v), (v, -x)]


OSC_STATE = [[0.0, 0.0, 0.0, 0.0], [1.0, 1.1, 1.2, 1.3]]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name,grid,tol", grid_fixtures(), ids=lambda v: v if isinstance(v, str) else "")
def test_propagate_grid_oscillator(name, grid, tol, mode):
    o = oracle.OracleIntegrator(hb.Program(sys_oscillator()), OSC_STATE, 4, mode=mode)
    ret = o.propagate_grid(grid)
    assert ret.shape == (1000, 2, 4)
    assert np.all(o.prop_outcome == hb.taylor_outcome.time_limit)
    assert np.array_equal(o.t_hi, grid[-1])
    amp = 1.0 + np.arange(4) / 10.0
    assert approximately(ret[:, 0, :], amp * np.sin(grid), tol)
    assert approximately(ret[:, 1, :], amp * np.cos(grid), tol)


# ---- continuous output: the batch block of test/c_output.cpp:289-420 ----

def cout_fixture(batch_size=4, seed=7):
    """Harmonic oscillator, ICs x = i / 100, v = 1 + i / 100, final times 10 + i / 100, a random grid per lane
    (test/c_output.cpp:330-366; numpy's generator instead of mt19937)."""
    ic = np.array([[i / 100.0 for i in range(batch_size)], [1 + i / 100.0 for i in range(batch_size)]])
    final_tm = np.array([10.0 + i / 100.0 for i in range(batch_size)])
    rng = np.random.default_rng(seed)
    n_points = 10
    grid = np.zeros((n_points, batch_size))
    for i in range(batch_size):
        grid[1:-1, i] = np.sort(rng.uniform(1e-6, 10.0 + i / 100.0 - 1e-6, n_points - 2))
        grid[-1, i] = final_tm[i]
    return ic, final_tm, grid


@pytest.mark.parametrize("mode", MODES)
def test_continuous_output_oscillator(mode):
    ic, final_tm, grid = cout_fixture()
    P = hb.Program(sys_oscillator())
    o = oracle.OracleIntegrator(P, ic, 4, mode=mode)
    co = o.propagate_until_cout(final_tm)
    assert co is not None and np.all(o.prop_outcome == hb.taylor_outcome.time_limit)
    assert np.array_equal(o.t_hi, final_tm)
    lb, ub = co.get_bounds()
    assert np.all(lb == 0) and np.array_equal(ub, final_tm)
    assert co.t_hi.shape[0] == co.get_n_steps() + 2
    o2 = oracle.OracleIntegrator(P, ic, 4, mode=mode)
    grid_out = o2.propagate_grid(grid)
    for k in range(grid.shape[0]):
        assert approximately(co(grid[k]), grid_out[k], 100.0)
    # Closed form: x = x0 cos t + v0 sin t, v = -x0 sin t + v0 cos t.
    t = np.linspace(0.05, 9.9, 37)[:, None] * np.ones(4)[None, :]
    for k in range(t.shape[0]):
        s = co(t[k])
        assert approximately(s[0], ic[0] * np.cos(t[k]) + ic[1] * np.sin(t[k]), 1e5)
        assert approximately(s[1], -ic[0] * np.sin(t[k]) + ic[1] * np.cos(t[k]), 1e5)


# ---- 6. test/two_body_batch.cpp:60-190: batch vs single-lane agreement, Keplerian elements conserved ----

@pytest.mark.parametrize("mode", MODES)
def test_two_body_kepler_conservation(mode):
    from common import check_kepler_conservation, sys_two_body_symmetric, two_body_kepler_fixture
    kep, st = two_body_kepler_fixture()
    P = hb.Program(sys_two_body_symmetric())
    assert (P.n_eq, P.order) == (12, 20)
    o = oracle.OracleIntegrator(P, st, 4, mode=mode)
    for _ in range(200):
        prev, t_prev = o.state.copy(), o.t_hi.copy()
        o.step()
        # The same step taken by a one-lane integrator (the reference compares with its scalar class).
        for i in range(4):
            s = oracle.OracleIntegrator(P, prev[:, i:i + 1], 1, time=t_prev[i], mode=mode)
            s.step()
            assert s.step_outcome[0] == o.step_outcome[i]
            assert approximately(s.last_h[0], o.last_h[i], 1e4)
            assert approximately(s.state[:, 0], o.state[:, i], 1e5)
        check_kepler_conservation(o.state, kep, approximately)


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.abs(b)))


@pytest.mark.parametrize("mode", MODES)
def test_tutorial_adaptive(mode):
    """doc/tut_adaptive.rst (tutorial/adaptive_basic.cpp): the scalar pendulum as a batch of 3 identical lanes."""
    g = golden("tut_adaptive.json")
    P = hb.Program(sys_pendulum())
    assert P.order == g["order"]
    n = 3
    ic = np.array([[g["x0"]] * n, [g["v0"]] * n])
    o = oracle.OracleIntegrator(P, ic, n, mode=mode)
    # step(): printed with 6 digits, then time and state with 16-17 digits (the reference's LLVM build may contract and
    # order its sums differently from any of the oracle's modes: a few ulp).
    o.step()
    fs = g["first_step"]
    assert [int(x) for x in o.step_outcome] == [OC[fs["outcome"]]] * n
    assert sig_digits_equal(o.last_h, [fs["h"]] * n)
    assert _rel(o.t_hi, [fs["time"]] * n) < 1e-14
    assert _rel(o.state, np.array(fs["state"])[:, None] * np.ones(n)) < 1e-14
    assert np.all(o.state == o.state[:, :1])  # identical lanes stay identical
    # step_backward(), step(0.01), step(-0.02)
    o.step(backward=True)
    assert [int(x) for x in o.step_outcome] == [OC[g["step_backward"]["outcome"]]] * n
    assert sig_digits_equal(o.last_h, [g["step_backward"]["h"]] * n)
    for r in g["clamped_steps"]:
        o.step(r["limit"])
        assert [int(x) for x in o.step_outcome] == [OC[r["outcome"]]] * n and np.all(o.last_h == r["h"])
    # Reset, propagate_for(5), propagate_until(20), propagate_until(0): 24 / 72 / 97 steps.
    o.state[:] = ic
    o.t_hi[:] = 0
    o.t_lo[:] = 0
    for r, tf in zip(g["propagate"], (5.0, 20.0, 0.0)):
        o.propagate_until(tf)
        assert [int(x) for x in o.prop_outcome] == [OC[r["outcome"]]] * n
        assert [int(x) for x in o.n_steps] == [r["n_steps"]] * n
        assert sig_digits_equal(o.min_h, [r["min_h"]] * n) and sig_digits_equal(o.max_h, [r["max_h"]] * n)
        assert np.all(o.t_hi == r["time"])
    # Back at t = 0 after 193 steps: the reference prints the initial condition to an ulp or two; the oracle's modes
    # land within 30 ulp of that (observed: 2e-15 .. 6e-15 relative).
    assert _rel(o.state, np.array(g["state_back_at_0"])[:, None] * np.ones(n)) < 5e-14
    # propagate_grid({0, 0.1, ..., 1.0}): x(0.4), v(0.4).
    o.state[:] = ic
    o.t_hi[:] = 0
    o.t_lo[:] = 0
    out = o.propagate_grid(np.array(g["grid"]["times"])[:, None] * np.ones(n))
    k = g["grid"]["index"]
    assert sig_digits_equal(out[k, 0], [g["grid"]["x"]] * n) and sig_digits_equal(out[k, 1], [g["grid"]["v"]] * n)


@pytest.mark.parametrize("mode", MODES)
def test_tutorial_dense_and_continuous_output(mode):
    """doc/tut_d_output.rst (tutorial/d_output.cpp): dense output after one step, continuous output of
    propagate_until(10) - 48 recorded steps, six printed samples."""
    g = golden("tut_d_output.json")
    P = hb.Program(sys_pendulum())
    ic = [[g["x0"]] * 2, [g["v0"]] * 2]
    o = oracle.OracleIntegrator(P, ic, 2, mode=mode)
    o.step(write_tc=True)
    assert np.all(o.tc[:, 0, :] == np.array(ic))           # "TC of order 0": the state at the start of the step
    # update_d_output(0.1): tau relative to the start of the step (t = 0).
    d = o.d_output(np.full(2, 0.1))
    assert sig_digits_equal(d[:, 0], g["d_output_at_0.1"]) and sig_digits_equal(d[:, 1], g["d_output_at_0.1"])
    # ... at the end of the step it is the current state ("rel. difference: 0").
    assert _rel(o.d_output(o.last_h.copy()), o.state) < 1e-15
    o.state[:] = ic
    o.t_hi[:] = 0
    o.t_lo[:] = 0
    co = o.propagate_until_cout(g["c_output"]["t_final"])
    assert co.get_n_steps() == g["c_output"]["n_steps"] == 48
    lb, ub = co.get_bounds()
    assert np.all(lb == 0) and np.all(ub == 10)
    for tm, x, v in g["c_output"]["samples"]:
        s = co(tm)
        assert sig_digits_equal(s[0], [x] * 2) and sig_digits_equal(s[1], [v] * 2), tm


@pytest.mark.parametrize("mode", MODES)
def test_tutorial_ensemble_members_as_lanes(mode):
    """doc/tut_ensemble.rst (tutorial/ensemble.cpp): ensemble_propagate_until(20) over 10 initial conditions; the members
    are independent integrations, here the 10 lanes of one batch. Member 9: final state as printed (17 digits), 124
    steps, min / max step."""
    g = golden("tut_ensemble.json")
    P = hb.Program(sys_pendulum())
    ics = np.array(g["ics"]).T.copy()
    o = oracle.OracleIntegrator(P, ics, g["n_iter"], mode=mode)
    o.propagate_until(g["t_final"])
    m = g["member"]
    assert np.all(o.t_hi == g["time"]) and int(o.prop_outcome[m]) == OC[g["outcome"]]
    assert int(o.n_steps[m]) == g["n_steps"]
    assert sig_digits_equal(o.min_h[m], g["min_h"]) and sig_digits_equal(o.max_h[m], g["max_h"])
    assert _rel(o.state[:, m], g["state"]) < 2e-14       # 124 steps; observed 3e-16 .. 1.4e-15 in the three modes


@pytest.mark.parametrize("mode", MODES)
def test_tutorial_runtime_parameters(mode):
    """doc/tut_param.rst (tutorial/pendulum_param.cpp): x' = v, v' = -par[0] / par[1] sin(x); after one period (for two
    values of par[0]) the state is back at (0.05, 0) - the reference prints velocities of 7.6e-17 and 2.2e-16."""
    g = golden("tut_param.json")
    x, v = hb.make_vars("x", "v")
    P = hb.Program([(x, v), (v, -hb.par[0] / hb.par[1] * hb.sin(x))])
    assert (P.order, P.n_pars) == (g["order"], 2)
    runs = g["runs"]
    o = oracle.OracleIntegrator(P, [[g["x0"]] * 2, [g["v0"]] * 2], 2, pars=np.array([r["pars"] for r in runs]).T.copy(),
                                mode=mode)
    o.propagate_until([r["t_final"] for r in runs])
    for i, r in enumerate(runs):
        assert abs(o.state[0, i] - r["state"][0]) < 1e-16
        assert abs(o.state[1, i] - r["state"][1]) < 2e-15


@pytest.mark.parametrize("mode", MODES)
def test_tutorial_time_dependent_rhs(mode):
    """doc/tut_nonauto.rst (tutorial/forced_damped_pendulum.cpp): x' = v, v' = cos(t) - 0.1 v - sin(x); x after every
    propagate_for(2), 25 printed values."""
    g = golden("tut_nonauto.json")
    x, v = hb.make_vars("x", "v")
    P = hb.Program([(x, v), (v, hb.cos(hb.time) - .1 * v - hb.sin(x))])
    o = oracle.OracleIntegrator(P, [[g["x0"]], [g["v0"]]], 1, mode=mode)
    for k, xr in enumerate(g["x"]):
        hi, lo = hb._dfloat_add(o.t_hi, o.t_lo, np.array([g["delta_t"]]), np.zeros(1))
        o.propagate_until(hi, lo)
        assert sig_digits_equal(o.state[0, 0], xr), k
    assert o.t_hi[0] == 50.


@pytest.mark.parametrize("mode", MODES)
def test_tutorial_custom_tolerance(mode):
    """doc/tut_adaptive_custom.rst (tutorial/adaptive_opt.cpp): tol = 1e-9 gives order 12; after 0 -> 10 -> 0 the
    reference prints [0.050000000001312848, 0.024999999997558649]: the 1e-12 deviation from the initial condition is
    the truncation error of this tolerance, reproduced here to 1e-15."""
    g = golden("tut_adaptive_custom.json")
    P = hb.Program(sys_pendulum(), tol=g["tol"])
    assert P.order == g["order"] == 12
    o = oracle.OracleIntegrator(P, [[g["x0"]], [g["v0"]]], 1, mode=mode)
    for tf in g["times"]:
        o.propagate_until(tf)
    assert np.max(np.abs(o.state[:, 0] - np.array(g["state_back_at_0"]))) < 1e-15
    assert abs(o.state[0, 0] - g["x0"]) > 1e-13       # (the deviation itself is there)
