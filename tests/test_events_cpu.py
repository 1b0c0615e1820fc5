"""Event detection on the CPU oracle: the batch blocks of test/batch_event_detection.cpp run through the product's Python
front end (callbacks, lock-step propagation: host logic) over oracle/taylor_oracle.c (jet with event equations, fast
exclusion check, root isolation, TOMS 748). This pins the oracle's event detection on the reference's own expectations
(trigger counts and orders, the golden pendulum periods of :777-780, event times to 1000 eps); tests/test_gpu_events.py
then holds the device kernels to the oracle. No GPU needed."""
import numpy as np
import pytest

import event_cases as ec
import heyoka_b200 as hb
import oracle


def make(*a, **k):
    return oracle.OracleEventIntegrator(*a, **k)


def test_linear_box():
    times = ec.case_linear_box(make)
    assert np.allclose(sorted(times), [1 / 8., 1 / 4., 1 / 2., 1.], rtol=1e-15)


def test_glancing_blow():
    ec.case_glancing_blow(make)


@pytest.mark.parametrize("tol", [0.0, ec.EPS / 100])
def test_multizero(tol):
    ec.case_multizero(make, tol=tol)


def test_multizero_backward_and_direction():
    ec.case_multizero(make, backward=True)
    ec.case_multizero_dir(make)


def test_nte_basic_golden_periods():
    times = ec.case_nte_basic(make)
    for i in range(4):
        # Half a period between consecutive zeros of the velocity.
        assert abs(times[i][1] - ec.PERIODS[i] / 2) < 1e-13


@pytest.mark.parametrize("tol", [0.0, ec.EPS / 100])
def test_te_basic(tol):
    ec.case_te_basic(make, tol=tol)


def test_directions():
    ec.case_nte_dir(make)
    ec.case_te_dir(make)


def test_te_identical_close_retrigger():
    ec.case_te_identical(make)
    ec.case_te_close(make)
    ec.case_te_retrigger(make)


def test_te_cooldowns():
    ec.case_te_custom_cooldown(make)
    ec.case_te_zero_cd(make)


def test_te_propagate_for():
    ec.case_te_propagate_for(make)


def test_te_damped_pendulum_and_boolean_callback():
    ec.case_te_damped_pendulum(make)
    ec.case_te_boolean_callback(make)


def test_te_step_end():
    ec.case_te_step_end(make)


def test_te_propagate_grid():
    ec.case_te_propagate_grid(make)


@pytest.mark.parametrize("terminal", [False, True])
def test_single_step_batch_vs_scalar(terminal):
    """:100-260: every batch element agrees with the same system integrated alone (batch of one), event times to 1000 eps
    and velocities to 10000 eps like the reference's comparison with its scalar integrator."""
    times, vels = ec.case_single_step(make, terminal)
    ic = [0.00, 0.01, 0.02, 0.03, 1.85, 1.86, 1.87, 1.88]
    pars = [0.10, 0.11, 0.12, 0.13]
    x, v = hb.make_vars("x", "v")
    sys = [(x, v), (v, hb.cos(hb.time) - hb.par[0] * v - hb.sin(x))]
    for i in range(4):
        t1, v1 = [], []
        if terminal:
            def cb(ta, d_sgn, k):
                t1.append(ta.time[0])
                v1.append(ta.state[1, 0])
                return True
            ta = make(sys, [ic[i], ic[4 + i]], 1, pars=[pars[i]],
                      t_events=[hb.t_event_batch(x + .1, callback=cb, direction=hb.event_direction.negative)])
        else:
            def cb(ta, tm, d_sgn, k):
                t1.append(tm)
                v1.append(ta.update_d_output([tm])[1, 0])
            ta = make(sys, [ic[i], ic[4 + i]], 1, pars=[pars[i]],
                      nt_events=[hb.nt_event_batch(x + .1, cb, direction=hb.event_direction.negative)])
        while ta.time[0] < 20:
            ta.step()
        assert len(t1) == len(times[i]) and len(t1) >= 1
        for a, b, c, d in zip(t1, times[i], v1, vels[i]):
            assert ec.approx(a, b, 1000.) and ec.approx(c, d, 10000.)


def test_event_validation_and_decomposition():
    x, v, sys = ec.pendulum_sys()
    y, = hb.make_vars("y")
    with pytest.raises(ValueError, match="an event function contains the variable 'y', which is not a state variable"):
        hb.Program(sys, events=[x + y])
    with pytest.raises(ValueError, match="non-finite cooldown"):
        hb.t_event_batch(v, cooldown=float("nan"))
    with pytest.raises(ValueError, match="empty callback"):
        hb.nt_event_batch(v, None)
    # An event that is a state variable refers to it directly; parameters of event equations count
    # (test/taylor_adaptive_batch.cpp:1015-1060).
    P = hb.Program(sys, events=[v, v - hb.par[3]])
    assert P.n_ev == 2 and P.n_pars == 4
    assert P.ev_defs()[0] == 1 and P.ev_defs()[1] >= P.n_eq


def test_te_cooldowns_property():
    ec.case_te_cooldowns_property(make)


def test_tutorial_events_golden():
    """doc/tut_events.rst: the event times and the grid output the reference prints with 16 digits."""
    from common import golden
    dev = ec.case_tutorial_events(make, golden("tut_events.json"))
    print(dev)


@pytest.mark.parametrize("case", [ec.case_step_count_te_stop_bug, ec.case_callback_ste, ec.case_propagate_grid_ste,
                                  ec.case_ev_inf_state, ec.case_event_cb_time, ec.case_ev_exception_callback,
                                  ec.case_events_error, ec.case_get_set_dtime, ec.case_reset_cooldowns,
                                  ec.case_param_deduction_from_events], ids=lambda f: f.__name__)
def test_reference_regression_cases(case):
    """Regression cases of test/taylor_adaptive_batch.cpp for the host loops of integrators with events."""
    case(make)
