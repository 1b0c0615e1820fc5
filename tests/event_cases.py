"""The batch blocks of the reference's event-detection tests (test/batch_event_detection.cpp), written once and run
twice: on the CPU oracle (tests/test_events_cpu.py, which pins the oracle's restatement of detect_events() and TOMS 748
on the reference's own expectations) and on the GPU (tests/test_gpu_events.py). `make` builds an integrator with the
signature of heyoka_b200.taylor_adaptive_batch."""
import numpy as np
import pytest

import heyoka_b200 as hb

EPS = float(np.finfo(np.float64).eps)
TO = hb.taylor_outcome
INF = float("inf")


def approx(a, b, tol_eps=100.0):
    """test/catch.hpp approximately(): |a - b| <= eps * tol * max(|a|, |b|)... (relative, eps-scaled)."""
    return abs(a - b) <= EPS * tol_eps * max(abs(a), abs(b))


def pendulum_sys():
    x, v = hb.make_vars("x", "v")
    return x, v, [(x, v), (v, -9.8 * hb.sin(x))]


PEND_IC = [0, 0.01, 0.02, 0.03, .25, .26, .27, .28]


def case_linear_box(make):
    """:262-327: an event exactly at the end of a step fires at the beginning of the next one."""
    x, = hb.make_vars("x")
    counter = [0]
    times = []

    def ncb(ta, tm, d_sgn, i):
        assert approx(tm, 1 / ta.pars[0, i])
        counter[0] += 1

    ta = make([(x, hb.par[0])], [0., 0., 0., 0.], 4, pars=[1., 2., 4., 8.], nt_events=[hb.nt_event_batch(x - 1., ncb)])
    lim = [1., 1 / 2., 1 / 4., 1 / 8.]
    ta.step(lim)
    assert counter[0] == 0 and all(r[0] == TO.time_limit for r in ta.step_res)
    ta.step(lim)
    assert counter[0] == 4 and all(r[0] == TO.time_limit for r in ta.step_res)

    counter[0] = 0

    def tcb(ta, d_sgn, i):
        counter[0] += 1
        times.append(ta.time[i])
        return True

    ta = make([(x, hb.par[0])], [0., 0., 0., 0.], 4, pars=[1., 2., 4., 8.],
              t_events=[hb.t_event_batch(x - 1., callback=tcb)])
    ta.step(lim)
    assert counter[0] == 0 and all(r[0] == TO.time_limit for r in ta.step_res)
    ta.step(lim)
    assert counter[0] == 4
    assert all(r[0] == 0 for r in ta.step_res)
    assert all(abs(r[1]) <= 100 * EPS for r in ta.step_res)
    return times


def case_glancing_blow(make):
    """:329-409: two spheres touching tangentially (double root of the distance polynomial) on batch index 1."""
    names = ["x0", "y0", "x1", "y1", "vx0", "vy0", "vx1", "vy1"]
    x0, y0, x1, y1, vx0, vy0, vx1, vy1 = hb.make_vars(*names)
    ic = [0., 0., 0., 0., 0., 0., 0., 0., -10., -10., -10., -10., 6., 2, 7., 8., 0., 0., 0., 0., 0., 0., 0., 0.,
          1., 1., 1., 1., 0., 0., 0., 0.]
    out = []
    for acc in (0., .1):
        counter = [0]

        def cb(ta, t, d_sgn, i, acc=acc, counter=counter):
            if acc == 0.:
                assert (t - 10.) ** 2 <= EPS
            assert i == 1
            counter[0] += 1

        zero = hb.expression(0.)
        sys = [(x0, vx0), (y0, vy0), (x1, vx1), (y1, vy1), (vx0, zero), (vy0, zero), (vx1, hb.expression(acc)), (vy1, zero)]
        ev = hb.nt_event_batch((x0 - x1) * (x0 - x1) + (y0 - y1) * (y0 - y1) - 4., cb)
        ta = make(sys, ic, 4, nt_events=[ev])
        for _ in range(20):
            ta.step([1.3] * 4)
            assert all(r[0] == TO.time_limit for r in ta.step_res)
        assert counter[0] <= 2
        out.append(counter[0])
    return out


def case_multizero(make, tol=0.0, backward=False):
    """:411-768: v^2 - 1e-10 and v fire in the same step in the order 0 1 0."""
    x, v, sys = pendulum_sys()
    counter = [0] * 4
    cur_time = [0.] * 4
    log = [[] for _ in range(4)]
    sgn = -1. if backward else 1.

    def cb0(ta, t, d_sgn, i):
        assert sgn * t > sgn * cur_time[i]
        assert sgn * ta.time[i] > sgn * t
        assert counter[i] % 3 in (0, 2)
        vel = ta.update_d_output([t] * 4)[1, i]
        assert abs(vel * vel - 1e-10) < EPS
        counter[i] += 1
        cur_time[i] = t
        log[i].append((0, t))

    def cb1(ta, t, d_sgn, i):
        assert sgn * t > sgn * cur_time[i]
        assert sgn * ta.time[i] > sgn * t
        assert counter[i] % 3 == 1
        vel = ta.update_d_output([t] * 4)[1, i]
        assert abs(vel) <= EPS * 100
        counter[i] += 1
        cur_time[i] = t
        log[i].append((1, t))

    ta = make(sys, PEND_IC, 4, tol=tol, nt_events=[hb.nt_event_batch(v * v - 1e-10, cb0), hb.nt_event_batch(v, cb1)])
    ta.propagate_until([sgn * 4.] * 4)
    for i in range(4):
        assert ta.propagate_res[i][0] == TO.time_limit
        assert counter[i] == 12
    return log, ta


def case_multizero_dir(make):
    """:560-690: the same with v only detected when going from positive to negative: 0 1 0 / 0 0 alternating."""
    x, v, sys = pendulum_sys()
    counter = [0] * 4
    cur_time = [0.] * 4

    def cb0(ta, t, d_sgn, i):
        assert t > cur_time[i] and ta.time[i] > t
        assert counter[i] == 0 or 2 <= counter[i] <= 9
        vel = ta.update_d_output([t] * 4)[1, i]
        assert abs(vel * vel - 1e-10) < EPS
        counter[i] += 1
        cur_time[i] = t

    def cb1(ta, t, d_sgn, i):
        assert t > cur_time[i] and ta.time[i] > t
        assert counter[i] in (1, 6)
        assert d_sgn == -1
        vel = ta.update_d_output([t] * 4)[1, i]
        assert abs(vel) <= EPS * 100
        counter[i] += 1
        cur_time[i] = t

    ta = make(sys, PEND_IC, 4, nt_events=[hb.nt_event_batch(v * v - 1e-10, cb0),
                                          hb.nt_event_batch(v, cb1, direction=hb.event_direction.negative)])
    ta.propagate_until([4.] * 4)
    for i in range(4):
        assert ta.propagate_res[i][0] == TO.time_limit
        assert counter[i] == 10
    return counter


PERIODS = [2.0149583072955119566777324135479727911105583481363, 2.015602866455777600694040810649276304933055944554756,
           2.0162731039077591887007722648120652760856018525920970125217,
           2.01696906642817313582861191326257261662145101139954930969969]


def case_nte_basic(make):
    """:770-813: the third zero of the velocity of a pendulum is one period after the first (GOLDEN: the periods are
    the reference's own numbers, to 1000 eps)."""
    x, v, sys = pendulum_sys()
    counter = [0] * 4
    times = [[] for _ in range(4)]

    def cb(ta, t, d_sgn, i):
        if counter[i] == 0:
            assert t == 0
        if counter[i] == 2:
            assert approx(t, PERIODS[i], 1000.)
        counter[i] += 1
        times[i].append(t)

    ta = make(sys, [-0.25, -0.26, -0.27, -0.28, 0., 0., 0., 0.], 4, nt_events=[hb.nt_event_batch(v, cb)])
    for _ in range(20):
        ta.step()
        assert all(r[0] == TO.success for r in ta.step_res)
    assert counter == [3] * 4
    return times


def _step_until_all_trigger(ta, limit, check):
    n_trig = 0
    mdt = [limit] * 4
    trig = [None] * 4
    while n_trig < 4:
        ta.step(mdt)
        for i in range(4):
            oc = ta.step_res[i][0]
            if oc > TO.success:
                check(oc)
                n_trig += 1
                mdt[i] = 0
                trig[i] = oc
            else:
                assert oc in (TO.success, TO.time_limit)
    return trig


def case_te_basic(make, tol=0.0):
    """:815-980: terminal event v with callbacks, non-terminal v^2 - 1e-10, forwards and backwards."""
    x, v, sys = pendulum_sys()
    c_nt, c_t = [0] * 4, [0] * 4
    cur_time = [0.] * 4
    direction = [True]
    ev_times = [[] for _ in range(4)]

    def ncb(ta, t, d_sgn, i):
        assert (t > cur_time[i]) if direction[0] else (t < cur_time[i])
        vel = ta.update_d_output([t] * 4)[1, i]
        assert abs(vel * vel - 1e-10) < EPS
        c_nt[i] += 1
        cur_time[i] = t

    def tcb(ta, d_sgn, i):
        t = ta.time[i]
        assert (t > cur_time[i]) if direction[0] else (t < cur_time[i])
        assert abs(ta.state[1, i]) < EPS * 100
        c_t[i] += 1
        cur_time[i] = t
        ev_times[i].append(t)
        return True

    ta = make(sys, PEND_IC, 4, tol=tol, nt_events=[hb.nt_event_batch(v * v - 1e-10, ncb)],
              t_events=[hb.t_event_batch(v, callback=tcb)])

    def ge0(oc):
        assert oc >= 0

    _step_until_all_trigger(ta, INF, ge0)
    assert c_nt == [1] * 4 and c_t == [1] * 4
    _step_until_all_trigger(ta, INF, ge0)
    assert c_nt == [3] * 4 and c_t == [2] * 4
    direction[0] = False
    _step_until_all_trigger(ta, -INF, ge0)
    assert c_nt == [5] * 4 and c_t == [3] * 4
    _step_until_all_trigger(ta, -INF, ge0)
    assert c_nt == [7] * 4 and c_t == [4] * 4
    return ev_times


def case_nte_dir(make):
    """:982-1018: direction filter; the events met forwards are met again, in reverse, backwards."""
    x, v, sys = pendulum_sys()
    fwd = [True]
    tlist = [[] for _ in range(4)]
    pos = [0] * 4

    def cb(ta, t, d_sgn, i):
        assert d_sgn == 1
        if fwd[0]:
            tlist[i].append(t)
        elif pos[i] < len(tlist[i]):
            # (The reference compares relatively; the event at t = 0 comes back as a few 1e-16, which only an absolute
            # comparison can accept - whether it is met again at all depends on the last bits of the trajectory.)
            ref = tlist[i][len(tlist[i]) - 1 - pos[i]]
            assert abs(ref - t) <= 100 * EPS * max(abs(ref), abs(t), 1.)
            pos[i] += 1

    ta = make(sys, [-0.25, -0.26, -0.27, -0.28, 0., 0., 0., 0.], 4,
              nt_events=[hb.nt_event_batch(v, cb, direction=hb.event_direction.positive)])
    ta.propagate_until([20.] * 4)
    fwd[0] = False
    ta.propagate_until([0.] * 4)
    assert all(len(t) >= 9 for t in tlist) and all(p >= len(t) - 1 for p, t in zip(pos, tlist))
    return tlist


def case_te_identical(make):
    """:1074-1123: two identical terminal events: one of them stops the step, the other may fire right after."""
    x, v, sys = pendulum_sys()
    ta = make(sys, PEND_IC, 4, t_events=[hb.t_event_batch(v), hb.t_event_batch(v)])
    while True:
        ta.step()
        if any(r[0] > TO.success for r in ta.step_res):
            break
    first = []
    for i in range(4):
        oc = ta.step_res[i][0]
        assert oc > TO.success
        assert -oc - 1 in (0, 1)
        first.append(-oc - 1)
    ta.step()
    for i in range(4):
        oc = ta.step_res[i][0]
        if oc > TO.success:
            assert -oc - 1 in (0, 1) and -oc - 1 != first[i]
        else:
            assert oc == TO.success


def case_te_close(make):
    """:1125-1264: two terminal events 2 eps apart: order of firing, cooldowns, there and back."""
    x, v, sys = pendulum_sys()
    ta = make(sys, [0.1, 0.11, 0.12, 0.13, .25, .26, .27, .28], 4,
              t_events=[hb.t_event_batch(x), hb.t_event_batch(x - EPS * 2, callback=lambda ta, s, i: True)])

    def ge0(oc):
        assert oc >= 0

    def lt0(oc):
        assert oc < 0

    assert _step_until_all_trigger(ta, INF, ge0) == [1] * 4
    assert _step_until_all_trigger(ta, INF, lt0) == [-1] * 4
    ta.step()
    assert all(r[0] == TO.success for r in ta.step_res)
    assert _step_until_all_trigger(ta, -INF, lt0) == [-1] * 4
    assert _step_until_all_trigger(ta, -INF, ge0) == [1] * 4
    ta.step()
    assert all(r[0] == TO.success for r in ta.step_res)


def case_te_retrigger(make):
    """:1266-1315: an event a few eps away from the initial state fires immediately, and again one period later."""
    x, v, sys = pendulum_sys()
    ta = make(sys, [1., 1.01, 1.02, 1.03, 0., 0.01, 0.02, 0.03], 4, pars=[1., 1.01, 1.02, 1.03],
              t_events=[hb.t_event_batch(x - (hb.par[0] - EPS * 6))])
    ta.step()
    assert all(r[0] == -1 for r in ta.step_res)
    assert all(t != 0 for t in ta.time)

    def m1(oc):
        assert oc == -1

    _step_until_all_trigger(ta, INF, m1)
    ta.step()
    assert all(r[0] == -1 for r in ta.step_res)


def case_te_dir(make):
    """:1317-1399: direction filter on a terminal event, zero-length steps, cooldown after the trigger."""
    x, v, sys = pendulum_sys()

    def pos(ta, d_sgn, i):
        assert d_sgn == 1
        return True

    ic = [1., 1.01, 1.02, 1.03, 0., 0., 0., 0.]
    ta = make(sys, ic, 4, t_events=[hb.t_event_batch(v, callback=pos, direction=hb.event_direction.positive)])
    ta.step()
    assert all(r[0] == TO.success for r in ta.step_res)
    while True:
        ta.step()
        if all(r[0] == 0 for r in ta.step_res):
            break
    for i in range(4):
        assert approx(ta.state[0, i], -ic[i])

    def neg(ta, d_sgn, i):
        assert d_sgn == -1
        return True

    ta = make(sys, ic, 4, t_events=[hb.t_event_batch(v, callback=neg, direction=hb.event_direction.negative)])
    ta.step([0.] * 4)
    assert all(r[0] == TO.time_limit for r in ta.step_res)
    ta.step()
    assert all(r[0] == 0 for r in ta.step_res)
    ta.step()
    assert all(r[0] == TO.success for r in ta.step_res)
    while True:
        ta.step()
        if all(r[0] == 0 for r in ta.step_res):
            break
    for i in range(4):
        assert approx(ta.state[0, i], ic[i])


def case_te_custom_cooldown(make):
    """:1401-1438: with a cooldown of 0.1 the double root of v^2 - 4 eps fires once."""
    x, v, sys = pendulum_sys()
    ta = make(sys, PEND_IC, 4,
              t_events=[hb.t_event_batch(v * v - EPS * 4, callback=lambda ta, s, i: True, cooldown=1e-1)])

    def is0(oc):
        assert oc == 0

    _step_until_all_trigger(ta, INF, is0)


def case_te_propagate_for(make):
    """:1440-1477: 100 zero crossings of the velocity in 100 time units; without callback the first one stops."""
    x, v, sys = pendulum_sys()
    counter = [0] * 4

    def cb(ta, d_sgn, i):
        counter[i] += 1
        return True

    ta = make(sys, PEND_IC, 4, t_events=[hb.t_event_batch(v, callback=cb)])
    ta.propagate_for([100.] * 4)
    assert all(r[0] == TO.time_limit for r in ta.propagate_res)
    assert all(t == 100. for t in ta.time)
    assert counter == [100] * 4
    ta = make(sys, PEND_IC, 4, t_events=[hb.t_event_batch(v)])
    ta.propagate_for([100.] * 4)
    assert all(r[0] == -1 for r in ta.propagate_res)
    return ta


def case_te_damped_pendulum(make):
    """:1580-1645: the callback flips a parameter of the system at every zero of the velocity."""
    x, v = hb.make_vars("x", "v")
    zvt = [[] for _ in range(4)]

    def cb(ta, d_sgn, i):
        tm = ta.time[i]
        ta.pars[0, i] = 1. if ta.pars[0, i] == 0 else 0.
        zvt[i].append(tm)
        return True

    ta = make([(x, v), (v, -9.8 * hb.sin(x) - hb.par[0] * v)], [0.05, 0.051, 0.052, 0.053, 0.025, 0.0251, 0.0252, 0.0253],
              4, t_events=[hb.t_event_batch(v, callback=cb)])
    ta.propagate_until([100.] * 4)
    assert [len(z) for z in zvt] == [99] * 4
    ta.step()
    assert [len(z) for z in zvt] == [100] * 4
    return zvt


def case_te_boolean_callback(make):
    """:1647-1721: the fifth invocation of the callback returns false: propagate_until() stops there."""
    x, v, sys = pendulum_sys()
    c_t = [0] * 4
    cur_time = [0.] * 4
    direction = [True]

    def cb(ta, d_sgn, i):
        t = ta.time[i]
        assert (t > cur_time[i]) if direction[0] else (t < cur_time[i])
        assert abs(ta.state[1, i]) < EPS * 100
        c_t[i] += 1
        cur_time[i] = t
        return c_t[i] != 5

    ta = make(sys, PEND_IC, 4, t_events=[hb.t_event_batch(v, callback=cb)])
    while True:
        ta.step()
        if all(r[0] == 0 for r in ta.step_res):
            break
    ta.propagate_until([1000.] * 4)
    assert all(r[0] == -1 for r in ta.step_res)
    for i in range(4):
        c_t[i] = 0
    direction[0] = False
    while True:
        ta.step_backward()
        if all(r[0] == 0 for r in ta.step_res):
            break
    ta.propagate_until([-1000.] * 4)
    assert all(r[0] == -1 for r in ta.step_res)


def case_te_step_end(make):
    """:1723-1747: an event (time - 1) that falls exactly on the end of a clamped step."""
    x, v, sys = pendulum_sys()
    counter = [0] * 4

    def cb(ta, d_sgn, i):
        counter[i] += 1
        assert ta.time[i] == 1.
        return True

    ta = make(sys, PEND_IC, 4, t_events=[hb.t_event_batch(hb.time - 1., callback=cb)])
    ta.propagate_until([10.] * 4, max_delta_t=[0.005] * 4)
    assert counter == [1] * 4


def case_te_zero_cd(make):
    """:1749-1785: zero cooldown and a callback that stops."""
    x, v, sys = pendulum_sys()
    ta = make(sys, PEND_IC, 4, t_events=[hb.t_event_batch(v, callback=lambda ta, s, i: False, cooldown=0)])
    ta.propagate_until([10.] * 4)
    assert all(r[0] == -1 for r in ta.step_res)


def case_single_step(make, terminal):
    """:100-260 (the batch half): driven damped pendulum with time-dependent right-hand side and parameters; returns
    the trigger times and the velocities at the events for every batch index."""
    x, v = hb.make_vars("x", "v")
    sys = [(x, v), (v, hb.cos(hb.time) - hb.par[0] * v - hb.sin(x))]
    ic = [0.00, 0.01, 0.02, 0.03, 1.85, 1.86, 1.87, 1.88]
    pars = [0.10, 0.11, 0.12, 0.13]
    times, vels = [[] for _ in range(4)], [[] for _ in range(4)]
    if terminal:
        def cb(ta, d_sgn, i):
            times[i].append(ta.time[i])
            vels[i].append(ta.state[1, i])
            return True

        ta = make(sys, ic, 4, pars=pars,
                  t_events=[hb.t_event_batch(x + .1, callback=cb, direction=hb.event_direction.negative)])
    else:
        def cb(ta, tm, d_sgn, i):
            times[i].append(tm)
            vels[i].append(ta.update_d_output([tm] * 4)[1, i])

        ta = make(sys, ic, 4, pars=pars,
                  nt_events=[hb.nt_event_batch(x + .1, cb, direction=hb.event_direction.negative)])
    while np.any(ta.time < 20):
        ta.step()
        for r in ta.step_res:
            assert r[0] == TO.success or (terminal and r[0] == 0)
    if terminal:
        assert [len(t) for t in times] == [2, 1, 1, 1]  # ex_n_trig, :196
    return times, vels


def case_te_propagate_grid(make):
    """:1479-1578: propagate_grid() with a terminal event: with a callback that continues every grid point is reached,
    without callback everything after the first trigger stays NaN."""
    x, v, sys = pendulum_sys()
    counter = [0] * 4

    def cb(ta, d_sgn, i):
        counter[i] += 1
        return True

    ta = make(sys, PEND_IC, 4, t_events=[hb.t_event_batch(v, callback=cb)])
    grid = np.repeat(np.arange(101.)[:, None], 4, axis=1)
    out = ta.propagate_grid(grid)
    assert out.shape == (101, 2, 4)
    assert np.all(out.reshape(-1)[1:] != 0) and np.all(np.isfinite(out))
    assert counter == [100] * 4
    assert all(r[0] == TO.time_limit for r in ta.propagate_res)
    ta2 = make(sys, PEND_IC, 4, t_events=[hb.t_event_batch(v)])
    out2 = ta2.propagate_grid(grid)
    assert np.all(np.isnan(out2.reshape(-1)[8:]))
    assert all(r[0] == -1 for r in ta2.propagate_res)
    # "first step bug" (:1536-1578): the event triggers before the second grid point.
    grid3 = np.repeat((5 / 100. * np.arange(100.))[:, None], 4, axis=1)
    ic = [0.05, 0.051, 0.052, 0.053, 0.025, 0.0251, 0.0252, 0.0253]
    ta3 = make(sys, ic, 4, t_events=[hb.t_event_batch(v, callback=lambda ta, s, i: True)])
    out3 = ta3.propagate_grid(grid3)
    assert np.all(out3 != 0) and np.all(np.isfinite(out3))
    ta4 = make(sys, ic, 4, t_events=[hb.t_event_batch(v)])
    out4 = ta4.propagate_grid(grid3)
    assert np.all(np.isnan(out4.reshape(-1)[32:]))
    return out


def case_te_cooldowns_property(make):
    """te_cooldowns (the reference's get_te_cooldowns(), src/taylor_adaptive_batch.cpp:2212-2219; the values are set at
    :936-950): empty before any event, (0, cooldown) for the terminal event that stopped the propagation, untouched
    for an event that never triggers, cleared by reset_cooldowns(i) / reset_cooldowns()."""
    x, v, sys = pendulum_sys()
    ic = np.array(PEND_IC).reshape(2, 4)
    ta = make(sys, ic, 4, t_events=[hb.t_event_batch(v), hb.t_event_batch(x - 100.)])
    assert ta.te_cooldowns == [[None, None]] * 4
    ta.propagate_for(100.)
    # (The lock-step loop ends at the first iteration in which a lane is stopped by its terminal event: the lanes that
    # have not reached theirs yet end with `success` and no cooldown.)
    res = [r[0] for r in ta.propagate_res]
    assert res.count(-1) >= 3 and all(r in (-1, TO.success) for r in res)
    before = ta.te_cooldowns
    for lane, oc in zip(before, res):
        assert lane[1] is None
        if oc == -1:
            assert lane[0][0] == 0. and 0. < lane[0][1] < np.inf
        else:
            assert lane[0] is None
    ta.reset_cooldowns(2)
    after = ta.te_cooldowns
    assert after[2] == [None, None] and [after[i] for i in (0, 1, 3)] == [before[i] for i in (0, 1, 3)]
    ta.reset_cooldowns()
    assert ta.te_cooldowns == [[None, None]] * 4
    with pytest.raises(ValueError, match="No events were defined"):
        make(sys, ic, 4).te_cooldowns


def case_tutorial_events(make, golden, loose=1.0):
    """doc/tut_events.rst (tutorial/event_basic.cpp), GOLDEN: the event times the reference prints with 16 digits, the
    scalar integrators of the page as batches of 4 identical lanes. Returns the largest relative deviations seen
    (zero-velocity times, times of the second event of the two-event run, grid output of the terminal-event run).
    `loose` scales the tolerances (1 for the oracle, whose arithmetic is the reference's up to contraction; larger for
    the device, whose sin / cos differ from the host's by an ulp or two per call)."""
    g = golden
    x, v, sys = pendulum_sys()
    n = 4
    ic = np.array([[g["x0"]] * n, [g["v0"]] * n])
    dev = {"nt": 0., "two": 0., "grid": 0.}

    def rel(a, b):
        return abs(a - b) / abs(b) if b != 0 else abs(a)

    # 1. Non-terminal event v = 0, dense output at the event time from inside the callback.
    times = [[] for _ in range(n)]
    xs = [[] for _ in range(n)]

    def cb(ta, t, d_sgn, i):
        xs[i].append(ta.update_d_output([t] * n)[0, i])
        times[i].append(t)

    ta = make(sys, ic, n, nt_events=[hb.nt_event_batch(v, cb)])
    ta.propagate_until(g["t_final"])
    ref = g["nt_zero_velocity"]
    for i in range(n):
        assert len(times[i]) == len(ref["times"]) == 5 and times[i][0] == 0.
        for t, rt in zip(times[i], ref["times"]):
            dev["nt"] = max(dev["nt"], rel(t, rt))
        for xv, rx in zip(xs[i], ref["x_at_events"]):
            assert abs(xv - rx) < 1e-9           # printed as +-0.05: the amplitude comes back at every turning point
    assert dev["nt"] < 1e-14 * loose, dev
    # 2. Same with direction = positive: t = 0, T, 2T.
    times = [[] for _ in range(n)]
    ta = make(sys, ic, n, nt_events=[hb.nt_event_batch(v, lambda ta, t, d_sgn, i: times[i].append(t),
                                                       direction=hb.event_direction.positive)])
    ta.propagate_until(g["t_final"])
    ref = g["nt_zero_velocity_positive_direction"]["times"]
    for i in range(n):
        assert len(times[i]) == 3 and times[i][0] == 0.
        assert all(rel(t, rt) < 1e-14 * loose for t, rt in zip(times[i], ref))
    # 3. Two events, v and v*v - 1e-12: callbacks in chronological order.
    seq = [[] for _ in range(n)]
    ta = make(sys, ic, n, nt_events=[hb.nt_event_batch(v, lambda ta, t, d_sgn, i: seq[i].append((0, t))),
                                     hb.nt_event_batch(v * v - 1e-12, lambda ta, t, d_sgn, i: seq[i].append((1, t)))])
    ta.propagate_until(g["t_final"])
    ref = g["nt_two_events"]["sequence"]
    for i in range(n):
        assert [e for e, _ in seq[i]] == [r["event"] for r in ref]
        for (e, t), r in zip(seq[i], ref):
            if e == 0:
                assert rel(t, r["t"]) < 1e-14 * loose
            else:
                # (v*v - 1e-12 crosses zero 2e-6 away from the turning point with a slope of 1e-6, and the event
                # polynomial - coefficients ~1e-2 - is evaluated to ~1e-18: the root moves by ~1e-12 with the rounding
                # of the evaluation. Observed on the oracle: 1.3e-12.)
                dev["two"] = max(dev["two"], abs(t - r["t"]))
    assert dev["two"] < 2e-11 * loose, dev
    # 4. Terminal event toggling the damping parameter; single steps up to the first stop, then propagate_grid().
    tg = g["terminal_damping_toggle"]

    def tcb(ta, d_sgn, i):
        ta.pars[0, i] = 1. if ta.pars[0, i] == 0 else 0.
        return True

    ta = make([(x, v), (v, -9.8 * hb.sin(x) - hb.par[0] * v)], np.array([[tg["x0"]] * n, [tg["v0"]] * n]), n,
              t_events=[hb.t_event_batch(v, callback=tcb)])
    while True:
        ta.step()
        if any(r[0] != TO.success for r in ta.step_res):
            break
    assert [r[0] for r in ta.step_res] == [tg["first_stop_event_index"]] * n   # terminal_event_0 (continuing)
    ta.propagate_until(1.)
    out = ta.propagate_grid(np.array(tg["grid"])[:, None] * np.ones(n))
    ref = np.array(tg["grid_output"])
    for i in range(n):
        dev["grid"] = max(dev["grid"], float(np.max(np.abs(out[:, :, i] - ref) / np.abs(ref))))
    assert dev["grid"] < 1e-12 * loose, dev
    assert np.all(ta.time == tg["final_time"])
    return dev


# ---- regression cases of test/taylor_adaptive_batch.cpp that exercise the front ends' host loops with events ----
def case_step_count_te_stop_bug(make):
    """:1819-1844 "propagate step count te stop bug": a terminal event without callback stops propagate_until() and
    propagate_grid() in the first step; that step is counted."""
    x, v, sys = pendulum_sys()
    ic = [0., 0., 0.5, 0.5001]
    ta = make(sys, ic, 2, t_events=[hb.t_event_batch(x - 1e-6)])
    ta.propagate_until([10., 10.])
    assert [r[3] for r in ta.propagate_res] == [1, 1]
    ta = make(sys, ic, 2, t_events=[hb.t_event_batch(x - 1e-6)])
    ta.propagate_grid([0., 0., 1., 1., 2., 2.])
    assert [r[3] for r in ta.propagate_res] == [1, 1]
    # :1847-1862 "set_time alias bug": set_time(get_time()) leaves the time alone.
    ta = make(sys, ic, 2, t_events=[hb.t_event_batch(x - 1e-6)])
    ta.set_time(ta.time)
    assert list(ta.time) == [0., 0.]


def case_callback_ste(make):
    """:1944-1980 "callback ste": a stopping terminal event in the first iteration of propagate_until(kw::callback): the
    step callback still runs once; the stopped lane reports -1, the others success, one step each."""
    x, v, sys = pendulum_sys()
    ta = make(sys, [-1., -0.0001, -1., -1., 0.025, 0.026, 0.027, 0.028], 4, t_events=[hb.t_event_batch(x)])
    n_invoked = [0]

    def pcb(t):
        n_invoked[0] += 1
        return True

    ta.propagate_until(10., callback=pcb)
    assert [r[0] for r in ta.propagate_res] == [TO.success, -1, TO.success, TO.success]
    assert [r[3] for r in ta.propagate_res] == [1, 1, 1, 1]
    assert n_invoked[0] == 1


def case_propagate_grid_ste(make):
    """:2011-2046 "propagate_grid ste": when propagate_grid() runs into a stopping terminal event, the grid points inside
    the last step taken are still produced, the later ones stay NaN."""
    x, v = hb.make_vars("x", "v")
    ta = make([(x, v), (v, -x)], [0., 0., 1., 1.], 2, t_events=[hb.t_event_batch(hb.time - .1)])
    grid = [0., 0., .1 - 2e-6, 10., .1 - 1e-6, 20., .1 + 1e-6, 30.]
    out = np.asarray(ta.propagate_grid(grid))          # [n_pts, dim, batch]
    assert out.shape == (4, 2, 2)
    assert [r[0] for r in ta.propagate_res] == [-1, -1]
    res = out.reshape(-1)                              # the reference's flat layout [pt][var][lane]
    nan = [bool(np.isnan(r)) for r in res]
    assert nan == [False, False, False, False, False, True, False, True, False, True, False, True, True, True, True, True]


def case_ev_inf_state(make):
    """:1456-1471 "ev inf state": a non-finite state in one lane is reported for that lane only; the others detect their
    terminal event."""
    x, = hb.make_vars("x")
    ta = make([(x, hb.expression(1.))], [0., 0., 0., 0.], 4, t_events=[hb.t_event_batch(x - 5.)])
    ta.state[0, 2] = np.inf
    ta.step([10.] * 4)
    assert [r[0] for r in ta.step_res] == [-1, -1, TO.err_nf_state, -1]


def case_event_cb_time(make):
    """:1560-1640 "event cb time": event callbacks that alter the time coordinate are an error naming the first altered
    batch index; every callback of the step has run by then, each with the event's own (finite) time."""
    x, v = hb.make_vars("x", "v")
    counts = [0, 0]

    def mk(k, new_t0):
        def cb(ta, t, d_sgn, i):
            assert np.isfinite(t)
            counts[k] += 1
            ta.set_time([new_t0, ta.time[1]])
        return cb

    ta = make([(x, v), (v, -x)], [0., 0., 1., -1.], 2,
              nt_events=[hb.nt_event_batch(x - 1e-5, mk(0, -10.)), hb.nt_event_batch(x - 1e-5, mk(1, -10.))])
    with pytest.raises(RuntimeError, match="The invocation of one or more event callbacks resulted in the alteration of "
                                           "the time coordinate of the integrator at the batch index 0 - this is not "
                                           "supported"):
        ta.step()
    assert counts == [1, 1]
    counts[:] = [0, 0]
    ta = make([(x, v), (v, -x)], [0., 0., 1., 1.], 2,
              nt_events=[hb.nt_event_batch(x - 1e-5, mk(0, -np.inf)), hb.nt_event_batch(x - 1e-5, mk(1, np.nan))])
    with pytest.raises(RuntimeError, match="at the batch index 0 - this is not supported"):
        ta.step()
    assert counts == [2, 2]


def case_ev_exception_callback(make):
    """:1473-1558 "ev exception callback": exceptions raised by event callbacks in several batch elements are collected
    into one RuntimeError naming every batch index (the callbacks of the other elements still run; a non-terminal
    callback that raises keeps the terminal one of that element from running); a single exception is re-raised as is."""
    x, v, sys = pendulum_sys()

    def raise0(ta, t, d_sgn, i):
        raise ValueError("hello world 0")

    def raise1(ta, d_sgn, i):
        raise ValueError("hello world 1")

    ta = make(sys, PEND_IC, 4, nt_events=[hb.nt_event_batch(v * v - 1e-10, raise0)],
              t_events=[hb.t_event_batch(v, callback=raise1)])
    with pytest.raises(RuntimeError) as ei:
        ta.propagate_until([4.] * 4)
    msg = str(ei.value)
    assert "Two or more exceptions were raised during the execution of event callbacks in a batch integrator" in msg
    assert "Batch index #0" not in msg and "hello world 1" not in msg and "hello world 0" in msg
    assert all("Batch index #%d" % i in msg for i in (1, 2, 3))
    ta = make(sys, PEND_IC, 4, nt_events=[hb.nt_event_batch(v * v - 1e-10, lambda ta, t, d_sgn, i: None)],
              t_events=[hb.t_event_batch(v, callback=raise1)])
    with pytest.raises(RuntimeError) as ei:
        ta.propagate_until([4.] * 4)
    msg = str(ei.value)
    assert "Batch index #0" not in msg and "hello world 1" in msg
    assert all("Batch index #%d" % i in msg for i in (1, 2, 3))
    # A single exception comes back as it was raised.
    ta = make(sys, [0, 0., 0., 0.03, .25, .25, .25, .28], 4,
              nt_events=[hb.nt_event_batch(v * v - 1e-10, lambda ta, t, d_sgn, i: None)],
              t_events=[hb.t_event_batch(v, callback=raise1)])
    with pytest.raises(ValueError, match="hello world 1"):
        ta.propagate_until([4.] * 4)


def case_events_error(make):
    """:1395-1453 "events error": messages of reset_cooldowns() / the event getters with and without events."""
    x, v, sys = pendulum_sys()
    ic = [0.1, 0.2, 0., 0.]
    msg = "Cannot reset the cooldowns at batch index 2: the batch size for this integrator is only 2"
    ta = make(sys, ic, 2, t_events=[hb.t_event_batch(x)])
    assert ta.with_events()
    with pytest.raises(ValueError, match=msg):
        ta.reset_cooldowns(2)
    # No terminal events: nothing to reset, the calls still work; the cooldown lists are empty.
    ta = make(sys, ic, 2, nt_events=[hb.nt_event_batch(x, lambda ta, t, d_sgn, i: None)])
    assert ta.with_events() and ta.te_cooldowns == [[], []]
    ta.reset_cooldowns()
    ta.reset_cooldowns(0)
    ta.reset_cooldowns(1)
    with pytest.raises(ValueError, match=msg):
        ta.reset_cooldowns(2)
    assert len(ta.get_nt_events()) == 1 and len(ta.get_t_events()) == 0
    ta = make(sys, ic, 2)
    assert not ta.with_events()
    for call in (ta.get_t_events, ta.get_nt_events, ta.reset_cooldowns, lambda: ta.reset_cooldowns(2),
                 lambda: ta.te_cooldowns):
        with pytest.raises(ValueError, match="No events were defined for this integrator"):
            call()


def case_get_set_dtime(make):
    """:1864-1941 "get_set_dtime": the double-length time through dtime / set_dtime(): sizes, normalisation, the checks
    (finite components, |hi| >= |lo|) made before anything is touched."""
    x, v, sys = pendulum_sys()
    # (An event that never triggers: the CPU run of this case goes through the oracle-backed front end.)
    ta = make(sys, [0, 0.01, 0.1, 0.11], 2, nt_events=[hb.nt_event_batch(x - 100., lambda ta, t, d_sgn, i: None)])
    ta.step()
    hi, lo = ta.dtime
    assert hi[0] != 0 and hi[1] != 0 and lo[0] == 0 and lo[1] == 0
    for _ in range(50):
        ta.step()
    with pytest.raises(ValueError, match=r"Invalid number of new times specified in a Taylor integrator in batch mode: "
                                         r"the batch size is 2, but the number of specified times is \(0, 1\)"):
        ta.set_dtime([], [1.])
    hi, lo = (np.array(a) for a in ta.dtime)
    ta.set_dtime(hi, lo)
    assert np.array_equal(ta.dtime[0], hi) and np.array_equal(ta.dtime[1], lo)
    ta.set_dtime([3., -7.], [2., 5.])
    assert list(ta.dtime[0]) == [5., -2.] and list(ta.dtime[1]) == [0., 0.]
    ta.set_dtime([3., -3.], [EPS, EPS])
    assert list(ta.dtime[0]) == [3., -3.] and list(ta.dtime[1]) == [EPS, EPS]
    ta.set_dtime(4., 3.)
    assert list(ta.dtime[0]) == [7., 7.] and list(ta.dtime[1]) == [0., 0.]
    ta.set_dtime(3., EPS)
    assert list(ta.dtime[0]) == [3., 3.] and list(ta.dtime[1]) == [EPS, EPS]
    ta.set_dtime([3., 4.], [1., 2.])
    finite = "The components of the double-length representation of the time coordinate must both be finite"
    order = "must not be smaller in magnitude than the second component"
    for args, msg in (((INF, 1.), finite), ((1., INF), finite), ((3., 4.), order), (([1., INF], [1., 2.]), finite),
                      (([1., .1], [INF, 2.]), finite), (([1., 2.], [1., 3.]), order), (([4., 4.], [8., 3.]), order)):
        with pytest.raises(ValueError, match=msg):
            ta.set_dtime(*args)
    assert list(ta.dtime[0]) == [4., 6.] and list(ta.dtime[1]) == [0., 0.]       # untouched by the failed calls


def case_reset_cooldowns(make):
    """:1726-1751 "reset cooldowns": a terminal event whose callback returns False stops the propagation and leaves a
    cooldown behind; reset_cooldowns() clears every one."""
    x, v, sys = pendulum_sys()
    ta = make(sys, PEND_IC, 4, t_events=[hb.t_event_batch(v, callback=lambda ta, d_sgn, i: False)])
    ta.propagate_until([100.] * 4)
    assert any(cd is not None for lane in ta.te_cooldowns for cd in lane)
    ta.reset_cooldowns()
    assert all(cd is None for lane in ta.te_cooldowns for cd in lane)


def case_param_deduction_from_events(make):
    """:1017-1049 "param deduction from events": the number of runtime parameters is deduced from the event equations
    too (the largest par[] index anywhere, plus one, per batch element)."""
    x, v, sys = pendulum_sys()
    ic = [0.05, 0.06, 0.025, 0.026]
    ncb = lambda ta, t, d_sgn, i: None  # noqa: E731
    ta = make(sys, ic, 2, t_events=[hb.t_event_batch(v - hb.par[0])])
    assert np.asarray(ta.pars).size == 2
    ta = make(sys, ic, 2, nt_events=[hb.nt_event_batch(v - hb.par[1], ncb)])
    assert np.asarray(ta.pars).size == 4
    ta = make(sys, ic, 2, t_events=[hb.t_event_batch(v - hb.par[10])], nt_events=[hb.nt_event_batch(v - hb.par[1], ncb)])
    assert np.asarray(ta.pars).size == 22
    ta.step()                # (and the integrator runs with them: v - par[10] = v crosses zero in the first step)
    assert [r[0] for r in ta.step_res] == [-1, -1] and np.all(np.isfinite(ta.state))
