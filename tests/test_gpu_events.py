"""Event detection on the GPU (include/heyoka_b200.h section E; csrc/ev_kernels.cuh): the batch blocks of
test/batch_event_detection.cpp on the device, and the device against the CPU oracle step by step (same events in the same
order, event times / Taylor coefficients of the event equations / cooldowns / states to the tolerances written below)."""
import numpy as np
import pytest

import event_cases as ec
import heyoka_b200 as hb
import oracle

pytestmark = pytest.mark.gpu


def make(*a, **k):
    return hb.taylor_adaptive_batch(*a, **k)


def test_linear_box_gpu():
    times = ec.case_linear_box(make)
    assert np.allclose(sorted(times), [1 / 8., 1 / 4., 1 / 2., 1.], rtol=1e-15)


def test_glancing_blow_gpu():
    ec.case_glancing_blow(make)


@pytest.mark.parametrize("tol", [0.0, ec.EPS / 100])
def test_multizero_gpu(tol):
    ec.case_multizero(make, tol=tol)


def test_multizero_backward_and_direction_gpu():
    ec.case_multizero(make, backward=True)
    ec.case_multizero_dir(make)


def test_nte_basic_golden_periods_gpu():
    ec.case_nte_basic(make)


@pytest.mark.parametrize("tol", [0.0, ec.EPS / 100])
def test_te_basic_gpu(tol):
    ec.case_te_basic(make, tol=tol)


def test_directions_gpu():
    ec.case_nte_dir(make)
    ec.case_te_dir(make)


def test_te_identical_close_retrigger_gpu():
    ec.case_te_identical(make)
    ec.case_te_close(make)
    ec.case_te_retrigger(make)


def test_te_cooldowns_gpu():
    ec.case_te_custom_cooldown(make)
    ec.case_te_zero_cd(make)


def test_te_propagate_for_gpu():
    ec.case_te_propagate_for(make)


def test_te_damped_pendulum_and_boolean_callback_gpu():
    ec.case_te_damped_pendulum(make)
    ec.case_te_boolean_callback(make)


def test_te_step_end_gpu():
    ec.case_te_step_end(make)


def test_te_propagate_grid_gpu():
    ec.case_te_propagate_grid(make)


@pytest.mark.parametrize("terminal", [False, True])
def test_single_step_gpu_vs_oracle(terminal):
    """test/batch_event_detection.cpp:100-260 on both: same number of triggers, times to 1000 eps, velocities to 10000
    eps (the reference's own bars between its batch and scalar integrators)."""
    tg, vg = ec.case_single_step(make, terminal)
    to, vo = ec.case_single_step(lambda *a, **k: oracle.OracleEventIntegrator(*a, **k), terminal)
    for i in range(4):
        assert len(tg[i]) == len(to[i]) >= 1
        for a, b, c, d in zip(tg[i], to[i], vg[i], vo[i]):
            assert ec.approx(a, b, 1000.) and ec.approx(c, d, 10000.)


@pytest.mark.parametrize("batch", [5, 257])
@pytest.mark.parametrize("backward", [False, True])
def test_event_step_parity_vs_oracle(batch, backward):
    """Lock-step comparison over 60 steps of a pendulum batch with two terminal (one with an explicit cooldown, one
    directional) and two non-terminal events: identical event lists (lane, index, kind, derivative sign), event times to
    2e-12 of the step + 1e-15 (the reference's own bar between its two integrators is 1000 eps, :158), Taylor coefficients of the event equations to 1e-13 (scaled by h^order), cooldown state,
    outcomes, step sizes to 1e-12, states to 1e-12."""
    x, v, sys = ec.pendulum_sys()
    rng = np.random.default_rng(7)
    st = np.stack([rng.uniform(-0.5, 0.5, batch), rng.uniform(-1.0, 1.0, batch)])

    def build(mk):
        return mk(sys, st, batch, t_events=[hb.t_event_batch(v, callback=lambda ta, s, i: True),
                                            hb.t_event_batch(x - 0.1, callback=lambda ta, s, i: True, cooldown=0.05,
                                                             direction=hb.event_direction.positive)],
                  nt_events=[hb.nt_event_batch(v * v - 1e-2, lambda ta, t, s, i: None),
                             hb.nt_event_batch(x * v + 0.05 * hb.cos(hb.time), lambda ta, t, s, i: None,
                                               direction=hb.event_direction.negative)])

    g = build(make)
    o = build(lambda *a, **k: oracle.OracleEventIntegrator(*a, **k))
    assert g._b.kernel_info()["tape"] == "hbm"
    n_events = 0
    for it in range(60):
        if backward:
            g.step_backward()
            o.step_backward()
        else:
            g.step()
            o.step()
        eg, eo = g._b.events(), o._b.events()
        assert [e[:4] for e in eg] == [e[:4] for e in eo], it
        h = o.last_h
        for a, b in zip(eg, eo):
            assert abs(a[4] - b[4]) <= 2e-12 * abs(h[a[0]]) + 1e-15
            assert abs(a[5] - b[5]) <= 1e-10 * max(abs(b[5]), 1.0)
        n_events += len(eg)
        assert [r[0] for r in g.step_res] == [r[0] for r in o.step_res], it
        assert np.max(np.abs(g.last_h / o.last_h - 1)) < 1e-12
        assert np.max(np.abs(g.state - o.state)) < 1e-12
        assert np.array_equal(g.time, o.time) or np.max(np.abs(g.time - o.time)) < 1e-12
        # Taylor coefficients of the event equations, scaled by h^order.
        tg, to_ = g._b.tc_events(4), o._b.tc_events(4)
        pw = np.arange(g.get_order() + 1)[None, :, None]
        hh = np.maximum(np.abs(h), 1e-300)[None, None, :] ** pw
        assert np.max(np.abs(tg - to_) * np.minimum(hh, 1.0)) < 1e-12
        ag, sg, cg = g._b.cooldowns(2)
        ao, so, co = o._b.cooldowns(2)
        assert np.array_equal(ag, ao), it
        assert np.allclose(sg[ag == 1], so[ao == 1], rtol=1e-9, atol=1e-13)
        assert np.allclose(cg[ag == 1], co[ao == 1], rtol=1e-6, atol=1e-15)
    assert n_events > batch  # (every lane met several events)


def test_many_lanes_few_events():
    """65,537 lanes, a terminal and a non-terminal event: most (event, lane) pairs are discarded by the fast exclusion
    check; the records that come back are exactly those of the lanes that cross."""
    x, v, sys = ec.pendulum_sys()
    batch = 65537
    rng = np.random.default_rng(3)
    st = np.stack([rng.uniform(0.05, 0.3, batch), np.zeros(batch)])
    fired = np.zeros(batch, dtype=np.int64)

    def cb(ta, d_sgn, i):
        fired[i] += 1
        return True

    ta = make(sys, st, batch, t_events=[hb.t_event_batch(x, callback=cb)],
              nt_events=[hb.nt_event_batch(v + 2.0, lambda ta, t, s, i: None)])
    for _ in range(8):
        ta.step()
    # x crosses zero after a quarter period (~0.5), then every half period.
    assert np.all(fired <= 3) and np.all(fired >= 1)
    idx = np.arange(0, batch, 4099)
    o = oracle.OracleEventIntegrator(sys, st[:, idx[:16]], len(idx[:16]), t_events=[hb.t_event_batch(x, callback=lambda ta, s, i: True)],
                                     nt_events=[hb.nt_event_batch(v + 2.0, lambda ta, t, s, i: None)])
    for _ in range(8):
        o.step()
    assert np.max(np.abs(o.state - ta.state[:, idx[:16]])) < 1e-12


def test_event_api_errors_gpu():
    x, v, sys = ec.pendulum_sys()
    ta = make(sys, ec.PEND_IC, 4, t_events=[hb.t_event_batch(v)])
    with pytest.raises(NotImplementedError, match="lock-step"):
        ta._b.propagate_until(np.full(4, 1.0))
    ta.reset_cooldowns()
    ta.reset_cooldowns(2)
    with pytest.raises(ValueError, match="Cannot reset the cooldowns at batch index 4: the batch size for this integrator is only 4"):
        ta.reset_cooldowns(4)
    tb = make(sys, ec.PEND_IC, 4)
    with pytest.raises(ValueError, match="No events"):
        tb.reset_cooldowns()


def test_continuous_output_with_events_gpu():
    """propagate_until(c_output=True) of an integrator WITH events (src/taylor_adaptive_batch.cpp:1320-1346 inside the
    lock-step loop at :1372-1527): every iteration of the host loop is recorded on the device (hy_cout_rec_*). Harmonic
    oscillators x = A sin t: the non-terminal event v = 0 fires at pi / 2 + k pi in every lane; the continuous output
    reproduces A sin t / A cos t over the whole range, its bounds are the initial and the final time, one recorded step per
    iteration. With a terminal event x = -A / 2 (t = 7 pi / 6) the recording ends at the event time."""
    from test_oracle_golden import sys_oscillator
    x, v = hb.make_vars("x", "v")
    amp = np.array([1.0, 1.1, 1.2, 1.3])
    ic = np.stack([np.zeros(4), amp])
    hits = []
    ev = hb.nt_event_batch(v, lambda ta, t, d_sgn, lane: hits.append((lane, t)))
    ta = hb.taylor_adaptive_batch(sys_oscillator(), ic, 4, nt_events=[ev])
    tf = np.array([10.0, 10.5, 11.0, 11.5])
    calls = []
    co = ta.propagate_until(tf, c_output=True, callback=lambda t: calls.append(1) or True)
    assert co is not None and np.array_equal(ta.time, tf)
    assert co.get_n_steps() == len(calls)
    lb, ub = co.get_bounds()
    assert np.all(lb == 0) and np.array_equal(ub, tf)
    for lane in range(4):
        got = sorted(t for ln, t in hits if ln == lane)
        exp = [np.pi / 2 + k * np.pi for k in range(4) if np.pi / 2 + k * np.pi < tf[lane]]
        assert np.allclose(got, exp, rtol=1e-14)
    for tq in (0.0, 0.3, 2.0, 5.5, 9.9):
        s = co(tq)
        assert np.max(np.abs(s[0] - amp * np.sin(tq))) < 1e-13 and np.max(np.abs(s[1] - amp * np.cos(tq))) < 1e-13
    # The same propagation without events: the recordings agree to rounding (the step sizes need not be the same: the
    # event equations take part in the step-size estimate).
    ref = hb.taylor_adaptive_batch(sys_oscillator(), ic, 4)
    co_ref = ref.propagate_until(tf, c_output=True)
    assert np.max(np.abs(co(4.2) - co_ref(4.2))) < 1e-13
    # A stopping terminal event: the recording ends where the integrator stopped.
    te = hb.t_event_batch(x + 0.5, direction=hb.event_direction.negative)
    tb = hb.taylor_adaptive_batch(sys_oscillator(), np.stack([np.zeros(4), np.ones(4)]), 4, t_events=[te])
    co2 = tb.propagate_until(20.0, c_output=True)
    assert co2 is not None
    t_ev = 7 * np.pi / 6
    assert np.allclose(tb.time, t_ev, rtol=1e-14) and np.array_equal(co2.get_bounds()[1], tb.time)
    assert np.max(np.abs(co2(3.0)[0] - np.sin(3.0))) < 1e-13
    assert all(r[0] == -1 for r in tb.propagate_res)  # outcome = -index - 1 of the stopping terminal event
