"""Full-size configurations of BASELINE.json, checked through size-independent properties (the oracle
is too slow at these sizes) plus a random sample of lanes against the oracle."""
import numpy as np
import pytest

import heyoka_b200 as hb
import oracle
from common import (OUTER_SS_G, OUTER_SS_MASSES, nbody_rel_err, outer_ss_batch_state, sys_outer_ss, sys_two_body,
                    two_body_batch_state)

pytestmark = pytest.mark.gpu


def nbody_energy(st, masses, G):
    """Total energy per lane; st is [6 * n_bodies, batch] with (x, y, z, vx, vy, vz) per body."""
    n = len(masses)
    pos = st.reshape(n, 6, -1)[:, :3]
    vel = st.reshape(n, 6, -1)[:, 3:]
    e = sum(0.5 * masses[i] * np.sum(vel[i] ** 2, axis=0) for i in range(n))
    for i in range(n):
        for j in range(i + 1, n):
            e -= G * masses[i] * masses[j] / np.sqrt(np.sum((pos[i] - pos[j]) ** 2, axis=0))
    return e


def test_two_body_single_step_2pow24():
    """two_body_step_batch: 16,777,216 lanes, one step. Circular orbits of radius a around a unit mass:
    after the step the massless body is still on the circle, has advanced by h * a^-3/2 radians, and the
    chosen h scales like the orbital period."""
    batch = 1 << 24
    st = two_body_batch_state(batch)
    a = st[6].copy()
    P = hb.Program(sys_two_body())
    b = hb.Batch(P, batch)
    b.upload(st, None, np.zeros(batch), np.zeros(batch))
    b.step()
    new, t_hi, t_lo, h = b.download()
    oc, _ = b.step_res()
    assert np.all(oc == hb.taylor_outcome.success)
    assert np.array_equal(t_hi, h) and np.all(t_lo == 0)
    r = np.sqrt(new[6] ** 2 + new[7] ** 2)
    assert np.max(np.abs(r / a - 1)) < 1e-14
    ang = np.arctan2(new[7], new[6])
    assert np.max(np.abs(ang - h * a ** -1.5)) < 1e-14
    assert np.all(new[:6] == 0) and np.all(new[8] == 0)
    # h / period is the same for every lane (scale invariance of the estimator in the relative-tolerance
    # regime a > 1). Not to rounding: h comes from the order-19/20 coefficients, whose RELATIVE accuracy is
    # ~1e-9 (they are tiny and built from cancelling sums); their contribution to the state is < 1 ulp.
    big = a > 1.
    ratio = h[big] * a[big] ** -1.5
    assert np.ptp(ratio) / np.mean(ratio) < 1e-7
    # a random sample of lanes against the oracle
    idx = np.random.default_rng(0).choice(batch, 64, replace=False)
    o = oracle.OracleIntegrator(P, st[:, idx], 64, mode=oracle.FMA)
    o.step()
    assert np.max(np.abs(o.last_h / h[idx] - 1)) < 1e-13
    assert np.max(np.abs(o.state - new[:, idx])) < 1e-13


def test_outer_ss_2pow20_energy_and_sample():
    """outer_ss_long_term_batch at the full batch of 1,048,576 lanes, short horizon (5 years): every lane
    conserves energy to rounding, lands exactly on the final time, and a random sample of lanes matches
    the oracle step for step."""
    batch = 1 << 20
    st = outer_ss_batch_state(batch)
    P = hb.Program(sys_outer_ss(), high_accuracy=True)
    b = hb.Batch(P, batch)
    b.upload(st, None, np.zeros(batch), np.zeros(batch))
    b.propagate_until(5.0)
    new, t_hi, t_lo, _ = b.download()
    oc, mn, mx, ns = b.prop_res()
    assert np.all(oc == hb.taylor_outcome.time_limit)
    assert np.all(t_hi == 5.0)
    assert ns.min() >= 5 and ns.max() <= 30
    e0 = nbody_energy(st, OUTER_SS_MASSES, OUTER_SS_G)
    e1 = nbody_energy(new, OUTER_SS_MASSES, OUTER_SS_G)
    assert np.max(np.abs(e1 / e0 - 1)) < 5e-15 * 20
    idx = np.random.default_rng(1).choice(batch, 48, replace=False)
    o = oracle.OracleIntegrator(P, st[:, idx], 48, mode=oracle.FMA)
    o.propagate_until(5.0)
    assert np.array_equal(o.n_steps, ns[idx])
    assert nbody_rel_err(new[:, idx], o.state) < 1e-12


def test_nbody32_shard_8192_energy_and_sample():
    """model::nbody N = 32, 65,536 ICs sharded over 8 GPUs = 8192 lanes per GPU (BASELINE.json configs[2]); the tape
    (6624 u variables x 21 orders per lane) lives in HBM. Energy conservation and exact landing on t_final for every
    lane, a random sample of lanes against the oracle (identical step counts)."""
    from common import N32_MASSES, nbody32_batch_state, sys_nbody32
    batch, t_final = 8192, 1.0
    st = nbody32_batch_state(batch)
    P = hb.Program(sys_nbody32())
    b = hb.Batch(P, batch)
    assert b.kernel_info()["tape"] == "nbody-cta"
    b.upload(st, None, np.zeros(batch), np.zeros(batch))
    b.propagate_until(np.full(batch, t_final))
    new, t_hi, t_lo, _ = b.download()
    oc, min_h, max_h, ns = b.prop_res()
    assert np.all(oc == hb.taylor_outcome.time_limit) and np.all(t_hi == t_final) and np.all(np.abs(t_lo) < 1e-15)
    # (Random phases: a few lanes start with two planets close to each other and need hundreds of steps.)
    assert ns.min() >= 1 and ns.max() <= 5000
    e0, e1 = nbody_energy(st, N32_MASSES, 1.0), nbody_energy(new, N32_MASSES, 1.0)
    de = np.abs(e1 / e0 - 1)
    # (The close-encounter lanes lose digits in the potential energy itself: bounded separately.)
    assert np.median(de) < 1e-14 and np.quantile(de, 0.99) < 1e-13 and np.max(de) < 1e-9
    # Every lane against the oracle's 8-lane port for the step counts, a sample for the states.
    o = oracle.OracleIntegrator(P, st, batch, mode=oracle.FMA, width=8)
    o.propagate_until(t_final, lockstep=False, n_threads=8)
    assert np.array_equal(ns, o.n_steps)
    idx = np.random.default_rng(1).choice(batch, 64, replace=False)
    assert nbody_rel_err(new[:, idx], o.state[:, idx]) < 1e-12


def test_ffnn_262144_step_and_sample():
    """model::ffnn right-hand side (3 x 64 tanh, order 15), 262,144 lanes (BASELINE.json configs[4]): 10476 u
    variables per lane, tape in HBM. One step and a short propagation; every lane lands on t_final; a random sample
    of lanes against the oracle."""
    from common import FFNN_TOL, ffnn_batch_state, sys_ffnn
    batch = 1 << 18
    st = ffnn_batch_state(batch)
    P = hb.Program(sys_ffnn(), tol=FFNN_TOL)
    b = hb.Batch(P, batch)
    b.upload(st, None, np.zeros(batch), np.zeros(batch))
    b.step()
    new, t_hi, _, h = b.download()
    oc, _ = b.step_res()
    assert np.all(oc == hb.taylor_outcome.success) and np.array_equal(t_hi, h) and np.all(h > 0)
    idx = np.random.default_rng(2).choice(batch, 32, replace=False)
    o = oracle.OracleIntegrator(P, st[:, idx], 32, mode=oracle.FMA)
    o.step()
    assert np.max(np.abs(o.last_h / h[idx] - 1)) < 1e-11
    assert np.max(np.abs(o.state - new[:, idx])) < 1e-13
    b.propagate_until(np.full(batch, 0.5))
    new, t_hi, t_lo, _ = b.download()
    oc, _, _, ns = b.prop_res()
    assert np.all(oc == hb.taylor_outcome.time_limit) and np.all(t_hi == 0.5) and np.all(np.abs(t_lo) < 1e-15)
    assert np.all(np.isfinite(new))
    o.propagate_until(0.5)
    assert np.array_equal(ns[idx], o.n_steps)
    assert np.max(np.abs(o.state - new[:, idx])) < 1e-12


def test_outer_ss_long_horizon_1000yr():
    """Long horizon (1000 yr of the outer Solar System, ~1370 steps per lane; the bench propagates 20 yr):
    4096 lanes on the GPU, a sample of 24 against the oracle - identical step counts, final state to 1e-9 (per-body
    vector norms; 1.4e3 steps of 1e-16-level differences in pow / sqrt), energy conserved to 1e-13 in every lane, every
    lane lands exactly on t = 1000."""
    batch = 4096
    st = outer_ss_batch_state(batch, seed=5)
    P = hb.Program(sys_outer_ss(), high_accuracy=True)
    ta = hb.taylor_adaptive_batch(sys_outer_ss(), st, batch, high_accuracy=True)
    e0 = nbody_energy(st, OUTER_SS_MASSES, OUTER_SS_G)
    ta.propagate_until(1000.0)
    assert all(r[0] == hb.taylor_outcome.time_limit for r in ta.propagate_res)
    assert np.all(ta.time == 1000.0)
    n_steps = np.array([r[3] for r in ta.propagate_res])
    assert n_steps.min() > 1000 and n_steps.max() < 2000
    e1 = nbody_energy(ta.state, OUTER_SS_MASSES, OUTER_SS_G)
    assert np.max(np.abs(e1 / e0 - 1)) < 1e-13
    idx = np.random.default_rng(1).choice(batch, 24, replace=False)
    o = oracle.OracleIntegrator(P, st[:, idx], len(idx), mode=oracle.FMA)
    o.propagate_until(1000.0, lockstep=False)
    assert [int(s) for s in o.n_steps] == [int(n_steps[i]) for i in idx]
    assert nbody_rel_err(ta.state[:, idx], o.state) < 1e-9
