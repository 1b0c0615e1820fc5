"""Shared helpers for the parity tests: systems named by BASELINE.json's configs + tolerances."""
import json
import os

import numpy as np

import heyoka_b200 as hb

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EPS = np.finfo(np.float64).eps


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def approx(a, b, tol_eps):
    """test/test_utils.hpp:68-80 `approximately`: |a - b| <= eps * tol * max(|a|, |b|) (absolute near zero)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300)
    return np.all(np.abs(a - b) <= EPS * tol_eps * scale)


def sig_digits_equal(val, printed, digits=6):
    """A value printed with `digits` significant digits matches `val`."""
    val, printed = np.asarray(val, dtype=np.float64), np.asarray(printed, dtype=np.float64)
    tol = 0.6 * 10.0 ** (np.floor(np.log10(np.maximum(np.abs(printed), 1e-300))) - (digits - 1))
    return np.all(np.abs(val - printed) <= tol + 1e-300)


def decimals_equal(val, printed, decimals=6):
    """A value printed in fixed notation with `decimals` decimals matches `val`."""
    val, printed = np.asarray(val, dtype=np.float64), np.asarray(printed, dtype=np.float64)
    return np.all(np.abs(val - printed) <= 0.6 * 10.0 ** (-decimals))


def nbody_rel_err(a, b):
    """Final-state agreement for N-body states [6 n_bodies, batch]: per body and lane, the error of the position
    (velocity) vector relative to the norm of that vector. (Component-wise relative error is meaningless for
    components that pass through zero, e.g. the Sun's barycentric coordinates ~1e-5 AU.)"""
    a = np.asarray(a).reshape(-1, 2, 3, np.asarray(a).shape[-1])
    b = np.asarray(b).reshape(a.shape)
    return float(np.max(np.linalg.norm(a - b, axis=2) / np.linalg.norm(b, axis=2)))


# ---- systems ------------------------------------------------------------------------------------
def sys_pendulum():
    x, v = hb.make_vars("x", "v")
    return [(x, v), (v, -9.8 * hb.sin(x))]


def sys_tutorial():
    """x' = v, v' = cos(t) - par[0]*v - sin(x) (tutorial/batch_mode.cpp:52-62)."""
    x, v = hb.make_vars("x", "v")
    return [(x, v), (v, hb.cos(hb.time) - hb.par[0] * v - hb.sin(x))]


OUTER_SS_MASSES = [1.00000597682, 1 / 1047.355, 1 / 3501.6, 1 / 22869., 1 / 19314., 7.4074074e-09]
OUTER_SS_G = 0.01720209895 * 0.01720209895 * 365 * 365

# benchmark/outer_ss_long_term_batch.cpp:75-94 (positions in AU, velocities in AU/day, scaled by 365 below).
_OUTER_SS_IC = [
    # Sun.
    -4.06428567034226e-3, -6.08813756435987e-3, -1.66162304225834e-6, +6.69048890636161e-6, -6.33922479583593e-6,
    -3.13202145590767e-9,
    # Jupiter.
    +3.40546614227466e+0, +3.62978190075864e+0, +3.42386261766577e-2, -5.59797969310664e-3, +5.51815399480116e-3,
    -2.66711392865591e-6,
    # Saturn.
    +6.60801554403466e+0, +6.38084674585064e+0, -1.36145963724542e-1, -4.17354020307064e-3, +3.99723751748116e-3,
    +1.67206320571441e-5,
    # Uranus.
    +1.11636331405597e+1, +1.60373479057256e+1, +3.61783279369958e-1, -3.25884806151064e-3, +2.06438412905916e-3,
    -2.17699042180559e-5,
    # Neptune.
    -3.01777243405203e+1, +1.91155314998064e+0, -1.53887595621042e-1, -2.17471785045538e-4, -3.11361111025884e-3,
    +3.58344705491441e-5,
    # Pluto.
    -2.13858977531573e+1, +3.20719104739886e+1, +2.49245689556096e+0, -1.76936577252484e-3, -2.06720938381724e-3,
    +6.58091931493844e-4,
]


def outer_ss_ic():
    ic = np.array(_OUTER_SS_IC).reshape(6, 6).copy()
    ic[:, 3:] *= 365.
    return ic.reshape(-1)


def sys_outer_ss():
    return hb.model.nbody(6, masses=OUTER_SS_MASSES, Gconst=OUTER_SS_G)


def outer_ss_batch_state(batch, perturb=1e-3, seed=42):
    """Perturbed copies of the outer Solar System (rule of benchmark/outer_ss_long_term_batch.cpp:103-108:
    x += |x| * U(-1, 1) * perturb, applied in memory order var-major / batch-minor; numpy RNG instead of
    std::mt19937 since only the distribution matters for throughput and parity uses the same arrays)."""
    rng = np.random.default_rng(seed)
    ic = outer_ss_ic()
    st = np.repeat(ic[:, None], batch, axis=1)
    st += np.abs(st) * rng.uniform(-1., 1., st.shape) * perturb
    return st


def sys_two_body():
    return hb.model.nbody(2, masses=[1., 0.])


def two_body_batch_state(batch, seed=7):
    """Circular orbits of radius a ~ U(0.5, 2) around a unit mass at rest (benchmark/two_body_step_batch.cpp:41-55)."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(0.5, 2.0, batch)
    st = np.zeros((12, batch))
    st[6] = a              # x_1
    st[10] = a ** -0.5     # vy_1
    return st


def kep_to_cart(a, e, inc, om, Om, nu, mu):
    """test/test_utils.hpp:204-233 (Keplerian elements -> Cartesian state)."""
    p = a * (1 - e * e)
    r = p / (1 + e * np.cos(nu))
    x_p, y_p = r * np.cos(nu), r * np.sin(nu)
    h = np.sqrt(mu * p)
    vx_p, vy_p = -mu / h * np.sin(nu), mu / h * (e + np.cos(nu))
    cO, sO, co, so, ci, si = np.cos(Om), np.sin(Om), np.cos(om), np.sin(om), np.cos(inc), np.sin(inc)
    R = np.array([[cO * co - sO * so * ci, -cO * so - sO * co * ci], [sO * co + cO * so * ci, -sO * so + cO * co * ci],
                  [so * si, co * si]])
    pos = R @ np.array([x_p, y_p])
    vel = R @ np.array([vx_p, vy_p])
    return pos, vel


# ---- BASELINE.json configs[2]: model::nbody N = 32 (SURVEY.md 8(d) "N32") ----
N32_MASSES = np.array([1.0] + [1e-4 * (1 + i / 32.0) for i in range(1, 32)])


def sys_nbody32():
    return hb.model.nbody(32, masses=N32_MASSES, Gconst=1.0)


def nbody32_batch_state(batch, seed=1234):
    """Planets on a = 1 + 0.5 i, e ~ U(0, 0.05), inc ~ U(0, 0.05), angles ~ U(0, 2 pi) (the formulas of
    kep_to_cart(), vectorised over the planets); the star is placed so that the centre of mass is at rest at the
    origin. One generator per lane (seed + lane)."""
    st = np.zeros((32 * 6, batch))
    a = 1.0 + 0.5 * np.arange(1, 32)
    mu = 1.0 + N32_MASSES[1:]
    for lane in range(batch):
        rng = np.random.default_rng(seed + lane)
        e, inc = rng.uniform(0, 0.05, 31), rng.uniform(0, 0.05, 31)
        om, Om, nu = rng.uniform(0, 2 * np.pi, (3, 31))
        p = a * (1 - e * e)
        r = p / (1 + e * np.cos(nu))
        xp, yp = r * np.cos(nu), r * np.sin(nu)
        h = np.sqrt(mu * p)
        vxp, vyp = -mu / h * np.sin(nu), mu / h * (e + np.cos(nu))
        cO, sO, co, so, ci, si = np.cos(Om), np.sin(Om), np.cos(om), np.sin(om), np.cos(inc), np.sin(inc)
        R = np.array([[cO * co - sO * so * ci, -cO * so - sO * co * ci], [sO * co + cO * so * ci, -sO * so + cO * co * ci],
                      [so * si, co * si]])  # [3, 2, planets]
        pos = R[:, 0] * xp + R[:, 1] * yp
        vel = R[:, 0] * vxp + R[:, 1] * vyp
        body = st[6:, lane].reshape(31, 6)
        body[:, :3] = pos.T
        body[:, 3:] = vel.T
        st[6:, lane] = body.reshape(-1)
        for k in range(6):
            st[k, lane] = -np.sum(N32_MASSES[1:] * st[6 + k::6, lane]) / N32_MASSES[0]
    return st


# ---- BASELINE.json configs[4]: model::ffnn right-hand side, 3 x 64 tanh (SURVEY.md 8(d) "NN") ----
FFNN_TOL = 1e-12  # -> order 15


def sys_ffnn(seed=11):
    """x' = ffnn(x): 4 inputs, hidden layers {64, 64, 64} with tanh, 4 linear outputs; weights and biases are numbers
    ~ N(0, 1 / fan_in) in the layout of src/model/ffnn.cpp:75-78 ([W01, W12, W23, W34 | b1..b4], row-major)."""
    rng = np.random.default_rng(seed)
    sizes = [4, 64, 64, 64, 4]
    w = [rng.normal(0, 1 / np.sqrt(sizes[i]), sizes[i] * sizes[i + 1]) for i in range(4)]
    b = [rng.normal(0, 0.1, sizes[i + 1]) for i in range(4)]
    xs = hb.make_vars("x0", "x1", "x2", "x3")
    out = hb.model.ffnn(list(xs), [64, 64, 64], 4, ["tanh", "tanh", "tanh", "identity"], nn_wb=np.concatenate(w + b))
    return [(xs[i], out[i]) for i in range(4)]


def ffnn_batch_state(batch, seed=13):
    return np.random.default_rng(seed).uniform(-1, 1, (4, batch))


# ---- test/two_body_batch.cpp:60-190: equal-mass two-body problem written by hand, Keplerian elements ----
def sys_two_body_symmetric():
    """The system of test/two_body_batch.cpp:63-80 (variables in the reference's order)."""
    vx0, vx1, vy0, vy1, vz0, vz1, x0, x1, y0, y1, z0, z1 = hb.make_vars("vx0", "vx1", "vy0", "vy1", "vz0", "vz1",
                                                                        "x0", "x1", "y0", "y1", "z0", "z1")
    x01, y01, z01 = x1 - x0, y1 - y0, z1 - z0
    r01_m3 = hb.pow(x01 * x01 + y01 * y01 + z01 * z01, hb.expression(-3.) / hb.expression(2.))
    return [(vx0, x01 * r01_m3), (vx1, -x01 * r01_m3), (vy0, y01 * r01_m3), (vy1, -y01 * r01_m3),
            (vz0, z01 * r01_m3), (vz1, -z01 * r01_m3), (x0, vx0), (x1, vx1), (y0, vy0), (y1, vy1), (z0, vz0), (z1, vz1)]


def cart_to_kep(x, v, mu):
    """test/test_utils.hpp:172-202 (Cartesian state -> a, e, i, omega, Omega, nu)."""
    x, v = np.asarray(x, dtype=float), np.asarray(v, dtype=float)
    h = np.cross(x, v)
    e_v = np.cross(v, h) / mu - x / np.linalg.norm(x)
    n = np.array([-h[1], h[0], 0.0])
    nu = np.arccos(np.dot(e_v, x) / (np.linalg.norm(e_v) * np.linalg.norm(x)))
    if np.dot(x, v) < 0:
        nu = 2 * np.pi - nu
    inc = np.arccos(h[2] / np.linalg.norm(h))
    e = np.linalg.norm(e_v)
    Om = np.arccos(n[0] / np.linalg.norm(n))
    if n[1] < 0:
        Om = 2 * np.pi - Om
    om = np.arccos(np.dot(n, e_v) / (np.linalg.norm(n) * np.linalg.norm(e_v)))
    if e_v[2] < 0:
        om = 2 * np.pi - om
    a = 1 / (2 / np.linalg.norm(x) - np.dot(v, v) / mu)
    return np.array([a, e, inc, om, Om, nu])


def two_body_kepler_fixture(batch=4, seed=5):
    """Random elements a ~ U(0.1, 10), e ~ U(0.1, 0.5), i ~ U(0.1, 3.0) (3.13 in the reference: the node of a
    nearly equatorial retrograde orbit is ill-conditioned at the 1e4 eps of the check), angles ~ U(0.1, 6.28); the two bodies sit at
    +-x, +-v of the Keplerian orbit with mu = 1/4 (test/two_body_batch.cpp:83-112)."""
    rng = np.random.default_rng(seed)
    kep = np.column_stack([rng.uniform(0.1, 10, batch), rng.uniform(0.1, 0.5, batch), rng.uniform(0.1, 3.0, batch),
                           rng.uniform(0.1, 6.28, batch), rng.uniform(0.1, 6.28, batch), rng.uniform(0.1, 6.28, batch)])
    st = np.zeros((12, batch))
    for i in range(batch):
        pos, vel = kep_to_cart(*kep[i], 0.25)
        for k in range(3):
            st[2 * k, i], st[2 * k + 1, i] = vel[k], -vel[k]
            st[6 + 2 * k, i], st[6 + 2 * k + 1, i] = pos[k], -pos[k]
    return kep, st


def check_kepler_conservation(st, kep, approx_fn, tol_mul=1e4):
    """test/two_body_batch.cpp:166-190 for every lane of st[12, batch]."""
    for i in range(st.shape[1]):
        for body in (0, 1):
            k = cart_to_kep(st[6 + body:12:2, i], st[body:6:2, i], 0.25)
            assert approx_fn(k[0], kep[i, 0], tol_mul) and approx_fn(k[1], kep[i, 1], tol_mul)
            assert approx_fn(k[2], kep[i, 2], tol_mul) and approx_fn(k[4], kep[i, 4], tol_mul)
            assert approx_fn(abs(np.cos(k[3])), abs(np.cos(kep[i, 3])), tol_mul)
