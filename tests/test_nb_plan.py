"""The N-body kernel's host logic, without a GPU: the planner (heyoka_b200/csrc/nb_plan.cpp) and the two-orders-at-a-time
arithmetic of heyoka_b200/csrc/nb_core.hpp, emulated on plain arrays (tests/cpp/nb_emul.cpp) and compared against the
oracle's one-order-at-a-time jet. The same nb_core.hpp is what k_nb (nb_kernel.cuh) runs on the device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import heyoka_b200 as hb
import oracle
from common import (nbody32_batch_state, outer_ss_batch_state, sys_ffnn, sys_nbody32, sys_outer_ss, sys_pendulum,
                    sys_two_body, sys_two_body_symmetric, two_body_batch_state)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "heyoka_b200", "csrc")
LIBDIR = os.path.join(ROOT, "heyoka_b200", "lib")
SO = os.path.join(ROOT, "build", "libnb_emul.so")


def _emul():
    src = os.path.join(ROOT, "tests", "cpp", "nb_emul.cpp")
    deps = [src, os.path.join(CSRC, "nb_core.hpp"), os.path.join(CSRC, "nb_plan.hpp"), os.path.join(CSRC, "nb_desc.hpp"),
            os.path.join(CSRC, "nb_plan.cpp"),
            os.path.join(LIBDIR, "libheyoka_b200.so")]
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in deps):
        # -ffp-contract=off: the only fused operations are the explicit fma() calls, like -fmad=false on the device.
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "include"),
                        "-I" + CSRC, src, "-o", SO, "-L" + LIBDIR, "-lheyoka_b200", "-Wl,-rpath," + LIBDIR], check=True)
    lib = C.CDLL(SO)
    lib.nb_emul_plan.restype = C.c_int
    lib.nb_emul_jet.restype = C.c_int
    return lib


def _plan(lib, P):
    out = (C.c_uint32 * 6)()
    why = C.create_string_buffer(256)
    lib.nb_emul_plan(P._h, out, why, C.c_size_t(256))
    return list(out), why.value.decode()


def _emul_jet(lib, P, state_lanes, n_pairs, tt, lt):
    """Jets of `lt` lanes run by one emulated team of `tt` threads. state_lanes: [n_eq, lt]."""
    n_ord = 2 * ((P.order + 1) // 2)
    coef = np.zeros((lt, P.n_eq, P.order + 1))
    u_idx = np.zeros(n_pairs * 8, dtype=np.uint32)
    u_rows = np.zeros((n_pairs, lt, 8, n_ord))
    st = np.ascontiguousarray(state_lanes, dtype=np.float64)
    assert st.shape == (P.n_eq, lt)
    rc = lib.nb_emul_jet(P._h, C.c_uint32(tt), C.c_uint32(lt), st.ctypes.data_as(C.POINTER(C.c_double)),
                         coef.ctypes.data_as(C.POINTER(C.c_double)), u_idx.ctypes.data_as(C.POINTER(C.c_uint32)),
                         u_rows.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0, rc
    return coef, u_idx, u_rows


def test_plan_shapes():
    lib = _emul()
    # 6 bodies: 15 pair interactions, 18 positions, 45 products, 18 sums of 5 terms in one level.
    info, why = _plan(lib, hb.Program(sys_outer_ss(), high_accuracy=True))
    assert info == [1, 15, 18, 90, 18, 1], (info, why)  # outputs: m_k and the rescaled n_k of every pair
    # Two bodies, one of them massless: one pair interaction, the accelerations of the massive body are the number 0.
    info, why = _plan(lib, hb.Program(sys_two_body()))
    assert info == [1, 1, 6, 3, 6, 1], (info, why)
    # 32 bodies: 496 pair interactions; 31-term sums are nested (4 partial sums + 1 per acceleration).
    info, why = _plan(lib, hb.Program(sys_nbody32()))
    assert info[:4] == [1, 496, 96, 496 * 6 + 96 * 4] and info[4] == 96 * 5 and info[5] == 2, (info, why)
    # Programs that are not N-body-shaped are refused (they run on the generic cooperative kernel).
    for s in (sys_pendulum(), sys_ffnn(), sys_two_body_symmetric()):
        info, why = _plan(lib, hb.Program(s))
        assert info[0] == 0 and why != "", (info, why)


@pytest.mark.parametrize("name,tt,lt", [("outer_ss", 32, 2), ("outer_ss", 32, 1), ("two_body", 32, 32), ("two_body", 32, 4),
                                        ("nbody32", 512, 1), ("outer_ss_odd_order", 32, 2)])
def test_two_order_blocks_match_oracle(name, tt, lt):
    """Every coefficient the blocked evaluation produces (coordinate differences, r^2, r^alpha, products, state
    variables) equals the oracle's sequential+FMA jet bit for bit, for the team shapes the kernel uses (threads per
    team, lanes per team): the pair threads, the output-slot layout and the per-thread role table of the sums."""
    lib = _emul()
    if name == "outer_ss":
        P, st = hb.Program(sys_outer_ss(), high_accuracy=True), outer_ss_batch_state(lt)
    elif name == "outer_ss_odd_order":
        P, st = hb.Program(sys_outer_ss(), tol=1e-12), outer_ss_batch_state(lt)
        assert P.order % 2 == 1
    elif name == "two_body":
        P, st = hb.Program(sys_two_body()), two_body_batch_state(lt)
    else:
        P, st = hb.Program(sys_nbody32()), nbody32_batch_state(lt)
    info, why = _plan(lib, P)
    assert info[0] == 1, why
    coef, u_idx, u_rows = _emul_jet(lib, P, st, info[1], tt, lt)
    for lane in range(lt):
        tape = oracle.jet(P, st, lane=lane, mode=oracle.FMA)  # [order + 1, n_uvars]; u variables: orders < order
        assert np.array_equal(coef[lane], tape[:, :P.n_eq].T), np.max(np.abs(coef[lane] - tape[:, :P.n_eq].T))
        ref = tape[:P.order, u_idx].T.reshape(info[1], 8, P.order)
        got = u_rows[:, lane, :, :P.order]
        assert np.array_equal(got, ref), np.argwhere(got != ref)[:5]
