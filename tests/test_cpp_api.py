"""Builds and runs the C++ API test program (tests/cpp/test_batch_api.cpp) against the in-tree library."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "test_batch_api")


def _build():
    src = os.path.join(ROOT, "tests", "cpp", "test_batch_api.cpp")
    lib = os.path.join(ROOT, "heyoka_b200", "lib")
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(
            os.path.join(lib, "libheyoka_b200.so"))):
        subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), src, "-o", EXE, "-L" + lib,
                        "-lheyoka_b200", "-Wl,-rpath," + lib], check=True)
    return EXE


def test_cpp_argument_validation():
    res = subprocess.run([_build(), "cpu"], capture_output=True, text=True)
    assert res.returncode == 0 and "ALL PASSED (cpu)" in res.stdout, res.stdout + res.stderr


@pytest.mark.gpu
def test_cpp_api_on_gpu():
    res = subprocess.run([_build(), "gpu"], capture_output=True, text=True, timeout=1200)
    assert res.returncode == 0 and "ALL PASSED (gpu)" in res.stdout, res.stdout[-3000:] + res.stderr[-2000:]


def test_planner_host_logic():
    """tests/cpp/test_plan.cpp: tape sizes of the shared-memory / tensor-memory plans, superinstructions, levels and
    the bank-conflict-free row layout for the 6-body program (host only)."""
    src = os.path.join(ROOT, "tests", "cpp", "test_plan.cpp")
    lib = os.path.join(ROOT, "heyoka_b200", "lib")
    exe = os.path.join(ROOT, "build", "test_plan")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "heyoka_b200", "csrc"), src, "-o", exe, "-L" + lib, "-lheyoka_b200",
                    "-Wl,-rpath," + lib], check=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0 and "ALL PASSED" in res.stdout, res.stdout + res.stderr
