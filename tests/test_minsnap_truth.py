"""K1's arithmetic (csrc/minsnap_core.cuh, compiled for the host by tests/devtools/host_probe_minsnap.cpp) against an
EXTENDED-PRECISION solve of the reference's KKT system (mpmath, 50 digits) -- including wide-volume missions on which the
reference's own LAPACK branches lose digits (cond(K) up to 1e17, lstsq rank-truncates).  Shows that the
reduced block-tridiagonal form computes THE minimiser of the reference's problem, not an approximation of its solver."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import minsnap_np

mpmath = pytest.importorskip("mpmath")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def k1_host(tmp_path_factory):
    so = tmp_path_factory.mktemp("k1") / "libk1_host.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-I", os.path.join(ROOT, "uav-autonomous-control_b200", "csrc"),
                    os.path.join(ROOT, "tests", "devtools", "host_probe_minsnap.cpp"), "-o", str(so)], check=True)
    L = ctypes.CDLL(str(so))
    dp = ctypes.POINTER(ctypes.c_double)
    L.probe_solve.argtypes = [dp, ctypes.c_double, ctypes.c_int, ctypes.c_double, dp, dp, ctypes.c_int]

    def solve(w, v, big=False):
        w = np.ascontiguousarray(w, float)
        S = len(w) - 1
        c, t = np.empty((8 * S, 3)), np.empty(S)
        rc = L.probe_solve(w.ctypes.data_as(dp), float(v), S, 1.5, c.ctypes.data_as(dp), t.ctypes.data_as(dp), int(big or S > 8))
        assert rc == 0
        return c, t
    return solve


def _truth(w, v):
    T = minsnap_np.segment_times(w, v)
    K, rhs = minsnap_np.kkt_system(w, T)
    mpmath.mp.dps = 50
    A = mpmath.matrix(K.tolist())
    n = 8 * len(T)
    LU, perm = mpmath.mp.LU_decomp(A)                         # one factorisation, three right-hand sides
    cols = [mpmath.mp.U_solve(LU, mpmath.mp.L_solve(LU, mpmath.matrix(rhs[:, j].tolist()), perm)) for j in range(3)]
    return np.array([[float(cols[j][i]) for j in range(3)] for i in range(n)])


def _missions():
    rng = np.random.default_rng(123)
    out = []
    g = np.load(os.path.join(ROOT, "tests", "golden", "planning.npz"))
    out.append(("lab_course course v=3", g["waypoints"][1:], 3.0))
    out.append(("config-2 mission", g["c2_waypoints"][5], float(g["c2_velocity"][5])))
    for k in range(4):                                   # wide volume, slow: segment durations up to ~30 s
        out.append((f"wide volume {k}", rng.uniform([0, 0, -6], [24, 14, 0], (5, 3)), float(rng.uniform(0.5, 1.0))))
    out.append(("two splines, very unequal durations", np.array([[0.0, 0, 0], [0.2, 0, 0], [20.0, 3, -2]]), 1.0))
    out.append(("12 splines", rng.uniform([0, 0, -6], [24, 14, 0], (13, 3)), 2.0))
    return out


def test_reduced_form_reaches_extended_precision_truth(k1_host):
    worst = 0.0
    for name, w, v in _missions():
        truth = _truth(w, v)
        scale = np.abs(truth).max()
        c, _ = k1_host(w, v)
        err = np.abs(c - truth).max() / scale
        lu = np.abs(minsnap_np.solve_coeffs(w, v, "solve")[0] - truth).max() / scale
        ls = np.abs(minsnap_np.solve_coeffs(w, v, "lstsq")[0] - truth).max() / scale
        print(f"{name:40s} K1 {err:.1e}   reference solve {lu:.1e}   reference lstsq {ls:.1e}")
        assert err < 1e-11, name                          # the 1e-9 contract with two digits to spare, on every mission
        worst = max(worst, err)
    assert worst < 1e-11
