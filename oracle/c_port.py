"""ctypes front-end of oracle/oracle_c.c (the plain-C fp64 twin of the NumPy oracle).  TEST INFRASTRUCTURE ONLY.

Same algorithm as ``minsnap_np`` / ``flight_np`` / ``freebody`` (which cite the reference line by line), fast enough
to re-fly thousands of whole missions as a checker for the CUDA path and to give bench.py an optimised multi-core CPU
figure next to the NumPy-style one.  Pinned by tests/test_oracle_c.py.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build
from .flight_np import Vehicle

_lib = None


class CVehicle(C.Structure):
    _fields_ = [("g", C.c_double), ("dt", C.c_double), ("mass", C.c_double), ("inertia", C.c_double * 3), ("arm", C.c_double),
                ("kf", C.c_double), ("kappa", C.c_double), ("min_thrust", C.c_double), ("max_thrust", C.c_double),
                ("tau_rise", C.c_double), ("tau_fall", C.c_double), ("max_ascent", C.c_double), ("max_descent", C.c_double),
                ("max_speed_xy", C.c_double), ("max_horiz_accel", C.c_double), ("max_tilt", C.c_double), ("gains", C.c_double * 11),
                ("integral_limit", C.c_double), ("ground_on", C.c_double), ("ground_z", C.c_double)]


def lib():
    global _lib
    if _lib is None:
        path = _build.LIB if os.path.exists(_build.LIB) else _build.build()
        L = C.CDLL(path)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.oracle_minsnap_solve.argtypes = [dp, C.c_int, C.c_double, C.c_double, dp, dp]
        L.oracle_minsnap_solve_batch.argtypes = [dp, dp, C.c_int, C.c_int, C.c_double, dp, dp, C.c_int]
        L.oracle_sample_table.argtypes = [dp, dp, C.c_int, C.c_double, dp, C.c_int]
        L.oracle_sample_count.argtypes = [C.c_double, C.c_double]
        L.oracle_closed_loop.argtypes = [C.POINTER(CVehicle), dp, C.c_int, dp, C.c_int, C.c_int, dp, C.c_int, dp, dp, C.c_int, dp, dp, dp, C.c_int, dp]
        L.oracle_closed_loop.restype = None
        L.oracle_closed_loop_batch.argtypes = [C.POINTER(CVehicle), C.c_int, C.c_int, dp, C.c_int, dp, C.c_int, C.c_int, dp, C.c_int, dp, dp,
                                               C.c_int, dp, dp, C.c_int]
        L.oracle_closed_loop_batch.restype = None
        _lib = L
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def c_vehicle(v: Vehicle, ground_z=None) -> CVehicle:
    c = CVehicle(g=v.g, dt=v.dt, mass=v.mass, arm=v.arm, kf=v.kf, kappa=v.kappa, min_thrust=v.min_thrust, max_thrust=v.max_thrust,
                 tau_rise=v.tau_rise, tau_fall=v.tau_fall, max_ascent=v.max_ascent, max_descent=v.max_descent, max_speed_xy=v.max_speed_xy,
                 max_horiz_accel=v.max_horiz_accel, max_tilt=v.max_tilt, integral_limit=10.0)
    if ground_z is not None:
        c.ground_on, c.ground_z = 1.0, float(ground_z)
    c.inertia[:] = list(np.asarray(v.inertia, dtype=float))
    c.gains[:] = [getattr(v, n) for n in Vehicle.GAIN_NAMES]
    return c


def solve_coeffs(waypoints, velocity, factor=1.5):
    w = np.ascontiguousarray(waypoints, dtype=float)
    S = len(w) - 1
    c, t = np.empty((8 * S, 3)), np.empty(S)
    if lib().oracle_minsnap_solve(_dp(w), S, float(velocity), float(factor), _dp(c), _dp(t)) != 0:
        raise np.linalg.LinAlgError("singular minimum-snap KKT system")
    return c, t


def solve_batch(waypoints, velocity, threads=1, factor=1.5):
    w = np.ascontiguousarray(waypoints, dtype=float)
    B, S = w.shape[0], w.shape[1] - 1
    v = np.ascontiguousarray(velocity, dtype=float)
    c, t = np.empty((B, 8 * S, 3)), np.empty((B, S))
    bad = lib().oracle_minsnap_solve_batch(_dp(w), _dp(v), B, S, float(factor), _dp(c), _dp(t), int(threads))
    return c, t, bad


def sample_table(coeffs, T, dt):
    c, T = np.ascontiguousarray(coeffs, dtype=float), np.ascontiguousarray(T, dtype=float)
    n = sum(lib().oracle_sample_count(float(t), float(dt)) for t in T)
    tab = np.empty((n, 11))
    got = lib().oracle_sample_table(_dp(c), _dp(T), len(T), float(dt), _dp(tab), n)
    assert got == n
    return tab


def mission_table(waypoints, velocity, dt):
    """Take-off table + course table without obstacles in the loop (main.py:73-84 when no midpoint is inserted)."""
    w = np.asarray(waypoints, dtype=float)
    return np.vstack([sample_table(*solve_coeffs(p, velocity), dt) for p in (w[:2], w[1:])])


def closed_loop(veh: Vehicle, table, start, *, freq=10, n_ticks=None, obstacles=None, goal=None, wind=None, thrust_frame_lag=1, log_stride=0,
                ground_z=None):
    table = np.ascontiguousarray(table, dtype=float)
    n_ticks = freq * len(table) if n_ticks is None else int(n_ticks)
    obs = None if obstacles is None else np.ascontiguousarray(obstacles, dtype=float).reshape(-1, 6)
    start = np.ascontiguousarray(start, dtype=float)
    goal_a = None if goal is None else np.ascontiguousarray(goal, dtype=float)
    wind_a = None if wind is None else np.ascontiguousarray(wind, dtype=float)
    m, X, om = np.empty(8), np.empty(13), np.empty(4)
    log = np.empty((n_ticks // log_stride, 13)) if log_stride else None
    cv = c_vehicle(veh, ground_z)
    lib().oracle_closed_loop(C.byref(cv), _dp(table), len(table), _dp(start), freq, n_ticks, _dp(obs), 0 if obs is None else len(obs), _dp(goal_a),
                             _dp(wind_a), int(thrust_frame_lag), _dp(m), _dp(X), _dp(om), int(log_stride), _dp(log))
    return dict(X=X, omega=om, final_dist=m[0], collision=bool(m[1]), rmse=m[2], mean_err=m[3], max_err=m[4], first_collision_tick=int(m[6]),
                periods=int(m[7]), log=log)


def closed_loop_batch(vehicles, table, start, *, freq=10, n_ticks=None, obstacles=None, goal=None, wind=None, thrust_frame_lag=1, threads=1,
                      ground_z=None):
    """B drones (list of Vehicle, or one Vehicle with ``wind`` (B, 3)) on one table; returns metrics (B, 8), X (B, 13)."""
    table = np.ascontiguousarray(table, dtype=float)
    n_ticks = freq * len(table) if n_ticks is None else int(n_ticks)
    single = isinstance(vehicles, Vehicle)
    B = len(wind) if single else len(vehicles)
    arr = (CVehicle * (1 if single else B))(*([c_vehicle(vehicles, ground_z)] if single else [c_vehicle(v, ground_z) for v in vehicles]))
    obs = None if obstacles is None else np.ascontiguousarray(obstacles, dtype=float).reshape(-1, 6)
    start = np.ascontiguousarray(start, dtype=float)
    goal_a = None if goal is None else np.ascontiguousarray(goal, dtype=float)
    wind_a = None if wind is None else np.ascontiguousarray(wind, dtype=float)
    m, X = np.empty((B, 8)), np.empty((B, 13))
    lib().oracle_closed_loop_batch(arr, 0 if single else 1, B, _dp(table), len(table), _dp(start), freq, n_ticks, _dp(obs),
                                   0 if obs is None else len(obs), _dp(goal_a), _dp(wind_a), int(thrust_frame_lag), _dp(m), _dp(X), int(threads))
    return m, X
