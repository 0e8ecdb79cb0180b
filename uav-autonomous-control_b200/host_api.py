"""Host-buffer entry point: NumPy (or pinned torch CPU) arrays in, metrics out, through
``uavb_fly_mission_host`` (include/uavb.h) -- the call a reference-side ctypes binding would make
(INTEGRATION.md).  Copies, planning (K1), the rollout (K2) and the synchronisation all happen inside
the C call; nothing is computed on the CPU.

Replaces for B drones what ``uav_ac/main.py:73-120`` and the loop of
``tests/integration/test_mujoco_trajectory_tracking.py:26-36`` do for one.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _native as nat


def _host_ptr(x, dtype, shape, name):
    """(pointer, keep-alive) of a C-contiguous host array; accepts numpy arrays and CPU torch tensors."""
    if x is None:
        return ctypes.c_void_p(None), None
    if hasattr(x, "data_ptr"):                      # torch tensor (pinned memory gives the fastest copies)
        if x.is_cuda:
            raise nat.UavbError(f"{name} must live in host memory for the *_host entry points")
        import torch
        want = {np.float32: torch.float32, np.float64: torch.float64}[dtype]
        if x.dtype != want or not x.is_contiguous() or tuple(x.shape) != tuple(shape):
            raise ValueError(f"{name} must be a contiguous {want} tensor of shape {tuple(shape)}")
        return ctypes.c_void_p(x.data_ptr()), x
    a = np.ascontiguousarray(x, dtype=dtype)
    if a.shape != tuple(shape):
        raise ValueError(f"{name} must have shape {tuple(shape)}, got {a.shape}")
    return ctypes.c_void_p(a.ctypes.data), a


def minsnap_solve_host(waypoints, velocity, *, start_end_time_factor: float = 1.5, coeffs_out=None, times_out=None, status_out=None):
    """B minimum-snap solves, host arrays in and out, through ``uavb_minsnap_solve_f64_host``: waypoints (B, S+1, 3) f64 and
    velocity (B,) f64 in; coeffs (B, 8 S, 3), times (B, S) f64 and status (B,) i32 out (preallocated -- e.g. pinned -- buffers are
    filled in place).  Replaces B calls of ``MinimumSnap._compute_spline_parameters`` (uav_ac/planning/minimum_snap.py:138-153)."""
    shape = tuple(waypoints.shape)
    if len(shape) != 3 or shape[2] != 3 or shape[1] < 2:
        raise ValueError("waypoints must have shape (B, S+1 >= 2, 3)")
    B, S = shape[0], shape[1] - 1
    wp_p, wp_k = _host_ptr(waypoints, np.float64, shape, "waypoints")
    v_p, v_k = _host_ptr(velocity, np.float64, (B,), "velocity")
    coeffs = coeffs_out if coeffs_out is not None else np.empty((B, 8 * S, 3), dtype=np.float64)
    times = times_out if times_out is not None else np.empty((B, S), dtype=np.float64)
    status = status_out if status_out is not None else np.empty((B,), dtype=np.int32)
    c_p, c_k = _host_ptr(coeffs, np.float64, (B, 8 * S, 3), "coeffs_out")
    t_p, t_k = _host_ptr(times, np.float64, (B, S), "times_out")
    if hasattr(status, "data_ptr"):
        s_p = ctypes.c_void_p(status.data_ptr())
    else:
        if status.dtype != np.int32 or not status.flags.c_contiguous or status.shape != (B,):
            raise ValueError("status_out must be a C-contiguous int32 array of shape (B,)")
        s_p = ctypes.c_void_p(status.ctypes.data)
    for given, kept, name in ((coeffs_out, c_k, "coeffs_out"), (times_out, t_k, "times_out")):
        if given is not None and kept is not given:
            raise ValueError(f"{name} must be C-contiguous float64 of the right shape")
    nat.check(nat.lib().uavb_minsnap_solve_f64_host(wp_p, v_p, int(B), int(S), float(start_end_time_factor), c_p, t_p, s_p), "uavb_minsnap_solve_f64_host")
    return (c_k if coeffs_out is None else coeffs_out), (t_k if times_out is None else times_out), status


def fly_mission_host(waypoints, velocity: float, B: int, *, n_takeoff_waypoints: int = 2, frequency: int = 10, n_ticks: int = 0,
                     vehicle: Optional[nat.Vehicle] = None, mc_mass=None, mc_inertia=None, mc_gains=None, mc_wind=None,
                     obstacles=None, start=None, goal=None, thrust_frame_lag: int = 1, start_end_time_factor: float = 1.5,
                     want_state: bool = False, metrics_out=None, state_out=None, correct: bool = True):
    """Plan one mission and fly it with B drones; host arrays in, host arrays out.

    waypoints (n, 3) f64; mc_mass (B,), mc_inertia (3, B), mc_gains (11, B), mc_wind (3, B) f32 SoA;
    obstacles (n_obs, 6): the collision flag tests their fp32 copy, the planner's obstacle-correction loop
    (minimum_snap.py:63-95; ``correct=False`` skips it) the fp64 values.  Returns (metrics (B, 8) f32, state (13, B) f32 | None, n_ticks).
    ``metrics_out`` / ``state_out`` may be preallocated (e.g. pinned) buffers.
    """
    wp = np.ascontiguousarray(waypoints, dtype=np.float64)
    if wp.ndim != 2 or wp.shape[1] != 3 or wp.shape[0] < 2:
        raise ValueError("waypoints must have shape (n >= 2, 3)")
    m = nat.MissionHost()
    keep = [wp]
    m.B, m.n_waypoints, m.n_takeoff_waypoints = int(B), wp.shape[0], int(n_takeoff_waypoints)
    m.waypoints = ctypes.c_void_p(wp.ctypes.data)
    m.velocity, m.start_end_time_factor = float(velocity), float(start_end_time_factor)
    m.frequency, m.n_ticks, m.thrust_frame_lag = int(frequency), int(n_ticks), int(thrust_frame_lag)
    m.veh = vehicle if vehicle is not None else nat.default_vehicle()
    for field, arr, shape in (("mc_mass", mc_mass, (B,)), ("mc_inertia", mc_inertia, (3, B)), ("mc_gains", mc_gains, (nat.N_GAINS, B)),
                              ("mc_wind", mc_wind, (3, B))):
        p, k = _host_ptr(arr, np.float32, shape, field)
        setattr(m, field, p)
        keep.append(k)
    if obstacles is not None and np.size(obstacles) > 0:
        obs = np.ascontiguousarray(obstacles, dtype=np.float32).reshape(-1, 6)
        keep.append(obs)
        m.aabbs, m.n_obs = ctypes.c_void_p(obs.ctypes.data), obs.shape[0]
        obs64 = np.ascontiguousarray(obstacles, dtype=np.float64).reshape(-1, 6)     # the planner tests the fp64 boxes, like the reference
        keep.append(obs64)
        m.plan_aabbs = ctypes.c_void_p(obs64.ctypes.data)
    m.no_correction = 0 if correct else 1
    for field, arr in (("start", start), ("goal", goal)):
        p, k = _host_ptr(arr, np.float64, (3,), field)
        setattr(m, field, p)
        keep.append(k)
    metrics = metrics_out if metrics_out is not None else np.empty((B, nat.N_METRICS), dtype=np.float32)
    mp, mk = _host_ptr(metrics, np.float32, (B, nat.N_METRICS), "metrics_out")
    if mk is not metrics and metrics_out is not None:
        raise ValueError("metrics_out must be a C-contiguous float32 array of shape (B, 8)")
    state = state_out if state_out is not None else (np.empty((nat.STATE_DIM, B), dtype=np.float32) if want_state else None)
    sp, sk = _host_ptr(state, np.float32, (nat.STATE_DIM, B), "state_out")
    ticks = ctypes.c_int(0)
    nat.check(nat.lib().uavb_fly_mission_host(ctypes.byref(m), mp, sp, ctypes.byref(ticks)), "uavb_fly_mission_host")
    return (mk if metrics_out is None else metrics_out), (sk if state_out is None else state_out), ticks.value
