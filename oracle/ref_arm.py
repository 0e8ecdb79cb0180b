"""The reference's own classes as the CPU arm.  TEST INFRASTRUCTURE ONLY (bench.py `cpu_baseline` / `--impl reference`, tests).

Imports the byte-compiled reference from oracle/_ref (oracle/build_ref.py) with a stub ``mujoco`` module and drives

    TrajectoryController.step()  (uav_ac/main.py:37-61 -> CascadedController, Quad.set_propeller_speed)
    simulation.step()            -> oracle.freebody.freebody_step (MuJoCo is not installed; parity unpinned at that boundary)

exactly like the loop of the reference's integration test (tests/integration/test_mujoco_trajectory_tracking.py:27-31) and
like tests/golden/make_golden.py::fly, which produced the committed golden logs.  Everything except the rigid-body step is
executed by the reference's objects: MinimumSnap (both LAPACK branches), _generate_mission_trajectory, CascadedController,
Quad, TrajectoryController.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import time
import types
import unittest.mock

import numpy as np

from . import build_ref
from .freebody import freebody_step

FREQ = 10
GAIN_NAMES = ("kp_xy", "kd_xy", "kp_z", "kd_z", "ki_z", "kp_roll", "kp_pitch", "kp_yaw", "kp_p", "kp_q", "kp_r")
_NS = None


def available() -> bool:
    return build_ref.available()


class _RefFinder(importlib.abc.MetaPathFinder):
    """Imports `uav_ac[.x[.y]]` from the byte-code files oracle/build_ref.py wrote (<module>.bin, .pyc format)."""

    def find_spec(self, name, path=None, target=None):
        if name != "uav_ac" and not name.startswith("uav_ac."):
            return None
        base = os.path.join(build_ref.OUT, *name.split("."))
        for cand, pkg in ((os.path.join(base, "__init__.bin"), True), (base + ".bin", False)):
            if os.path.exists(cand):
                loader = importlib.machinery.SourcelessFileLoader(name, cand)
                return importlib.util.spec_from_file_location(name, cand, loader=loader, submodule_search_locations=[base] if pkg else None)
        return None


def load() -> types.SimpleNamespace:
    """The reference classes (cached).  Raises when oracle/_ref has not been built."""
    global _NS
    if _NS is None:
        if not available():
            raise RuntimeError("oracle/_ref is not built: run oracle/build_ref.py where /root/reference exists")
        sys.modules.setdefault("mujoco", unittest.mock.MagicMock())
        if "uav_ac" in sys.modules and "oracle/_ref" not in str(getattr(sys.modules["uav_ac"].__spec__, "origin", "")):
            raise RuntimeError("uav_ac is already imported from somewhere else in this process")
        if not any(isinstance(f, _RefFinder) for f in sys.meta_path):
            sys.meta_path.insert(0, _RefFinder())
        from uav_ac.control.controller import CascadedController
        from uav_ac.main import TrajectoryController, _generate_mission_trajectory
        from uav_ac.planning.minimum_snap import MinimumSnap
        from uav_ac.quadrotor.quad import Quad
        assert sys.modules["uav_ac.quadrotor.quad"].__spec__.origin.startswith(build_ref.OUT), "uav_ac was imported from elsewhere"
        _NS = types.SimpleNamespace(CascadedController=CascadedController, TrajectoryController=TrajectoryController, MinimumSnap=MinimumSnap,
                                    Quad=Quad, generate_mission_trajectory=_generate_mission_trajectory)
    return _NS


def make_quad(gain_scale=None, mass_scale: float = 1.0, inertia_scale=None):
    """Quad with the arguments of mujoco_sim._create_quad for lab_course.xml, optionally Monte-Carlo perturbed."""
    ns = load()
    quad = ns.Quad(g=9.81, dt=0.001, mass=0.5, inertia=np.array([0.0023, 0.0023, 0.0046]), arm_length=0.120208, force_coefficient=1.0,
                   drag_to_thrust=0.016, thrust_limits=np.array([0.1, 4.5]), motor_time_constants=np.array([0.0125, 0.025]),
                   flight_limits=np.array([3.0, 2.0, 3.0, 12.0, 0.7]))
    if gain_scale is not None:
        for g, s in zip(GAIN_NAMES, gain_scale):
            setattr(quad, g, getattr(quad, g) * float(s))
    quad.m *= mass_scale
    if inertia_scale is not None:
        quad.i_x, quad.i_y, quad.i_z = quad.i_x * inertia_scale[0], quad.i_y * inertia_scale[1], quad.i_z * inertia_scale[2]
    return quad


def mission_table(waypoints, obstacles, velocity: float, dt: float = 0.01):
    """uav_ac.main._generate_mission_trajectory (main.py:64-91): take-off + course tables, default lstsq branch."""
    return load().generate_mission_trajectory(np.asarray(waypoints, float), None if obstacles is None else np.asarray(obstacles, float), velocity, dt)


def fly(table, start, *, n_ticks=None, gain_scale=None, mass_scale=1.0, inertia_scale=None, obstacles=None, goal=None, lag=1):
    """n_ticks ticks (default: the whole table) of the reference closed loop; returns final state and mission metrics."""
    ns = load()
    quad = make_quad(gain_scale, mass_scale, inertia_scale)
    quad.X[0:3] = start
    ctrl = ns.CascadedController(quad.g, quad.dt * FREQ)
    tc = ns.TrajectoryController(ctrl, quad, table, FREQ)
    n_ticks = FREQ * len(table) if n_ticks is None else int(n_ticks)
    inertia = np.array([quad.i_x, quad.i_y, quad.i_z])
    R_stale = quad.R()
    errs, collided, first_hit = [], False, -1
    for k in range(n_ticks):
        row = min(k // FREQ, len(table) - 1)
        tc.step()
        R_now = quad.R()
        quad.X = freebody_step(quad.X, quad.omega, R_stale if lag else R_now, g=quad.g, dt=quad.dt, mass=quad.m, inertia=inertia, kf=quad.kf,
                               arm=quad.l, kappa=quad.kappa)
        R_stale = R_now
        if obstacles is not None and not collided:
            for box in obstacles:
                if ns.MinimumSnap.is_collision_cuboid(*quad.position, box):
                    collided, first_hit = True, k
                    break
        if (k + 1) % FREQ == 0:
            errs.append(np.linalg.norm(quad.position - table[row][:3]))
    errs = np.array(errs) if errs else np.zeros(1)
    return dict(X=quad.X.copy(), omega=quad.omega.copy(), collision=collided, first_collision_tick=first_hit, periods=len(errs),
                final_dist=float(np.linalg.norm(quad.position - goal)) if goal is not None else 0.0, mean_err=float(errs.mean()),
                rmse=float(np.sqrt(np.mean(errs ** 2))), max_err=float(errs.max()))


def solve_lstsq(waypoints, velocity: float, method: str = "lstsq"):
    """MinimumSnap._compute_spline_parameters (minimum_snap.py:138-153) of the reference: coefficients [8 S, 3] and times."""
    ms = load().MinimumSnap(np.asarray(waypoints, float), None, float(velocity), 0.01)
    ms._compute_spline_parameters(method)
    return np.asarray(ms.coeffs), np.asarray(ms.times)


def timed_ticks(seed: int, ticks: int, table, start, obstacles, goal):
    """Worker of the CPU baseline: fly `ticks` ticks of one Monte-Carlo-perturbed lab_course rollout; (ticks, seconds)."""
    rng = np.random.default_rng(seed)
    gs, ms, is_ = rng.uniform(0.8, 1.2, 11), rng.uniform(0.9, 1.1), rng.uniform(0.9, 1.1, 3)
    t0 = time.perf_counter()
    fly(table, start, n_ticks=ticks, gain_scale=gs, mass_scale=ms, inertia_scale=is_, obstacles=obstacles, goal=goal)
    return ticks, time.perf_counter() - t0
