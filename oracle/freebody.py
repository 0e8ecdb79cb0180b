"""Free rigid-body 1 kHz step in NED/FRD, fp64.  TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.

The reference has no integrator of its own: ``MujocoSimulation.step`` applies four rotor forces and
calls the third-party MuJoCo engine (``/root/reference/uav_ac/simulation/mujoco_sim.py:144-151``,
``:232-251``; model ``uav_ac/simulation/models/lab_course.xml:3,98-101,116-119``).  MuJoCo 3.11.0
(``uv.lock:128-129``) is not vendored and cannot be installed in the build container, so this file
restates its *documented* algorithm for this model -- one free-joint body, inertial frame equal to
the body frame, diagonal inertia, ``integrator="Euler"``, no damping -- expressed directly in the
NED/FRD frame that ``mujoco_to_ned_state`` (``mujoco_sim.py:20-45``) converts to:

    f_i   = kf * w_i^2                                    (mujoco_sim.py:235)
    F     = R_thrust @ [0, 0, -sum f] + wind              (mujoco_sim.py:238; NED: thrust is -z body)
    tau   = [ l (f0+f3-f1-f2), l (f0+f1-f2-f3), kappa (-f0+f1-f2+f3) ]
                                                          (sites/spins lab_course.xml:116-119, checked
                                                           against reference tests/unit/quadrotor/test_quad.py:85-112)
    vdot  = g e_z + F / m ;  wdot = (tau - w x (I w)) / I
    v+ = v + dt vdot ; w+ = w + dt wdot                   (semi-implicit Euler: velocities first)
    p+ = p + dt v+
    q+ = normalise(q) * [cos(a/2), sin(a/2) w+/|w+|],  a = dt |w+|    (mju_quatIntegrate, body-frame rate)
    q  = normalise(q+)                                    (mujoco_sim.py:36-42)

``R_thrust`` is the body rotation MuJoCo last computed.  In the headless loop of
``tests/integration/test_mujoco_trajectory_tracking.py:27-31`` that is the rotation of the PREVIOUS
tick's state (``data.xmat`` is filled by the forward pass inside ``mj_step`` before integration and
is not refreshed afterwards), which the caller models with ``thrust_frame_lag=1``.

Not modelled: MuJoCo's soft contacts (ground at take-off, obstacles).  ``wind`` is an extension of
the batched path (BASELINE.json configs[3]); it is zero for every reference-equivalent run.
"""
from __future__ import annotations

import math

import numpy as np

MJ_MINVAL = 1e-15  # mjMINVAL: below this |w| the rotation axis defaults to x and the angle to 0


def quat_to_rot(q: np.ndarray) -> np.ndarray:
    """R = I + 2 S^2 + 2 q0 S with S = skew(q1,q2,q3), q normalised first (quad.py:133-155)."""
    q = np.asarray(q, dtype=float)
    q = q / math.sqrt(float(q @ q))
    q0, q1, q2, q3 = q
    return np.array([
        [1 - 2 * (q2 * q2 + q3 * q3), 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2)],
        [2 * (q1 * q2 + q0 * q3), 1 - 2 * (q1 * q1 + q3 * q3), 2 * (q2 * q3 - q0 * q1)],
        [2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), 1 - 2 * (q1 * q1 + q2 * q2)],
    ])


def rotor_wrench(omega: np.ndarray, kf: float, arm: float, kappa: float):
    """Collective thrust and FRD body torque of the four rotors (D1)."""
    f = kf * np.asarray(omega, dtype=float) ** 2
    tau = np.array([
        arm * (f[0] + f[3] - f[1] - f[2]),
        arm * (f[0] + f[1] - f[2] - f[3]),
        kappa * (-f[0] + f[1] - f[2] + f[3]),
    ])
    return float(f.sum()), tau


def freebody_step(X: np.ndarray, omega: np.ndarray, R_thrust: np.ndarray, *, g: float, dt: float, mass: float,
                  inertia: np.ndarray, kf: float, arm: float, kappa: float, wind: np.ndarray | None = None) -> np.ndarray:
    """One tick X_k -> X_{k+1}; X = [p(3), q(4) scalar first, v(3) world, w(3) body] (quad.py:75-80)."""
    X = np.asarray(X, dtype=float)
    p, q, v, w = X[0:3], X[3:7], X[7:10], X[10:13]
    I = np.asarray(inertia, dtype=float)
    thrust, tau = rotor_wrench(omega, kf, arm, kappa)
    F = R_thrust @ np.array([0.0, 0.0, -thrust])
    if wind is not None:
        F = F + np.asarray(wind, dtype=float)
    vdot = np.array([0.0, 0.0, g]) + F / mass
    Iw = I * w
    wdot = (tau - np.array([w[1] * Iw[2] - w[2] * Iw[1], w[2] * Iw[0] - w[0] * Iw[2], w[0] * Iw[1] - w[1] * Iw[0]])) / I
    v1 = v + dt * vdot
    w1 = w + dt * wdot
    p1 = p + dt * v1
    wn = math.sqrt(float(w1 @ w1))
    if wn < MJ_MINVAL:
        axis, ang = np.array([1.0, 0.0, 0.0]), 0.0
    else:
        axis, ang = w1 / wn, dt * wn
    s, c = math.sin(0.5 * ang), math.cos(0.5 * ang)
    dq = np.array([c, s * axis[0], s * axis[1], s * axis[2]])
    qn = q / math.sqrt(float(q @ q))
    a0, a1, a2, a3 = qn
    b0, b1, b2, b3 = dq
    q1 = np.array([
        a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3,
        a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2,
        a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1,
        a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0,
    ])
    q1 = q1 / math.sqrt(float(q1 @ q1))
    out = np.empty(13)
    out[0:3], out[3:7], out[7:10], out[10:13] = p1, q1, v1, w1
    return out
