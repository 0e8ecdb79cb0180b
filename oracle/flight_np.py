"""NumPy fp64 restatement of the reference closed loop.  TEST INFRASTRUCTURE ONLY.

Restates, one drone at a time in plain Python/NumPy (the reference's own style and cost profile):

* the vehicle model, gains, rotor allocation and motor lag of ``/root/reference/uav_ac/quadrotor/quad.py``
  (cited ``quad:LINE``),
* the cascaded controller of ``/root/reference/uav_ac/control/controller.py`` (``ctl:LINE``),
* the outer/inner scheduler of ``/root/reference/uav_ac/main.py:10-61`` (``main:LINE``),
* the headless loop and metrics of ``/root/reference/tests/integration/test_mujoco_trajectory_tracking.py:26-36``,
* the point-in-AABB collision flag of ``minimum_snap.py:327-357`` applied to the body origin
  (the batched path's substitute for MuJoCo contacts, BASELINE.json north_star),

around the free-body step of ``oracle/freebody.py`` (parity unpinned at that boundary, see there).

Pinned by tests/test_oracle_golden.py: every stage function is compared with outputs of the
reference's own ``CascadedController`` / ``Quad`` / ``TrajectoryController`` objects recorded in
tests/golden/stages.npz and tests/golden/closed_loop_*.npz.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .freebody import freebody_step, quat_to_rot

INTEGRAL_ERROR_LIMIT = 10.0  # ctl:10
TWO_PI = 2.0 * math.pi

# rows = rotors, columns = [p_bar, q_bar, r_bar]; the 4th (collective) column of quad:163-166 multiplies 0
MIXER = np.array([[1.0, 1.0, 1.0], [-1.0, 1.0, -1.0], [-1.0, -1.0, 1.0], [1.0, -1.0, -1.0]])


@dataclass
class Vehicle:
    """Constants of ``Quad.__init__`` (quad:36-73) with the lab_course.xml defaults."""
    g: float = 9.81
    dt: float = 0.001
    mass: float = 0.5
    inertia: np.ndarray = field(default_factory=lambda: np.array([0.0023, 0.0023, 0.0046]))
    arm: float = 0.120208
    kf: float = 1.0
    kappa: float = 0.016
    min_thrust: float = 0.1
    max_thrust: float = 4.5
    tau_rise: float = 0.0125
    tau_fall: float = 0.025
    max_ascent: float = 3.0
    max_descent: float = 2.0
    max_speed_xy: float = 3.0
    max_horiz_accel: float = 12.0
    max_tilt: float = 0.7
    # gains from the response parameters (quad:54-73, second_order_gains quad:124-127)
    kp_xy: float = 1 / 0.25 ** 2
    kd_xy: float = 2 * 0.875 / 0.25
    kp_z: float = 1 / 0.2 ** 2
    kd_z: float = 2 * 0.8 / 0.2
    ki_z: float = 0.1
    kp_roll: float = 1 / 0.07
    kp_pitch: float = 1 / 0.07
    kp_yaw: float = 1 / 0.25
    kp_p: float = 1 / 0.008
    kp_q: float = 1 / 0.008
    kp_r: float = 1 / 0.09

    GAIN_NAMES = ("kp_xy", "kd_xy", "kp_z", "kd_z", "ki_z", "kp_roll", "kp_pitch", "kp_yaw", "kp_p", "kp_q", "kp_r")

    def perturbed(self, gain_scale: np.ndarray, mass_scale: float, inertia_scale: np.ndarray) -> "Vehicle":
        """Monte-Carlo copy: 11 gains, mass and the three inertias multiplied (BASELINE configs[2])."""
        kw = {k: getattr(self, k) for k in self.__dataclass_fields__}
        for n, s in zip(self.GAIN_NAMES, gain_scale):
            kw[n] = kw[n] * float(s)
        kw["mass"] = self.mass * float(mass_scale)
        kw["inertia"] = np.asarray(self.inertia, dtype=float) * np.asarray(inertia_scale, dtype=float)
        return Vehicle(**kw)

    def with_values(self, gains: np.ndarray, mass: float, inertia: np.ndarray) -> "Vehicle":
        """Copy with the 11 gains, the mass and the three inertias SET (e.g. to the fp32-rounded values a kernel launch was given)."""
        kw = {k: getattr(self, k) for k in self.__dataclass_fields__}
        for n, g in zip(self.GAIN_NAMES, gains):
            kw[n] = float(g)
        kw["mass"] = float(mass)
        kw["inertia"] = np.asarray(inertia, dtype=float).copy()
        return Vehicle(**kw)


# ----------------------------------------------------------------------------- attitude helpers
def euler_from_quat(q):
    """phi, theta, psi exactly as ``Quad.phi/theta/psi`` (quad:189-213): raw quaternion, no normalisation."""
    q0, q1, q2, q3 = q
    phi = math.atan2(2 * (q0 * q1 + q2 * q3), 1 - 2 * (q1 ** 2 + q2 ** 2))
    theta = math.asin(min(1.0, max(-1.0, 2 * (q0 * q2 - q3 * q1))))
    psi = math.atan2(2 * (q0 * q3 + q1 * q2), 1 - 2 * (q2 ** 2 + q3 ** 2))
    return phi, theta, psi


def wrap_to_pi(a):
    """ctl:170-173 (Python ``%``: result takes the sign of the divisor)."""
    return (a + math.pi) % TWO_PI - math.pi


def wrap_to_2pi(a):
    """ctl:175-178."""
    return a % TWO_PI


# ----------------------------------------------------------------------------- controller stages
def altitude(veh: Vehicle, z, z_vel, des_z, R22, integral, dt_outer):
    """Collective thrust command and the updated integrator (ctl:26-56)."""
    climb = min(max(des_z[1], -veh.max_ascent), veh.max_descent)
    err = des_z[0] - z
    err_dot = climb - z_vel
    integral = integral + err * dt_outer
    integral = min(max(integral, -INTEGRAL_ERROR_LIMIT), INTEGRAL_ERROR_LIMIT)
    acc = veh.kp_z * err + veh.ki_z * integral + veh.kd_z * err_dot + des_z[2] - veh.g
    acc = acc / R22
    c = -veh.mass * acc
    c = min(max(c, 4 * veh.min_thrust), 4 * veh.max_thrust)
    return c, integral


def lateral(veh: Vehicle, pos_xy, vel_xy, des_x, des_y, thrust_cmd):
    """Commanded R02, R12 (ctl:58-97)."""
    p_des = np.array([des_x[0], des_y[0]])
    v_des = np.array([des_x[1], des_y[1]])
    a_ff = np.array([des_x[2], des_y[2]])
    vmag = math.sqrt(float(v_des @ v_des))
    if vmag > veh.max_speed_xy:
        v_des = (v_des / vmag) * veh.max_speed_xy
    acc = veh.kp_xy * (p_des - pos_xy) + veh.kd_xy * (v_des - vel_xy) + a_ff
    amag = math.sqrt(float(acc @ acc))
    if amag > veh.max_horiz_accel:
        acc = (acc / amag) * veh.max_horiz_accel
    acc_z = -thrust_cmd / veh.mass
    return np.clip(acc / acc_z, -veh.max_tilt, veh.max_tilt)


def roll_pitch(veh: Vehicle, bxy_cmd, R):
    """p_c, q_c from the tilt error (ctl:132-154)."""
    bdot = np.array([veh.kp_roll, veh.kp_pitch]) * (bxy_cmd - np.array([R[0, 2], R[1, 2]]))
    M = np.array([[R[1, 0], -R[0, 0]], [R[1, 1], -R[0, 1]]]) / R[2, 2]
    return M @ bdot


def yaw_rate(veh: Vehicle, q, psi_des, q_cmd):
    """Body yaw rate from the Euler yaw error (ctl:156-168)."""
    phi, theta, psi = euler_from_quat(q)
    err = wrap_to_pi(wrap_to_2pi(psi_des) - psi)
    return (veh.kp_yaw * err * math.cos(theta) - q_cmd * math.sin(phi)) / math.cos(phi)


def body_rate(veh: Vehicle, pqr, pqr_cmd):
    """Moments: I kp (cmd - w) + w x (I w) (ctl:115-130)."""
    I = np.asarray(veh.inertia, dtype=float)
    kp = np.array([veh.kp_p, veh.kp_q, veh.kp_r])
    Iw = I * pqr
    gyro = np.array([pqr[1] * Iw[2] - pqr[2] * Iw[1], pqr[2] * Iw[0] - pqr[0] * Iw[2], pqr[0] * Iw[1] - pqr[1] * Iw[0]])
    return I * kp * (pqr_cmd - pqr) + gyro


def allocate(veh: Vehicle, thrust_cmd, moment):
    """Rotor forces that keep the collective and scale the moments into the limits (quad:105-122)."""
    c_bar = min(max(thrust_cmd, 4 * veh.min_thrust), 4 * veh.max_thrust)
    bars = np.array([moment[0] / veh.arm, moment[1] / veh.arm, -moment[2] / veh.kappa])
    mf = MIXER @ bars / 4
    coll = c_bar / 4
    lim = np.ones(4)
    for i in range(4):
        if mf[i] > 0:
            lim[i] = (veh.max_thrust - coll) / mf[i]
        elif mf[i] < 0:
            lim[i] = (veh.min_thrust - coll) / mf[i]
    s = min(max(float(lim.min()), 0.0), 1.0)
    return np.clip(coll + s * mf, veh.min_thrust, veh.max_thrust)


def motor_lag(veh: Vehicle, omega, forces):
    """omega_cmd = sqrt(f/kf); first-order lag with rise/fall time constants (quad:88-103)."""
    cmd = np.sqrt(forces / veh.kf)
    tau = np.where(cmd > omega, veh.tau_rise, veh.tau_fall)
    return omega + (1 - np.exp(-veh.dt / tau)) * (cmd - omega), cmd


# ----------------------------------------------------------------------------- closed loop
def closed_loop(veh: Vehicle, table: np.ndarray, start, *, freq: int = 10, n_ticks: int | None = None,
                obstacles=None, goal=None, wind=None, thrust_frame_lag: int = 1, log_stride: int = 0, ground_z=None,
                traj_gate_z=None, traj_interval: float = 0.05):
    """Headless mission (integration test :26-31 + main:37-61) on the free-body model.

    Tick k: (1) if k % freq == 0 outer loop on X_k with table row ``idx`` then idx=min(idx+1,N-1);
    (2) body-rate loop; (3) allocation + motor lag; (4) rotor wrench with the stale thrust frame;
    (5-6) free-body step; (7) sticky AABB flag on the body origin; after each outer period the
    tracking error |p - row[:3]| of the row used in that period is recorded.

    Returns a dict: X (13,), omega (4,), integral, errors (per period), collision (bool),
    first_collision_tick, final_dist, mean_err, rmse, max_err, and ``log`` (n_log, 13) every
    ``log_stride`` ticks (state after the tick) when log_stride > 0.
    ``ground_z`` (optional): unilateral floor -- after the free-body step a position below it (NED: z > ground_z) is put
    back onto it and a downward velocity set to zero.  ``traj_gate_z``: also return ``traj_ticks`` / ``traj`` -- the flown-path list
    of MujocoSimulation._record_actual_trajectory (ms:201-218): data.time += dt per tick, a sample when z <= gate and
    time >= next, next = time + traj_interval.
    """
    N = len(table)
    if n_ticks is None:
        n_ticks = freq * N
    dt_outer = veh.dt * freq
    X = np.zeros(13)
    X[0:3] = start
    X[3] = 1.0
    omega = np.zeros(4)
    integral = 0.0
    thrust_cmd = 0.0
    pqr_cmd = np.zeros(3)
    idx = 0
    R_stale = quat_to_rot(X[3:7])
    errors = []
    collided = False
    first_hit = -1
    log = []
    row = table[0]
    sim_time, traj_next, traj_ticks, traj = 0.0, 0.0, [], []
    for k in range(n_ticks):
        if k % freq == 0:
            row = table[idx]
            R = quat_to_rot(X[3:7])
            thrust_cmd, integral = altitude(veh, X[2], X[9], row[[2, 5, 8]], R[2, 2], integral, dt_outer)
            bxy = lateral(veh, X[0:2], X[7:9], row[[0, 3, 6]], row[[1, 4, 7]], thrust_cmd)
            pq = roll_pitch(veh, bxy, R)
            r_c = yaw_rate(veh, X[3:7], row[9], pq[1])
            pqr_cmd = np.array([pq[0], pq[1], r_c])
            idx = min(idx + 1, N - 1)
        moment = body_rate(veh, X[10:13], pqr_cmd)
        forces = allocate(veh, thrust_cmd, moment)
        omega, _ = motor_lag(veh, omega, forces)
        R_now = quat_to_rot(X[3:7])
        R_use = R_stale if thrust_frame_lag else R_now
        X = freebody_step(X, omega, R_use, g=veh.g, dt=veh.dt, mass=veh.mass, inertia=veh.inertia, kf=veh.kf,
                          arm=veh.arm, kappa=veh.kappa, wind=wind)
        R_stale = R_now
        if ground_z is not None and X[2] > ground_z:
            X[2] = ground_z
            X[9] = min(X[9], 0.0)
        if traj_gate_z is not None:
            sim_time += veh.dt
            if not (X[2] > traj_gate_z) and not (sim_time < traj_next):
                traj_ticks.append(k)
                traj.append(X[0:3].copy())
                traj_next = sim_time + traj_interval
        if obstacles is not None and not collided:
            for box in obstacles:
                if box[0] <= X[0] <= box[1] and box[2] <= X[1] <= box[3] and box[4] <= X[2] <= box[5]:
                    collided, first_hit = True, k
                    break
        if log_stride and (k + 1) % log_stride == 0:
            log.append(X.copy())
        if (k + 1) % freq == 0:
            errors.append(math.sqrt(float(np.sum((X[0:3] - row[0:3]) ** 2))))
    errors = np.asarray(errors)
    out = dict(X=X, omega=omega, integral=integral, errors=errors, collision=collided, first_collision_tick=first_hit,
               mean_err=float(errors.mean()) if len(errors) else 0.0,
               rmse=float(math.sqrt(np.mean(errors ** 2))) if len(errors) else 0.0,
               max_err=float(errors.max()) if len(errors) else 0.0,
               log=np.asarray(log) if log_stride else None)
    if traj_gate_z is not None:
        out["traj_ticks"], out["traj"] = np.asarray(traj_ticks, dtype=int), np.asarray(traj).reshape(-1, 3)
    if goal is not None:
        out["final_dist"] = float(math.sqrt(np.sum((X[0:3] - np.asarray(goal, dtype=float)) ** 2)))
    return out
