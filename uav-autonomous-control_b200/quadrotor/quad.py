"""Batched ``Quad``: same constructor, attributes and methods as ``uav_ac/quadrotor/quad.py:4-250``
with a leading batch dimension, computed by the sm_100a stage kernels (``uavb_stage_f32``).

State ``X`` is a float32 CUDA tensor of shape (B, 13) laid out like the reference row vector
``[x y z q0 q1 q2 q3 x_dot y_dot z_dot p q r]`` (NED / FRD, scalar-first quaternion).  Mass, inertia
and every gain may be a Python float (all drones alike) or a (B,) tensor (Monte-Carlo vehicles).
Like the reference class it holds no integrator: the rigid-body step belongs to the simulation.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native as nat, _stages as st


class Quad:
    GAIN_NAMES = nat.GAIN_NAMES

    def __init__(self, g: float, dt: float, mass, inertia, arm_length: float, force_coefficient: float, drag_to_thrust: float,
                 thrust_limits, motor_time_constants, flight_limits, batch: int = 1, device=None):
        """Arguments as ``Quad.__init__`` (quad.py:11-51); ``batch`` and ``device`` are the batched additions.
        ``mass`` may be a (B,) tensor and ``inertia`` a (B, 3) tensor."""
        if not torch.cuda.is_available():
            raise nat.UavbError("no CUDA device visible: the batched Quad has no CPU implementation")
        nat.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.batch = int(batch)
        self.g = g
        self.dt = dt
        self.l = float(arm_length)
        self.m = mass if isinstance(mass, torch.Tensor) else float(mass)
        self.kf = float(force_coefficient)
        self.kappa = float(drag_to_thrust)
        if isinstance(inertia, torch.Tensor) and inertia.dim() == 2:
            self.i_x, self.i_y, self.i_z = inertia[:, 0].contiguous(), inertia[:, 1].contiguous(), inertia[:, 2].contiguous()
        else:
            self.i_x, self.i_y, self.i_z = (float(v) for v in np.asarray(inertia, dtype=float))
        self.min_thrust, self.max_thrust = (float(v) for v in np.asarray(thrust_limits, dtype=float))
        (self.max_ascent_rate, self.max_descent_rate, self.max_speed_xy, self.max_horiz_accel,
         self.max_tilt_angle) = (float(v) for v in np.asarray(flight_limits, dtype=float))

        # controller response parameters (quad.py:54-73)
        self.tau_xy, self.zeta_xy = 0.25, 0.875
        self.tau_altitude, self.zeta_altitude = 0.2, 0.8
        self.tau_roll = self.tau_pitch = 0.07
        self.tau_yaw = 0.25
        self.tau_p = self.tau_q = 0.008
        self.tau_r = 0.09
        self.kp_xy, self.kd_xy = Quad.second_order_gains(self.tau_xy, self.zeta_xy)
        self.kp_z, self.kd_z = Quad.second_order_gains(self.tau_altitude, self.zeta_altitude)
        self.ki_z = 0.1
        self.kp_roll = 1 / self.tau_roll
        self.kp_pitch = 1 / self.tau_pitch
        self.kp_yaw = 1 / self.tau_yaw
        self.kp_p = 1 / self.tau_p
        self.kp_q = 1 / self.tau_q
        self.kp_r = 1 / self.tau_r

        self.X = torch.zeros((self.batch, 13), dtype=torch.float32, device=self.device)
        self.X[:, 3] = 1.0
        self.motor_rise_time_constant, self.motor_fall_time_constant = (float(v) for v in np.asarray(motor_time_constants, dtype=float))
        self.omega = torch.zeros((self.batch, 4), dtype=torch.float32, device=self.device)
        self.omega_command = torch.zeros((self.batch, 4), dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------ C-ABI vehicle description
    def vehicle_struct(self, B: int, gain_overrides: dict):
        """(struct uavb_vehicle, per-drone override tensors) for a stage / rollout call."""
        if B != self.batch:
            raise ValueError(f"batch size {B} does not match this Quad ({self.batch})")
        v = nat.Vehicle()
        v.g, v.dt = float(self.g), float(self.dt)
        v.arm, v.kf, v.kappa = self.l, self.kf, self.kappa
        v.min_thrust, v.max_thrust = self.min_thrust, self.max_thrust
        v.tau_rise, v.tau_fall = self.motor_rise_time_constant, self.motor_fall_time_constant
        v.max_ascent, v.max_descent, v.max_speed_xy = self.max_ascent_rate, self.max_descent_rate, self.max_speed_xy
        v.max_horiz_accel, v.max_tilt, v.integral_limit = self.max_horiz_accel, self.max_tilt_angle, 10.0
        mc = {}
        if isinstance(self.m, torch.Tensor):
            mc["mc_mass"] = st.as_batch(self.m, B, None, self.device)
            v.mass = float(self.m.double().mean())
        else:
            v.mass = self.m
        inertia = (self.i_x, self.i_y, self.i_z)
        if any(isinstance(i, torch.Tensor) for i in inertia):
            mc["mc_inertia"] = torch.stack([st.as_batch(i, B, None, self.device) for i in inertia]).contiguous()
            v.inertia[:] = [float(torch.as_tensor(i).double().mean()) for i in inertia]
        else:
            v.inertia[:] = list(inertia)
        gains = [gain_overrides.get(n, getattr(self, n)) for n in self.GAIN_NAMES]
        if any(isinstance(gv, torch.Tensor) for gv in gains):
            mc["mc_gains"] = torch.stack([st.as_batch(gv, B, None, self.device) for gv in gains]).contiguous()
            v.gains[:] = [float(torch.as_tensor(gv).double().mean()) for gv in gains]
        else:
            v.gains[:] = [float(gv) for gv in gains]
        return v, mc

    # ------------------------------------------------------------------ actuation (quad.py:88-122)
    def set_propeller_speed(self, thrust_cmd, moment_cmd):
        """Collective thrust (B,) and body moments (B, 3) -> rotor speeds through the allocation and the
        asymmetric first-order motor lag (quad.py:88-103).  Updates ``omega`` and ``omega_command``."""
        B = self.batch
        thrust = st.as_batch(thrust_cmd, B, None, self.device)
        moment = st.soa(st.as_batch(moment_cmd, B, 3, self.device))
        omega, cmd = st.soa(self.omega), torch.empty((4, B), dtype=torch.float32, device=self.device)
        st.run(nat.STAGE_PROPELLER, self, self.dt, B, thrust=thrust, moment=moment, omega=omega, omega_cmd=cmd)
        self.omega, self.omega_command = omega.t().contiguous(), cmd.t().contiguous()

    def _allocate_rotor_forces(self, thrust_cmd, moment_cmd) -> torch.Tensor:
        """Rotor forces (B, 4) that keep the feasible collective and scale the moments into the limits (quad.py:105-122)."""
        B = self.batch
        forces = torch.empty((4, B), dtype=torch.float32, device=self.device)
        st.run(nat.STAGE_ALLOCATE, self, self.dt, B, thrust=st.as_batch(thrust_cmd, B, None, self.device),
               moment=st.soa(st.as_batch(moment_cmd, B, 3, self.device)), forces=forces)
        return forces.t().contiguous()

    @staticmethod
    def second_order_gains(time_constant: float, damping_ratio: float) -> tuple[float, float]:
        """quad.py:124-127."""
        return 1 / time_constant ** 2, 2 * damping_ratio / time_constant

    # ------------------------------------------------------------------ attitude (quad.py:129-155, 189-213)
    def _attitude(self, want_rot: bool, want_euler: bool):
        B = self.batch
        rot = torch.empty((9, B), dtype=torch.float32, device=self.device) if want_rot else None
        eul = torch.empty((3, B), dtype=torch.float32, device=self.device) if want_euler else None
        st.run(nat.STAGE_ATTITUDE, self, self.dt, B, X=st.soa(self.X), rot_out=rot, euler_out=eul)
        return (rot.t().reshape(B, 3, 3).contiguous() if want_rot else None), (eul.t().contiguous() if want_euler else None)

    def R(self) -> torch.Tensor:
        """Rotation matrices (B, 3, 3) from the state quaternions (normalised first, quad.py:133-155)."""
        return self._attitude(True, False)[0]

    @staticmethod
    def quat_to_rot(quaternion: torch.Tensor) -> torch.Tensor:
        """(B, 4) or (4,) scalar-first quaternions -> rotation matrices, via the same stage kernel."""
        q = torch.as_tensor(quaternion, dtype=torch.float32)
        single = q.dim() == 1
        q = q.reshape(-1, 4)
        if not q.is_cuda:
            q = q.cuda()
        tmp = Quad(9.81, 0.001, 1.0, [1.0, 1.0, 1.0], 1.0, 1.0, 1.0, [0.0, 1.0], [1.0, 1.0], [1.0] * 5, batch=q.shape[0], device=q.device)
        tmp.X[:, 3:7] = q
        R = tmp.R()
        return R[0] if single else R

    @staticmethod
    def propeller_coeffs() -> np.ndarray:
        """Mixer sign matrix (quad.py:157-166): rows = rotors, columns = [p_bar, q_bar, r_bar, c_bar]."""
        return np.array([[1.0, 1.0, 1.0, 1.0], [-1.0, 1.0, -1.0, 1.0], [-1.0, -1.0, 1.0, 1.0], [1.0, -1.0, -1.0, 1.0]])

    # ------------------------------------------------------------------ state accessors (quad.py:168-250)
    @property
    def x(self): return self.X[:, 0]
    @property
    def y(self): return self.X[:, 1]
    @property
    def z(self): return self.X[:, 2]
    @property
    def position(self): return self.X[:, 0:3]
    @property
    def quaternion(self): return self.X[:, 3:7]
    @property
    def euler_angles(self): return self._attitude(False, True)[1]
    @property
    def phi(self): return self.euler_angles[:, 0]
    @property
    def theta(self): return self.euler_angles[:, 1]
    @property
    def psi(self): return self.euler_angles[:, 2]
    @property
    def x_vel(self): return self.X[:, 7]
    @property
    def y_vel(self): return self.X[:, 8]
    @property
    def z_vel(self): return self.X[:, 9]
    @property
    def velocity(self): return self.X[:, 7:10]
    @property
    def p(self): return self.X[:, 10]
    @property
    def q(self): return self.X[:, 11]
    @property
    def r(self): return self.X[:, 12]
    @property
    def body_angular_velocity(self): return self.X[:, 10:13]
