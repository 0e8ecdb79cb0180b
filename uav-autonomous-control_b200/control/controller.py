"""Batched ``CascadedController``: same methods and argument meaning as
``uav_ac/control/controller.py:4-191`` with a leading batch dimension; every method is one launch of
the sm_100a stage kernel (``uavb_stage_f32``), sharing its arithmetic with the persistent rollout.
"""
from __future__ import annotations

import math

import torch

from .. import _native as nat, _stages as st


class _ArgumentsOnly:
    """Stand-in vehicle for controller methods that read nothing but their arguments (roll_pitch_controller): the lab_course
    constants with the gains the caller passed."""

    def __init__(self, batch: int, device):
        self.batch, self.device = int(batch), device

    def vehicle_struct(self, B: int, gain_overrides: dict):
        v = nat.default_vehicle()
        mc = {}
        gains = [gain_overrides.get(n, v.gains[k]) for k, n in enumerate(nat.GAIN_NAMES)]
        if any(isinstance(gv, torch.Tensor) for gv in gains):
            mc["mc_gains"] = torch.stack([st.as_batch(gv, B, None, self.device) for gv in gains]).contiguous()
            v.gains[:] = [float(torch.as_tensor(gv).double().mean()) for gv in gains]
        else:
            v.gains[:] = [float(gv) for gv in gains]
        return v, mc


class CascadedController:
    """Cascaded controller (Lupashin et al.) for B drones."""

    INTEGRAL_ERROR_LIMIT = 10.0   # controller.py:10

    def __init__(self, g: float, dt: float):
        """:param g: gravity acceleration  :param dt: time step of the outer control loop (controller.py:12-20)"""
        self.g = g
        self.dt = dt
        self.integral_error = 0      # becomes a (B,) tensor at the first altitude() call

    def reset(self) -> None:
        """Clear the state accumulated across control cycles (controller.py:22-24)."""
        if isinstance(self.integral_error, torch.Tensor):
            self.integral_error.zero_()
        else:
            self.integral_error = 0

    def _integral(self, quad) -> torch.Tensor:
        if not isinstance(self.integral_error, torch.Tensor):
            self.integral_error = torch.full((quad.batch,), float(self.integral_error), dtype=torch.float32, device=quad.device)
        return self.integral_error

    @staticmethod
    def _target(quad, des_x=None, des_y=None, des_z=None, yaw=None) -> torch.Tensor:
        """[10][B] table-row layout (x y z vx vy vz ax ay az yaw) from per-axis [pos, vel, acc] triples."""
        B = quad.batch
        t = torch.zeros((10, B), dtype=torch.float32, device=quad.device)
        for axis, des in enumerate((des_x, des_y, des_z)):
            if des is not None:
                d = st.as_batch(des, B, 3, quad.device)
                t[axis], t[3 + axis], t[6 + axis] = d[:, 0], d[:, 1], d[:, 2]
        if yaw is not None:
            t[9] = st.as_batch(yaw, B, None, quad.device)
        return t

    @staticmethod
    def _rot(quad, rot_mat):
        if rot_mat is None:
            return None
        r = torch.as_tensor(rot_mat, dtype=torch.float32, device=quad.device)
        if r.dim() == 2:
            r = r.unsqueeze(0).expand(quad.batch, 3, 3)
        return r.reshape(quad.batch, 9).t().contiguous()

    def altitude(self, quad, des_z, rot_mat, kp_z, kd_z, ki_z) -> torch.Tensor:
        """Collective thrust command (B,) (controller.py:26-56).  des_z = [z, z_dot, z_ddot] per drone."""
        B = quad.batch
        thrust = torch.empty((B,), dtype=torch.float32, device=quad.device)
        st.run(nat.STAGE_ALTITUDE, quad, self.dt, B, gains=dict(kp_z=kp_z, kd_z=kd_z, ki_z=ki_z), X=st.soa(quad.X),
               target=self._target(quad, des_z=des_z), rot=self._rot(quad, rot_mat), integral=self._integral(quad), thrust=thrust)
        return thrust

    def lateral(self, quad, des_x, des_y, thrust_cmd, kp_xy, kd_xy) -> torch.Tensor:
        """Commanded rotation-matrix entries [R02, R12] (B, 2) (controller.py:58-97)."""
        B = quad.batch
        bxy = torch.empty((2, B), dtype=torch.float32, device=quad.device)
        st.run(nat.STAGE_LATERAL, quad, self.dt, B, gains=dict(kp_xy=kp_xy, kd_xy=kd_xy), X=st.soa(quad.X),
               target=self._target(quad, des_x=des_x, des_y=des_y), thrust=st.as_batch(thrust_cmd, B, None, quad.device), bxy=bxy)
        return bxy.t().contiguous()

    def reduced_attitude(self, quad, bxy_cmd, psi_des, rot_mat, kp_roll, kp_pitch, kp_yaw) -> torch.Tensor:
        """pqr_cmd (B, 3): roll/pitch rates from the tilt error, yaw rate from the heading error (controller.py:99-113)."""
        pq_cmd = self.roll_pitch_controller(bxy_cmd, rot_mat, kp_roll, kp_pitch, quad=quad)
        r_cmd = self.yaw_controller(quad, psi_des, kp_yaw, pq_cmd[:, 1])
        return torch.cat((pq_cmd, r_cmd[:, None]), dim=1)

    def body_rate_controller(self, quad, pqr_cmd, kp_p, kp_q, kp_r) -> torch.Tensor:
        """Body moments (B, 3): I kp (cmd - w) + w x (I w) (controller.py:115-130)."""
        B = quad.batch
        moment = torch.empty((3, B), dtype=torch.float32, device=quad.device)
        st.run(nat.STAGE_BODY_RATE, quad, self.dt, B, gains=dict(kp_p=kp_p, kp_q=kp_q, kp_r=kp_r), X=st.soa(quad.X),
               pqr_cmd=st.soa(st.as_batch(pqr_cmd, B, 3, quad.device)), moment=moment)
        return moment.t().contiguous()

    def roll_pitch_controller(self, bxy_cmd, rot_mat, kp_roll, kp_pitch, quad=None) -> torch.Tensor:
        """[p_cmd, q_cmd] (B, 2) from the commanded and actual [R02, R12] (controller.py:132-154).
        The reference method takes no vehicle: it reads only its arguments.  So does this one -- the batch size and the device
        come from ``bxy_cmd`` / ``rot_mat`` ((B, 2) / (B, 3, 3); plain (2,) / (3, 3) inputs are one drone, as in the reference) --
        and ``quad`` is an optional hint for them (reduced_attitude passes it)."""
        if quad is None:
            bt, rt = torch.as_tensor(bxy_cmd), torch.as_tensor(rot_mat)
            B = bt.shape[0] if bt.dim() == 2 else (rt.shape[0] if rt.dim() == 3 else 1)
            device = next((t.device for t in (bt, rt) if t.is_cuda), torch.device("cuda", torch.cuda.current_device()))
            quad = _ArgumentsOnly(B, device)
        B = quad.batch
        pqr = torch.zeros((3, B), dtype=torch.float32, device=quad.device)
        st.run(nat.STAGE_ROLL_PITCH, quad, self.dt, B, gains=dict(kp_roll=kp_roll, kp_pitch=kp_pitch), X=st.soa(getattr(quad, "X", None)),
               bxy=st.soa(st.as_batch(bxy_cmd, B, 2, quad.device)), rot=self._rot(quad, rot_mat), pqr_cmd=pqr)
        return pqr[:2].t().contiguous()

    def yaw_controller(self, quad, psi_des, kp_yaw, q_cmd) -> torch.Tensor:
        """Body yaw-rate command (B,) from the Euler yaw error (controller.py:156-168)."""
        B = quad.batch
        pqr = torch.zeros((3, B), dtype=torch.float32, device=quad.device)
        pqr[1] = st.as_batch(q_cmd, B, None, quad.device)
        st.run(nat.STAGE_YAW, quad, self.dt, B, gains=dict(kp_yaw=kp_yaw), X=st.soa(quad.X), target=self._target(quad, yaw=psi_des), pqr_cmd=pqr)
        return pqr[2].contiguous()

    # ------------------------------------------------------------------ scalar helpers (controller.py:170-191)
    @staticmethod
    def wrap_to_pi(angle):
        return (angle + math.pi) % (2 * math.pi) - math.pi

    @staticmethod
    def wrap_to_2pi(angle):
        return angle % (2 * math.pi)

    @staticmethod
    def _pd(kp, kd, error, error_dot, target):
        return kp * error + kd * error_dot + target

    @staticmethod
    def _pid(kp, kd, ki, error, error_dot, i_error, target):
        return kp * error + ki * i_error + kd * error_dot + target
