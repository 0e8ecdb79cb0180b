"""Batched ``MinimumSnap``: the reference planner's surface (``uav_ac/planning/minimum_snap.py:6-395``)
over the CUDA kernels K1 (solve), table geometry, K3 (sampled table + yaw rules) and the sampled-point
AABB sweep of the obstacle-correction loop.

``MinimumSnap(path, obstacles, velocity, dt).get_trajectory()`` with a (S+1, 3) path returns the same
(N, 11) NumPy table ``[pos3 vel3 acc3 yaw spline_id]`` as the reference.  A (B, S+1, 3) array or a list
of per-mission arrays plans B missions at once (ragged after midpoint insertion); ``get_trajectory()``
then returns the packed (rows, 11) CUDA tensor and ``row_offsets`` delimits the missions.

The coefficients are the unique minimiser the reference obtains from its KKT system (K1 computes it
from the reduced block-tridiagonal form), so ``method`` ("lstsq" / "solve") is accepted for
signature compatibility and does not change the result.  The solver never forms the reference's constraint system; the
``A`` and ``b`` attributes are materialised on first access (on the device, in the reference's row order) once a plan exists.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Set

import numpy as np
import torch

from .. import _native as nat, kernels


class MinimumSnap:
    START_END_TIME_FACTOR = 1.5            # minimum_snap.py:10
    MIN_HORIZONTAL_SPEED_FOR_YAW = 1e-3    # minimum_snap.py:11

    def __init__(self, path, obstacles, velocity=1.0, dt=0.01):
        """
        :param path: waypoints (S+1, 3); batched: (B, S+1, 3) or a list of (S_b+1, 3) arrays
        :param obstacles: (n, 6) AABBs [xmin xmax ymin ymax zmin zmax] or None
        :param velocity: average velocity (scalar, or one per mission)
        :param dt: time step between rows of the generated trajectory
        """
        if not torch.cuda.is_available():
            raise nat.UavbError("no CUDA device visible: the batched MinimumSnap has no CPU implementation")
        nat.lib()
        self.coord_obstacles = obstacles
        self._single = not isinstance(path, (list, tuple)) and np.ndim(path) == 2
        self._paths: List[np.ndarray] = [np.array(path, dtype=float)] if self._single else [np.array(p, dtype=float) for p in path]
        for p in self._paths:
            if p.ndim != 2 or p.shape[1] != 3 or p.shape[0] < 2:
                raise ValueError("every path must have shape (S+1 >= 2, 3)")
        self.velocity = velocity
        self.dt = dt
        self.n_coeffs = 8
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.reset()

    # ------------------------------------------------------------------ reference attribute surface
    @property
    def waypoints(self):
        return self._paths[0] if self._single else self._paths

    @waypoints.setter
    def waypoints(self, value):
        if self._single:
            self._paths = [np.array(value, dtype=float)]
        else:
            self._paths = [np.array(p, dtype=float) for p in value]

    def reset(self):
        self.times = []
        self.spline_id = []
        self.nb_splines = None
        self.positions = []
        self.velocities = []
        self.accelerations = []
        self.yaws = []
        self.jerks = []
        self.snap = []
        self.full_trajectory = None
        self.row_counter = 0
        self._A = None
        self._b = None
        self.coeffs = None
        self.row_offsets = None
        self.status = None
        self.correction_rounds = 0         # plan rounds of the last obstacle-correction run (1 = nothing was hit)

    def get_trajectory(self):
        self._generate_collision_free_trajectory()
        return self.full_trajectory

    def _constraint_system(self):
        """A and b of the reference (minimum_snap.py:171-255) for the current waypoints and times; None before a plan exists
        (the reference allocates them in _setup, :288-293, during _compute_spline_parameters).  Single mission: NumPy
        (6S+2, 8S) / (6S+2, 3); batch: one tensor per mission (a stacked tensor when every mission has the same S)."""
        if self._A is None and self.coeffs is not None:
            times = self._times_dev
            off = self._seg_offsets.cpu().numpy()
            groups = {}
            for b_, p_ in enumerate(self._paths):
                groups.setdefault(len(p_) - 1, []).append(b_)
            A_list, b_list = [None] * len(self._paths), [None] * len(self._paths)
            for S, members in groups.items():
                wp = torch.tensor(np.stack([self._paths[m] for m in members]), dtype=torch.float64, device=self.device)
                tt = torch.stack([times[off[m]:off[m] + S] for m in members]).contiguous()
                A, b = kernels.minsnap_constraints(wp, tt)
                for k, m in enumerate(members):
                    A_list[m], b_list[m] = A[k], b[k]
            if self._single:
                self._A, self._b = A_list[0].cpu().numpy(), b_list[0].cpu().numpy()
            elif len(groups) == 1:
                self._A, self._b = torch.stack(A_list), torch.stack(b_list)
            else:
                self._A, self._b = A_list, b_list
        return self._A, self._b

    @property
    def A(self):
        return self._constraint_system()[0]

    @property
    def b(self):
        return self._constraint_system()[1]

    # ------------------------------------------------------------------ planning on the device
    def _velocity_tensor(self) -> torch.Tensor:
        B = len(self._paths)
        v = torch.as_tensor(self.velocity, dtype=torch.float64, device=self.device)
        return (v.expand(B) if v.dim() == 0 else v.reshape(B)).contiguous()

    def _pack(self):
        counts = np.array([len(p) for p in self._paths])
        offs = np.concatenate(([0], np.cumsum(counts))).astype(np.int32)
        wp = torch.tensor(np.concatenate(self._paths), dtype=torch.float64, device=self.device)
        return wp, torch.tensor(offs, device=self.device), counts - 1

    def _compute_spline_parameters(self, method="lstsq"):
        """Time allocation + coefficients for every mission (minimum_snap.py:138-153, 311-321), one K1 launch."""
        wp, wp_offs, splines = self._pack()
        if int(splines.max()) > nat.MAX_SPLINES:
            raise ValueError(f"too many splines ({int(splines.max())}; the solver accepts at most {nat.MAX_SPLINES} per mission)")
        B = len(self._paths)
        vel = self._velocity_tensor()
        if len(set(splines.tolist())) == 1:
            S = int(splines[0])
            c, t, status = kernels.minsnap_solve(wp.reshape(B, S + 1, 3), vel, self.START_END_TIME_FACTOR)
            c, t = c.reshape(B * S, 8, 3), t.reshape(-1)
        else:
            c, t, status = kernels.minsnap_solve_ragged(wp, wp_offs, vel, self.START_END_TIME_FACTOR)
        self._seg_offsets = (wp_offs - torch.arange(B + 1, dtype=torch.int32, device=self.device)).contiguous()
        self._coeffs_dev, self._times_dev, self.status = c, t, status
        self.nb_splines = int(splines[0]) if self._single else splines.tolist()
        if bool((status != 0).any()):
            bad = int((status != 0).nonzero()[0])
            # the reference's "solve" branch raises LinAlgError here (singular KKT matrix); lstsq returns garbage
            raise np.linalg.LinAlgError(f"mission {bad}: zero-length spline or non-positive velocity (singular minimum-snap system)")
        if self._single:
            self.coeffs = c.reshape(-1, 3).cpu().numpy()
            self.times = t.cpu().numpy().tolist()
        else:
            self.coeffs, self.times = c, t

    def _generate_trajectory(self, method="lstsq"):
        """Sampled (N, 11) table(s): rows at j*dt for j < ceil(T_i/dt) per spline, yaw hold/unwrap rules (minimum_snap.py:97-136)."""
        self._compute_spline_parameters(method)
        return self._sample_tables()

    def _sample_tables(self):
        """Table geometry + K3 over the current coefficients."""
        B = len(self._paths)
        rows, yaw0, total = kernels.table_meta(self._coeffs_dev, self._times_dev, self._seg_offsets, self.dt)
        roff = torch.zeros(B + 1, dtype=torch.int32, device=self.device)
        roff[1:] = torch.cumsum(total, 0)
        table = kernels.minsnap_sample(self._coeffs_dev, self._times_dev, self._seg_offsets, rows, roff, self.dt)
        self.row_offsets, self._table_dev = roff, table
        if self._single:
            tab = table.cpu().numpy()
            self.positions, self.velocities, self.accelerations = tab[:, 0:3], tab[:, 3:6], tab[:, 6:9]
            self.yaws, self.spline_id = tab[:, 9], tab[:, 10]
            self.full_trajectory = tab
        else:
            self.full_trajectory = table
        return self.full_trajectory

    def _generate_collision_free_trajectory(self):
        """Per obstacle: plan, mark the splines that own a sampled point inside the box, insert a midpoint waypoint in
        each, re-plan until clean (minimum_snap.py:63-95).  The whole loop runs on the device for all missions at once
        (kernels.minsnap_correct: every mission walks the obstacle list with its own cursor, only missions that were hit
        are planned again); the host reads the final waypoints back, like the reference's ``self.waypoints`` after the loop."""
        if self.coord_obstacles is None:
            self._generate_trajectory()
            return
        obstacles = np.asarray(self.coord_obstacles, dtype=float).reshape(-1, 6)
        if len(obstacles) == 0:            # an empty obstacle array leaves full_trajectory = None, like the reference
            return
        paths = self._paths
        self.reset()
        self._paths = paths
        obs = torch.tensor(obstacles, dtype=torch.float64, device=self.device)
        try:
            c, t, seg_off, status, wp, n_wp, self.correction_rounds = kernels.plan_collision_free(
                self._paths, self._velocity_tensor(), self.dt, obs, self.START_END_TIME_FACTOR, device=self.device)
        except kernels.TooManySplines as exc:
            raise RuntimeError(f"obstacle correction did not converge: {exc}") from None
        if bool((status != 0).any()):
            bad = int((status != 0).nonzero()[0])
            raise np.linalg.LinAlgError(f"mission {bad}: zero-length spline or non-positive velocity (singular minimum-snap system)")
        wp_h, n_h = wp.cpu().numpy(), n_wp.cpu().numpy()
        self._paths = [wp_h[b, :n_h[b]].copy() for b in range(len(n_h))]
        self._coeffs_dev, self._times_dev, self._seg_offsets, self.status = c, t, seg_off, status
        splines = (n_h - 1).tolist()
        self.nb_splines = int(splines[0]) if self._single else splines
        if self._single:
            self.coeffs = c.reshape(-1, 3).cpu().numpy()
            self.times = t.cpu().numpy().tolist()
        else:
            self.coeffs, self.times = c, t
        self._sample_tables()

    def trajectories(self) -> List[np.ndarray]:
        """Per-mission (N_b, 11) NumPy tables of a batched plan."""
        tab, off = self._table_dev.cpu().numpy(), self.row_offsets.cpu().numpy()
        return [tab[off[b]:off[b + 1]] for b in range(len(off) - 1)]

    # ------------------------------------------------------------------ small host-side helpers of the reference API
    @staticmethod
    def _calculate_yaws(velocities) -> np.ndarray:
        """Heading profile of a velocity sequence (minimum_snap.py:126-136), computed by the K3 yaw kernels."""
        v = torch.tensor(np.asarray(velocities, dtype=float).reshape(-1, 3), dtype=torch.float64, device=torch.device("cuda", torch.cuda.current_device()))
        n = v.shape[0]
        out = torch.empty(n, dtype=torch.float64, device=v.device)
        offs = torch.tensor([0, n], dtype=torch.int32, device=v.device)
        nat.check(nat.lib().uavb_minsnap_yaw_profile_f64(nat.ptr(v), nat.ptr(offs), 1, n, nat.ptr(out), nat.stream_ptr(v.device)),
                  "uavb_minsnap_yaw_profile_f64")
        return out.cpu().numpy()

    @staticmethod
    def polynom(n_coeffs, order, t):
        """k-th derivative of the ascending monomial basis at t (minimum_snap.py:258-286): entry i = i!/(i-k)! t^(i-k)."""
        i = np.arange(n_coeffs, dtype=float)
        fall = np.ones(n_coeffs)
        expo = i.copy()
        for _ in range(order):
            fall = fall * expo
            expo = np.where(expo > 0, expo - 1, expo)
        return fall * np.power(float(t), expo)

    @staticmethod
    def is_collision_cuboid(x: float, y: float, z: float, cuboid_params) -> bool:
        """Inclusive point-in-AABB (minimum_snap.py:327-357)."""
        x_min, x_max, y_min, y_max, z_min, z_max = cuboid_params
        return bool(x_min <= x <= x_max and y_min <= y <= y_max and z_min <= z <= z_max)

    @staticmethod
    def insert_midpoints_at_indexes(points, indexes: Sequence[int] | Set[int]) -> np.ndarray:
        """Insert (p[i-1] + p[i]) / 2 before every listed index (minimum_snap.py:359-391)."""
        points = np.asarray(points, dtype=float)
        result = []
        for i in range(len(points)):
            if i in indexes:
                result.append((points[i - 1] + points[i]) / 2)
            result.append(points[i])
        return np.array(result)
