"""Mission loop of the reference (``uav_ac/main.py``) with a leading batch dimension.

``TrajectoryController`` keeps the reference's name, constructor and ``step()/reset()`` contract
(main.py:10-61) and drives B drones per call through the stage kernels; ``main()`` flies the
laboratory course for a Monte-Carlo batch in one launch of the persistent rollout kernel and prints
the mission report of main.py:115-120 as fractions over the batch.

Precision of the method-level path: ``quad.X`` and the table handed to ``TrajectoryController`` are float32 tensors (the stage
kernels' interface), so positions round at ~1e-6 m near |p| = 20 m every tick and the path holds ~1e-3 m over a mission -- it
exists for method-by-method parity with the reference classes.  The fused rollout (``BatchedSimulation.rollout`` /
``kernels.rollout``) keeps fp64 set-points and an fp64 position fold and is the one that meets the 1e-4 m budget.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as nat, _stages as st, sharding, utils
from .control.controller import CascadedController
from .planning.minimum_snap import MinimumSnap
from .quadrotor.quad import Quad
from .simulation.batched_sim import BatchedSimulation


class TrajectoryController:
    """Drive the cascaded controller from a time-parameterised trajectory, B drones at once."""

    def __init__(self, controller: CascadedController, quad: Quad, trajectory, inner_loop_frequency: int):
        """``trajectory`` is the (N, 11) table of ``MinimumSnap.get_trajectory()`` shared by all drones, or (B, N, 11)."""
        self.controller = controller
        self.quad = quad
        t = torch.as_tensor(np.asarray(trajectory) if not isinstance(trajectory, torch.Tensor) else trajectory)
        self.trajectory = t.to(device=quad.device, dtype=torch.float32).contiguous()
        self.inner_loop_frequency = inner_loop_frequency
        self.trajectory_index = 0
        self.inner_step = 0
        self.thrust_cmd = torch.zeros((quad.batch,), dtype=torch.float32, device=quad.device)
        self.pqr_cmd = torch.zeros((quad.batch, 3), dtype=torch.float32, device=quad.device)

    def reset(self) -> None:
        """Restart trajectory tracking from its initial state (main.py:29-35)."""
        self.controller.reset()
        self.trajectory_index = 0
        self.inner_step = 0
        self.thrust_cmd.zero_()
        self.pqr_cmd.zero_()

    def step(self) -> None:
        """One inner body-rate control cycle for every drone (main.py:37-45)."""
        if self.inner_step % self.inner_loop_frequency == 0:
            self._update_outer_loop()
        q = self.quad
        B = q.batch
        omega, cmd = st.soa(q.omega), torch.empty((4, B), dtype=torch.float32, device=q.device)
        st.run(nat.STAGE_INNER, q, self.controller.dt, B, X=st.soa(q.X), thrust=self.thrust_cmd, pqr_cmd=st.soa(self.pqr_cmd), omega=omega,
               omega_cmd=cmd)
        q.omega, q.omega_command = omega.t().contiguous(), cmd.t().contiguous()
        self.inner_step += 1

    def _update_outer_loop(self) -> None:
        """main.py:47-61: altitude, lateral, reduced attitude on the current table row; index clamps at the end."""
        q = self.quad
        B = q.batch
        n = self.trajectory.shape[-2]
        row = self.trajectory[..., self.trajectory_index, :10]
        target = (row.t() if row.dim() == 2 else row[:, None].expand(10, B)).contiguous()
        pqr = torch.empty((3, B), dtype=torch.float32, device=q.device)
        st.run(nat.STAGE_OUTER, q, self.controller.dt, B, X=st.soa(q.X), target=target, integral=self.controller._integral(q),
               thrust=self.thrust_cmd, pqr_cmd=pqr)
        self.pqr_cmd = pqr.t().contiguous()
        self.trajectory_index = min(self.trajectory_index + 1, n - 1)


def _trajectory_after_takeoff(trajectory, takeoff_waypoint) -> np.ndarray:
    """Rows of the table from the one closest to the take-off waypoint onwards (main.py:64-70; used for display only)."""
    table = np.asarray(trajectory)
    first = int(np.argmin(((table[:, :3] - np.asarray(takeoff_waypoint)) ** 2).sum(axis=1)))
    return table[first:]


def _generate_mission_trajectory(waypoints, obstacles, velocity: float, dt: float) -> np.ndarray:
    """Two independent plans stacked (main.py:73-84): the vertical take-off (start -> first waypoint), then the course."""
    legs = (waypoints[:2], waypoints[1:])
    return np.vstack([MinimumSnap(leg, obstacles, velocity, dt).get_trajectory() for leg in legs])


def main(batch: int = 100_000) -> None:
    """The mission of ``uav_ac/main.py:87-120`` for a Monte-Carlo batch (headless: MuJoCo's viewer is out of scope)."""
    cfg, cfg_flight = utils.get_config()
    bcfg = utils.get_batch_config()
    frequency = cfg.getint("frequency")
    velocity = cfg_flight.getfloat("velocity")
    min_distance_target = cfg_flight.getfloat("min_dist_target")

    from . import kernels
    simulation = BatchedSimulation(batch)
    lo_g, hi_g = (float(v) for v in bcfg.get("gain_scale").split(","))
    lo_m, hi_m = (float(v) for v in bcfg.get("mass_scale").split(","))
    lo_i, hi_i = (float(v) for v in bcfg.get("inertia_scale").split(","))
    scales = kernels.mc_uniform(bcfg.getint("seed"), batch, [lo_g] * 11 + [lo_m] + [lo_i] * 3, [hi_g] * 11 + [hi_m] + [hi_i] * 3)
    quad = simulation.quad
    gains = {name: scales[k] * float(getattr(quad, name)) for k, name in enumerate(Quad.GAIN_NAMES)}
    quad.m = scales[11] * quad.m
    quad.i_x, quad.i_y, quad.i_z = scales[12] * quad.i_x, scales[13] * quad.i_y, scales[14] * quad.i_z
    result = simulation.rollout(velocity, frequency, gains=gains)
    torch.cuda.synchronize()
    report = sharding.summarize(result.metrics, min_distance_target)
    print(f"{report['rollouts']} flights finished on average {report['mean_final_dist']:.2f} m away from the goal "
          f"({100 * report['reached_fraction']:.2f} % reached).")
    if report["collision_fraction"] > 0:
        print(f"At least one collision occurred in {100 * report['collision_fraction']:.2f} % of the flights.")


if __name__ == "__main__":
    main()
