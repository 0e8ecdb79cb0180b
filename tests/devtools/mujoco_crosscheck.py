#!/usr/bin/env python
"""Cross-check of the restated rigid-body step against the REAL MuJoCo engine -- for anyone who has MuJoCo installed.

NOT RUNNABLE IN THE BUILD CONTAINER OR ON THE GPU BOX (MuJoCo 3.x is absent there and cannot be installed), so this script
has never been executed by the authors of this repository; it exists because parity of oracle/freebody.py with MuJoCo is
UNPINNED (DESIGN.md "Oracle") and only a machine with MuJoCo can pin it.

    python tests/devtools/mujoco_crosscheck.py --reference /path/to/UAV-Autonomous-control [--velocity 2.0]

It runs the reference's own headless loop (tests/integration/test_mujoco_trajectory_tracking.py:11-36: MujocoSimulation +
TrajectoryController + CascadedController + Quad), records quad.X after every outer period, flies the same table with the
NumPy oracle (oracle/flight_np.py on oracle/freebody.py) and prints the deviations.  Expected: ~1e-2 m during the
first 0.1 s (MuJoCo's soft ground contact at take-off, which the free-body model does not have), 1e-4 .. 1e-5 m afterwards.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="checkout of Mdhvince/UAV-Autonomous-control")
    ap.add_argument("--velocity", type=float, default=2.0)
    ap.add_argument("--frequency", type=int, default=10)
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    try:
        import mujoco  # noqa: F401
    except ImportError:
        raise SystemExit("MuJoCo is not installed: this cross-check needs `pip install mujoco` (the reference pins 3.11.0)")
    from uav_ac.control.controller import CascadedController
    from uav_ac.main import TrajectoryController, _generate_mission_trajectory
    from uav_ac.simulation.mujoco_sim import MujocoSimulation
    from oracle import flight_np

    sim = MujocoSimulation()
    quad = sim.quad
    dt_outer = quad.dt * args.frequency
    table = _generate_mission_trajectory(sim.mission_waypoints, sim.obstacles, args.velocity, dt_outer)
    tc = TrajectoryController(CascadedController(quad.g, dt_outer), quad, table, args.frequency)
    X_mj = []
    for _ in table:
        for _ in range(args.frequency):
            tc.step()
            sim.step()
        X_mj.append(quad.X.copy())
    X_mj = np.array(X_mj)

    out = flight_np.closed_loop(flight_np.Vehicle(), table, sim.mission_waypoints[0], freq=args.frequency, obstacles=sim.obstacles,
                                goal=sim.goal_position, log_stride=args.frequency)
    X_or = out["log"]
    dp = np.linalg.norm(X_mj[:, :3] - X_or[:, :3], axis=1)
    early = int(round(0.1 / dt_outer))
    print(f"rows {len(table)}  MuJoCo collision flag {sim.collision_detected}  oracle AABB flag {out['collision']}")
    print(f"max |dpos| first 0.1 s : {dp[:early].max():.3e} m   (ground contact phase, expected ~1e-2)")
    print(f"max |dpos| 0.1 s .. 2 s: {dp[early:int(2 / dt_outer)].max():.3e} m")
    print(f"max |dpos| after 2 s   : {dp[int(2 / dt_outer):].max():.3e} m   (expected 1e-4 .. 1e-5)")
    print(f"final distance to goal : MuJoCo {np.linalg.norm(X_mj[-1, :3] - sim.goal_position):.5f} m, oracle {out['final_dist']:.5f} m")
    for lag in (1, 0):
        o = flight_np.closed_loop(flight_np.Vehicle(), table, sim.mission_waypoints[0], freq=args.frequency, thrust_frame_lag=lag, log_stride=args.frequency)
        d = np.linalg.norm(X_mj[int(2 / dt_outer):, :3] - o["log"][int(2 / dt_outer):, :3], axis=1).max()
        print(f"thrust_frame_lag={lag}: max |dpos| after 2 s {d:.3e} m   (the headless loop is expected to match lag=1, SURVEY 3.2)")


if __name__ == "__main__":
    main()
