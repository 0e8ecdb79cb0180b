"""Development probe: throughput of the warp-per-mission RRT* kernel on random lab-volume missions."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from uav_ac_b200.planning.rrt import RRTStar
from uav_ac_b200.simulation.scene import LAB_COURSE_OBSTACLES as OBS, PLANNING_BOUNDS as LIM
rng = np.random.default_rng(0)
for B in [int(x) for x in (sys.argv[1:] or ["2048", "16384"])]:
    s = np.round(rng.uniform(LIM[0] + [0.5, 0.5, 0.3], [3.0, 13.5, -0.5], (B, 3)), 2)
    g = np.round(rng.uniform([21.0, 0.5, -5.5], LIM[1] - [0.5, 0.5, 0.5], (B, 3)), 2)
    r = RRTStar(LIM, s, g, 1.5, 1500, OBS, seed=1)
    r.run()                                   # warm-up (allocations)
    torch.cuda.synchronize(); t0 = time.perf_counter(); r.run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    it = r.stats[:, 0].astype(float)
    print(f"B={B}: {dt * 1e3:.1f} ms -> {B / dt:.0f} missions/s, found {(r.status == 0).mean():.3f}, iterations/mission mean {it.mean():.0f}, "
          f"tree iterations/s {it.sum() / dt / 1e6:.2f} M, mean cost {np.mean(r.cost[r.status == 0]):.2f} m")
