"""Development probe: fixed per-call cost of the host-buffer entry point (tiny batch / tiny mission) next to a full-size call."""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from uav_ac_b200 import host_api
from uav_ac_b200.simulation.scene import LAB_COURSE_WAYPOINTS
for B, nt in ((32, 10), (32, 10760), (100000, 10)):
    host_api.fly_mission_host(LAB_COURSE_WAYPOINTS, 3.0, B, n_ticks=nt)
    ts = []
    for i in range(20):
        t0 = time.perf_counter(); host_api.fly_mission_host(LAB_COURSE_WAYPOINTS, 3.0, B, n_ticks=nt); ts.append(time.perf_counter() - t0)
    print(B, nt, f"median {np.median(ts) * 1e3:.3f} ms  min {min(ts) * 1e3:.3f} ms")
