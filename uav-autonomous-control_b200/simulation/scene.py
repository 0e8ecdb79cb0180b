"""lab_course scene constants in NED, transcribed from the reference's single source of truth
``uav_ac/simulation/models/lab_course.xml`` (MuJoCo ENU -> NED is diag(1,-1,-1),
``uav_ac/simulation/mujoco_sim.py:10``).  Values match what ``MujocoSimulation`` exposes as
``mission_waypoints`` / ``obstacles`` / ``goal_position`` (reference tests
``tests/unit/simulation/test_mujoco_sim.py:10,32,40-50,246``).
"""
import numpy as np

# start (lab_course.xml:98), waypoint_00..06 (:78-84), goal (:96)
LAB_COURSE_WAYPOINTS = np.array([
    [1.0, 7.0, -0.021], [1.0, 7.0, -1.3], [4.0, 7.0, -1.3], [7.5, 4.0, -3.0], [11.0, 7.0, -3.5],
    [14.0, 10.0, -2.5], [17.0, 10.0, -3.2], [20.5, 7.0, -1.4], [23.0, 7.0, -2.0]])
# obstacle_00..03 (lab_course.xml:37,53,54,67) as [xmin xmax ymin ymax zmin zmax] (mujoco_sim.py:282-300)
LAB_COURSE_OBSTACLES = np.array([
    [3.7, 4.3, 4.0, 10.0, -3.4, -2.8], [10.7, 11.3, 4.0, 10.0, -2.2, 0.0],
    [13.3, 14.7, 6.3, 7.7, -6.0, 0.0], [20.2, 20.8, 4.0, 10.0, -3.3, -2.7]])
LAB_COURSE_START = LAB_COURSE_WAYPOINTS[0].copy()
LAB_COURSE_GOAL = LAB_COURSE_WAYPOINTS[-1].copy()
PLANNING_BOUNDS = np.array([[0.0, 0.0, -6.0], [24.0, 14.0, 0.0]])   # lab_course.xml:8
