"""Pins the CPU oracle (oracle/*.py) on outputs of the REFERENCE's own classes.

tests/golden/*.npz were produced by tests/golden/make_golden.py, which imports
/root/reference/uav_ac/{planning/minimum_snap,control/controller,quadrotor/quad,main}.py and records
what MinimumSnap / CascadedController / Quad / TrajectoryController return.  The known-answer vectors
of the reference's own unit tests (tests/unit/planning/test_minimum_snap.py,
tests/unit/control/test_controller.py, tests/unit/quadrotor/test_quad.py) are restated here against
the oracle as well.  No GPU and no /root/reference needed at run time.

The rigid-body step (oracle/freebody.py) stands in for MuJoCo in BOTH the golden generator and the
oracle, so the closed-loop comparisons pin controller / allocation / motor lag / scheduler, not the
MuJoCo boundary (parity unpinned there; only the reference's hover / gravity-sign expectations apply).
"""
import math

import numpy as np
import pytest

from oracle import flight_np, freebody, minsnap_np
from helpers import GOAL, normwise


# ------------------------------------------------------------------------------------------ planning
def test_basis_rows_match_reference_polynom(golden):
    g = golden["planning"]
    for k in range(7):
        for j, t in enumerate(g["polynom_t"]):
            np.testing.assert_allclose(minsnap_np.basis_row(k, float(t)), g["polynom"][k, j], rtol=1e-15, atol=0)


def test_basis_rows_known_answers_of_reference_unit_tests():
    """tests/unit/planning/test_minimum_snap.py:22-61: exact integer vectors at t = 0 and t = 3."""
    at0 = {0: [1, 0, 0, 0, 0, 0, 0, 0], 1: [0, 1, 0, 0, 0, 0, 0, 0], 2: [0, 0, 2, 0, 0, 0, 0, 0], 3: [0, 0, 0, 6, 0, 0, 0, 0],
           4: [0, 0, 0, 0, 24, 0, 0, 0], 5: [0, 0, 0, 0, 0, 120, 0, 0], 6: [0, 0, 0, 0, 0, 0, 720, 0]}
    for k, want in at0.items():
        np.testing.assert_array_equal(minsnap_np.basis_row(k, 0.0), want)
    np.testing.assert_array_equal(minsnap_np.basis_row(0, 3.0), [1, 3, 9, 27, 81, 243, 729, 2187])
    np.testing.assert_array_equal(minsnap_np.basis_row(1, 3.0), [0, 1, 6, 27, 108, 405, 1458, 5103])
    np.testing.assert_array_equal(minsnap_np.basis_row(2, 3.0), [0, 0, 2, 18, 108, 540, 2430, 10206])
    np.testing.assert_array_equal(minsnap_np.basis_row(4, 3.0), [0, 0, 0, 0, 24, 360, 3240, 22680])
    np.testing.assert_array_equal(minsnap_np.basis_row(6, 3.0), [0, 0, 0, 0, 0, 0, 720, 15120])


@pytest.mark.parametrize("tag", ["v2", "v3"])
@pytest.mark.parametrize("name", ["takeoff", "course"])
def test_constraints_hessian_times_and_coefficients_match_reference(golden, tag, name):
    g = golden["planning"]
    wp = g["waypoints"][:2] if name == "takeoff" else g["waypoints"][1:]
    v = float(tag[1])
    T = minsnap_np.segment_times(wp, v)
    np.testing.assert_allclose(T, g[f"{tag}_{name}_times"], rtol=1e-15)
    A, b = minsnap_np.constraint_system(wp, T)
    np.testing.assert_allclose(A, g[f"{tag}_{name}_A"], rtol=1e-14, atol=0)          # same row order as minimum_snap.py:171-255
    np.testing.assert_array_equal(b, g[f"{tag}_{name}_b"])
    np.testing.assert_allclose(minsnap_np.snap_hessian(T), g[f"{tag}_{name}_Q"], rtol=1e-14, atol=0)
    for method in ("solve", "lstsq"):
        c, _ = minsnap_np.solve_coeffs(wp, v, method)
        assert normwise(c, g[f"{tag}_{name}_coeffs_{method}"]) < 1e-12            # same LAPACK call on the same matrix


def test_lab_course_times_are_the_survey_anchors(golden):
    """SURVEY 8(a) P2 values measured from the reference."""
    wp = golden["planning"]["waypoints"]
    np.testing.assert_allclose(minsnap_np.segment_times(wp[:2], 3.0), [0.6395], rtol=1e-12)
    np.testing.assert_allclose(minsnap_np.segment_times(wp[1:], 3.0),
                               [1.5, 1.637749133, 1.545603083, 1.452966315, 1.026861453, 1.649579071, 1.285496013], rtol=1e-9)
    np.testing.assert_allclose(minsnap_np.segment_times(wp[:2], 2.0), [0.95925], rtol=1e-12)


def test_random_and_ragged_missions_match_reference(golden):
    g = golden["planning"]
    for i in range(0, 64, 4):
        c, T = minsnap_np.solve_coeffs(g["c2_waypoints"][i], float(g["c2_velocity"][i]), "solve")
        assert normwise(c, g["c2_coeffs_solve"][i]) < 1e-12
        np.testing.assert_allclose(T, g["c2_times"][i], rtol=1e-15)
    for S in (1, 2, 3, 5, 8, 12):
        c, T = minsnap_np.solve_coeffs(g[f"rag{S}_waypoints"][0], float(g[f"rag{S}_velocity"][0]), "solve")
        assert normwise(c, g[f"rag{S}_coeffs_solve"][0]) < 1e-11
        np.testing.assert_allclose(T, g[f"rag{S}_times"][0], rtol=1e-15)


@pytest.mark.parametrize("tag", ["v2", "v3"])
def test_mission_table_matches_reference_generate_mission_trajectory(golden, tag):
    g = golden["planning"]
    tab = minsnap_np.mission_table(g["waypoints"], g["obstacles"], float(tag[1]), 0.01)
    ref = g[f"{tag}_table"]
    assert tab.shape == ref.shape                                                 # 1613 / 1076 rows (np.arange counts)
    np.testing.assert_allclose(tab, ref, rtol=0, atol=1e-9)


def test_sampled_table_of_a_random_mission_matches_get_trajectory(golden):
    g = golden["planning"]
    tab = minsnap_np.plan_table(g["c2_waypoints"][3], None, float(g["c2_velocity"][3]), 0.01)[0]
    assert tab.shape == g["c2_table3"].shape
    np.testing.assert_allclose(tab, g["c2_table3"], rtol=0, atol=1e-9)


def test_yaw_profiles_match_reference_calculate_yaws(golden):
    g = golden["planning"]
    for v, y in zip(g["yaw_vel"], g["yaw_out"]):
        np.testing.assert_allclose(minsnap_np.yaw_profile(v), y, rtol=0, atol=1e-14)
    # reference unit tests :93-136: +x line -> 0, +y line -> pi/2, all-invalid -> zeros
    assert np.all(minsnap_np.yaw_profile(np.tile([1.0, 0, 0], (5, 1))) == 0.0)
    np.testing.assert_allclose(minsnap_np.yaw_profile(np.tile([0.0, 2.0, 0], (5, 1))), np.pi / 2)
    assert np.all(minsnap_np.yaw_profile(np.zeros((5, 3))) == 0.0)


def test_aabb_truth_table_and_midpoint_insertion_match_reference(golden):
    g = golden["planning"]
    got = [minsnap_np.point_in_cuboid(*p, g["aabb_box"]) for p in g["aabb_pts"]]
    np.testing.assert_array_equal(got, g["aabb_hit"])
    np.testing.assert_array_equal(minsnap_np.insert_midpoints(g["mid_points"], [1, 3]), g["mid_out_13"])
    np.testing.assert_array_equal(minsnap_np.insert_midpoints(g["mid_points"], {2}), g["mid_out_2"])
    # reference test :7-19
    pts = np.array([[0.0, 0, 0], [2, 0, 0], [4, 0, 0], [6, 0, 0]])
    np.testing.assert_array_equal(minsnap_np.insert_midpoints(pts, [1, 3]),
                                  [[0, 0, 0], [1, 0, 0], [2, 0, 0], [4, 0, 0], [5, 0, 0], [6, 0, 0]])


def test_obstacle_correction_loop_matches_reference(golden):
    """minimum_snap.py:63-95: the midpoint-insertion loop ends on the same waypoints, coefficients and table."""
    g = golden["planning"]
    tab, w, c, T = minsnap_np.plan_table(g["fix_waypoints_in"], g["fix_obstacles"], 1.5, 0.01)
    np.testing.assert_allclose(w, g["fix_waypoints_out"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(T, g["fix_times"], rtol=1e-15)
    assert normwise(c, g["fix_coeffs"]) < 1e-11
    assert tab.shape == g["fix_table"].shape
    np.testing.assert_allclose(tab, g["fix_table"], rtol=0, atol=1e-8)
    assert minsnap_np.plan_table(g["fix_waypoints_in"], np.zeros((0, 6)), 1.5, 0.01)[0] is None     # empty-obstacle quirk (P10)


def test_kkt_optimality_and_waypoint_interpolation():
    """Properties the reference tests assert (:64-76, :139-168): the solution interpolates the waypoints,
    is C^4 at the junctions, and the reduced gradient N^T Q c vanishes (N = null space of A)."""
    wp = np.array([[0.0, 0, 0], [2, 1, -1], [4, -1, -2], [6, 0, -1]])
    T = minsnap_np.segment_times(wp, 1.0)
    c, _ = minsnap_np.solve_coeffs(wp, 1.0, "solve")
    A, b = minsnap_np.constraint_system(wp, T)
    assert np.abs(A @ c - b).max() < 1e-9
    Q = minsnap_np.snap_hessian(T)
    _, s, Vt = np.linalg.svd(A)
    N = Vt[np.sum(s > 1e-10 * s[0]):].T
    assert np.abs(N.T @ Q @ c).max() < 1e-6
    np.testing.assert_allclose(T, [1.5 * math.sqrt(6), 3.0, 1.5 * math.sqrt(6)], rtol=1e-15)


# ------------------------------------------------------------------------------------------ controller / vehicle stages
def test_gains_and_wraps_match_reference(golden):
    g = golden["stages"]
    veh = flight_np.Vehicle()
    np.testing.assert_allclose([getattr(veh, n) for n in veh.GAIN_NAMES], g["gains"], rtol=1e-15)
    np.testing.assert_allclose([flight_np.wrap_to_pi(a) for a in g["wrap_in"]], g["wrap_pi"], rtol=0, atol=1e-15)
    np.testing.assert_allclose([flight_np.wrap_to_2pi(a) for a in g["wrap_in"]], g["wrap_2pi"], rtol=0, atol=1e-15)
    # reference tests/unit/control/test_controller.py:24-50
    assert flight_np.wrap_to_pi(3 * math.pi) == pytest.approx(-math.pi) or flight_np.wrap_to_pi(3 * math.pi) == pytest.approx(math.pi)
    assert flight_np.wrap_to_2pi(-math.pi / 2) == pytest.approx(3 * math.pi / 2)


def test_outer_loop_stages_match_reference_controller(golden):
    g = golden["stages"]
    veh = flight_np.Vehicle()
    for i in range(len(g["X"])):
        X = g["X"][i]
        R = freebody.quat_to_rot(X[3:7])
        np.testing.assert_allclose(R, g["R"][i], rtol=0, atol=1e-14)
        np.testing.assert_allclose(flight_np.euler_from_quat(X[3:7]), g["euler"][i], rtol=0, atol=1e-14)
        c, integ = flight_np.altitude(veh, X[2], X[9], g["des"][i, 2], R[2, 2], g["integ0"][i], 0.01)
        assert c == pytest.approx(g["thrust"][i], rel=1e-13, abs=1e-13) and integ == pytest.approx(g["integ1"][i], rel=1e-14, abs=1e-14)
        bxy = flight_np.lateral(veh, X[0:2], X[7:9], g["des"][i, 0], g["des"][i, 1], c)
        np.testing.assert_allclose(bxy, g["bxy"][i], rtol=1e-12, atol=1e-13)
        pq = flight_np.roll_pitch(veh, bxy, R)
        np.testing.assert_allclose(pq, g["pq"][i], rtol=1e-11, atol=1e-12)
        r = flight_np.yaw_rate(veh, X[3:7], g["psi_des"][i], pq[1])
        assert r == pytest.approx(g["r_c"][i], rel=1e-11, abs=1e-12)


def test_inner_loop_stages_match_reference_quad(golden):
    g = golden["stages"]
    veh = flight_np.Vehicle()
    for i in range(len(g["X"])):
        m = flight_np.body_rate(veh, g["X"][i, 10:13], g["pqr_cmd"][i])
        if i % 3 == 0:
            m = m * 0.02
        np.testing.assert_allclose(m, g["moment"][i], rtol=1e-12, atol=1e-14)
        f = flight_np.allocate(veh, g["thrust_cmd"][i], g["moment"][i])
        np.testing.assert_allclose(f, g["forces"][i], rtol=1e-12, atol=1e-13)
        om, cmd = flight_np.motor_lag(veh, g["omega0"][i], f)
        np.testing.assert_allclose(om, g["omega1"][i], rtol=1e-13, atol=0)
        np.testing.assert_allclose(cmd, g["omega_cmd"][i], rtol=1e-13, atol=0)


def test_allocation_closed_forms_of_reference_unit_tests():
    """tests/unit/quadrotor/test_quad.py:72-140: sum f = thrust, moments reproduced, limits respected with sum f kept."""
    veh = flight_np.Vehicle()
    l, k = veh.arm, veh.kappa
    f = flight_np.allocate(veh, 6.0, np.array([0.02, -0.01, 0.004]))
    assert f.sum() == pytest.approx(6.0)
    assert l * (f[0] + f[3] - f[1] - f[2]) == pytest.approx(0.02)
    assert l * (f[0] + f[1] - f[2] - f[3]) == pytest.approx(-0.01)
    assert k * (-f[0] + f[1] - f[2] + f[3]) == pytest.approx(0.004)
    f = flight_np.allocate(veh, 6.0, np.array([5.0, -4.0, 1.0]))                   # far beyond the limits: moments scaled
    assert f.min() >= veh.min_thrust - 1e-12 and f.max() <= veh.max_thrust + 1e-12 and f.sum() == pytest.approx(6.0)
    assert flight_np.allocate(veh, 100.0, np.zeros(3)).sum() == pytest.approx(4 * veh.max_thrust)      # collective clip
    assert flight_np.allocate(veh, -5.0, np.zeros(3)).sum() == pytest.approx(4 * veh.min_thrust)
    # motor lag response (:143-169)
    om, _ = flight_np.motor_lag(veh, np.zeros(4), np.full(4, 1.0))
    np.testing.assert_allclose(om, 1 - math.exp(-0.001 / 0.0125))
    om, _ = flight_np.motor_lag(veh, np.full(4, 2.0), np.full(4, 1.0))
    np.testing.assert_allclose(om, 2.0 - (1 - math.exp(-0.001 / 0.025)))


def test_controller_closed_forms_of_reference_unit_tests():
    """tests/unit/control/test_controller.py:77-183."""
    veh = flight_np.Vehicle()
    c, integ = flight_np.altitude(veh, -1.0, 0.0, np.array([-1.0, 0.0, 0.0]), 1.0, 0.0, 0.01)
    assert c == pytest.approx(veh.mass * veh.g) and integ == 0.0                  # hover thrust = m g at the set-point
    integ = 0.0
    for _ in range(10_000):
        _, integ = flight_np.altitude(veh, 0.0, 0.0, np.array([50.0, 0.0, 0.0]), 1.0, integ, 0.01)
    assert integ == flight_np.INTEGRAL_ERROR_LIMIT                                # clamp at 10
    b = flight_np.lateral(veh, np.zeros(2), np.zeros(2), np.array([100.0, 0, 0]), np.array([-100.0, 0, 0]), veh.mass * veh.g)
    assert np.abs(b).max() <= veh.max_tilt
    m = flight_np.body_rate(veh, np.zeros(3), np.array([0.1, -0.2, 0.3]))
    np.testing.assert_allclose(m, veh.inertia * np.array([veh.kp_p, veh.kp_q, veh.kp_r]) * np.array([0.1, -0.2, 0.3]))
    w = np.array([1.0, 2.0, 3.0])
    np.testing.assert_allclose(flight_np.body_rate(veh, w, w), np.cross(w, veh.inertia * w))           # gyroscopic term only


# ------------------------------------------------------------------------------------------ rigid body (MuJoCo boundary)
def test_hover_invariance_and_gravity_sign():
    """Reference expectations at the MuJoCo boundary (tests/unit/simulation/test_mujoco_sim.py:150-174):
    hover at omega = sqrt(m g / 4 kf) keeps position and velocity to 1e-6 over 100 steps; rotors off => falls (NED +z)."""
    veh = flight_np.Vehicle()
    X = np.zeros(13); X[0:3] = (1.0, 7.0, -1.0); X[3] = 1.0
    om = np.full(4, math.sqrt(veh.mass * veh.g / (4 * veh.kf)))
    Y = X.copy()
    for _ in range(100):
        Y = freebody.freebody_step(Y, om, freebody.quat_to_rot(Y[3:7]), g=veh.g, dt=veh.dt, mass=veh.mass, inertia=veh.inertia,
                                   kf=veh.kf, arm=veh.arm, kappa=veh.kappa)
    assert np.abs(Y - X).max() < 1e-6
    Z = freebody.freebody_step(X, np.zeros(4), np.eye(3), g=veh.g, dt=veh.dt, mass=veh.mass, inertia=veh.inertia, kf=veh.kf,
                               arm=veh.arm, kappa=veh.kappa)
    assert Z[9] == pytest.approx(veh.g * veh.dt) and Z[2] > X[2]


def test_rotor_torque_signs_match_reference_mixer():
    th, tau = freebody.rotor_wrench(np.sqrt([1.0, 2.0, 3.0, 4.0]), 1.0, 0.120208, 0.016)
    assert th == pytest.approx(10.0)
    np.testing.assert_allclose(tau, [0.120208 * (1 + 4 - 2 - 3), 0.120208 * (1 + 2 - 3 - 4), 0.016 * (-1 + 2 - 3 + 4)])


# ------------------------------------------------------------------------------------------ closed loop
def _compare_closed_loop(ref, out, n_rows):
    assert len(out["errors"]) == n_rows
    np.testing.assert_allclose(out["X"], ref["X"][-1], rtol=0, atol=1e-9)
    np.testing.assert_allclose(out["omega"], ref["omega"][-1], rtol=0, atol=1e-9)
    assert out["integral"] == pytest.approx(float(ref["integral"][-1]), abs=1e-10)
    np.testing.assert_allclose(out["errors"], ref["errors"], rtol=0, atol=1e-9)
    assert out["collision"] == bool(ref["collision"]) and out["first_collision_tick"] == int(ref["first_collision_tick"])
    for k in ("final_dist", "mean_err", "rmse", "max_err"):
        assert out[k] == pytest.approx(float(ref[k]), abs=1e-9)


@pytest.mark.parametrize("v", [2, 3])
def test_closed_loop_matches_reference_objects(golden, v):
    """The oracle loop reproduces TrajectoryController + CascadedController + Quad driven tick by tick
    (golden generator) to 1e-9 over the whole lab_course mission, and meets the reference integration
    thresholds (tests/integration/test_mujoco_trajectory_tracking.py:34-36)."""
    g, ref = golden["planning"], golden[f"closed_loop_v{v}"]
    tab = g[f"v{v}_table"]
    out = flight_np.closed_loop(flight_np.Vehicle(), tab, g["waypoints"][0], obstacles=g["obstacles"], goal=GOAL, log_stride=10)
    _compare_closed_loop(ref, out, len(tab))
    np.testing.assert_allclose(out["log"], ref["X"], rtol=0, atol=1e-9)
    assert out["final_dist"] < 0.5 and out["mean_err"] < 0.5 and not out["collision"]
    if v == 2:                                                                    # SURVEY 8(c) anchors
        assert out["final_dist"] == pytest.approx(0.01371, abs=2e-5) and out["mean_err"] == pytest.approx(0.02673, abs=2e-5)


def test_closed_loop_variants_match_reference_objects(golden):
    g, var = golden["planning"], golden["closed_loop_variants"]
    tab = g["v3_table"]

    def sub(prefix):
        return {k[len(prefix):]: var[k] for k in var.files if k.startswith(prefix)}

    ref = sub("nolag_")
    _compare_closed_loop(ref, flight_np.closed_loop(flight_np.Vehicle(), tab, g["waypoints"][0], obstacles=g["obstacles"], goal=GOAL,
                                                    thrust_frame_lag=0), len(tab))
    ref = sub("mc1_")
    veh = flight_np.Vehicle().perturbed(ref["gain_scale"], float(ref["mass_scale"]), ref["inertia_scale"])
    _compare_closed_loop(ref, flight_np.closed_loop(veh, tab, g["waypoints"][0], obstacles=g["obstacles"], goal=GOAL), len(tab))
    ref = sub("wind_")
    _compare_closed_loop(ref, flight_np.closed_loop(flight_np.Vehicle(), tab, g["waypoints"][0], obstacles=g["obstacles"], goal=GOAL,
                                                    wind=ref["force"]), len(tab))
    ref = sub("hit_")
    out = flight_np.closed_loop(flight_np.Vehicle(), tab, g["waypoints"][0], obstacles=ref["obstacles"], goal=GOAL)
    _compare_closed_loop(ref, out, len(tab))
    assert out["collision"] and out["first_collision_tick"] == 4086


def test_actual_trajectory_list_matches_the_reference_recorder(golden):
    """D5: the viewer's flown-path list.  tests/golden/actual_trajectory.npz was recorded by the reference's OWN
    MujocoSimulation._record_actual_trajectory (mujoco_sim.py:201-218) called after every tick of the reference objects' closed loop;
    the oracle's restatement of its two tests (take-off gate on z, 0.05 s of accumulated data.time) must pick the same ticks."""
    g, ref = golden["planning"], golden["actual_trajectory"]
    out = flight_np.closed_loop(flight_np.Vehicle(), g["v3_table"], g["waypoints"][0], goal=GOAL, traj_gate_z=float(ref["takeoff_z"]),
                                traj_interval=float(ref["interval"]))
    np.testing.assert_array_equal(out["traj_ticks"], ref["ticks"])
    np.testing.assert_allclose(out["traj"], ref["positions"], rtol=0, atol=1e-9)
    assert set(np.diff(ref["ticks"]).tolist()) >= {50, 51}                        # accumulated fp64 time: not a fixed 50-tick stride
    assert ref["ticks"][0] > 400 and (np.diff(ref["ticks"]) > 51).any()            # gated during take-off and again where the course dips below it


def test_ground_floor_switch(golden):
    """SURVEY 7.3's documented choice as a switch: ground_z puts a unilateral floor under the free body.  Off (default) the drone
    sags through its start height while the rotors spin up (<= 1.5 cm); on, it never goes below it and the rest of
    the flight is unchanged to a few millimetres.  NumPy and C oracles agree."""
    from oracle import c_port
    g = golden["planning"]
    tab, start = g["v3_table"][:120], g["waypoints"][0]
    free = flight_np.closed_loop(flight_np.Vehicle(), tab, start, goal=GOAL, log_stride=1)
    held = flight_np.closed_loop(flight_np.Vehicle(), tab, start, goal=GOAL, log_stride=1, ground_z=float(start[2]))
    sag = float((free["log"][:, 2] - start[2]).max())
    assert 0.002 < sag < 0.016
    assert float((held["log"][:, 2] - start[2]).max()) <= 0.0
    assert np.abs(held["log"][-1, :3] - free["log"][-1, :3]).max() < 0.02
    c = c_port.closed_loop(flight_np.Vehicle(), tab, start, goal=GOAL, ground_z=float(start[2]))
    np.testing.assert_allclose(c["X"], held["X"], rtol=0, atol=1e-10)
