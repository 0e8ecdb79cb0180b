"""K2 parity through the C ABI: the persistent fp32 rollout kernel against closed-loop logs produced
by the reference's own TrajectoryController / CascadedController / Quad objects (tests/golden/
closed_loop_*.npz; physics = oracle/freebody.py, parity unpinned at the MuJoCo boundary) and against
the NumPy oracle on seeded Monte-Carlo inputs.

Tolerances (BASELINE.json north_star): closed-loop states 1e-4 m / 1e-4 rad over the lab_course
mission with fp32 rollouts; collision flags bit-exact.  Velocities and body rates are held to
1e-3 m/s and 1e-3 rad/s (not named by north_star; measured ~3e-6 and ~3e-5)."""
import numpy as np
import pytest

from helpers import GOAL, lab_course_plan, mc_arrays, rotation_angle

pytestmark = pytest.mark.gpu
POS_TOL, ANG_TOL, VEL_TOL, RATE_TOL = 1e-4, 1e-4, 1e-3, 1e-3


def _fly(dev, plan, B, n_ticks, dtype=None, **kw):
    import torch
    from uav_ac_b200 import kernels
    from uav_ac_b200.simulation.scene import LAB_COURSE_START
    start = kw.pop("start", None)
    if start is None:
        start = torch.tensor(LAB_COURSE_START, dtype=torch.float64, device=dev)
    goal = kw.pop("goal", torch.tensor(GOAL, dtype=torch.float64, device=dev))
    res = kernels.rollout(plan, B, n_ticks, start=start, goal=goal, dtype=dtype or torch.float32, **kw)
    torch.cuda.synchronize()
    return res


def _check_log(log, gold, pos_tol=POS_TOL, ang_tol=ANG_TOL):
    """log [n, 13, B] -> compare rollout 0 with the golden per-period states [n, 13]."""
    X = log[:, :, 0].double().cpu().numpy()
    ref = gold["X"]
    assert X.shape == ref.shape
    dp = np.abs(X[:, 0:3] - ref[:, 0:3]).max()
    da = rotation_angle(X[:, 3:7], ref[:, 3:7]).max()
    dv = np.abs(X[:, 7:10] - ref[:, 7:10]).max()
    dw = np.abs(X[:, 10:13] - ref[:, 10:13]).max()
    print(f"max |dpos| {dp:.2e} m, attitude {da:.2e} rad, |dvel| {dv:.2e} m/s, |drate| {dw:.2e} rad/s")
    assert dp < pos_tol and da < ang_tol and dv < VEL_TOL and dw < RATE_TOL
    return dp, da


def _check_metrics(m, gold, prefix="", tol=2e-5):
    m = m.double().cpu().numpy()
    g = lambda k: float(gold[prefix + k])
    assert abs(m[0] - g("final_dist")) < POS_TOL
    assert m[1] == float(bool(gold[prefix + "collision"]))                  # bit-exact flag
    assert abs(m[2] - g("rmse")) < tol and abs(m[3] - g("mean_err")) < tol and abs(m[4] - g("max_err")) < POS_TOL
    assert m[5] == 0 and m[6] == float(gold[prefix + "first_collision_tick"])
    assert m[7] == len(gold[prefix + "errors"])


@pytest.mark.parametrize("v", [2, 3])
def test_lab_course_fp32_rollout_matches_reference_closed_loop(cuda, golden, v):
    import torch
    from uav_ac_b200.simulation.scene import LAB_COURSE_OBSTACLES
    gold = golden[f"closed_loop_v{v}"]
    plan = lab_course_plan(cuda, float(v))
    n_rows = int(plan.total_rows.item())
    assert n_rows == len(gold["X"])                                          # 1613 / 1076 table rows
    obs = torch.tensor(LAB_COURSE_OBSTACLES, dtype=torch.float32, device=cuda)
    res = _fly(cuda, plan, 1, 10 * n_rows, obstacles=obs, log_stride=10)
    _check_log(res.log, gold)
    _check_metrics(res.metrics[0], gold)
    # the reference integration thresholds (tests/integration/test_mujoco_trajectory_tracking.py:34-36)
    m = res.metrics[0].cpu().numpy()
    assert m[0] < 0.5 and m[3] < 0.5 and m[1] == 0.0
    # first 2000 ticks at full rate, incl. rotor speeds' effect on the state
    fine = _fly(cuda, plan, 1, 2000, log_stride=1).log[:, :, 0].double().cpu().numpy()
    ref = gold["fine"][:, :13]
    assert np.abs(fine[:, :3] - ref[:, :3]).max() < POS_TOL and rotation_angle(fine[:, 3:7], ref[:, 3:7]).max() < ANG_TOL


@pytest.mark.parametrize("v", [2, 3])
def test_lab_course_fp64_rollout_matches_reference_closed_loop(cuda, golden, v):
    import torch
    gold = golden[f"closed_loop_v{v}"]
    plan = lab_course_plan(cuda, float(v))
    res = _fly(cuda, plan, 1, 10 * len(gold["X"]), dtype=torch.float64, log_stride=10)
    dp, da = _check_log(res.log, gold, 1e-6, 1e-6)
    _check_metrics(res.metrics[0], gold, tol=1e-7)


def test_variants_thrust_frame_wind_montecarlo_and_collision(cuda, golden):
    import torch
    from uav_ac_b200.simulation.scene import LAB_COURSE_OBSTACLES
    var = golden["closed_loop_variants"]
    plan = lab_course_plan(cuda, 3.0)
    n = 10 * int(plan.total_rows.item())
    obs = torch.tensor(LAB_COURSE_OBSTACLES, dtype=torch.float32, device=cuda)

    def sub(prefix):
        return {k[len(prefix):]: var[k] for k in var.files if k.startswith(prefix)}

    res = _fly(cuda, plan, 1, n, obstacles=obs, log_stride=10, thrust_frame_lag=0)     # viewer path: fresh thrust frame
    _check_log(res.log, sub("nolag_")); _check_metrics(res.metrics[0], sub("nolag_"))
    for j in range(3):                                                                 # Monte-Carlo gains / mass / inertia
        g = sub(f"mc{j}_")
        mc = mc_arrays(cuda, 1, g["gain_scale"][None], [float(g["mass_scale"])], g["inertia_scale"][None])
        res = _fly(cuda, plan, 1, n, obstacles=obs, log_stride=10, **mc)
        _check_log(res.log, g); _check_metrics(res.metrics[0], g)
    g = sub("wind_")
    res = _fly(cuda, plan, 1, n, obstacles=obs, log_stride=10, **mc_arrays(cuda, 1, wind=g["force"][None]))
    _check_log(res.log, g); _check_metrics(res.metrics[0], g)
    g = sub("hit_")                                                                    # an AABB on the course: flag and first tick exact
    res = _fly(cuda, plan, 1, n, obstacles=torch.tensor(g["obstacles"], dtype=torch.float32, device=cuda), log_stride=10)
    _check_log(res.log, g); _check_metrics(res.metrics[0], g)
    assert res.metrics[0, 1].item() == 1.0 and res.metrics[0, 6].item() == 4086.0


def test_batch_is_deterministic_and_position_independent(cuda, golden):
    """The same rollout at every thread position gives bit-identical results; a batch mixing the golden
    Monte-Carlo variants reproduces each of them wherever it sits in the grid."""
    import torch
    var = golden["closed_loop_variants"]
    plan = lab_course_plan(cuda, 3.0)
    n = 10 * int(plan.total_rows.item())
    B = 1000
    res = _fly(cuda, plan, B, n)
    assert bool((res.state == res.state[:, :1]).all()) and bool((res.metrics == res.metrics[:1]).all())
    gs, ms, ins = np.ones((B, 11)), np.ones(B), np.ones((B, 3))
    where = {0: 3, 1: 517, 2: 999}
    for j, b in where.items():
        gs[b], ms[b], ins[b] = var[f"mc{j}_gain_scale"], float(var[f"mc{j}_mass_scale"]), var[f"mc{j}_inertia_scale"]
    res2 = _fly(cuda, plan, B, n, **mc_arrays(cuda, B, gs, ms, ins))
    for j, b in where.items():
        Xf = res2.state[:, b].double().cpu().numpy()
        ref = var[f"mc{j}_X"][-1]
        assert np.abs(Xf[:3] - ref[:3]).max() < POS_TOL and rotation_angle(Xf[3:7], ref[3:7]) < ANG_TOL
        assert abs(res2.metrics[b, 0].item() - float(var[f"mc{j}_final_dist"])) < POS_TOL
    # untouched slots equal the nominal flight (fp32 defaults are the rounded fp64 defaults)
    assert np.abs((res2.state[:, 0] - res.state[:, 0]).cpu().numpy()).max() < 1e-5


def test_chunked_launches_resume_bit_exactly(cuda):
    import torch
    plan = lab_course_plan(cuda, 3.0)
    n = 10 * int(plan.total_rows.item())
    B = 256
    rng = np.random.default_rng(3)
    mc = mc_arrays(cuda, B, rng.uniform(0.8, 1.2, (B, 11)), rng.uniform(0.9, 1.1, B), rng.uniform(0.9, 1.1, (B, 3)))
    whole = _fly(cuda, plan, B, n, **mc)
    first = _fly(cuda, plan, B, 4003, want_carry=True, **mc)                    # split in the middle of an outer period
    second = _fly(cuda, plan, B, n - 4003, carry=first.carry, resume=True, **mc)
    assert torch.equal(whole.state, second.state)
    assert torch.equal(whole.metrics, second.metrics)


def test_holds_last_row_after_the_table_ends(cuda, golden):
    """index = min(index+1, N-1) (main.py:61): flying past the table keeps tracking the last row."""
    plan = lab_course_plan(cuda, 3.0)
    n = 10 * int(plan.total_rows.item())
    res = _fly(cuda, plan, 1, n + 5000)
    m = res.metrics[0].cpu().numpy()
    assert m[7] == (n + 5000) // 10 and m[5] == 0
    assert m[0] < 0.02                                                          # settles on the goal while holding the last row


def test_montecarlo_batch_matches_numpy_oracle_on_a_sample(cuda, golden):
    """BASELINE configs[2] at reduced size: per-rollout perturbed gains / mass / inertia from the
    counter-based generator; a strided sample is re-flown by the NumPy oracle with the same fp32 inputs."""
    import torch
    from oracle import flight_np
    from uav_ac_b200 import kernels, _native as nat
    g = golden["planning"]
    plan = lab_course_plan(cuda, 3.0)
    n = 10 * int(plan.total_rows.item())
    B = 20000
    sc = kernels.mc_uniform(11, B, [0.8] * 11 + [0.9] * 4, [1.2] * 11 + [1.1] * 4)            # [15, B] multiplicative scales
    v = nat.default_vehicle()
    base = torch.tensor(list(v.gains) + [v.mass] + list(v.inertia), dtype=torch.float32, device=cuda)[:, None]
    vals = (sc * base).contiguous()
    obs = torch.tensor(g["obstacles"], dtype=torch.float32, device=cuda)
    res = _fly(cuda, plan, B, n, obstacles=obs, mc_gains=vals[:11].contiguous(), mc_mass=vals[11].contiguous(),
               mc_inertia=vals[12:15].contiguous())
    met = res.metrics.cpu().numpy()
    assert np.isfinite(met).all() and (met[:, 5] == 0).all()
    vals = vals.double().cpu().numpy()
    tab = g["v3_table"]
    worst = 0.0
    for b in range(0, B, B // 6):
        veh = flight_np.Vehicle()
        kw = {k: getattr(veh, k) for k in veh.__dataclass_fields__}
        for k, name in enumerate(flight_np.Vehicle.GAIN_NAMES):
            kw[name] = vals[k, b]
        kw["mass"], kw["inertia"] = vals[11, b], vals[12:15, b]
        ref = flight_np.closed_loop(flight_np.Vehicle(**kw), tab, g["waypoints"][0], obstacles=g["obstacles"], goal=GOAL)
        Xf = res.state[:, b].double().cpu().numpy()
        dp = np.abs(Xf[:3] - ref["X"][:3]).max()
        worst = max(worst, dp)
        assert dp < POS_TOL and rotation_angle(Xf[3:7], ref["X"][3:7]) < ANG_TOL
        assert met[b, 1] == float(ref["collision"])
        assert abs(met[b, 3] - ref["mean_err"]) < 2e-5 and abs(met[b, 0] - ref["final_dist"]) < POS_TOL
    print(f"worst final-position difference over the sample: {worst:.2e} m")


def test_random_missions_wind_and_obstacle_sets_match_numpy_oracle(cuda):
    """BASELINE configs[3] at reduced size: per-rollout waypoint sets (vertical take-off + 4-spline course),
    constant wind, per-rollout AABB sets; collision flags exact outside the 1e-4 m ambiguity band."""
    import torch
    from oracle import flight_np, minsnap_np
    from uav_ac_b200 import kernels
    B, S = 4096, 4
    wp, vel = kernels.mc_missions(21, B, S)
    ground = wp[:, 0].clone()
    ground[:, 2] = -0.021
    tk = torch.stack((ground, wp[:, 0]), dim=1).contiguous()                  # vertical take-off to the first waypoint
    plan = kernels.plan_missions([(tk, vel), (wp, vel)], 0.01)
    wind = kernels.mc_uniform(22, B, [-0.08] * 3, [0.08] * 3)
    # 8 obstacle sets of 5 boxes scattered in the flight volume
    rng = np.random.default_rng(8)
    ctr = rng.uniform([2, 2, -5], [22, 12, -1], (8, 5, 3))
    half = rng.uniform(0.3, 1.2, (8, 5, 3))
    boxes = np.stack((ctr[..., 0] - half[..., 0], ctr[..., 0] + half[..., 0], ctr[..., 1] - half[..., 1], ctr[..., 1] + half[..., 1],
                      ctr[..., 2] - half[..., 2], ctr[..., 2] + half[..., 2]), axis=-1).astype(np.float32)
    sets = torch.tensor(rng.integers(0, 8, B), dtype=torch.int32, device=cuda)
    n_ticks = int(plan.total_rows.max().item()) * 10
    res = _fly(cuda, plan, B, n_ticks, start=ground.contiguous(), goal=wp[:, -1].contiguous(), mc_wind=wind,
               obstacles=torch.tensor(boxes, device=cuda), obstacle_set=sets)
    met = res.metrics.cpu().numpy()
    assert np.isfinite(met[:, :5]).all()
    assert 0.02 < met[:, 1].mean() < 0.98                                      # both outcomes are exercised
    wpn, veln, windn, setn = wp.cpu().numpy(), vel.cpu().numpy(), wind.double().cpu().numpy(), sets.cpu().numpy()
    checked = ambiguous = 0
    for b in range(0, B, B // 10):
        tab = np.vstack([minsnap_np.sample_table(*minsnap_np.solve_coeffs(w, veln[b], "solve"), 0.01)
                         for w in (np.stack((ground[b].cpu().numpy(), wpn[b, 0])), wpn[b])])
        veh = flight_np.Vehicle()
        ob = boxes[setn[b]].astype(float)
        ref = flight_np.closed_loop(veh, tab, ground[b].cpu().numpy(), obstacles=ob, goal=wpn[b, -1], wind=windn[:, b], n_ticks=n_ticks,
                                    log_stride=1)
        if met[b, 5] != 0 or ref["max_err"] > 5.0:
            continue                                                          # diverged flights amplify rounding: not comparable
        Xf = res.state[:, b].double().cpu().numpy()
        assert np.abs(Xf[:3] - ref["X"][:3]).max() < POS_TOL and rotation_angle(Xf[3:7], ref["X"][3:7]) < ANG_TOL
        # signed distance of the fp64 path to the nearest box face decides whether the flag is ambiguous
        P = ref["log"][:, :3]
        gap = np.inf
        for q in ob:
            d = np.maximum.reduce([q[0] - P[:, 0], P[:, 0] - q[1], q[2] - P[:, 1], P[:, 1] - q[3], q[4] - P[:, 2], P[:, 2] - q[5]])
            gap = min(gap, np.abs(d).min())
        if gap < POS_TOL:
            ambiguous += 1
            continue
        assert met[b, 1] == float(ref["collision"]) and met[b, 6] == ref["first_collision_tick"]
        checked += 1
    print(f"collision flags: {checked} exact, {ambiguous} inside the {POS_TOL} m ambiguity band")
    assert checked >= 5
    # the library flew these per-rollout missions one drone per thread; the two-drones-per-thread kernel (pair_kernel_only) flies the
    # same batch to the same answers (other bits: another formulation of the quaternion map), flags included
    pair = _fly(cuda, plan, B, n_ticks, start=ground.contiguous(), goal=wp[:, -1].contiguous(), mc_wind=wind,
                obstacles=torch.tensor(boxes, device=cuda), obstacle_set=sets, pair_kernel_only=True)
    sane = (res.metrics[:, 5] == 0) & (res.metrics[:, 4] < 5.0) & (pair.metrics[:, 5] == 0)
    dp = (pair.state[:3] - res.state[:3]).abs().max(dim=0).values[sane]
    print(f"pair kernel vs one-drone-per-thread kernel: max |dpos| {float(dp.max()):.2e} m over {int(sane.sum())} sane rollouts")
    assert float(dp.max()) < POS_TOL
    assert float((pair.metrics[sane, 1] != res.metrics[sane, 1]).float().mean()) < 2e-3      # a flag may flip only within rounding of a box face


def test_state_log_layout_and_stride(cuda):
    import torch
    plan = lab_course_plan(cuda, 3.0)
    B, n = 96, 3000
    a = _fly(cuda, plan, B, n, log_stride=1)
    b = _fly(cuda, plan, B, n, log_stride=50)
    assert a.log.shape == (n, 13, B) and b.log.shape == (n // 50, 13, B)
    assert torch.equal(a.log[49::50], b.log)                                   # sample s is the state after tick (s+1)*stride
    assert torch.equal(a.log[-1], a.state)


def test_argument_errors_are_reported_not_ignored(cuda):
    import torch
    from uav_ac_b200 import kernels, _native as nat
    plan = lab_course_plan(cuda, 3.0)
    with pytest.raises(nat.UavbError):
        kernels.rollout(plan, 4, 100, start=torch.zeros(3, dtype=torch.float64))            # host tensor: no CPU path
    with pytest.raises(nat.UavbError):
        kernels.rollout(plan, 4, 100, start=torch.zeros(3, dtype=torch.float64, device=cuda), frequency=0)
    with pytest.raises(ValueError):
        kernels.rollout(plan, 4, 100, start=torch.zeros(3, dtype=torch.float64, device=cuda), mc_mass=torch.ones(3, device=cuda))


def test_host_buffer_entry_point_matches_device_path_and_reference(cuda, golden):
    """uavb_fly_mission_host (the call a reference-side binding makes): NumPy arrays in, metrics out; same result as the
    device-pointer path bit for bit, and the reference closed-loop metrics within the stated tolerances."""
    import torch
    from uav_ac_b200 import host_api
    from uav_ac_b200.simulation.scene import LAB_COURSE_OBSTACLES, LAB_COURSE_WAYPOINTS
    gold = golden["closed_loop_v3"]
    var = golden["closed_loop_variants"]
    B = 300
    gs, ms, ins = np.ones((B, 11)), np.ones(B), np.ones((B, 3))
    gs[7], ms[7], ins[7] = var["mc0_gain_scale"], float(var["mc0_mass_scale"]), var["mc0_inertia_scale"]
    mc = mc_arrays(cuda, B, gs, ms, ins)
    met, state, n_ticks = host_api.fly_mission_host(LAB_COURSE_WAYPOINTS, 3.0, B, obstacles=LAB_COURSE_OBSTACLES, want_state=True,
                                                    mc_mass=mc["mc_mass"].cpu().numpy(), mc_inertia=mc["mc_inertia"].cpu().numpy(),
                                                    mc_gains=mc["mc_gains"].cpu().numpy())
    assert n_ticks == 10 * len(gold["X"]) and met.shape == (B, 8) and state.shape == (13, B)
    plan = lab_course_plan(cuda, 3.0)
    res = _fly(cuda, plan, B, n_ticks, obstacles=torch.tensor(LAB_COURSE_OBSTACLES, dtype=torch.float32, device=cuda), **mc)
    assert np.array_equal(met, res.metrics.cpu().numpy()) and np.array_equal(state, res.state.cpu().numpy())
    _check_metrics(torch.tensor(met[0]), gold)
    assert abs(met[7, 0] - float(var["mc0_final_dist"])) < POS_TOL and abs(met[7, 3] - float(var["mc0_mean_err"])) < 2e-5
    # single table (no take-off split) and a bounded number of ticks
    met1, _, n1 = host_api.fly_mission_host(LAB_COURSE_WAYPOINTS[1:], 3.0, 2, n_takeoff_waypoints=0, n_ticks=500)
    assert n1 == 500 and met1[0, 7] == 50 and np.isfinite(met1).all()
    with pytest.raises(Exception):
        host_api.fly_mission_host(np.array([[0.0, 0, 0], [0, 0, 0], [1, 0, 0]]), 3.0, 2, n_takeoff_waypoints=2)   # zero-length take-off


def test_montecarlo_batch_matches_c_oracle_on_every_rollout(cuda, golden):
    """BASELINE configs[2] at reduced size with EVERY rollout checked: 2 048 perturbed vehicles (counter-based generator)
    fly the whole lab_course mission on the GPU and in the C twin of the oracle (fp64, same fp32-rounded inputs)."""
    import os
    import torch
    from oracle import c_port, flight_np
    from uav_ac_b200 import kernels, _native as nat
    g = golden["planning"]
    plan = lab_course_plan(cuda, 3.0)
    n = 10 * int(plan.total_rows.item())
    B = 2048
    sc = kernels.mc_uniform(77, B, [0.8] * 11 + [0.9] * 4, [1.2] * 11 + [1.1] * 4)
    v = nat.default_vehicle()
    base = torch.tensor(list(v.gains) + [v.mass] + list(v.inertia), dtype=torch.float32, device=cuda)[:, None]
    vals = (sc * base).contiguous()
    obs = torch.tensor(g["obstacles"], dtype=torch.float32, device=cuda)
    res = _fly(cuda, plan, B, n, obstacles=obs, mc_gains=vals[:11].contiguous(), mc_mass=vals[11].contiguous(), mc_inertia=vals[12:15].contiguous())
    met, X = res.metrics.double().cpu().numpy(), res.state.double().cpu().numpy().T
    vals = vals.double().cpu().numpy()
    vehs = []
    for b in range(B):
        kw = {k: getattr(flight_np.Vehicle(), k) for k in flight_np.Vehicle.__dataclass_fields__}
        for k, name in enumerate(flight_np.Vehicle.GAIN_NAMES):
            kw[name] = vals[k, b]
        kw["mass"], kw["inertia"] = vals[11, b], vals[12:15, b]
        vehs.append(flight_np.Vehicle(**kw))
    tab = c_port.mission_table(g["waypoints"], 3.0, 0.01)                      # solve-branch coefficients, like K1
    m_ref, X_ref = c_port.closed_loop_batch(vehs, tab, g["waypoints"][0], obstacles=g["obstacles"], goal=GOAL, threads=os.cpu_count() or 1)
    dp = np.abs(X[:, :3] - X_ref[:, :3]).max(axis=1)
    da = rotation_angle(X[:, 3:7], X_ref[:, 3:7])
    print(f"2048 rollouts: max |dpos| {dp.max():.2e} m (median {np.median(dp):.2e}), max attitude {da.max():.2e} rad")
    assert dp.max() < POS_TOL and da.max() < ANG_TOL
    assert np.array_equal(met[:, 1], m_ref[:, 1])                              # collision flags bit-exact
    assert np.abs(met[:, 0] - m_ref[:, 0]).max() < POS_TOL and np.abs(met[:, 3] - m_ref[:, 3]).max() < 2e-5
    assert np.abs(met[:, 2] - m_ref[:, 2]).max() < 2e-5 and np.abs(met[:, 4] - m_ref[:, 4]).max() < POS_TOL
    assert (met[:, 7] == m_ref[:, 7]).all() and (met[:, 5] == 0).all()


def test_default_slicing_of_a_large_batch_is_bit_identical_to_one_slice(cuda):
    """A batch with more work groups than resident CTAs is cut into slices of DECREASING length by default, and a pair's later slices
    load the per-rollout constants its first slice cached (rollout_kernels.cu, RolloutDev::vehp_cache): neither may change a bit
    against the same batch flown as one slice or as equal slices -- with Monte-Carlo overrides, an odd batch size, boxes, and a
    tick count that is not a multiple of the outer period."""
    import torch
    from uav_ac_b200.simulation.scene import LAB_COURSE_OBSTACLES
    plan = lab_course_plan(cuda, 3.0)
    B, n = 90_001, 2507                                              # 1 407 work groups of 64 drones > 148 x 8 resident CTAs
    rng = np.random.default_rng(21)
    mc = mc_arrays(cuda, B, rng.uniform(0.8, 1.2, (B, 11)), rng.uniform(0.9, 1.1, B), rng.uniform(0.9, 1.1, (B, 3)))
    obs = torch.tensor(LAB_COURSE_OBSTACLES, dtype=torch.float32, device=cuda)
    runs = [_fly(cuda, plan, B, n, n_slices=k, obstacles=obs, want_carry=True, **mc) for k in (0, 1, 6)]
    for r in runs[1:]:
        assert torch.equal(r.metrics, runs[0].metrics) and torch.equal(r.state, runs[0].state)
        assert torch.equal(r.carry.view(torch.int32)[:50], runs[0].carry.view(torch.int32)[:50])
    assert bool(torch.isfinite(runs[0].state).all())


def test_default_slicing_of_per_rollout_missions_is_bit_identical_to_one_slice(cuda):
    """The same for the one-drone-per-thread kernel that flies per-rollout missions: 80 001 rollouts are more work groups (of 32) than
    the 148 x 16 resident CTAs, so the default launch cuts slices of decreasing length; one slice and four equal slices must give
    the same bits."""
    import torch
    from uav_ac_b200 import kernels
    B, n = 80_001, 1803
    wp, vel = kernels.mc_missions(17, B, 4)
    plan = kernels.plan_missions([(wp, vel)], 0.01)
    wind = kernels.mc_uniform(5, B, [-0.08] * 3, [0.08] * 3)
    runs = [kernels.rollout(plan, B, n, start=wp[:, 0].contiguous(), goal=wp[:, -1].contiguous(), mc_wind=wind, n_slices=k) for k in (0, 1, 4)]
    torch.cuda.synchronize()
    for r in runs[1:]:
        assert torch.equal(r.metrics, runs[0].metrics) and torch.equal(r.state, runs[0].state)
    assert bool(torch.isfinite(runs[0].state).all())


def test_time_sliced_schedule_is_bit_identical_to_a_single_slice(cuda, monkeypatch):
    """The persistent work queue ((slice, group) items, state parked in the carry block between slices) must not change
    a single bit: 5 000 per-rollout missions with wind and obstacle sets flown as one slice, as 7 slices and as 60 slices."""
    import torch
    from uav_ac_b200 import kernels
    B = 5000
    wp, vel = kernels.mc_missions(3, B, 4)
    ground = wp[:, 0].clone()
    ground[:, 2] = -0.021
    plan = kernels.plan_missions([(torch.stack((ground, wp[:, 0]), dim=1).contiguous(), vel), (wp, vel)], 0.01)
    wind = kernels.mc_uniform(4, B, [-0.08] * 3, [0.08] * 3)
    rng = np.random.default_rng(1)
    ctr, half = rng.uniform([2, 2, -5], [22, 12, -1], (4, 5, 3)), rng.uniform(0.3, 1.2, (4, 5, 3))
    boxes = torch.tensor(np.stack((ctr[..., 0] - half[..., 0], ctr[..., 0] + half[..., 0], ctr[..., 1] - half[..., 1], ctr[..., 1] + half[..., 1],
                                   ctr[..., 2] - half[..., 2], ctr[..., 2] + half[..., 2]), axis=-1).astype(np.float32), device=cuda)
    sets = (torch.arange(B, device=cuda) % 4).to(torch.int32)
    n = 6000 + 7                                                     # not a multiple of the outer period
    runs = []
    for chunks in (1, 7, 60):
        r = kernels.rollout(plan, B, n, start=ground.contiguous(), goal=wp[:, -1].contiguous(), mc_wind=wind, obstacles=boxes, obstacle_set=sets,
                            want_carry=True, n_slices=chunks)
        torch.cuda.synchronize()
        runs.append(r)
    for r in runs[1:]:
        assert torch.equal(r.metrics, runs[0].metrics) and torch.equal(r.state, runs[0].state) and torch.equal(r.carry.view(torch.int32)[:50], runs[0].carry.view(torch.int32)[:50])
    assert float(runs[0].metrics[:, 1].mean()) > 0.02                # collisions happen, so first-hit ticks cross slice boundaries


def test_precomputed_set_point_table_is_bit_identical_to_on_the_fly_evaluation(cuda):
    """Shared missions read one 56-byte row per outer period (uavb_rollout_targets_f64) instead of evaluating the polynomials per
    drone; both forms must give the same bits, also past the end of the table (hold-last-row) and with an over-long table."""
    import torch
    from uav_ac_b200 import kernels
    from uav_ac_b200.simulation.scene import LAB_COURSE_WAYPOINTS as W
    for v in (2.0, 3.0):
        plan = lab_course_plan(cuda, v)
        n = 10 * int(plan.total_rows.item()) + 3000
        B = 300
        rng = np.random.default_rng(int(v))
        mc = mc_arrays(cuda, B, rng.uniform(0.8, 1.2, (B, 11)), rng.uniform(0.9, 1.1, B), rng.uniform(0.9, 1.1, (B, 3)))
        a = _fly(cuda, plan, B, n, use_targets=True, **mc)
        b = _fly(cuda, plan, B, n, use_targets=False, **mc)
        assert plan.targets.shape == (int(plan.total_rows.item()), 56)
        assert torch.equal(a.metrics, b.metrics) and torch.equal(a.state, b.state)
        wp = torch.tensor(W, dtype=torch.float64, device=cuda)
        vel = torch.tensor([v], dtype=torch.float64, device=cuda)
        longer = kernels.plan_missions([(wp[None, :2].contiguous(), vel), (wp[None, 1:].contiguous(), vel)], 0.01, shared=True,
                                       table_rows=int(plan.total_rows.item()) + 77)
        c = _fly(cuda, longer, B, n, **mc)
        assert torch.equal(c.metrics, b.metrics) and torch.equal(c.state, b.state)


def test_state_log_is_identical_across_slice_counts(cuda, monkeypatch):
    """A state log written slice by slice lands in the same places with the same bits, for strides that do and do not divide
    the slice length, and its last sample is the final state."""
    import torch
    plan = lab_course_plan(cuda, 3.0)
    B, n = 700, 2350
    rng = np.random.default_rng(9)
    mc = mc_arrays(cuda, B, rng.uniform(0.8, 1.2, (B, 11)), rng.uniform(0.9, 1.1, B), rng.uniform(0.9, 1.1, (B, 3)))
    for stride in (1, 7, 50):
        logs = []
        for chunks in (1, 9):
            r = _fly(cuda, plan, B, n, log_stride=stride, n_slices=chunks, **mc)
            logs.append(r)
        assert logs[0].log.shape == (n // stride, 13, B)
        assert torch.equal(logs[0].log, logs[1].log) and torch.equal(logs[0].state, logs[1].state) and torch.equal(logs[0].metrics, logs[1].metrics)
        if n % stride == 0:
            assert torch.equal(logs[0].log[-1], logs[0].state)


def test_state_log_through_tensor_stores_equals_per_thread_stores(cuda):
    """The staged TMA log (uavb_rollout_args.log_tma = 0, batches that are a multiple of 4) against the per-thread streaming stores
    (log_tma = -1): same samples, same places, same bits -- for whole and ragged last warps, strides that do and do not divide
    the staging depth or the slice length, odd sample counts, and several slices; a batch that is not a multiple of 4 takes the
    per-thread path by itself."""
    import torch
    plan = lab_course_plan(cuda, 3.0)
    rng = np.random.default_rng(17)
    for B, n, stride, slices in ((96, 1001, 1, 1), (700, 2350, 7, 9), (1028, 1500, 50, 3), (4, 333, 3, 2), (131072, 400, 1, 0), (98, 500, 1, 1)):
        mc = mc_arrays(cuda, B, rng.uniform(0.8, 1.2, (B, 11)), rng.uniform(0.9, 1.1, B), rng.uniform(0.9, 1.1, (B, 3)))
        a = _fly(cuda, plan, B, n, log_stride=stride, n_slices=slices, **mc)
        b = _fly(cuda, plan, B, n, log_stride=stride, n_slices=slices, log_tma=-1, **mc)
        assert a.log.shape == (n // stride, 13, B)
        assert torch.equal(a.log, b.log) and torch.equal(a.state, b.state) and torch.equal(a.metrics, b.metrics), (B, n, stride, slices)
        assert bool(torch.isfinite(a.log).all())


def test_actual_trajectory_list_matches_the_reference_recorder(cuda, golden):
    """D5 (mujoco_sim.py:201-218) in the persistent rollout: the gated 20 Hz position list of every drone against the list the
    reference's own _record_actual_trajectory produced (tests/golden/actual_trajectory.npz: 191 samples, first after tick 509,
    50 / 51 ticks apart, a 671-tick gap where the course dips below the take-off altitude).  A sample taken one tick off would
    sit ~3 mm away, so count + 1e-4 m pin the ticks.  The list does not depend on the slice count, and recording it leaves
    the flight that of the metrics-only rollout (same numerics, another compiled kernel: equal to a few ulps)."""
    import torch
    ref = golden["actual_trajectory"]
    plan = lab_course_plan(cuda, 3.0)
    n = 10 * int(plan.total_rows.item())
    B = 70
    kw = dict(traj_max_samples=256, traj_gate_z=float(ref["takeoff_z"]), traj_interval=float(ref["interval"]))
    a = _fly(cuda, plan, B, n, **kw)
    assert a.traj_count.tolist() == [len(ref["ticks"])] * B
    got = a.traj[:len(ref["ticks"]), :, 0].double().cpu().numpy()
    assert np.abs(got - ref["positions"]).max() < 1e-4
    assert bool((a.traj == a.traj[:, :, :1]).all()) and float(a.traj[len(ref["ticks"]):].abs().max()) == 0.0
    b = _fly(cuda, plan, B, n, n_slices=9, **kw)
    assert torch.equal(a.traj, b.traj) and torch.equal(a.traj_count, b.traj_count)
    plain = _fly(cuda, plan, B, n)
    assert float((a.state - plain.state).abs().max()) < 1e-5 and float((a.metrics - plain.metrics).abs().max()) < 1e-5
    # a list shorter than the flight keeps its first samples and still counts them all; per-rollout vehicles differ
    rng = np.random.default_rng(3)
    mc = mc_arrays(cuda, B, rng.uniform(0.8, 1.2, (B, 11)), rng.uniform(0.9, 1.1, B), rng.uniform(0.9, 1.1, (B, 3)))
    c = _fly(cuda, plan, B, n, traj_max_samples=16, traj_gate_z=float(ref["takeoff_z"]), **mc)
    assert int(c.traj_count.min()) > 150 and len(set(c.traj_count.tolist())) > 1
    # through the simulation object
    from uav_ac_b200.simulation.batched_sim import BatchedSimulation
    r = BatchedSimulation(3).rollout(3.0, record_actual_trajectory=True)
    assert r.traj_count.tolist() == [len(ref["ticks"])] * 3 and np.abs(r.traj[:len(ref["ticks"]), :, 1].double().cpu().numpy() - ref["positions"]).max() < 1e-4


def test_ground_floor_switch_matches_the_oracle(cuda, golden):
    """uavb_rollout_args.ground_on: the unilateral floor in the fp32 pair kernel and the fp64 kernel against the C
    oracle with the same switch -- with and without obstacles (the floor rides on the obstacle culling), Monte-Carlo vehicles."""
    import torch
    from oracle import c_port, flight_np
    g = golden["planning"]
    tab, start = g["v3_table"], g["waypoints"][0]
    plan = lab_course_plan(cuda, 3.0)
    n = 10 * len(tab)
    B = 130
    rng = np.random.default_rng(21)
    gs, ms_, is_ = rng.uniform(0.8, 1.2, (B, 11)), rng.uniform(0.9, 1.1, B), rng.uniform(0.9, 1.1, (B, 3))
    mc = mc_arrays(cuda, B, gs, ms_, is_)
    vehs = [flight_np.Vehicle().with_values(mc["mc_gains"][:, i].double().cpu().numpy(), float(mc["mc_mass"][i]), mc["mc_inertia"][:, i].double().cpu().numpy())
            for i in range(B)]
    gz = float(start[2])
    m_ref, X_ref = c_port.closed_loop_batch(vehs, tab, start, obstacles=g["obstacles"], goal=GOAL, threads=8, ground_z=gz)
    obs = torch.tensor(g["obstacles"], dtype=torch.float32, device=cuda)
    for kw in (dict(obstacles=obs), dict()):
        r = _fly(cuda, plan, B, n, ground_z=gz, **kw, **mc)
        X = r.state.double().cpu().numpy().T
        assert np.abs(X[:, :3] - X_ref[:, :3]).max() < 1e-4 and rotation_angle(X[:, 3:7], X_ref[:, 3:7]).max() < 1e-4
    r64 = _fly(cuda, plan, 4, n, dtype=torch.float64, ground_z=gz, obstacles=obs)
    ref0 = c_port.closed_loop(flight_np.Vehicle(), tab, start, obstacles=g["obstacles"], goal=GOAL, ground_z=gz)
    assert np.abs(r64.state[:, 0].cpu().numpy() - ref0["X"]).max() < 1e-7
    # the floor holds: the logged altitude never goes below the start height; without it the drone sags while the rotors spin up
    held = _fly(cuda, plan, 8, 400, log_stride=1, ground_z=gz)
    free = _fly(cuda, plan, 8, 400, log_stride=1)
    assert float((held.log[:, 2] - gz).max()) <= 1e-6 and 0.002 < float((free.log[:, 2] - gz).max()) < 0.016
