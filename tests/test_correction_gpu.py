"""The obstacle-correction loop of MinimumSnap._generate_collision_free_trajectory (uav_ac/planning/minimum_snap.py:63-95) run
entirely on the device (csrc/minsnap_correct.cu) -- against the NumPy oracle's restatement of the reference loop
(oracle/minsnap_np.plan_table, itself pinned on the reference's own result in tests/test_oracle_golden.py) and against the
reference golden."""
from __future__ import annotations

import numpy as np
import pytest

from helpers import normwise

pytestmark = pytest.mark.gpu


DT = 0.02


def _polyline_distance(p, w):
    """Distance of the points p (n, 3) from the polyline through the waypoints w."""
    best = np.full(len(p), np.inf)
    for a, b in zip(w[:-1], w[1:]):
        ab = b - a
        t = np.clip(((p - a) @ ab) / (ab @ ab), 0.0, 1.0)
        best = np.minimum(best, np.linalg.norm(p - (a + t[:, None] * ab), axis=1))
    return best


def _random_cases(rng, n):
    """Missions of 4-6 waypoints with turns and 1-3 small boxes placed where the minimum-snap curve BULGES away from the
    straight legs (a box on a straight leg can never be cleared: the inserted midpoints converge onto it and the reference loops
    forever).  The boxes stay clear of the polyline, so inserting midpoints pulls the curve out of them."""
    from oracle import minsnap_np
    cases = []
    while len(cases) < n:
        k = int(rng.integers(4, 7))
        steps = rng.uniform([2.0, -3.0, -0.6], [4.0, 3.0, 0.6], (k - 1, 3))
        w = np.cumsum(np.concatenate((rng.uniform([2, 2, -4], [6, 8, -1], (1, 3)), steps)), axis=0)
        v = float(rng.uniform(1.5, 3.0))
        tab = minsnap_np.plan_table(w, None, v, DT, method="solve")[0]
        dist = _polyline_distance(tab[:, :3], w)
        cand = np.flatnonzero(dist > 0.12)
        if len(cand) == 0:
            continue
        boxes = []
        for r in rng.choice(cand, size=min(len(cand), int(rng.integers(1, 4))), replace=False):
            h = rng.uniform(0.25, 0.55, 3) * dist[r]
            c = tab[r, :3]
            boxes.append([c[0] - h[0], c[0] + h[0], c[1] - h[1], c[1] + h[1], c[2] - h[2], c[2] + h[2]])
        boxes = np.array(boxes)
        if len(cases) % 5 == 4:
            boxes = boxes + 50.0                                  # every fifth mission: boxes far away, nothing to correct
        cases.append((w, boxes, v))
    return cases


def test_device_correction_loop_matches_the_oracle_on_random_missions(cuda):
    """400 random missions x 1-3 boxes each (per-mission obstacle sets): the waypoints after insertion are bit-identical to
    the oracle's, the spline counts match, the coefficients agree to 1e-9 norm-wise, and missions that need nothing keep their
    waypoints.  The oracle's table is sampled from ITS coefficients, so a sample within rounding of a box face could decide
    differently; such cases are counted, not hidden."""
    import torch
    from oracle import minsnap_np
    from uav_ac_b200 import _native as nat, kernels
    rng = np.random.default_rng(2024)
    cases = _random_cases(rng, 400)
    n_obs = 3
    far = np.array([1e6, 1e6 + 1, 1e6, 1e6 + 1, 1e6, 1e6 + 1.0])
    obs = np.stack([np.concatenate((b, np.tile(far, (n_obs - len(b), 1)))) for _, b, _ in cases])      # [B, 4, 6], padded with far boxes
    vel = torch.tensor([v for _, _, v in cases], dtype=torch.float64, device=cuda)
    c, t, seg_off, status, wp, n_wp, rounds = kernels.plan_collision_free([w for w, _, _ in cases], vel, DT, torch.tensor(obs, device=cuda), device=cuda)
    assert rounds >= 2
    wp_h, n_h, off, ch, st_h = wp.cpu().numpy(), n_wp.cpu().numpy(), seg_off.cpu().numpy(), c.cpu().numpy(), status.cpu().numpy()
    differ, grown, worst, endless = 0, 0, 0.0, 0
    for b, (w, boxes, v) in enumerate(cases):
        try:
            _, w_ref, c_ref, _ = minsnap_np.plan_table(w, boxes, v, DT, method="solve", max_waypoints=nat.MAX_SPLINES + 1)
        except RuntimeError:                                     # the reference itself would never finish this one
            endless += 1
            assert st_h[b] == nat.SOLVE_TOO_MANY
            continue
        assert st_h[b] == 0
        grown += len(w_ref) > len(w)
        if n_h[b] != len(w_ref) or not np.array_equal(wp_h[b, :n_h[b]], w_ref):
            differ += 1
            continue
        assert off[b + 1] - off[b] == len(w_ref) - 1
        worst = max(worst, normwise(ch[off[b]:off[b + 1]].reshape(-1, 3), c_ref))
    print(f"correction loop: {grown} of {len(cases)} missions grew, {rounds} plan rounds, {endless} endless in the reference too, "
          f"{differ} decided differently at a box face, worst coefficient error {worst:.2e}")
    assert grown >= 300 and differ <= 2 and worst < 1e-9


def test_device_correction_loop_reference_golden_and_limits(cuda, golden):
    """The reference's own corrected mission (tests/unit/planning/test_minimum_snap.py:171-183 scenario) through the kernel-level
    entry points, the shared / per-mission obstacle forms, and the capacity report."""
    import torch
    from uav_ac_b200 import _native as nat, kernels
    g = golden["planning"]
    w_in, boxes = g["fix_waypoints_in"], torch.tensor(g["fix_obstacles"], dtype=torch.float64, device=cuda)
    vel = torch.tensor([1.5, 1.5, 1.5], dtype=torch.float64, device=cuda)
    far = w_in + np.array([100.0, 0, 0])
    c, t, seg_off, status, wp, n_wp, rounds = kernels.plan_collision_free([w_in, far, w_in], vel, 0.01, boxes, device=cuda)
    assert n_wp.tolist() == [len(g["fix_waypoints_out"]), len(far), len(g["fix_waypoints_out"])] and status.tolist() == [0, 0, 0]
    np.testing.assert_array_equal(wp[0, :int(n_wp[0])].cpu().numpy(), g["fix_waypoints_out"])
    assert normwise(c[:int(seg_off[1])].reshape(-1, 3).cpu().numpy(), g["fix_coeffs"]) < 1e-8           # golden: the reference's lstsq branch
    assert torch.equal(c[int(seg_off[2]):], c[:int(seg_off[1])])                                          # same mission, same bits
    # no obstacles: a plain plan, one round, K1's result (another instantiation of the same solver: the compiler contracts a few
    # multiply-adds differently, so equal to rounding, not bit for bit)
    c0, t0, off0, st0, _, n0, r0 = kernels.plan_collision_free([w_in], vel[:1], 0.01, None, device=cuda)
    ck, tk, _ = kernels.minsnap_solve(torch.tensor(w_in[None], dtype=torch.float64, device=cuda), vel[:1])
    assert r0 == 1 and normwise(c0.reshape(-1, 3).cpu().numpy(), ck.reshape(-1, 3).cpu().numpy()) < 1e-13 and torch.equal(t0, tk.reshape(-1))
    # a box around a waypoint: the reference would loop forever; here the mission is reported once it outgrows the solver's limit
    trap = torch.tensor([[w_in[1, 0] - 0.05, w_in[1, 0] + 0.05, w_in[1, 1] - 0.05, w_in[1, 1] + 0.05, w_in[1, 2] - 0.05, w_in[1, 2] + 0.05]],
                        dtype=torch.float64, device=cuda)
    with pytest.raises(kernels.TooManySplines):
        kernels.plan_collision_free([w_in], vel[:1], 0.01, trap, device=cuda)
    wpf, nf = kernels.fixed_pitch([w_in], 6, cuda)
    _, _, st, _ = kernels.minsnap_correct(wpf, nf, vel[:1], 0.01, trap)
    assert st.tolist() == [nat.SOLVE_TOO_MANY]


def test_fused_planners_fly_the_corrected_mission(cuda, golden):
    """the fused paths (uavb_fly_mission_host, BatchedSimulation.rollout / plan_missions) must plan like
    _generate_mission_trajectory(waypoints, obstacles, ...) -- with the correction loop -- not only use the obstacles for the
    collision flag.  Scene: the reference's corrected scenario as the course after a vertical take-off."""
    import torch
    from oracle import flight_np, minsnap_np
    from uav_ac_b200 import host_api, kernels
    g = golden["planning"]
    course, boxes = g["fix_waypoints_in"], g["fix_obstacles"]
    W = np.vstack((course[0] + np.array([0.0, 0.0, 1.0]), course))                  # start 1 m below (NED) the first course waypoint
    v = 1.5
    tab = minsnap_np.mission_table(W, boxes, v, 0.01, method="solve")
    tab_plain = minsnap_np.mission_table(W, None, v, 0.01, method="solve")
    assert len(tab) != len(tab_plain)                                                # the correction changes the mission
    ref = flight_np.closed_loop(flight_np.Vehicle(), tab, W[0], obstacles=boxes, goal=W[-1])
    # C ABI, host buffers
    met, state, n_ticks = host_api.fly_mission_host(W, v, 3, obstacles=boxes, want_state=True)
    assert n_ticks == 10 * len(tab)
    assert np.abs(state[:3, 0] - ref["X"][:3]).max() < 1e-4 and met[0, 1] == float(ref["collision"]) and abs(met[0, 0] - ref["final_dist"]) < 1e-4
    assert (state == state[:, :1]).all()
    _, _, n_plain = host_api.fly_mission_host(W, v, 1, obstacles=boxes, correct=False)
    assert n_plain == 10 * len(tab_plain)
    # tensor-level planner: shared mission and per-rollout missions give the same flight
    wp = torch.tensor(W, dtype=torch.float64, device=cuda)
    vel = torch.tensor([v], dtype=torch.float64, device=cuda)
    obs64 = torch.tensor(boxes, dtype=torch.float64, device=cuda)
    kw = dict(start=wp[0].contiguous(), goal=wp[-1].contiguous(), obstacles=obs64.float())
    shared = kernels.plan_missions([(wp[None, :2].contiguous(), vel), (wp[None, 1:].contiguous(), vel)], 0.01, shared=True, obstacles=obs64)
    assert int(shared.total_rows.item()) == len(tab)
    r_sh = kernels.rollout(shared, 2, n_ticks, **kw)
    np.testing.assert_array_equal(r_sh.state[:, 0].cpu().numpy(), state[:, 0])       # the host-buffer call flies exactly this plan
    B = 5
    wpB, velB = wp[None].expand(B, -1, -1).contiguous(), vel.expand(B).contiguous()
    per = kernels.plan_missions([(wpB[:, :2].contiguous(), velB), (wpB[:, 1:].contiguous(), velB)], 0.01, obstacles=obs64)
    assert per.total_rows.tolist() == [len(tab)] * B and per.seg_count.tolist() == [int(shared.n_seg_shared)] * B
    r_pr = kernels.rollout(per, B, n_ticks, **kw)
    assert float((r_pr.state[:3].double() - torch.tensor(ref["X"][:3], device=cuda)[:, None]).abs().max()) < 1e-4
    assert torch.equal(r_pr.metrics[:, 1], r_sh.metrics[:1, 1].expand(B))


def test_speculative_shared_plan_is_verified_after_the_fact(cuda, golden):
    """plan_missions(shared=True, obstacles=..., table_rows=N): no host round trip while planning; plan.verify() accepts the plan when
    the correction loop had nothing to do and the table has N rows, and refuses it otherwise."""
    import torch
    from uav_ac_b200 import kernels
    from uav_ac_b200.simulation.scene import LAB_COURSE_OBSTACLES, LAB_COURSE_WAYPOINTS
    f64 = dict(dtype=torch.float64, device=cuda)
    wp, vel, obs = torch.tensor(LAB_COURSE_WAYPOINTS, **f64), torch.tensor([3.0], **f64), torch.tensor(LAB_COURSE_OBSTACLES, **f64)
    tables = [(wp[None, :2].contiguous(), vel), (wp[None, 1:].contiguous(), vel)]
    sure = kernels.plan_missions(tables, 0.01, shared=True, obstacles=obs)
    n_rows = int(sure.total_rows.item())
    assert sure.correction_rounds == 1 and sure.report is None
    spec = kernels.plan_missions(tables, 0.01, shared=True, obstacles=obs, table_rows=n_rows)
    assert spec.report is not None
    spec.verify()
    assert spec.report is None and spec.status.tolist() == [0, 0]
    for name in ("seg_coeffs", "seg_rows", "seg_table", "seg_yaw0", "times", "targets"):
        assert torch.equal(getattr(spec, name), getattr(sure, name)), name
    with pytest.raises(RuntimeError, match="table rows"):
        kernels.plan_missions(tables, 0.01, shared=True, obstacles=obs, table_rows=n_rows + 1).verify()
    # the reference's corrected scenario: the loop inserts midpoints, so a speculative plan must be refused
    g = golden["planning"]
    course = torch.tensor(g["fix_waypoints_in"], **f64)
    boxes = torch.tensor(g["fix_obstacles"], **f64)
    v15 = torch.tensor([1.5], **f64)
    with pytest.raises(RuntimeError, match="inside a box"):
        kernels.plan_missions([(course[None].contiguous(), v15)], 0.01, shared=True, obstacles=boxes, table_rows=500).verify()


def test_every_bucket_and_the_size_limits(cuda):
    """Missions of 3 / 6 / 12 / 40 / 64 splines in one batch (all four work-list buckets; 64 = UAVB_MAX_SPLINES) with the boxes far away:
    one plan round, and every mission's coefficients are K1's for its own spline count.  Then the edges: an empty batch, a
    two-waypoint mission, a mission that is already at the capacity and gets hit."""
    import torch
    from uav_ac_b200 import _native as nat, kernels
    rng = np.random.default_rng(12)
    counts = [3, 6, 12, 40, 64, 4, 8, 16, 1]
    paths = [np.cumsum(rng.uniform([1.0, -1.5, -0.3], [2.5, 1.5, 0.3], (s + 1, 3)), axis=0) for s in counts]
    vel = torch.tensor(rng.uniform(1.5, 3.0, len(paths)), dtype=torch.float64, device=cuda)
    far = torch.tensor([[1e6, 1e6 + 1, 1e6, 1e6 + 1, 1e6, 1e6 + 1]], dtype=torch.float64, device=cuda)
    c, t, seg_off, status, wp, n_wp, rounds = kernels.plan_collision_free(paths, vel, 0.01, far, device=cuda)
    assert rounds == 1 and status.tolist() == [0] * len(paths) and n_wp.tolist() == [s + 1 for s in counts]
    for b, (s, p_) in enumerate(zip(counts, paths)):
        ck, tk, st = kernels.minsnap_solve(torch.tensor(p_[None], dtype=torch.float64, device=cuda), vel[b:b + 1])
        mine = c[int(seg_off[b]):int(seg_off[b + 1])].reshape(-1, 3).cpu().numpy()
        assert int(st[0]) == 0 and normwise(mine, ck.reshape(-1, 3).cpu().numpy()) < 1e-12, s
        assert torch.equal(t[int(seg_off[b]):int(seg_off[b + 1])], tk.reshape(-1))
    # empty batch: nothing to do, nothing touched
    e_wp = torch.empty((0, 6, 3), dtype=torch.float64, device=cuda)
    e_n = torch.empty((0,), dtype=torch.int32, device=cuda)
    ce, te, ste, re_ = kernels.minsnap_correct(e_wp, e_n, torch.empty((0,), dtype=torch.float64, device=cuda), 0.01, far)
    assert ce.shape == (0, 5, 8, 3) and re_ == 0
    # at the capacity already: a hit cannot be answered with a midpoint -> reported, waypoints untouched
    g_wp = np.array([[0.0, 0, -1], [4, 0, -1], [4, 4, -1], [8, 4, -1.5]])
    box = torch.tensor([[4.1, 4.6, 0.5, 1.2, -1.2, -0.8]], dtype=torch.float64, device=cuda)      # the reference's corrected scenario
    wpf, nf = kernels.fixed_pitch([g_wp], 4, cuda)
    before = wpf.clone()
    _, _, st2, _ = kernels.minsnap_correct(wpf, nf, torch.tensor([1.5], dtype=torch.float64, device=cuda), 0.01, box)
    assert st2.tolist() == [nat.SOLVE_TOO_MANY] and nf.tolist() == [4] and torch.equal(wpf, before)
    # a degenerate mission (repeated waypoint) is flagged and does not disturb its neighbours
    bad = np.array([[0.0, 0, 0], [1, 0, 0], [1, 0, 0], [2, 0, 0]])
    c3, _, off3, st3, _, _, _ = kernels.plan_collision_free([paths[0], bad, paths[1]], vel[:3].contiguous(), 0.01, far, device=cuda)
    assert st3.tolist() == [0, nat.SOLVE_DEGENERATE, 0]
    assert torch.equal(c3[:int(off3[1])], c[:int(seg_off[1])]) and bool(torch.isnan(c3[int(off3[1]):int(off3[2])]).all())
