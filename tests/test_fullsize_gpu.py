"""BASELINE.json configs at FULL size through size-independent properties (the oracle cannot fly 10^6 missions in a test):
determinism under re-sharding, sortedness / sanity of the metrics, chunked == one-launch, spot checks against the C twin of
the oracle on a strided sample.  Each test is sized to finish in seconds on one B200."""
import os

import numpy as np
import pytest

from helpers import GOAL, lab_course_plan, rotation_angle

pytestmark = pytest.mark.gpu


def _c4_inputs(kernels, torch, B, lo=0):
    wp, vel = kernels.mc_missions(31, B, 4, index_base=lo)
    ground = wp[:, 0].clone()
    ground[:, 2] = -0.021
    plan = kernels.plan_missions([(torch.stack((ground, wp[:, 0]), dim=1).contiguous(), vel), (wp, vel)], 0.01)
    wind = kernels.mc_uniform(32, B, [-0.08] * 3, [0.08] * 3, index_base=lo)
    return plan, ground.contiguous(), wp, vel, wind


def test_config3_one_million_rollouts_random_missions_wind_and_obstacle_sets(cuda):
    """configs[3]: 10^6 rollouts, per-rollout waypoint sets (vertical take-off + 4 splines), constant wind, 64 AABB sets of 6 boxes."""
    import torch
    from oracle import c_port, flight_np
    from uav_ac_b200 import kernels
    B = 1_000_000
    plan, start, wp, vel, wind = _c4_inputs(kernels, torch, B)
    rng = np.random.default_rng(8)
    ctr, half = rng.uniform([2, 2, -5], [22, 12, -1], (64, 6, 3)), rng.uniform(0.3, 1.2, (64, 6, 3))
    boxes = np.stack((ctr[..., 0] - half[..., 0], ctr[..., 0] + half[..., 0], ctr[..., 1] - half[..., 1], ctr[..., 1] + half[..., 1],
                      ctr[..., 2] - half[..., 2], ctr[..., 2] + half[..., 2]), axis=-1).astype(np.float32)
    boxes_t = torch.tensor(boxes, device=cuda)
    sets = (torch.arange(B, device=cuda) % 64).to(torch.int32)
    n = 10 * int(plan.total_rows.max().item())
    res = kernels.rollout(plan, B, n, start=start, goal=wp[:, -1].contiguous(), mc_wind=wind, obstacles=boxes_t, obstacle_set=sets)
    torch.cuda.synchronize()
    m = res.metrics
    assert bool(torch.isfinite(m[:, :5]).all()) and int((m[:, 5] != 0).sum()) == 0
    hit = m[:, 1] > 0
    assert 0.05 < float(hit.float().mean()) < 0.6
    assert bool((m[hit, 6] >= 0).all()) and bool((m[~hit, 6] == -1).all()) and bool((m[:, 6] < n).all())     # first-hit tick consistent with the flag
    assert bool((m[:, 7] == n // 10).all())                                                                    # every outer period accounted for
    assert bool((m[:, 3] <= m[:, 2] + 1e-6).all()) and bool((m[:, 2] <= m[:, 4] + 1e-6).all())                  # mean <= rms <= max tracking error
    assert float(m[:, 0].median()) < 0.05                                                                      # drones end on their goals
    # determinism under re-sharding: a slice of the global index range flown alone reproduces its rows bit for bit
    lo, k = 700_000, 1024
    plan2, start2, wp2, _, wind2 = _c4_inputs(kernels, torch, k, lo)
    sub = kernels.rollout(plan2, k, n, start=start2, goal=wp2[:, -1].contiguous(), mc_wind=wind2, obstacles=boxes_t,
                          obstacle_set=sets[lo:lo + k].contiguous())
    assert torch.equal(sub.metrics, m[lo:lo + k]) and torch.equal(sub.state, res.state[:, lo:lo + k])
    # strided spot check against the C twin of the oracle (fp64) with the same inputs
    idx = np.arange(0, B, B // 24)
    wpn, veln, windn = wp[idx].cpu().numpy(), vel[idx].cpu().numpy(), wind[:, idx].double().cpu().numpy().T
    mm, X = m[idx].double().cpu().numpy(), res.state[:, idx].double().cpu().numpy().T
    checked = 0
    for j, b in enumerate(idx):
        g0 = wpn[j, 0].copy(); g0[2] = -0.021
        tab = np.vstack([c_port.sample_table(*c_port.solve_coeffs(w, veln[j]), 0.01) for w in (np.stack((g0, wpn[j, 0])), wpn[j])])
        ref = c_port.closed_loop(flight_np.Vehicle(), tab, g0, obstacles=boxes[int(b) % 64].astype(float), goal=wpn[j, -1], wind=windn[j], n_ticks=n,
                                 log_stride=1)
        if ref["max_err"] > 5.0:
            continue                                                          # diverged flights amplify rounding
        assert np.abs(X[j, :3] - ref["X"][:3]).max() < 1e-4 and rotation_angle(X[j, 3:7], ref["X"][3:7]) < 1e-4
        gap = np.inf
        for q in boxes[int(b) % 64].astype(float):
            P = ref["log"][:, :3]
            d = np.maximum.reduce([q[0] - P[:, 0], P[:, 0] - q[1], q[2] - P[:, 1], P[:, 1] - q[3], q[4] - P[:, 2], P[:, 2] - q[5]])
            gap = min(gap, np.abs(d).min())
        if gap >= 1e-4:                                                       # outside the fp32 ambiguity band the flag is exact
            assert mm[j, 1] == float(ref["collision"]) and mm[j, 6] == ref["first_collision_tick"]
            checked += 1
    assert checked >= 12


def test_config4_share_sixty_seconds_chunked_with_carry(cuda):
    """configs[4], one GPU's share at reduced width (125 000 rollouts x 60 000 ticks): three chunked launches through the
    resumable carry reproduce a single launch bit for bit; the mission ends in a hold-last-row hover on the goal."""
    import torch
    from uav_ac_b200 import kernels, _native as nat
    from uav_ac_b200.simulation.scene import LAB_COURSE_OBSTACLES, LAB_COURSE_START
    B, ticks, chunk = 125_000, 60_000, 20_000
    plan = lab_course_plan(cuda, 3.0)
    v = nat.default_vehicle()
    base = torch.tensor(list(v.gains) + [v.mass] + list(v.inertia), dtype=torch.float32, device=cuda)[:, None]
    mc = (kernels.mc_uniform(5, B, [0.8] * 11 + [0.9] * 4, [1.2] * 11 + [1.1] * 4) * base).contiguous()
    kw = dict(start=torch.tensor(LAB_COURSE_START, dtype=torch.float64, device=cuda), goal=torch.tensor(GOAL, dtype=torch.float64, device=cuda),
              obstacles=torch.tensor(LAB_COURSE_OBSTACLES, dtype=torch.float32, device=cuda), mc_gains=mc[:11], mc_mass=mc[11], mc_inertia=mc[12:15])
    whole = kernels.rollout(plan, B, ticks, **kw)
    carry = torch.empty((nat.CARRY_WORDS, B), dtype=torch.float32, device=cuda)
    for k in range(ticks // chunk):
        part = kernels.rollout(plan, B, chunk, carry=carry, resume=k > 0, **kw)
    torch.cuda.synchronize()
    assert torch.equal(part.metrics, whole.metrics) and torch.equal(part.state, whole.state)
    m = whole.metrics
    assert bool((m[:, 7] == ticks // 10).all()) and int((m[:, 5] != 0).sum()) == 0 and int((m[:, 1] != 0).sum()) == 0
    assert float(m[:, 0].max()) < 5e-3                                         # 49 s of hover on the last row: every drone sits on the goal
    assert float(whole.state[7:13].abs().max()) < 1e-2                         # at rest


def test_config2_every_rollout_of_the_headline_batch_against_the_c_oracle(cuda):
    """BASELINE configs[2] at FULL size, checked exhaustively: all 10^5 Monte-Carlo lab_course rollouts x 10 760 ticks of the bench
    workload (same generator, same seed as bench.py) against oracle/oracle_c.c flown with the fp32-rounded vehicles the kernel was
    given (~7 s on 16 host cores).  North-star tolerances: 1e-4 m, 1e-4 rad, collision flags bit-exact; metrics to 2e-5."""
    import torch
    import bench_workloads as wl
    from oracle import c_port, flight_np, minsnap_np
    from uav_ac_b200 import _native as nat, kernels
    from uav_ac_b200.simulation.scene import LAB_COURSE_GOAL, LAB_COURSE_OBSTACLES, LAB_COURSE_START, LAB_COURSE_WAYPOINTS
    B = 100_000
    mc = wl.mc_vehicle_arrays(kernels, nat, cuda, B)
    plan, kw = wl.lab_course(kernels, cuda)
    n = 10 * int(plan.total_rows.item())
    kw["want_state"] = True
    res = kernels.rollout(plan, B, n, mc_gains=mc[:11], mc_mass=mc[11], mc_inertia=mc[12:15], **kw)
    torch.cuda.synchronize()
    mch = mc.double().cpu().numpy()
    base = flight_np.Vehicle()
    vehicles = [base.with_values(mch[:11, i], mch[11, i], mch[12:15, i]) for i in range(B)]
    tab = minsnap_np.mission_table(LAB_COURSE_WAYPOINTS, None, 3.0, 0.01, method="solve")
    assert len(tab) * 10 == n
    m_ref, X_ref = c_port.closed_loop_batch(vehicles, tab, np.asarray(LAB_COURSE_START, float), obstacles=LAB_COURSE_OBSTACLES, goal=LAB_COURSE_GOAL,
                                            threads=os.cpu_count() or 1)
    X, m = res.state.double().cpu().numpy().T, res.metrics.double().cpu().numpy()
    dpos = np.abs(X[:, :3] - X_ref[:, :3]).max(axis=1)
    dang = rotation_angle(X[:, 3:7], X_ref[:, 3:7])
    print(f"10^5 rollouts: max |dpos| {dpos.max():.2e} m, attitude {dang.max():.2e} rad (median {np.median(dang):.2e}), |dv| {np.abs(X[:, 7:10] - X_ref[:, 7:10]).max():.2e}, "
          f"|dw| {np.abs(X[:, 10:13] - X_ref[:, 10:13]).max():.2e}, collisions {int(m_ref[:, 1].sum())}")
    assert dpos.max() < 1e-4 and dang.max() < 1e-4
    assert np.abs(X[:, 7:10] - X_ref[:, 7:10]).max() < 1e-3 and np.abs(X[:, 10:13] - X_ref[:, 10:13]).max() < 1e-3
    assert np.array_equal(m[:, 1], m_ref[:, 1]) and np.array_equal(m[:, 6], m_ref[:, 6])            # flags and first-hit ticks, bit for bit
    assert np.abs(m[:, 0] - m_ref[:, 0]).max() < 1e-4 and np.abs(m[:, 2:5] - m_ref[:, 2:5]).max() < 1e-4 and np.array_equal(m[:, 7], m_ref[:, 7])
    assert int((m[:, 5] != 0).sum()) == 0


def test_correction_loop_leaves_no_sampled_point_inside_a_box_at_full_size(cuda):
    """10^5 random five-waypoint missions x 4 shared boxes through the device-side correction loop; size-independent property: for
    every mission that finished (status OK), the sampled (N, 11) table of its final plan -- K3, a different kernel from the loop's own
    sweep -- has no row inside any box, and missions that were never hit keep their five waypoints.  Missions whose waypoint sits in
    a box cannot be cleared (the reference loops forever) and are reported as UAVB_SOLVE_TOO_MANY."""
    import torch
    from uav_ac_b200 import _native as nat, kernels
    B, cap = 100_000, 33
    wp, vel = kernels.mc_missions(123, B, 4, device=cuda)
    rng = np.random.default_rng(5)
    ctr, half = rng.uniform([4, 3, -4.5], [20, 11, -1.5], (4, 3)), rng.uniform(0.15, 0.4, (4, 3))
    boxes = torch.tensor(np.stack((ctr[:, 0] - half[:, 0], ctr[:, 0] + half[:, 0], ctr[:, 1] - half[:, 1], ctr[:, 1] + half[:, 1],
                                   ctr[:, 2] - half[:, 2], ctr[:, 2] + half[:, 2]), axis=-1), dtype=torch.float64, device=cuda)
    w, n = kernels.fixed_pitch(wp, cap, cuda)
    c, t, status, rounds = kernels.minsnap_correct(w, n, vel, 0.01, boxes)
    ok = status == 0
    assert int(ok.sum()) > 0.97 * B and set(status.unique().tolist()) <= {0, nat.SOLVE_TOO_MANY}
    # a waypoint inside a box <=> cannot be cleared
    p = wp[:, :, None, :]                                                        # [B, 5, 1, 3]
    inside = ((boxes[None, None, :, 0] <= p[..., 0]) & (p[..., 0] <= boxes[None, None, :, 1]) & (boxes[None, None, :, 2] <= p[..., 1])
              & (p[..., 1] <= boxes[None, None, :, 3]) & (boxes[None, None, :, 4] <= p[..., 2]) & (p[..., 2] <= boxes[None, None, :, 5])).any(dim=2).any(dim=1)
    # (not every one: the last waypoint of a mission is never sampled -- np.arange excludes T -- so a box that holds only that point
    # within one step of its face may go unnoticed, exactly as in the reference)
    assert float((status[inside] == nat.SOLVE_TOO_MANY).float().mean()) > 0.9
    grown = n > 5
    print(f"correction loop at full size: {rounds} rounds, {int(grown.sum())} missions grew (up to {int(n.max())} waypoints), "
          f"{int((status != 0).sum())} cannot be cleared, {int(inside.sum())} of them with a waypoint inside a box")
    # K3 tables of the finished missions
    idx = ok.nonzero().squeeze(1)
    cp, tp, seg_off = kernels.pack_segments(c[idx].contiguous(), t[idx].contiguous(), n[idx].contiguous())
    rows, yaw0, total = kernels.table_meta(cp, tp, seg_off, 0.01)
    roff = torch.zeros(idx.numel() + 1, dtype=torch.int32, device=cuda)
    roff[1:] = torch.cumsum(total, 0)
    table = kernels.minsnap_sample(cp, tp, seg_off, rows, roff, 0.01)
    mask = torch.zeros(idx.numel(), dtype=torch.int64, device=cuda)
    for k in range(4):
        kernels.table_hits(table, roff, boxes[k].contiguous(), mask)
    assert int((mask != 0).sum()) == 0
    assert bool((n[ok & ~grown] == 5).all()) and torch.equal(w[~grown & ok, :5], wp[~grown & ok])
