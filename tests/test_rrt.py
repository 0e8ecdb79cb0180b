"""RRT* (SURVEY 8(f) rank 4).  CPU part: the NumPy oracle (oracle/rrt_np.py) reproduces the reference planner
(uav_ac/planning/rrt.py) bit for bit when both draw the same random numbers (tests/golden/rrt.npz was produced by the
reference itself with a patched np.random.uniform), and restates the reference's unit tests.  GPU part: the
warp-per-mission kernel reproduces the oracle -- and therefore the reference -- bit for bit, plus size-independent
properties on a large batch."""
import numpy as np
import pytest

from oracle import rrt_np


def test_oracle_reproduces_the_reference_paths(golden):
    g = golden["rrt"]
    for m in range(len(g["starts"])):
        out = rrt_np.rrt_star(g["limits"], g["starts"][m], g["goals"][m], float(g["step"]), int(g["iters"]), g["obstacles"], seed=int(g["seed"]), mission=m)
        np.testing.assert_array_equal(out["path"], g[f"path{m}"])
        np.testing.assert_array_equal(rrt_np.simplify_path(out["path"], g["obstacles"]), g[f"simple{m}"])
        assert out["cost"] == pytest.approx(float(g[f"cost{m}"]), rel=1e-14)
    out = rrt_np.rrt_star(g["limits"], g["starts"][0], g["goals"][0], float(g["step"]), int(g["iters"]), None, seed=int(g["seed"]), mission=0)
    np.testing.assert_array_equal(out["path"], g["path_free"])


def test_oracle_slab_test_matches_reference(golden):
    g = golden["rrt"]
    got = np.array([[rrt_np.segment_hits_cuboid(p, q, box) for box in g["obstacles"]] for p, q in zip(g["seg_p"], g["seg_q"])])
    np.testing.assert_array_equal(got, g["seg_hits"])


def test_oracle_known_answers_of_reference_unit_tests():
    """tests/unit/planning/test_rrt.py:7-40, 200-225."""
    assert rrt_np.path_cost(np.array([[1, 1, 1], [3, 3, 9], [11, 5, 5], [1, 1, 1]], dtype=float)) == pytest.approx(29.1, abs=0.1)
    path = np.array([[0., 0., 0.], [2., 0., 0.], [4., 0., 0.], [6., 0., 0.]])
    np.testing.assert_array_equal(rrt_np.simplify_path(path, None), [[0, 0, 0], [6, 0, 0]])
    obs = np.array([[5., 7., -1., 1., -1., 1.]])
    detour = rrt_np.simplify_path(np.array([[0., 0., 0.], [4., 2., 0.], [8., 2., 0.], [12., 0., 0.]]), obs)
    assert len(detour) > 2 and all(rrt_np.valid_connection(a, b, obs) for a, b in zip(detour[:-1], detour[1:]))
    thin = np.array([[4.9, 5.1, -5., 5., -5., 5.]])
    assert not rrt_np.valid_connection([0., 0, 0], [10., 0, 0], thin)                  # thin wall between the samples is detected
    assert rrt_np.valid_connection([0., 0, 0], [4., 0, 0], thin)
    u = [rrt_np.philox_u01(1, 2, it) for it in range(200)]
    assert all(0.0 <= x < 1.0 for t in u for x in t) and len({t[0] for t in u}) == 200


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_kernel_reproduces_oracle_and_reference_bit_for_bit(cuda, golden):
    from uav_ac_b200.planning.rrt import RRTStar
    g = golden["rrt"]
    B = len(g["starts"])
    r = RRTStar(g["limits"], g["starts"], g["goals"], float(g["step"]), int(g["iters"]), g["obstacles"], seed=int(g["seed"]))
    r.run()
    assert (r.status == 0).all()
    for m in range(B):
        np.testing.assert_array_equal(r.best_path[m], g[f"path{m}"])                  # the reference's own best_path
        np.testing.assert_array_equal(r.simplified_path[m], g[f"simple{m}"])
        assert r.cost[m] == pytest.approx(float(g[f"cost{m}"]), rel=1e-14)
    # single mission, drop-in surface: numpy best_path, simplify_path / _is_valid_connection on the device
    one = RRTStar(g["limits"], g["starts"][2], g["goals"][2], float(g["step"]), int(g["iters"]), g["obstacles"], seed=int(g["seed"]), index_base=2)
    one.run()
    np.testing.assert_array_equal(one.best_path, g["path2"])
    np.testing.assert_array_equal(one.simplify_path(one.best_path), g["simple2"])
    assert RRTStar.path_cost(one.best_path) == pytest.approx(one.cost, rel=1e-12)
    free = RRTStar(g["limits"], g["starts"][0], g["goals"][0], float(g["step"]), int(g["iters"]), None, seed=int(g["seed"]))
    free.run()
    np.testing.assert_array_equal(free.best_path, g["path_free"])
    # oracle on missions the goldens do not contain (other seed)
    for m in (0, 3):
        o = rrt_np.rrt_star(g["limits"], g["starts"][m], g["goals"][m], 1.2, 800, g["obstacles"], seed=5, mission=m)
        k = RRTStar(g["limits"], g["starts"][m], g["goals"][m], 1.2, 800, g["obstacles"], seed=5, index_base=m)
        k.run()
        np.testing.assert_array_equal(k.best_path, o["path"])


@pytest.mark.gpu
def test_segment_tests_match_reference_truth_table(cuda, golden):
    from uav_ac_b200.planning.rrt import RRTStar, segments_hit
    g = golden["rrt"]
    for k, box in enumerate(g["obstacles"]):
        np.testing.assert_array_equal(segments_hit(g["seg_p"], g["seg_q"], box[None]), g["seg_hits"][:, k])
    np.testing.assert_array_equal(segments_hit(g["seg_p"], g["seg_q"], g["obstacles"]), g["seg_hits"].any(axis=1))
    thin = np.array([4.9, 5.1, -5., 5., -5., 5.])
    assert RRTStar._segment_intersects_cuboid([0., 0, 0], [10., 0, 0], thin) and not RRTStar._segment_intersects_cuboid([0., 0, 0], [4., 0, 0], thin)


@pytest.mark.gpu
def test_two_thousand_missions_give_valid_paths(cuda, golden):
    """Properties the reference's planner guarantees, on 2 048 random start/goal pairs in the lab volume: the path starts and ends
    on the (rounded) start and goal, every edge is at most the neighbourhood radius long and misses every obstacle, nodes
    stay inside the planning bounds, the simplified path is a collision-free subsequence, results do not depend on batching."""
    from uav_ac_b200.planning.rrt import RRTStar, segments_hit
    g = golden["rrt"]
    rng = np.random.default_rng(11)
    B = 2048
    lim, obs = g["limits"], g["obstacles"]

    def free_points(n):
        pts = np.empty((0, 3))
        while len(pts) < n:
            c = np.round(rng.uniform(lim[0] + 0.3, lim[1] - 0.3, (2 * n, 3)), 2)
            inside = np.zeros(len(c), bool)
            for b in obs:
                inside |= (b[0] - 0.2 <= c[:, 0]) & (c[:, 0] <= b[1] + 0.2) & (b[2] - 0.2 <= c[:, 1]) & (c[:, 1] <= b[3] + 0.2) & (b[4] - 0.2 <= c[:, 2]) & (c[:, 2] <= b[5] + 0.2)
            pts = np.vstack((pts, c[~inside]))
        return pts[:n]
    starts, goals = free_points(B), free_points(B)
    r = RRTStar(lim, starts, goals, 1.5, 1200, obs, seed=99)
    r.run()
    ok = r.status == 0
    assert ok.mean() > 0.97                                                      # the rest ran out of iterations ("No path found")
    p_all, q_all = [], []
    for b in np.flatnonzero(ok):
        path = r.best_path[b]
        assert np.array_equal(path[0], starts[b]) and np.array_equal(path[-1], goals[b])
        assert (path >= lim[0] - 1e-9).all() and (path <= lim[1] + 1e-9).all()
        seg = np.linalg.norm(np.diff(path, axis=0), axis=1)
        assert seg.max() <= 1.5 * 1.5 + 1e-9 and r.cost[b] == pytest.approx(seg.sum(), rel=1e-12)
        sp = r.simplified_path[b]
        assert np.array_equal(sp[0], path[0]) and np.array_equal(sp[-1], path[-1]) and len(sp) <= len(path)
        idx = [int(np.flatnonzero((path == s).all(axis=1))[0]) for s in sp]
        assert idx == sorted(idx)                                                # a subsequence of the path
        p_all += [path[:-1], sp[:-1]]
        q_all += [path[1:], sp[1:]]
    assert not segments_hit(np.vstack(p_all), np.vstack(q_all), obs).any()      # every edge of every path misses every obstacle
    lo = 700
    sub = RRTStar(lim, starts[lo:lo + 8], goals[lo:lo + 8], 1.5, 1200, obs, seed=99, index_base=lo)
    sub.run()
    for k in range(8):
        np.testing.assert_array_equal(sub.best_path[k], r.best_path[lo + k])
