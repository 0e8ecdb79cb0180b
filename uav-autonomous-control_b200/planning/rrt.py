"""Batched ``RRTStar``: the reference planner's surface (``uav_ac/planning/rrt.py:7-301``) over the warp-per-mission
CUDA kernel ``uavb_rrt_star_f64``.

``RRTStar(space_limits, start, goal, max_distance, max_iterations, obstacles).run()`` with (3,) start / goal plans one
mission and leaves the (n, 3) NumPy ``best_path`` like the reference; (B, 3) arrays plan B missions at once
(``best_path`` is then a list).  The reference draws from NumPy's global generator; here the stream is Philox keyed by
``seed`` and the mission index, so a run is reproducible and independent of batch size.  The reference prints progress
and raises ``Exception("No path found")``; the batched class raises for a single mission and reports ``status`` per
mission for a batch.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _native as nat


def _dev():
    if not torch.cuda.is_available():
        raise nat.UavbError("no CUDA device visible: the batched RRTStar has no CPU implementation")
    nat.lib()
    return torch.device("cuda", torch.cuda.current_device())


def segments_hit(p, q, obstacles) -> np.ndarray:
    """hit[i] = segment p[i] -> q[i] intersects any obstacle (exact slab test, rrt.py:246-274), on the device."""
    dev = _dev()
    p_t = torch.as_tensor(np.asarray(p, dtype=float).reshape(-1, 3), dtype=torch.float64, device=dev).contiguous()
    q_t = torch.as_tensor(np.asarray(q, dtype=float).reshape(-1, 3), dtype=torch.float64, device=dev).contiguous()
    n = p_t.shape[0]
    hit = torch.zeros(n, dtype=torch.int32, device=dev)
    if obstacles is None or np.size(obstacles) == 0:
        return hit.cpu().numpy().astype(bool)
    boxes = torch.as_tensor(np.asarray(obstacles, dtype=float).reshape(-1, 6), dtype=torch.float64, device=dev).contiguous()
    nat.check(nat.lib().uavb_segments_hit_aabbs_f64(nat.ptr(p_t), nat.ptr(q_t), n, nat.ptr(boxes), boxes.shape[0], nat.ptr(hit), nat.stream_ptr(dev)),
              "uavb_segments_hit_aabbs_f64")
    return hit.cpu().numpy().astype(bool)


class RRTStar:
    """Rapidly-exploring Random Tree (RRT*) for B missions."""

    def __init__(self, space_limits, start, goal, max_distance, max_iterations, obstacles=None, seed: int = 0, index_base: int = 0,
                 max_path: int = 256):
        space_limits = np.asarray(space_limits, dtype=float)
        self.space_limits_lw, self.space_limits_up = space_limits[0], space_limits[1]
        self._single = np.ndim(start) == 1
        self.start = np.round(np.asarray(start, dtype=float), 2)
        self.goal = np.round(np.asarray(goal, dtype=float), 2)
        self.step_size = max_distance
        self.max_iterations = int(max_iterations)
        self.obstacles = obstacles
        self.epsilon = 0.15
        self.neighborhood_radius = 1.5 * max_distance
        self.seed, self.index_base, self.max_path = int(seed), int(index_base), int(max_path)
        self.best_path = None
        self.simplified_path = None
        self.cost = None
        self.status = None
        s2, g2 = self.start.reshape(-1, 3), self.goal.reshape(-1, 3)
        assert self.neighborhood_radius > self.step_size, "Neighborhood radius must be larger than step size"
        assert np.all((self.space_limits_lw[2] <= s2[:, 2]) & (s2[:, 2] <= self.space_limits_up[2])), \
            "The z location of the start must be within the z space limits"
        assert np.all((self.space_limits_lw[2] <= g2[:, 2]) & (g2[:, 2] <= self.space_limits_up[2])), \
            "The z location of the goal must be within the z space limits"

    def run(self):
        dev = _dev()
        f64 = dict(dtype=torch.float64, device=dev)
        start = torch.as_tensor(self.start.reshape(-1, 3), **f64).contiguous()
        goal = torch.as_tensor(self.goal.reshape(-1, 3), **f64).contiguous()
        B = start.shape[0]
        limits = torch.as_tensor(np.stack((self.space_limits_lw, self.space_limits_up)), **f64).contiguous()
        obs = None
        if self.obstacles is not None and np.size(self.obstacles) > 0:
            obs = torch.as_tensor(np.asarray(self.obstacles, dtype=float).reshape(-1, 6), **f64).contiguous()
        L = nat.lib()
        ws = torch.empty(int(L.uavb_rrt_workspace_bytes(B, self.max_iterations)), dtype=torch.uint8, device=dev)
        path = torch.empty((B, self.max_path, 3), **f64)           # rows [0, plen) / [0, slen) are written by the kernel
        simple = torch.empty((B, self.max_path, 3), **f64)
        plen = torch.zeros(B, dtype=torch.int32, device=dev)
        slen = torch.zeros(B, dtype=torch.int32, device=dev)
        cost = torch.empty(B, **f64)
        status = torch.empty(B, dtype=torch.int32, device=dev)
        stats = torch.empty((B, 2), dtype=torch.int32, device=dev)
        nat.check(L.uavb_rrt_star_f64(nat.ptr(limits), nat.ptr(start), nat.ptr(goal), B, float(self.step_size), self.max_iterations, nat.ptr(obs),
                                      0 if obs is None else obs.shape[0], self.seed, self.index_base, ctypes.c_void_p(ws.data_ptr()), nat.ptr(path),
                                      self.max_path, nat.ptr(plen), nat.ptr(simple), nat.ptr(slen), nat.ptr(cost), nat.ptr(status), nat.ptr(stats),
                                      nat.stream_ptr(dev)), "uavb_rrt_star_f64")
        torch.cuda.synchronize(dev)
        plen, slen = plen.cpu().numpy(), slen.cpu().numpy()
        # only the used prefix of the path buffers travels to the host (B x max_path x 24 bytes each otherwise)
        pmax, smax = int(plen.max(initial=0)), int(slen.max(initial=0))
        path, simple = path[:, :pmax].cpu().numpy(), simple[:, :smax].cpu().numpy()
        self.status, self.cost, self.stats = status.cpu().numpy(), cost.cpu().numpy(), stats.cpu().numpy()
        paths = [path[b, :plen[b]] for b in range(B)]
        simples = [simple[b, :slen[b]] for b in range(B)]
        if self._single:
            if self.status[0] == 1:
                raise Exception("No path found")
            if self.status[0] == 2:
                raise Exception(f"path longer than max_path={self.max_path}")
            self.best_path, self.simplified_path, self.cost = paths[0], simples[0], float(self.cost[0])
        else:
            self.best_path, self.simplified_path = paths, simples

    @staticmethod
    def path_cost(path):
        """Length of a polyline (rrt.py:86-93)."""
        path = np.asarray(path, dtype=float)
        return float(sum(np.linalg.norm(path[i + 1] - path[i]) for i in range(len(path) - 1)))

    def simplify_path(self, path: np.ndarray) -> np.ndarray:
        """Remove waypoints bypassed by a collision-free direct connection (rrt.py:97-118); connection tests on the device."""
        path = np.asarray(path, dtype=float)
        if len(path) <= 2:
            return path
        out, cur = [path[0]], 0
        while cur < len(path) - 1:
            cand = np.arange(len(path) - 1, cur + 1, -1)                       # farthest first; cur + 1 is always accepted
            nxt = cur + 1
            if len(cand):
                hit = segments_hit(np.repeat(path[cur][None], len(cand), 0), path[cand], self.obstacles)
                free = np.flatnonzero(~hit)
                if len(free):
                    nxt = int(cand[free[0]])
            out.append(path[nxt])
            cur = nxt
        return np.asarray(out)

    def _is_valid_connection(self, node, new_node) -> bool:
        """True when the segment misses every obstacle (rrt.py:232-243)."""
        if self.obstacles is None:
            return True
        return not bool(segments_hit(node, new_node, self.obstacles)[0])

    @staticmethod
    def _segment_intersects_cuboid(node1, node2, cuboid) -> bool:
        """Exact segment vs AABB test (rrt.py:246-274)."""
        return bool(segments_hit(node1, node2, np.asarray(cuboid, dtype=float).reshape(1, 6))[0])
