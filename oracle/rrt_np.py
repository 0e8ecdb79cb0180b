"""NumPy restatement of the reference RRT* planner with a counter-based random stream.  TEST INFRASTRUCTURE ONLY.

Follows ``/root/reference/uav_ac/planning/rrt.py`` (cited ``rrt:LINE``).  Two deliberate, result-preserving differences:

* nodes are identified by index instead of the reference's ``str(np.round(node, 2).tolist())`` dictionary key
  (rrt:164-166).  When ``_update_tree`` re-parents an existing key (rrt:201-206) the reference appends a duplicate to
  ``all_nodes``; duplicates never change an argmin (first index wins) or a rewire decision, so one entry per key with
  an updated parent is equivalent;
* random numbers come from ``philox_u01`` keyed by (seed, mission, iteration) -- the stream the CUDA kernel draws
  from -- instead of NumPy's global generator (rrt:122-131).  ``tests/golden/make_golden.py`` feeds the same numbers to
  the reference itself to pin this file on it.

Distances are formed as ``sqrt((dx*dx + dy*dy) + dz*dz)`` in separately rounded steps, the order the kernel uses.
"""
from __future__ import annotations

import math

import numpy as np

M32 = 0xFFFFFFFF


def philox4x32(counter, key):
    """Philox4x32-10 block (same constants as csrc/mc_kernels.cu)."""
    c0, c1, c2, c3 = counter
    k0, k1 = key
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & M32, p1 >> 32, p1 & M32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & M32, lo1, (hi0 ^ c3 ^ k1) & M32, lo0
        k0, k1 = (k0 + 0x9E3779B9) & M32, (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


def u01d(hi, lo):
    """[0, 1) with 53 random bits (mc_kernels.cu u01d)."""
    return ((hi >> 5) * 67108864.0 + (lo >> 6)) * 1.1102230246251565e-16


def philox_u01(seed: int, mission: int, it: int):
    """The four uniforms of iteration ``it``: goal-bias draw, x, y, z (stream id 0x5252 'RR')."""
    key = (seed & M32, (seed >> 32) & M32)
    a = philox4x32((mission & M32, (mission >> 32) & M32, 0x5252, (2 * it) & M32), key)
    b = philox4x32((mission & M32, (mission >> 32) & M32, 0x5252, (2 * it + 1) & M32), key)
    return u01d(a[0], a[1]), u01d(a[2], a[3]), u01d(b[0], b[1]), u01d(b[2], b[3])


def round2(x):
    """np.round(x, 2): rint(x * 100) / 100 (round-half-even on the scaled value)."""
    return np.rint(np.asarray(x, dtype=float) * 100.0) / 100.0


def dist(a, b) -> float:
    dx, dy, dz = float(a[0]) - float(b[0]), float(a[1]) - float(b[1]), float(a[2]) - float(b[2])
    return math.sqrt((dx * dx + dy * dy) + dz * dz)


def segment_hits_cuboid(p, q, box) -> bool:
    """Slab test, rrt:246-274."""
    t_min, t_max = 0.0, 1.0
    for ax in range(3):
        d = float(q[ax]) - float(p[ax])
        lo, hi = float(box[2 * ax]), float(box[2 * ax + 1])
        if abs(d) < 1e-12:
            if p[ax] < lo or p[ax] > hi:
                return False
            continue
        t_lo, t_hi = (lo - float(p[ax])) / d, (hi - float(p[ax])) / d
        if t_lo > t_hi:
            t_lo, t_hi = t_hi, t_lo
        t_min, t_max = max(t_min, t_lo), min(t_max, t_hi)
        if t_min > t_max:
            return False
    return True


def valid_connection(p, q, obstacles) -> bool:
    """rrt:232-243."""
    if obstacles is None:
        return True
    return not any(segment_hits_cuboid(p, q, box) for box in obstacles)


def path_cost(path) -> float:
    """rrt:86-93."""
    return sum(dist(path[i + 1], path[i]) for i in range(len(path) - 1))


def simplify_path(path, obstacles):
    """Greedy shortcutting, rrt:97-118."""
    path = np.asarray(path, dtype=float)
    if len(path) <= 2:
        return path
    out, cur = [path[0]], 0
    while cur < len(path) - 1:
        nxt = len(path) - 1
        while nxt > cur + 1:
            if valid_connection(path[cur], path[nxt], obstacles):
                break
            nxt -= 1
        out.append(path[nxt])
        cur = nxt
    return np.asarray(out)


def rrt_star(space_limits, start, goal, max_distance, max_iterations, obstacles=None, *, seed=0, mission=0, uniforms=None, epsilon=0.15):
    """``RRTStar(...).run()`` (rrt:37-79).  Returns dict(path, cost, nodes, parents, iterations); raises if no path.

    ``uniforms(it) -> (u0, ux, uy, uz)`` overrides the Philox stream (used to feed the reference's own numbers)."""
    lw, up = np.asarray(space_limits[0], dtype=float), np.asarray(space_limits[1], dtype=float)
    start, goal = round2(start), round2(goal)
    step, radius = float(max_distance), 1.5 * float(max_distance)
    draw = uniforms if uniforms is not None else (lambda it: philox_u01(seed, mission, it))
    nodes, parent = [start], [-1]

    def cost_to_come(i):                                  # rrt:168-179, walking towards the start
        c = 0.0
        while i != 0:
            c += dist(nodes[i], nodes[parent[i]])
            i = parent[i]
        return c

    def goal_index():
        for i, n in enumerate(nodes):
            if i and np.array_equal(n, goal):
                return i
        return -1

    def get_path(par):
        i = next(k for k, n in enumerate(nodes) if k and np.array_equal(n, goal))
        out = [nodes[i]]
        while i != 0:
            i = par[i]
            out.append(nodes[i])
        return np.array(out[::-1]).reshape(-1, 3)

    best_parent, old_cost, stall, used = None, math.inf, 0, 0
    for it in range(max_iterations):
        used = it + 1
        u0, ux, uy, uz = draw(it)
        if u0 < epsilon:                                  # rrt:122-131
            new = goal.copy()
        else:
            new = round2(lw + (up - lw) * np.array([ux, uy, uz]))
        d = [dist(new, n) for n in nodes]                 # rrt:133-138
        near = int(np.argmin(d))
        if d[near] > step:                                # rrt:140-148
            new = round2(nodes[near] + (new - nodes[near]) * step / d[near])
        nb = [i for i, n in enumerate(nodes) if dist(n, new) <= radius and valid_connection(n, new, obstacles)]   # rrt:150-156
        if not nb:
            continue
        costs = [cost_to_come(i) + dist(nodes[i], new) for i in nb]   # rrt:181-192
        best = nb[int(np.argmin(costs))]
        existing = next((i for i, n in enumerate(nodes) if np.array_equal(n, new)), -1)
        # rrt:194-213
        if not np.array_equal(nodes[best], new):
            if existing > 0:
                if not cost_to_come(existing) <= cost_to_come(best) + dist(new, nodes[best]):
                    parent[existing] = best
                idx = existing
            elif existing == 0:
                idx = 0                                   # the reference would fail on a node equal to the start; never re-parent it
            else:
                nodes.append(new)
                parent.append(best)
                idx = len(nodes) - 1
        else:
            idx = existing
        if idx <= 0:
            continue                                      # new node is the start itself (the reference raises KeyError here)
        # rrt:215-241, sequential: an accepted rewire changes the cost of later neighbours that descend from it
        new_cost = cost_to_come(idx)
        rewired = False
        for i in nb:
            if i == 0 or i == parent[idx]:
                continue
            if new_cost + dist(nodes[i], new) < cost_to_come(i):
                parent[i] = idx
                rewired = True
        gi = goal_index()
        if gi > 0:                                        # rrt:53-72
            cost = path_cost(get_path(parent))
            if rewired and cost > old_cost:
                raise RuntimeError("Cost increased after rewiring")
            if cost < old_cost:
                best_parent, old_cost, stall = list(parent), cost, 0
            else:
                stall += 1
            if stall >= max_iterations / 10:
                break
    if best_parent is None:
        raise RuntimeError("No path found")
    path = get_path(best_parent + [-1] * (len(nodes) - len(best_parent)))
    return dict(path=path, cost=path_cost(path), nodes=np.array(nodes), parents=np.array(best_parent), iterations=used)
