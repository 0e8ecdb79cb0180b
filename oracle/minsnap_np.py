"""NumPy fp64 restatement of the reference minimum-snap planner.  TEST INFRASTRUCTURE ONLY.

Follows ``/root/reference/uav_ac/planning/minimum_snap.py`` (cited per function as ``ms:LINE``).
The algorithm is the reference's: assemble the equality constraints ``A c = b`` and the snap
Hessian ``Q``, solve the dense KKT system ``[[Q, A^T], [A, 0]] [c; lam] = [0; b]`` with LAPACK
(``solve`` = LU with partial pivoting, ``lstsq`` = SVD), then sample the piecewise polynomial on
``np.arange(0, T_i, dt)``.  It is written vectorised and shares no code with the reference.

Pinned by tests/test_oracle_golden.py against tests/golden/*.npz (reference outputs generated in the
build container by tests/golden/make_golden.py) and against the known-answer vectors of
``tests/unit/planning/test_minimum_snap.py`` of the reference.
"""
from __future__ import annotations

import numpy as np

N_COEFFS = 8                       # ms:28
START_END_TIME_FACTOR = 1.5        # ms:10
MIN_HORIZONTAL_SPEED_FOR_YAW = 1e-3  # ms:11


def basis_row(order: int, t: float, n: int = N_COEFFS) -> np.ndarray:
    """k-th derivative of the ascending monomial basis at ``t`` (ms:258-286).

    entry i = i (i-1) ... (i-k+1) t^(i-k), zero for i < k.  The reference clamps the exponent at
    zero, so entries with i < k are 0 * t**0 = 0.
    """
    i = np.arange(n, dtype=float)
    fall = np.ones(n)
    expo = i.copy()
    for _ in range(order):
        fall = fall * expo
        expo = np.where(expo > 0, expo - 1, expo)
    return fall * np.power(float(t), expo)


def segment_times(waypoints: np.ndarray, velocity: float) -> np.ndarray:
    """T_i = |w_{i+1} - w_i| / velocity, x1.5 on the first and last spline (ms:311-321).

    With a single spline the factor is applied once (``i in (0, S-1)`` is one membership test).
    """
    w = np.asarray(waypoints, dtype=float)
    S = w.shape[0] - 1
    T = np.empty(S)
    for i in range(S):
        T[i] = np.linalg.norm(w[i + 1] - w[i]) / velocity
        if i in (0, S - 1):
            T[i] *= START_END_TIME_FACTOR
    return T


def constraint_system(waypoints: np.ndarray, T: np.ndarray):
    """Equality constraints in the reference's row order (ms:171-255, 293-309).

    rows 0..S-1     position at t=0 of spline i       = w_i
    rows S..2S-1    position at t=T_i of spline i     = w_{i+1}
    rows 2S..2S+2   vel/acc/jerk at the start         = 0
    rows 2S+3..+5   vel/acc/jerk at the end           = 0
    then, per junction s=1..S-1 and k=1..4:  d^k/dt^k spline s-1 (T_{s-1}) - d^k/dt^k spline s (0) = 0
    """
    w = np.asarray(waypoints, dtype=float)
    S = w.shape[0] - 1
    m = 2 * S + 6 + 4 * (S - 1)
    A = np.zeros((m, N_COEFFS * S))
    b = np.zeros((m, w.shape[1]))
    r = 0
    p0 = basis_row(0, 0.0)
    for i in range(S):
        A[r, 8 * i:8 * i + 8] = p0
        b[r] = w[i]
        r += 1
    for i in range(S):
        A[r, 8 * i:8 * i + 8] = basis_row(0, T[i])
        b[r] = w[i + 1]
        r += 1
    for k in (1, 2, 3):
        A[r, 0:8] = basis_row(k, 0.0)
        r += 1
    for k in (1, 2, 3):
        A[r, 8 * (S - 1):8 * S] = basis_row(k, T[S - 1])
        r += 1
    for s in range(1, S):
        for k in (1, 2, 3, 4):
            A[r, 8 * (s - 1):8 * s] = basis_row(k, T[s - 1])
            A[r, 8 * s:8 * s + 8] = -basis_row(k, 0.0)
            r += 1
    assert r == m
    return A, b


def snap_hessian(T: np.ndarray) -> np.ndarray:
    """Block-diagonal Hessian of the integral of snap^2, no factor 2 (ms:155-169)."""
    S = len(T)
    Q = np.zeros((8 * S, 8 * S))
    for s, dur in enumerate(T):
        for r in range(4, 8):
            fr = r * (r - 1) * (r - 2) * (r - 3)
            for c in range(4, 8):
                fc = c * (c - 1) * (c - 2) * (c - 3)
                e = r + c - 7
                Q[8 * s + r, 8 * s + c] = fr * fc * dur ** e / e
    return Q


def kkt_system(waypoints: np.ndarray, T: np.ndarray):
    """K = [[Q, A^T], [A, 0]], rhs = [0; b] (ms:138-146)."""
    A, b = constraint_system(waypoints, T)
    Q = snap_hessian(T)
    m = A.shape[0]
    K = np.block([[Q, A.T], [A, np.zeros((m, m))]])
    rhs = np.vstack((np.zeros((Q.shape[0], b.shape[1])), b))
    return K, rhs


def solve_coeffs(waypoints: np.ndarray, velocity: float, method: str = "solve"):
    """coeffs (8S, dims) with row 8i+j = coefficient of t^j of spline i, and times (S,) (ms:138-153)."""
    T = segment_times(waypoints, velocity)
    K, rhs = kkt_system(waypoints, T)
    if method == "lstsq":
        sol = np.linalg.lstsq(K, rhs, rcond=None)[0]
    else:
        sol = np.linalg.solve(K, rhs)
    return sol[:8 * len(T)], T


def sample_counts(T: np.ndarray, dt: float) -> np.ndarray:
    """len(np.arange(0, T_i, dt)) = ceil(T_i / dt) with the division in fp64 (ms:104)."""
    return np.array([len(np.arange(0.0, Ti, dt)) for Ti in T], dtype=np.int64)


def yaw_profile(velocities: np.ndarray) -> np.ndarray:
    """Heading of the horizontal velocity with hold-last-valid and unwrap (ms:126-136)."""
    v = np.asarray(velocities, dtype=float)
    speed = np.sqrt(v[:, 0] ** 2 + v[:, 1] ** 2)
    valid = np.flatnonzero(speed >= MIN_HORIZONTAL_SPEED_FOR_YAW)
    if valid.size == 0:
        return np.zeros(len(v))
    yv = np.unwrap(np.arctan2(v[valid, 1], v[valid, 0]))
    prev = np.searchsorted(valid, np.arange(len(v)), side="right") - 1
    prev = np.clip(prev, 0, valid.size - 1)
    return yv[prev]


def sample_table(coeffs: np.ndarray, T: np.ndarray, dt: float) -> np.ndarray:
    """(N, 11) table [pos3, vel3, acc3, yaw, spline_id] (ms:97-124)."""
    rows = []
    for i, Ti in enumerate(T):
        c = coeffs[8 * i:8 * i + 8]
        for t in np.arange(0.0, Ti, dt):
            rows.append(np.concatenate((basis_row(0, t) @ c, basis_row(1, t) @ c, basis_row(2, t) @ c, [0.0, i])))
    tab = np.asarray(rows)
    tab[:, 9] = yaw_profile(tab[:, 3:6])
    return tab


def point_in_cuboid(x, y, z, cuboid) -> bool:
    """Inclusive point-in-AABB, cuboid = [xmin,xmax,ymin,ymax,zmin,zmax] (ms:327-357)."""
    return bool(cuboid[0] <= x <= cuboid[1] and cuboid[2] <= y <= cuboid[3] and cuboid[4] <= z <= cuboid[5])


def insert_midpoints(points: np.ndarray, indexes) -> np.ndarray:
    """Insert (p[i-1]+p[i])/2 before every i in ``indexes`` (ms:359-391)."""
    out = []
    for i in range(len(points)):
        if i in indexes:
            out.append((points[i - 1] + points[i]) / 2)
        out.append(points[i])
    return np.array(out)


def plan_table(waypoints: np.ndarray, obstacles, velocity: float, dt: float, method: str = "lstsq", max_waypoints=None):
    """``MinimumSnap(...).get_trajectory()`` incl. the per-obstacle midpoint loop (ms:59-95).

    Returns (table, waypoints_after_insertion, coeffs, T).  ``obstacles=None`` skips the loop; an
    empty obstacle array returns ``None`` for the table exactly like the reference (the loop body
    never runs).  ``max_waypoints``: raise RuntimeError once the mission has grown past it (the reference
    itself never stops when a box contains a waypoint or keeps catching the inserted midpoints).
    """
    w = np.asarray(waypoints, dtype=float)

    def gen(wp):
        c, T = solve_coeffs(wp, velocity, method)
        return sample_table(c, T, dt), c, T

    if obstacles is None:
        tab, c, T = gen(w)
        return tab, w, c, T
    tab = c = T = None
    for box in obstacles:
        tab, c, T = gen(w)
        while True:
            inside = ((box[0] <= tab[:, 0]) & (tab[:, 0] <= box[1]) & (box[2] <= tab[:, 1]) & (tab[:, 1] <= box[3])
                      & (box[4] <= tab[:, 2]) & (tab[:, 2] <= box[5]))                 # point_in_cuboid on every row
            hit = {int(sid) + 1 for sid in np.unique(tab[inside, 10])}
            if not hit:
                break
            w = insert_midpoints(w, hit)
            if max_waypoints is not None and len(w) > max_waypoints:
                raise RuntimeError(f"correction loop grew past {max_waypoints} waypoints")
            tab, c, T = gen(w)
    return tab, w, c, T


def mission_table(waypoints: np.ndarray, obstacles, velocity: float, dt: float, method: str = "lstsq"):
    """Vertical take-off table followed by the course table (main.py:73-84)."""
    w = np.asarray(waypoints, dtype=float)
    tk = plan_table(w[:2], obstacles, velocity, dt, method)[0]
    co = plan_table(w[1:], obstacles, velocity, dt, method)[0]
    return np.vstack((tk, co))
