"""Drive tests/devtools/host_probe (development numerics probe) against tests/golden closed-loop logs."""
import struct, subprocess, sys
import numpy as np
sys.path.insert(0, ".")  # run from the repo root
from oracle import minsnap_np as M

def mission_blob(wp, v, dt=0.01, method="solve"):
    segs = []
    for tab_wp in (wp[:2], wp[1:]):
        c, T = M.solve_coeffs(tab_wp, v, method)
        rows = M.sample_counts(T, dt)
        tab = M.sample_table(c, T, dt)
        sp = np.hypot(tab[:, 3], tab[:, 4])
        valid = np.flatnonzero(sp >= 1e-3)
        yaw0 = float(np.arctan2(tab[valid[0], 4], tab[valid[0], 3])) if len(valid) else 0.0
        for i in range(len(T)):
            segs.append((c[8 * i:8 * i + 8].ravel(), int(rows[i]), 1 if i == 0 else 0, yaw0))
    return segs

def write_blob(path, segs, start, n_ticks, lag=1, mc=None):
    mc = np.concatenate((np.ones(15), np.zeros(3))) if mc is None else np.asarray(mc, float)
    with open(path, "wb") as f:
        f.write(struct.pack("i", len(segs)))
        f.write(np.concatenate([s[0] for s in segs]).astype("<f8").tobytes())
        f.write(np.array([s[1] for s in segs], "<i4").tobytes())
        f.write(np.array([s[2] for s in segs], "<i4").tobytes())
        f.write(np.array([s[3] for s in segs], "<f8").tobytes())
        f.write(np.asarray(start, "<f8").tobytes())
        f.write(struct.pack("ii", n_ticks, lag))
        f.write(mc.astype("<f8").tobytes())

def quat_angle(qa, qb):
    # rotation angle of conj(qa)*qb from its vector part (well conditioned for small angles)
    a0, a1, a2, a3 = qa.T; b0, b1, b2, b3 = qb.T
    vx = a0 * b1 - a1 * b0 - a2 * b3 + a3 * b2
    vy = a0 * b2 + a1 * b3 - a2 * b0 - a3 * b1
    vz = a0 * b3 - a1 * b2 + a2 * b1 - a3 * b0
    return 2 * np.arcsin(np.sqrt(vx * vx + vy * vy + vz * vz).clip(0, 1))

if __name__ == "__main__":
    g = np.load("tests/golden/planning.npz")
    for v in (2.0, 3.0):
        cl = np.load(f"tests/golden/closed_loop_v{int(v)}.npz")
        segs = mission_blob(g["waypoints"], v)
        n_ticks = 10 * sum(s[1] for s in segs)
        write_blob("/tmp/mission.bin", segs, g["waypoints"][0], n_ticks)
        for mode in ("f64", "f32"):
            out = subprocess.run(["/tmp/host_probe", "/tmp/mission.bin", "/tmp/out.bin", mode], capture_output=True, text=True)
            log = np.fromfile("/tmp/out.bin", "<f8").reshape(-1, 17)
            X = cl["X"]
            dp = np.abs(log[:, 0:3] - X[:, 0:3]).max()
            dv = np.abs(log[:, 7:10] - X[:, 7:10]).max()
            dw = np.abs(log[:, 10:13] - X[:, 10:13]).max()
            da = quat_angle(log[:, 3:7], X[:, 3:7]).max()
            dom = np.abs(log[:, 13:17] - cl["omega"]).max()
            print(f"v={v} {mode}: rows {len(log)} pos {dp:.3e} m  vel {dv:.3e}  rate {dw:.3e}  att {da:.3e} rad  omega {dom:.3e} | {out.stdout.strip()}")
        print("   golden mean/rmse/max", float(cl["mean_err"]), float(cl["rmse"]), float(cl["max_err"]))
