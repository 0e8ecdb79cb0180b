"""Development probe: are rollout results independent of time slicing and of sharding?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from uav_ac_b200 import _native as nat, kernels
dev = torch.device("cuda", 0)

def c4_inputs(B, lo=0):
    wp, vel = kernels.mc_missions(31, B, 4, index_base=lo)
    g = wp[:, 0].clone(); g[:, 2] = -0.021
    plan = kernels.plan_missions([(torch.stack((g, wp[:, 0]), dim=1).contiguous(), vel), (wp, vel)], 0.01)
    wind = kernels.mc_uniform(32, B, [-0.08] * 3, [0.08] * 3, index_base=lo)
    return plan, g.contiguous(), wp[:, -1].contiguous(), wind

rng = np.random.default_rng(8)
ctr, half = rng.uniform([2, 2, -5], [22, 12, -1], (64, 6, 3)), rng.uniform(0.3, 1.2, (64, 6, 3))
boxes = torch.tensor(np.stack((ctr[..., 0] - half[..., 0], ctr[..., 0] + half[..., 0], ctr[..., 1] - half[..., 1], ctr[..., 1] + half[..., 1],
                               ctr[..., 2] - half[..., 2], ctr[..., 2] + half[..., 2]), axis=-1).astype(np.float32), device=dev)
B = 200_000
sets = (torch.arange(B, device=dev) % 64).to(torch.int32)
plan, start, goal, wind = c4_inputs(B)
n = 10 * int(plan.total_rows.max().item())
def fly(plan, B, start, goal, wind, sets, **env):
    for k in ("UAVB_ROLLOUT_CHUNKS",):
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ[k] = str(v)
    r = kernels.rollout(plan, B, n, start=start, goal=goal, mc_wind=wind, obstacles=boxes, obstacle_set=sets)
    torch.cuda.synchronize()
    return r
ref = fly(plan, B, start, goal, wind, sets, UAVB_ROLLOUT_CHUNKS=1)
for env in (dict(UAVB_ROLLOUT_CHUNKS=3), dict(UAVB_ROLLOUT_CHUNKS=40), dict()):
    r = fly(plan, B, start, goal, wind, sets, **env)
    print(env, "metrics equal", bool(torch.equal(r.metrics, ref.metrics)), "state equal", bool(torch.equal(r.state, ref.state)),
          "max |dmetrics|", float((r.metrics - ref.metrics).abs().max()))
lo, m = 120_000, 512
plan2, s2, g2, w2 = c4_inputs(m, lo)
sub = fly(plan2, m, s2, g2, w2, sets[lo:lo + m].contiguous())
d = (sub.metrics - ref.metrics[lo:lo + m]).abs()
print("reshard: metrics equal", bool(torch.equal(sub.metrics, ref.metrics[lo:lo + m])), "per-column max diff", d.max(dim=0).values.tolist())
print("  plan coeffs equal", bool(torch.equal(plan2.seg_coeffs, plan.seg_coeffs.reshape(B, -1, 8, 3)[lo:lo + m].reshape(-1, 8, 3))),
      "rows equal", bool(torch.equal(plan2.seg_rows, plan.seg_rows.reshape(B, -1)[lo:lo + m].reshape(-1))),
      "yaw0 equal", bool(torch.equal(plan2.seg_yaw0, plan.seg_yaw0.reshape(B, -1)[lo:lo + m].reshape(-1))),
      "wind equal", bool(torch.equal(w2, wind[:, lo:lo + m])), "start equal", bool(torch.equal(s2, start[lo:lo + m])))
