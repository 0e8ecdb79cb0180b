"""Multi-GPU layout of the rollout path: one process per GPU, contiguous shards of the global rollout
index range, counter-based Monte-Carlo inputs keyed by the GLOBAL index, and one gather of the
per-rollout metrics at the end (the only collective -- rollouts never interact).

The reference runs exactly one drone in one process (uav_ac/main.py:87-120 prints final distance /
reached / collision for it); the gathered [B, 8] metrics tensor is the batched form of that report.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `total` units owned by `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("rank/world out of range")
    base, rem = divmod(int(total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world) from torchrun's environment; initialises the default group when world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"), rank=rank, world_size=world)
    return rank, local, world


def gather_metrics(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather the per-rollout metric rows of every shard into the global [total, C] tensor.

    Shards produced by shard_range differ by at most one row, so the shorter ones are padded by one
    row for the fixed-size collective and the padding is dropped afterwards.
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    width = local.shape[1]
    longest = -(-int(total) // world)
    buf = local
    if local.shape[0] < longest:
        buf = torch.cat((local, local.new_zeros((longest - local.shape[0], width))), dim=0)
    out = local.new_empty((world * longest, width))
    dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    parts = []
    for r in range(world):
        b, e = shard_range(total, r, world)
        parts.append(out[r * longest:r * longest + (e - b)])
    return torch.cat(parts, dim=0)


class MetricGather:
    """Double-buffered, asynchronous form of gather_metrics for a job that flies several batches back to back.

    launch(k) enqueues the all-gather of shard buffer k % 2 on the communicator's own stream, ordered after everything already
    enqueued on the current stream, and returns at once: the next batch's kernels (which write the OTHER shard buffer) run while
    the gather is in flight.  local(k) hands out shard buffer k % 2 after making the current stream wait for the gather that last
    read it; result(k) returns the global [total, C] tensor of batch k once its gather is ordered before the current stream.
    Equal shards (the weak-scaling case) are gathered straight into place; unequal ones go through gather_metrics' padding.
    """

    def __init__(self, rows_local: int, width: int, total: int, device, dtype=torch.float32, group=None):
        self.group, self.total, self.width = group, int(total), int(width)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rows = int(rows_local)
        self.equal = self.world == 1 or self.rows * self.world == self.total
        self.shard = [torch.empty((self.rows, width), dtype=dtype, device=device) for _ in range(2)]
        self.out = [torch.empty((self.total, width), dtype=dtype, device=device) if self.world > 1 and self.equal else None for _ in range(2)]
        self.work = [None, None]

    def local(self, k: int) -> torch.Tensor:
        w = self.work[k % 2]
        if w is not None:
            w.wait()                                   # stream-ordered: the current stream waits, the host does not
            self.work[k % 2] = None
        return self.shard[k % 2]

    def launch(self, k: int) -> None:
        if self.world == 1:
            return
        if self.equal:
            self.work[k % 2] = dist.all_gather_into_tensor(self.out[k % 2], self.shard[k % 2], group=self.group, async_op=True)
        else:
            self.out[k % 2] = gather_metrics(self.shard[k % 2], self.total, self.group)

    def result(self, k: int) -> torch.Tensor:
        if self.world == 1:
            return self.shard[k % 2]
        w = self.work[k % 2]
        if w is not None:
            w.wait()
            self.work[k % 2] = None
        return self.out[k % 2]


def summarize(metrics: torch.Tensor, min_dist_target: float = 0.5) -> dict:
    """Mission report of main.py:115-120 over a batch: reached fraction, collision fraction, error stats."""
    m = metrics.double()
    ok = m[:, 5] == 0
    return {
        "rollouts": int(m.shape[0]),
        "reached_fraction": float(((m[:, 0] < min_dist_target) & ok).double().mean()),
        "collision_fraction": float((m[:, 1] > 0).double().mean()),
        "mean_final_dist": float(m[ok, 0].mean()) if bool(ok.any()) else float("nan"),
        "mean_tracking_rmse": float(m[ok, 2].mean()) if bool(ok.any()) else float("nan"),
        "failed_fraction": float((~ok).double().mean()),
    }
