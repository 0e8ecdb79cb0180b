"""Multi-GPU host logic on CPU: shard ranges and the world_size-2 metric gather over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from uav_ac_b200 import sharding


def test_shard_ranges_partition_the_index_space():
    for total in (0, 1, 7, 100_000, 10_000_001):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_summarize_reports_the_reference_mission_outcome():
    """main.py:115-120 for a batch: reached iff final distance < min_dist_target, collision flag, failed rollouts excluded."""
    m = torch.zeros((4, 8))
    m[:, 0] = torch.tensor([0.1, 0.6, 0.2, 0.3])
    m[2, 1] = 1.0
    m[3, 5] = 1.0                                               # non-finite rollout
    m[:, 2] = torch.tensor([0.1, 0.2, 0.3, 9.0])
    s = sharding.summarize(m, 0.5)
    assert s["rollouts"] == 4 and s["reached_fraction"] == 0.5 and s["collision_fraction"] == 0.25 and s["failed_fraction"] == 0.25
    assert s["mean_tracking_rmse"] == pytest.approx(0.2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, _, w = sharding.init_from_env("gloo")
    b, e = sharding.shard_range(total, r, w)
    idx = torch.arange(b, e, dtype=torch.float32)
    local = torch.stack([idx * (k + 1) for k in range(8)], dim=1)           # row i = global index i times (k+1)
    full = sharding.gather_metrics(local, total)
    want = torch.arange(total, dtype=torch.float32)[:, None] * torch.arange(1, 9, dtype=torch.float32)[None, :]
    q.put((rank, bool(torch.equal(full, want)), tuple(full.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 11])
def test_gather_metrics_world_size_2_gloo(total):
    """Uneven shards (11 = 6 + 5) are padded for the fixed-size all_gather and trimmed afterwards;
    every rank ends with the global [total, 8] tensor in global rollout order."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, (total, 8)), (1, True, (total, 8))]


def _worker_async(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, _, w = sharding.init_from_env("gloo")
    b, e = sharding.shard_range(total, r, w)
    g = sharding.MetricGather(e - b, 8, total, torch.device("cpu"))
    ok = True
    for k in range(5):                                                        # five batches back to back through two buffers
        local = g.local(k)
        idx = torch.arange(b, e, dtype=torch.float32) + 1000.0 * k
        local.copy_(torch.stack([idx * (c + 1) for c in range(8)], dim=1))
        g.launch(k)
        if k >= 1:                                                            # batch k-1 is read while batch k is in flight
            want = (torch.arange(total, dtype=torch.float32) + 1000.0 * (k - 1))[:, None] * torch.arange(1, 9, dtype=torch.float32)[None, :]
            ok &= bool(torch.equal(g.result(k - 1), want))
    want = (torch.arange(total, dtype=torch.float32) + 4000.0)[:, None] * torch.arange(1, 9, dtype=torch.float32)[None, :]
    ok &= bool(torch.equal(g.result(4), want))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 11])
def test_async_double_buffered_gather_world_size_2_gloo(total):
    """MetricGather: equal shards gather straight into place asynchronously, unequal ones fall back to the padded gather;
    a batch's result stays intact while the next batch is being written and gathered."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_async, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
