#!/usr/bin/env python
"""Headline benchmark of the batched closed-loop flight path (BASELINE.json metric: drone-sim-steps/s).

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...          # the reference-style CPU path (oracle port) on the host cores

Workload (config.workload): BASELINE.json configs[2] -- 10^5 lab_course rollouts per GPU with
Monte-Carlo PID gains and mass/inertia perturbations, velocity 3.0 (config.ini) => 10 760 ticks each,
1.076e9 drone-sim-steps per GPU and step.  One "step" = one pass of the hot path over the batch:
min-snap solve of the mission (K1, take-off + course tables), table geometry, persistent rollout
(K2) of every drone over the whole mission, metrics written.  Weak scaling: every rank flies its own
10^5 rollouts (Monte-Carlo inputs keyed by the global rollout index) and the per-rollout metrics are
all-gathered over NCCL inside the timed region.

`value`  : steps/s with inputs resident in HBM (CUDA events around the K timed steps, max over ranks).
`e2e`    : same metric through the reference-facing C-ABI call uavb_fly_mission_host with HOST buffers: per step
           the waypoints and the Monte-Carlo arrays are copied from pinned host memory, the mission is planned and
           flown, the metrics are copied back and the call synchronises.
`roofline`: K2 is FP32-issue bound (no dense contraction => no tensor path, ~0 HBM bytes per tick in
           metrics-only mode); `achieved` = 269 algorithmic flop/tick x ticks / K2 time,
           `peak` = FP32 FMA rate measured in this run by uavb_measure_fma_peak.  `roofline_log` is the
           HBM roofline of the full-rate state-log mode (52 B/tick) against MEASURED_PEAKS.json.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_TICK = 269          # algorithmic flop per drone tick with 4 AABBs: 239 in the 1 kHz body + 298/10 from the 100 Hz loop (DESIGN.md "K2 work per tick")
K2_DRAM_BYTES_PER_LAUNCH = 8.65e6   # ncu: 6.18 MB read + 2.46 MB written per launch of the bench workload (metrics-only: ~0 B per tick)
LOG_BYTES_PER_TICK = 52      # 13 fp32 state words
ROLLOUTS_PER_GPU = 100_000   # BASELINE configs[2]
VELOCITY = 3.0               # config.ini:7
FREQUENCY = 10               # config.ini:2
METRIC = "drone-sim-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rollouts", type=int, default=ROLLOUTS_PER_GPU, help="rollouts per GPU (default: BASELINE configs[2])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the solves/s and log-mode side measurements")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="wall-clock budget of the CPU baseline sample")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ CPU arm
_TABLE_CACHE = {}


def lab_course_table(velocity):
    """(table, waypoints, obstacles) of the lab_course mission from the oracle planner (cached per process)."""
    if velocity not in _TABLE_CACHE:
        from oracle import minsnap_np
        from uav_ac_b200.simulation.scene import LAB_COURSE_OBSTACLES, LAB_COURSE_WAYPOINTS
        tab = minsnap_np.mission_table(LAB_COURSE_WAYPOINTS, LAB_COURSE_OBSTACLES, velocity, 0.01)
        _TABLE_CACHE[velocity] = (tab, LAB_COURSE_WAYPOINTS, LAB_COURSE_OBSTACLES)
    return _TABLE_CACHE[velocity]


def _cpu_worker(job):
    """Fly `ticks` ticks of a Monte-Carlo-perturbed lab_course mission with the NumPy oracle port."""
    seed, ticks = job
    import numpy as np
    from oracle import flight_np
    rng = np.random.default_rng(seed)
    veh = flight_np.Vehicle().perturbed(rng.uniform(0.8, 1.2, 11), rng.uniform(0.9, 1.1), rng.uniform(0.9, 1.1, 3))
    tab, wp, obs = lab_course_table(VELOCITY)
    t0 = time.perf_counter()
    flight_np.closed_loop(veh, tab, wp[0], obstacles=obs, goal=wp[-1], n_ticks=ticks)
    return ticks, time.perf_counter() - t0


def cpu_rollout_rate(seconds: float, cores: int | None = None):
    """Whole-machine rate of the oracle port (the reference's Python/NumPy style: one drone per
    process, small-array NumPy calls per tick) on a bounded sample of the same workload."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        ticks_probe = 400
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(i, ticks_probe) for i in range(cores)])       # also warms imports and the table
        per_core = ticks_probe / max(time.perf_counter() - t0, 1e-6)
        ticks = max(1000, int(per_core * seconds * 0.8))
        ticks = min(ticks, FREQUENCY * 1076)
        t0 = time.perf_counter()
        done = pool.map(_cpu_worker, [(1000 + i, ticks) for i in range(cores)])
        wall = time.perf_counter() - t0
    total = sum(d[0] for d in done)
    return total / wall, cores, f"{cores} processes x {ticks} ticks of Monte-Carlo lab_course rollouts (v={VELOCITY}), oracle/flight_np.py"


def cpu_rollout_rate_c(seconds: float):
    """The C twin of the oracle (oracle/oracle_c.c, -O2, one thread per host core) on the same bounded sample: what an
    optimised scalar CPU implementation of the path reaches on this host (reported next to the NumPy-style figure)."""
    import numpy as np
    from oracle import c_port, flight_np
    cores = os.cpu_count() or 1
    tab, wp, obs = lab_course_table(VELOCITY)
    rng = np.random.default_rng(7)
    def vehicles(n):
        return [flight_np.Vehicle().perturbed(rng.uniform(0.8, 1.2, 11), rng.uniform(0.9, 1.1), rng.uniform(0.9, 1.1, 3)) for _ in range(n)]
    t0 = time.perf_counter()
    c_port.closed_loop_batch(vehicles(cores), tab, wp[0], obstacles=obs, goal=wp[-1], threads=cores)
    per = max(time.perf_counter() - t0, 1e-4)                      # one full mission per core
    n = max(cores, min(4096, int(cores * seconds / per)))
    vs = vehicles(n)
    t0 = time.perf_counter()
    c_port.closed_loop_batch(vs, tab, wp[0], obstacles=obs, goal=wp[-1], threads=cores)
    wall = time.perf_counter() - t0
    return n * FREQUENCY * len(tab) / wall, cores, f"{n} whole Monte-Carlo lab_course missions (v={VELOCITY}) on {cores} threads, oracle/oracle_c.c"


def run_reference(args):
    """--impl reference: the reference-style CPU implementation (oracle port; the reference itself is
    Python + MuJoCo and cannot travel to the GPU box) on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = max(2.0, min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup)))
    vals = []
    for i in range(args.warmup + args.steps):
        v, cores, sample = cpu_rollout_rate(per_step)
        if i >= args.warmup:
            vals.append(v)
    value = statistics.mean(vals)
    ms = 1e3 * ROLLOUTS_PER_GPU * FREQUENCY * 1076 / value
    line = {
        "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "impl": "reference",
        "config": {"workload": workload_name(args.rollouts), "note": "bounded sample per step; ms_per_step extrapolated to the full batch"},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_name(rollouts):
    return (f"BASELINE configs[2]: {rollouts} lab_course rollouts per GPU, Monte-Carlo gains x U(0.8,1.2), mass/inertia x U(0.9,1.1), "
            f"v={VELOCITY} m/s, 10760 ticks each, 4 AABBs, metrics only")


# ------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).

    A thread polls NVML (nvidia-ml-py) every ~2 ms -- the timed region of the default run lasts tens of
    milliseconds, too short for `nvidia-smi -lms`; nvidia-smi is the fallback when NVML cannot be loaded."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
               "hw_power_brake_slowdown": 0x80}

    def __init__(self, index):
        import threading
        self.index, self.samples, self.stop_flag, self.thread = index, [], threading.Event(), None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(local):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if local < len(ids) and ids[local].isdigit():
                return int(ids[local])
        return local

    def _poll(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    why = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    why = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                self.samples.append((sm, why, pw))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        import threading
        if self.nvml is None:
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def stop(self):
        if self.nvml is None:
            return self._smi_once()
        self.stop_flag.set()
        self.thread.join(timeout=2)
        if not self.samples:
            return self._smi_once()
        reasons = set()
        for _, why, _ in self.samples:
            for name, bit in self.REASONS.items():
                if why & bit:
                    reasons.add(name)
        return {"sm_mhz": statistics.median(s[0] for s in self.samples), "sm_max_mhz": float(self.max_sm), "reasons": sorted(reasons),
                "power_w_max": max(s[2] for s in self.samples), "samples": len(self.samples), "source": "nvml, 2 ms poll during the timed region"}

    def _smi_once(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                 capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
            p = [x.strip() for x in out.split(",")]
            reasons = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]) if v.lower().startswith("active")]
            return {"sm_mhz": float(p[0]), "sm_max_mhz": float(p[1]), "reasons": reasons, "power_w_max": float(p[2]), "samples": 1,
                    "source": "nvidia-smi, one sample right after the timed region (NVML unavailable)"}
        except Exception as e:  # noqa: BLE001
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"clock sampling unavailable: {e}"]}


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from uav_ac_b200 import _native as nat, host_api, kernels, sharding
    from uav_ac_b200.simulation.scene import LAB_COURSE_GOAL, LAB_COURSE_OBSTACLES, LAB_COURSE_START, LAB_COURSE_WAYPOINTS

    rank, local, world = sharding.init_from_env("nccl")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nat.lib()
    B = args.rollouts
    total = B * world
    begin, end = sharding.shard_range(total, rank, world)
    assert end - begin == B

    # ---- host inputs (pinned): waypoints, velocity, Monte-Carlo scales as a user would hand them over
    veh = nat.default_vehicle()
    base = np.array(list(veh.gains) + [veh.mass] + list(veh.inertia), dtype=np.float32)
    lo, hi = [0.8] * 11 + [0.9] * 4, [1.2] * 11 + [1.1] * 4
    scales = kernels.mc_uniform(20261017, B, lo, hi, index_base=begin, device=dev)          # keyed by the global rollout index
    mc_dev = (scales * torch.tensor(base, device=dev)[:, None]).contiguous()                # [15, B] fp32: 11 gains, mass, 3 inertia
    mc_host = mc_dev.cpu().pin_memory()
    wp_host = torch.tensor(LAB_COURSE_WAYPOINTS, dtype=torch.float64).pin_memory()
    vel_host = torch.tensor([VELOCITY], dtype=torch.float64).pin_memory()
    obs = torch.tensor(LAB_COURSE_OBSTACLES, dtype=torch.float32, device=dev)
    start = torch.tensor(LAB_COURSE_START, dtype=torch.float64, device=dev)
    goal = torch.tensor(LAB_COURSE_GOAL, dtype=torch.float64, device=dev)
    metrics_host = torch.empty((B, nat.N_METRICS), dtype=torch.float32).pin_memory()
    wp_dev = wp_host.to(dev)
    vel_dev = vel_host.to(dev)
    result = kernels.RolloutResult(torch.empty((B, nat.N_METRICS), dtype=torch.float32, device=dev), None, None, None)
    n_ticks_holder = {}

    def hot_path(wp, vel, mc):
        """K1 (two tables) + table geometry + K2 over the shard; returns the per-rollout metrics."""
        n_ticks = n_ticks_holder.get("n")
        plan = kernels.plan_missions([(wp[None, :2].contiguous(), vel), (wp[None, 1:].contiguous(), vel)], FREQUENCY * veh.dt, shared=True,
                                     table_rows=None if n_ticks is None else n_ticks // FREQUENCY)
        if n_ticks is None:                                  # mission length is data dependent: read it once, outside the timed steps
            n_ticks = n_ticks_holder["n"] = FREQUENCY * int(plan.total_rows.item())
        kernels.rollout(plan, B, n_ticks, start=start, goal=goal, vehicle=veh, frequency=FREQUENCY, mc_gains=mc[:11], mc_mass=mc[11],
                        mc_inertia=mc[12:15], obstacles=obs, want_state=False, out=result, index_base=begin)
        return result.metrics

    def step_device():
        m = hot_path(wp_dev, vel_dev, mc_dev)
        return sharding.gather_metrics(m, total) if world > 1 else m

    mc_mass_h, mc_inertia_h, mc_gains_h = mc_host[11], mc_host[12:15], mc_host[:11]     # contiguous pinned views, SoA
    wp_np = np.ascontiguousarray(LAB_COURSE_WAYPOINTS, dtype=np.float64)

    def step_e2e():
        """The reference-facing call: uavb_fly_mission_host (C ABI, HOST buffers).  Inside the call: H2D of the
        waypoints and the Monte-Carlo arrays, K1 x2, table geometry, K2, D2H of the metrics, synchronisation."""
        host_api.fly_mission_host(wp_np, VELOCITY, B, n_takeoff_waypoints=2, frequency=FREQUENCY, vehicle=veh, mc_mass=mc_mass_h,
                                  mc_inertia=mc_inertia_h, mc_gains=mc_gains_h, obstacles=LAB_COURSE_OBSTACLES, start=LAB_COURSE_START,
                                  goal=LAB_COURSE_GOAL, metrics_out=metrics_host)

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)                    # > 126 MB L2

    def timed(fn, steps, warmup, sampler=None, wall=False):
        """K timed steps between barriers + synchronize; device time from CUDA events on the launching (current torch)
        stream, or host wall-clock for the synchronous host-buffer call (wall=True); max over ranks."""
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        wall_ms = 0.0
        for s in range(steps):
            flush.fill_(s & 0xFF)                                                           # flush L2 between timed iterations
            if wall:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                fn()
                wall_ms += (time.perf_counter() - t0) * 1e3
            else:
                ev[s][0].record()
                fn()
                ev[s][1].record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        ms = wall_ms if wall else sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), clocks

    W = max(args.warmup, 3)
    sampler = ClockSampler(local) if rank == 0 else None
    ms_dev, clocks = timed(step_device, args.steps, W, sampler)
    n_ticks = n_ticks_holder["n"]
    ms_e2e, _ = timed(step_e2e, args.steps, 2, wall=True)
    torch.cuda.synchronize()
    sim_steps = float(total) * n_ticks                                                       # whole job, one step
    value = sim_steps * args.steps / (ms_dev * 1e-3)
    e2e = sim_steps * args.steps / (ms_e2e * 1e-3)

    # ---- K2 alone: average launch duration with events on the launching stream
    plan = kernels.plan_missions([(wp_dev[None, :2].contiguous(), vel_dev), (wp_dev[None, 1:].contiguous(), vel_dev)], FREQUENCY * veh.dt, shared=True)
    kw = dict(start=start, goal=goal, vehicle=veh, frequency=FREQUENCY, mc_gains=mc_dev[:11], mc_mass=mc_dev[11], mc_inertia=mc_dev[12:15],
              obstacles=obs, want_state=False, out=result)
    k2 = []
    for i in range(2 + args.steps):
        flush_and_space(flush, i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); kernels.rollout(plan, B, n_ticks, **kw); b.record()
        torch.cuda.synchronize()
        if i >= 2:
            k2.append(a.elapsed_time(b))
    k2_ms = statistics.mean(k2)
    summary = sharding.summarize(result.metrics)

    line = None
    if rank == 0:
        fp32_peak, fp64_peak = nat.measure_fma_peak(local)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        achieved = FLOP_PER_TICK * float(B) * n_ticks / (k2_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(B), "rollouts_per_gpu": B, "ticks_per_rollout": n_ticks, "frequency": FREQUENCY,
                       "l2": "256 MB buffer written between timed iterations (L2 flush)", "fp64_parts": "K1 solve and 100 Hz set-point evaluation"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "steps/s", "h2d_bytes_per_step": int(mc_host.numel() * 4 + wp_host.numel() * 8 + 8),
                    "d2h_bytes_per_step": int(metrics_host.numel() * 4), "ms_per_step": ms_e2e / args.steps,
                    "call": "uavb_fly_mission_host (C ABI, pinned host buffers, synchronous); host wall-clock, max over ranks"},
            "gpu_launches": 6 * args.steps,        # own kernels per step: 2x minsnap_solve, table_meta, target_rows + target_heading, rollout_sliced (torch glue not counted)
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak if fp32_peak else None,
                         "traffic": K2_DRAM_BYTES_PER_LAUNCH, "traffic_unit": "bytes per launch, dram read+write (ncu --set full, profiles/r01_ncu_rollout_v10.md)",
                         "kernel": "rollout_sliced_kernel<MC,TABLE>", "kernel_ms": k2_ms, "flop_per_tick": FLOP_PER_TICK,
                         "peak_source": "uavb_measure_fma_peak in this run (MEASURED_PEAKS.json carries no fp32 figure)",
                         "fp64_peak_tflops": fp64_peak, "ticks_per_s_k2": float(B) * n_ticks / (k2_ms * 1e-3)},
            "mission_report": summary,
        }
        if not args.no_extras:
            line["roofline_log"] = log_mode_roofline(kernels, plan, kw, dev, n_ticks, peaks, flush)
            line["solves"] = solve_rate(kernels, dev, flush, peaks)
    if world > 1:
        dist.barrier()
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            v, cores, sample = cpu_rollout_rate(args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample}
            try:
                vc, cc, sc = cpu_rollout_rate_c(min(args.cpu_seconds, 8.0))
                line["cpu_baseline_c"] = {"value": vc, "unit": "steps/s", "cores": cc, "kind": "port", "sample": sc,
                                          "note": "optimised C restatement; the reference itself is Python/NumPy (cpu_baseline)"}
            except Exception as e:  # noqa: BLE001  (no C compiler on the box: the NumPy figure stands alone)
                line["cpu_baseline_c"] = {"unavailable": str(e)}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def flush_and_space(flush, i):
    """L2 flush (256 MB written) repeated a few times: besides evicting L2 it keeps the GPU busy for ~0.3 ms, so the host has
    enqueued the start event, the launch and the stop event before the GPU reaches them -- the event pair then brackets the
    kernel(s) of the call, not the host's launch latency (which matters for the sub-millisecond kernels)."""
    for k in range(8):
        flush.fill_((i + k) & 0xFF)


def log_mode_roofline(kernels, plan, kw, dev, n_ticks, peaks, flush):
    """Full-rate state log (52 B/tick): HBM roofline of the logging epilogue (north_star)."""
    import torch
    Bl = 151552                                           # two full waves of 148 SMs x 8 CTAs x 64 drones
    ticks = 400                                           # 151552 x 400 x 52 B = 3.15 GB of log per launch
    from uav_ac_b200 import _native as nat
    kw = dict(kw)
    veh = nat.default_vehicle()
    base = torch.tensor(list(veh.gains) + [veh.mass] + list(veh.inertia), dtype=torch.float32, device=dev)[:, None]
    mc = (kernels.mc_uniform(20261017, Bl, [0.8] * 11 + [0.9] * 4, [1.2] * 11 + [1.1] * 4, device=dev) * base).contiguous()
    kw.update(mc_gains=mc[:11], mc_mass=mc[11], mc_inertia=mc[12:15])
    kw["out"] = None
    kw["want_metrics"] = False
    log = torch.empty((ticks, 13, Bl), dtype=torch.float32, device=dev)
    res = kernels.RolloutResult(None, None, log, None)
    kw["out"] = res
    ms = []
    for i in range(4):
        flush_and_space(flush, i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); kernels.rollout(plan, Bl, ticks, log_stride=1, **kw); b.record()
        torch.cuda.synchronize()
        if i >= 1:
            ms.append(a.elapsed_time(b))
    t = statistics.mean(ms) * 1e-3
    gbs = LOG_BYTES_PER_TICK * float(Bl) * ticks / t / 1e9
    peak = peaks.get("hbm_gbs", 6650.0)
    return {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": 3.18e9, "traffic_unit": "bytes per launch, dram read+write (ncu --set full, profiles/r01_ncu_rollout_log_v2.md)",
            "kernel": "rollout_sliced_kernel<MC,TABLE,LOG>", "kernel_ms": t * 1e3, "ticks_per_s": float(Bl) * ticks / t,
            "rollouts": Bl, "ticks": ticks,
            "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)",
            "note": "full-rate state log: rollouts x ticks x 52 B written per launch (write-only traffic against the read+write copy peak)"}


def solve_rate(kernels, dev, flush, peaks):
    """BASELINE configs[1]: 10^6 random 5-waypoint missions, K1 only (second half of the headline metric)."""
    import torch
    Bm = 1_000_000
    wp, vel = kernels.mc_missions(99, Bm, 4, device=dev)
    ms = []
    for i in range(6):
        flush_and_space(flush, i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); kernels.minsnap_solve(wp, vel); b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ms.append(a.elapsed_time(b))
    t = statistics.mean(ms) * 1e-3
    byts = 928.0 * Bm                                      # 128 B in + 800 B out per solve
    peak = peaks.get("hbm_gbs", 6650.0)
    return {"metric": "min-snap solves/s", "value": Bm / t, "unit": "solves/s", "missions": Bm, "splines": 4, "kernel_ms": t * 1e3,
            "roofline": {"bound": "hbm", "achieved": byts / t / 1e9, "peak": peak, "unit": "GB/s", "frac": byts / t / 1e9 / peak,
                         "traffic": None, "kernel": "minsnap_solve_kernel<4,kStagePair,6>", "note": "includes output allocation by torch (cached allocator)"}}


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, written to the process's original stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    a = parse()
    # Libraries (NCCL's version banner, torchrun warnings) may print to stdout; the contract is ONE JSON line there.
    # Everything else is sent to stderr for the lifetime of the run.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
