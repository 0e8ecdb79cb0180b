"""Method-level calls into ``uavb_stage_f32`` (include/uavb.h) for the batched drop-in classes.

User-facing tensors are batch-first like the reference arrays with a leading batch dimension
(``X`` is (B, 13), a rotation matrix (B, 3, 3)); the kernel wants SoA with the drone index fastest,
so operands are transposed on the device per call.  This is the unit-granularity API -- the
throughput path is ``kernels.rollout`` (K2), which never leaves registers between ticks.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _native as nat

_F32 = torch.float32


def as_batch(x, B: int, width: Optional[int], device) -> torch.Tensor:
    """Scalar / (width,) / (B,) / (B, width) -> float32 (B,) or (B, width) on the device."""
    t = torch.as_tensor(x, dtype=_F32, device=device) if not (isinstance(x, torch.Tensor) and x.is_cuda) else x.to(_F32)
    if width is None:
        if t.dim() == 0:
            t = t.expand(B)
        if t.shape != (B,):
            raise ValueError(f"expected a scalar or shape ({B},), got {tuple(t.shape)}")
    else:
        if t.dim() == 1 and t.shape[0] == width:
            t = t.unsqueeze(0).expand(B, width)
        if t.shape != (B, width):
            raise ValueError(f"expected shape ({width},) or ({B}, {width}), got {tuple(t.shape)}")
    return t.contiguous()


def soa(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """(B, k) batch-first -> [k][B] contiguous."""
    if t is None:
        return None
    return t.t().contiguous() if t.dim() == 2 else t.contiguous()


def run(stage: int, quad, dt_outer: float, B: int, *, gains: Optional[dict] = None, **arrays) -> None:
    """Launch one stage.  ``dt_outer`` is CascadedController.dt (the integrator step of the ALTITUDE / OUTER stages; the vehicle
    stages never read it and pass their own dt).  ``arrays`` are SoA float32 device tensors keyed by the struct field names;
    ``gains`` overrides vehicle gains by name with floats or (B,) tensors (the reference passes gains
    as method arguments)."""
    a = nat.StageArgs()
    a.B, a.stage, a.dt_outer = int(B), int(stage), float(dt_outer)
    veh, mc = quad.vehicle_struct(B, gains or {})
    a.veh = veh
    keep = []
    for name, t in mc.items():
        setattr(a, name, nat.ptr(t, _F32, name))
        keep.append(t)
    dev = None
    for name, t in arrays.items():
        if name == "n_obs":
            a.n_obs = int(t)
            continue
        if t is None:
            continue
        setattr(a, name, nat.ptr(t, _F32, name))
        dev = t.device
        keep.append(t)
    with torch.cuda.device(dev):                      # the library works on the CURRENT device: make it the tensors' device
        nat.check(nat.lib().uavb_stage_f32(ctypes.byref(a), nat.stream_ptr(dev)), "uavb_stage_f32")
