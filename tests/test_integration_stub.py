"""The ctypes stub printed in INTEGRATION.md (what a maintainer of the reference would add as uav_ac/batched.py) is
extracted from the document and executed: struct mirrors must have the C layout, and on a GPU box `fly_batch` must
return the same metrics as the library's own Python path."""
import ctypes
import os
import re
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_stub():
    from uav_ac_b200 import _native as nat
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(\"\"\"Fly B Monte-Carlo.*?)```", text, flags=re.S).group(1)
    code = code.replace('C.CDLL("libuavb.so")', f'C.CDLL({nat.LIB_PATH!r})')
    mod = types.ModuleType("uav_ac_batched_stub")
    exec(compile(code, "INTEGRATION.md:uav_ac/batched.py", "exec"), mod.__dict__)
    return mod, nat


def _fake_simulation():
    """What MujocoSimulation exposes for lab_course.xml (mujoco_sim.py:258-325), built without MuJoCo."""
    from oracle import flight_np
    from uav_ac_b200.simulation.scene import LAB_COURSE_OBSTACLES, LAB_COURSE_WAYPOINTS
    v = flight_np.Vehicle()
    quad = types.SimpleNamespace(g=v.g, dt=v.dt, m=v.mass, l=v.arm, kf=v.kf, kappa=v.kappa, i_x=v.inertia[0], i_y=v.inertia[1], i_z=v.inertia[2],
                                 min_thrust=v.min_thrust, max_thrust=v.max_thrust, motor_rise_time_constant=v.tau_rise,
                                 motor_fall_time_constant=v.tau_fall, max_ascent_rate=v.max_ascent, max_descent_rate=v.max_descent,
                                 max_speed_xy=v.max_speed_xy, max_horiz_accel=v.max_horiz_accel, max_tilt_angle=v.max_tilt,
                                 **{n: getattr(v, n) for n in v.GAIN_NAMES})
    return types.SimpleNamespace(quad=quad, mission_waypoints=LAB_COURSE_WAYPOINTS, obstacles=LAB_COURSE_OBSTACLES)


def test_stub_struct_mirrors_match_the_library():
    stub, nat = _load_stub()
    assert ctypes.sizeof(stub.Vehicle) == ctypes.sizeof(nat.Vehicle) and ctypes.sizeof(stub.Mission) == ctypes.sizeof(nat.MissionHost)
    assert [f[0] for f in stub.Vehicle._fields_] == [f[0] for f in nat.Vehicle._fields_]
    assert [f[0] for f in stub.Mission._fields_] == [f[0] for f in nat.MissionHost._fields_]
    v = stub.vehicle_from_quad(_fake_simulation().quad)
    d = nat.default_vehicle()
    assert bytes(v) == bytes(d)                                       # the Quad of lab_course.xml is the library's default vehicle


@pytest.mark.gpu
def test_stub_fly_batch_matches_the_library_path(cuda):
    import torch
    from uav_ac_b200 import host_api
    from uav_ac_b200.simulation.scene import LAB_COURSE_OBSTACLES, LAB_COURSE_WAYPOINTS
    stub, nat = _load_stub()
    B = 64
    rng = np.random.default_rng(0)
    gs, ms, ins = rng.uniform(0.8, 1.2, (B, 11)), rng.uniform(0.9, 1.1, B), rng.uniform(0.9, 1.1, (B, 3))
    metrics, n_ticks = stub.fly_batch(_fake_simulation(), 3.0, 10, B, gain_scale=gs, mass_scale=ms, inertia_scale=ins)
    v = nat.default_vehicle()
    want, _, n2 = host_api.fly_mission_host(LAB_COURSE_WAYPOINTS, 3.0, B, obstacles=LAB_COURSE_OBSTACLES,
                                            mc_gains=(np.array(list(v.gains))[None] * gs).T.astype(np.float32),
                                            mc_mass=(v.mass * ms).astype(np.float32), mc_inertia=(np.array(list(v.inertia))[None] * ins).T.astype(np.float32))
    assert n_ticks == n2 == 10760 and np.array_equal(metrics, want)
    assert (metrics[:, 0] < 0.5).all() and (metrics[:, 1] == 0).all()
