"""C-ABI boundary and host-side logic on a GPU-less machine (no compute calls need a device).

* libuavb.so loads and exports every symbol include/uavb.h declares (and nothing in the header is
  missing from the ctypes table);
* the ctypes mirrors of the ABI structs have the C layout;
* argument validation answers UAVB_EINVAL with a message, and -- when no CUDA device is visible --
  every compute entry point fails loudly with UAVB_ENODEVICE: there is no CPU path behind the ABI;
* the product package never imports the oracle.
"""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def nat():
    from uav_ac_b200 import _native
    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _native


def _header_functions():
    src = open(os.path.join(ROOT, "include", "uavb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(uavb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(nat):
    declared = _header_functions()
    assert len(declared) >= 17
    assert sorted(nat.SYMBOLS) == declared                       # the ctypes table and the header agree
    L = nat.lib()
    for name in declared:
        assert getattr(L, name) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", nat.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (uavb_[a-z0-9_]+)", out))
    assert set(declared) <= exported


def test_version_and_error_string(nat):
    L = nat.lib()
    assert L.uavb_version() == nat.ABI_VERSION == 210
    import uav_ac_b200
    assert uav_ac_b200.__version__ == "0.2.1"
    assert isinstance(L.uavb_last_error(), bytes)
    assert L.uavb_device_count() >= 0


def test_struct_layouts_match_the_header(nat, tmp_path):
    """Compile a 10-line C program against include/uavb.h and compare sizeof/offsetof with the ctypes mirrors."""
    probe = tmp_path / "probe.c"
    probe.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "uavb.h"\nint main(void){'
                     'printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(uavb_vehicle), sizeof(uavb_rollout_args), sizeof(uavb_stage_args),'
                     'offsetof(uavb_vehicle, gains), offsetof(uavb_rollout_args, veh), offsetof(uavb_rollout_args, seg_coeffs),'
                     'offsetof(uavb_rollout_args, log_out), offsetof(uavb_stage_args, euler_out));return 0;}\n')
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(probe), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(nat.Vehicle), ctypes.sizeof(nat.RolloutArgs), ctypes.sizeof(nat.StageArgs), nat.Vehicle.gains.offset,
            nat.RolloutArgs.veh.offset, nat.RolloutArgs.seg_coeffs.offset, nat.RolloutArgs.log_out.offset, nat.StageArgs.euler_out.offset]
    assert got == want


def test_vehicle_defaults_are_the_lab_course_constants(nat):
    """lab_course.xml:3,8-13,100,116-119 and Quad.__init__ gains (quad.py:54-73); SURVEY appendix A."""
    v = nat.default_vehicle()
    assert (v.g, v.dt, v.mass) == (9.81, 0.001, 0.5)
    assert list(v.inertia) == [0.0023, 0.0023, 0.0046]
    assert (v.arm, v.kf, v.kappa, v.min_thrust, v.max_thrust) == (0.120208, 1.0, 0.016, 0.1, 4.5)
    assert (v.tau_rise, v.tau_fall) == (0.0125, 0.025)
    assert (v.max_ascent, v.max_descent, v.max_speed_xy, v.max_horiz_accel, v.max_tilt) == (3.0, 2.0, 3.0, 12.0, 0.7)
    assert list(v.gains) == pytest.approx([16.0, 7.0, 25.0, 8.0, 0.1, 1 / 0.07, 1 / 0.07, 4.0, 125.0, 125.0, 1 / 0.09], rel=1e-15)
    assert v.integral_limit == 10.0
    from oracle import flight_np
    o = flight_np.Vehicle()
    assert list(v.gains) == [getattr(o, n) for n in o.GAIN_NAMES]


def test_argument_validation_and_no_cpu_path(nat):
    import torch
    L = nat.lib()
    null = ctypes.c_void_p(None)
    one = ctypes.c_void_p(8)                                     # non-NULL dummy; validation happens before any dereference
    assert L.uavb_minsnap_solve_f64(null, null, 1, 4, 1.5, null, null, null, null) == -1
    assert b"NULL" in L.uavb_last_error()
    assert L.uavb_minsnap_solve_f64(one, one, 1, 0, 1.5, one, one, null, null) == -1          # S out of range
    assert L.uavb_minsnap_solve_f64(one, one, 1, 65, 1.5, one, one, null, null) == -1
    assert L.uavb_rollout_f32(None, null) == -1
    a = nat.RolloutArgs()
    a.B, a.n_ticks, a.inner_per_outer = 4, 10, 0
    assert L.uavb_rollout_f32(ctypes.byref(a), null) == -1 and b"inner_per_outer" in L.uavb_last_error()
    s = nat.StageArgs()
    s.B, s.stage = 1, 99
    assert L.uavb_stage_f32(ctypes.byref(s), null) == -1
    # round-2 entry points: the correction loop, the shared planner, the constraint system
    assert L.uavb_minsnap_correct_f64(one, one, one, 1, 66, 1.5, 0.01, null, 0, 0, one, one, one, None, null) == -1      # max_wp > UAVB_MAX_SPLINES + 1
    assert b"max_wp" in L.uavb_last_error()
    assert L.uavb_minsnap_correct_f64(null, one, one, 1, 6, 1.5, 0.01, null, 0, 0, one, one, one, None, null) == -1 and b"NULL" in L.uavb_last_error()
    assert L.uavb_minsnap_correct_f64(one, one, one, 1, 6, 1.5, 0.01, null, 2, 0, one, one, one, None, null) == -1       # n_obs > 0 without cuboids
    assert L.uavb_minsnap_correct_f64(one, one, one, 1, 6, 1.5, 0.0, null, 0, 0, one, one, one, None, null) == -1        # dt <= 0
    assert L.uavb_minsnap_correct_f64(one, one, one, 1, 6, 1.5, 0.01, one, 2, 6, one, one, one, None, null) == -1        # cuboid_stride < 6 n_obs
    assert L.uavb_minsnap_constraints_f64(one, one, 1, 0, one, one, null) == -1
    n_seg = ctypes.c_int()
    assert L.uavb_plan_shared_f64(0, None, None, one, 1.5, 0.01, null, 0, 8, one, one, one, one, one, ctypes.byref(n_seg), None, None, None, None, null) == -1
    a = nat.RolloutArgs()
    a.B, a.n_ticks, a.inner_per_outer, a.n_seg_shared, a.dt_outer, a.log_stride = 4, 10, 10, 1, 0.01, 1
    a.veh = nat.default_vehicle()
    for f in ("seg_coeffs", "seg_rows", "seg_table", "seg_yaw0", "start", "log_out", "traj_out"):
        setattr(a, f, one)
    a.traj_max_samples = 4
    assert L.uavb_rollout_f32(ctypes.byref(a), null) == -1 and b"traj_out" in L.uavb_last_error()                # flown-path list together with a state log
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is visible: the ENODEVICE leg only applies to GPU-less hosts")
    assert L.uavb_device_count() == 0
    assert L.uavb_minsnap_solve_f64(one, one, 1, 4, 1.5, one, one, null, null) == -3          # UAVB_ENODEVICE
    assert b"no CPU path" in L.uavb_last_error()
    assert L.uavb_minsnap_solve_f64_host(one, one, 1, 4, 1.5, one, one, null) == -3
    assert L.uavb_mc_uniform_f32(1, 0, 0, 4, 1, one, one, one, null) == -3
    fp32, fp64 = ctypes.c_double(), ctypes.c_double()
    assert L.uavb_measure_fma_peak(0, ctypes.byref(fp32), ctypes.byref(fp64)) == -3
    from uav_ac_b200 import kernels
    with pytest.raises(nat.UavbError):
        kernels.minsnap_solve(torch.zeros((1, 5, 3), dtype=torch.float64), torch.ones(1, dtype=torch.float64))
    with pytest.raises(nat.UavbError):
        kernels.mc_missions(1, 8)


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "uav-autonomous-control_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"
                assert "oracle/" not in text.replace("oracle/freebody.py states", "") or f.endswith((".cuh", ".cu")), f
    code = ("import sys; sys.path.insert(0, %r); import uav_ac_b200, uav_ac_b200.kernels, uav_ac_b200.sharding; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True)


def test_config_ini_keeps_the_reference_keys():
    """uav_ac/config.ini:2,7,9 verbatim: frequency, velocity, min_dist_target."""
    import configparser
    cfg = configparser.ConfigParser(inline_comment_prefixes="#")
    cfg.read(os.path.join(ROOT, "uav-autonomous-control_b200", "config.ini"))
    assert cfg["DEFAULT"].getint("frequency") == 10
    assert cfg["SIM_FLIGHT"].getfloat("velocity") == 3.0 and cfg["SIM_FLIGHT"].getfloat("min_dist_target") == 0.5
