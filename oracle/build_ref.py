"""Recipe for oracle/_ref: the REFERENCE's own hot-path classes, byte-compiled where they lie.  TEST INFRASTRUCTURE ONLY.

    python oracle/build_ref.py            (called by __graft_entry__.build(); needs /root/reference)

The reference is pure Python, so "building" it means compiling its modules to sourceless byte-code (the .pyc format, written
as `oracle/_ref/uav_ac/<pkg>/<module>.bin` -- not `*.pyc`, which file synchronisers and .gitignore rules routinely drop -- and
imported through a small finder in oracle/ref_arm.py).  No reference SOURCE is written into this
repository: only compiler output goes to oracle/_ref/, which is git-ignored (not gpurun-ignored, so it travels to the GPU
box like the built .so files).  Compiled: uav_ac/{__init__,main,utils}.py, control/controller.py, planning/minimum_snap.py,
planning/rrt.py, quadrotor/quad.py, simulation/{__init__,mujoco_sim}.py -- everything `uav_ac.main` imports.  The one
third-party piece, `mujoco` (not installed here or on the box), is stubbed at import time by oracle/ref_arm.py exactly as
the golden generator does (tests/golden/make_golden.py); its rigid-body step is oracle/freebody.py.

Used by: bench.py (`cpu_baseline` and `--impl reference`, kind "reference") and tests/test_ref_arm.py.
"""
from __future__ import annotations

import hashlib
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
MODULES = ["uav_ac/__init__.py", "uav_ac/main.py", "uav_ac/utils.py", "uav_ac/control/__init__.py", "uav_ac/control/controller.py",
           "uav_ac/planning/__init__.py", "uav_ac/planning/minimum_snap.py", "uav_ac/planning/rrt.py", "uav_ac/quadrotor/__init__.py",
           "uav_ac/quadrotor/quad.py", "uav_ac/simulation/__init__.py", "uav_ac/simulation/mujoco_sim.py"]


def available() -> bool:
    return os.path.exists(os.path.join(OUT, "STAMP")) and all(os.path.exists(os.path.join(OUT, m[:-3] + ".bin")) for m in MODULES)


def build(force: bool = False) -> str | None:
    """Compile the reference modules into oracle/_ref (None when /root/reference is absent and nothing was built before)."""
    if not os.path.isdir(os.path.join(REF, "uav_ac")):
        return OUT if available() else None
    h = hashlib.sha256(sys.version.encode())
    for m in MODULES:
        with open(os.path.join(REF, m), "rb") as f:
            h.update(m.encode()); h.update(f.read())
    stamp = os.path.join(OUT, "STAMP")
    if not force and os.path.exists(stamp) and open(stamp).read().split()[0] == h.hexdigest() and all(
            os.path.exists(os.path.join(OUT, m[:-3] + ".bin")) for m in MODULES):
        return OUT
    for m in MODULES:
        dst = os.path.join(OUT, m[:-3] + ".bin")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile: the path shown in tracebacks points back at the reference, not at a file of this repository
        py_compile.compile(os.path.join(REF, m), cfile=dst, dfile=os.path.join(REF, m), doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(stamp, "w") as f:
        f.write(f"{h.hexdigest()} python {sys.version_info.major}.{sys.version_info.minor} from {REF}\n")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
