"""CPU oracle for the closed-loop flight hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (NumPy fp64 + a plain-C fp64 twin in ``oracle_c.c``) of the
reference algorithm for the path named in BASELINE.json:north_star:

    MinimumSnap solve  ->  cascaded controller  ->  rotor allocation / motor lag  ->  rigid-body step

It exists to CHECK the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``uav-autonomous-control_b200/`` imports, links or executes anything in here; the product path fails
loudly when the CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):

* planning (P1-P11), control (C0-C11), allocation, motor lag, scheduler: PINNED.  The restatement is
  compared with golden vectors produced by running the reference's own Python classes
  (``/root/reference/uav_ac/...``) in the build container -- ``tests/golden/make_golden.py`` is the
  generating script, ``tests/golden/*.npz`` the committed output -- and with every known-answer
  vector in the reference's own unit tests (``tests/unit/planning/test_minimum_snap.py``,
  ``tests/unit/control/test_controller.py``, ``tests/unit/quadrotor/test_quad.py``).
* rigid-body step (D1-D3, ``freebody.py``): PARITY UNPINNED.  The reference delegates it to the
  third-party MuJoCo 3.11.0 engine (``uv.lock:128-129``; call site
  ``uav_ac/simulation/mujoco_sim.py:147``), which is absent from ``/root/reference`` and cannot be
  installed here.  ``freebody.py`` restates MuJoCo's documented semi-implicit Euler update of one
  free-joint body and is anchored only on the reference tests that pin that boundary (hover
  invariance, gravity sign, frame conversion, integration-test thresholds).
"""
