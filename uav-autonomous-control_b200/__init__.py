"""B200-native batched closed-loop flight path of UAV-Autonomous-control.

Drop-in (batched) counterparts of the reference's hot-path classes -- ``MinimumSnap``,
``CascadedController``, ``Quad``, ``TrajectoryController``, ``BatchedSimulation`` (in place of
``MujocoSimulation``) and ``utils.get_config`` -- over hand-written sm_100a CUDA kernels behind the
C ABI of ``include/uavb.h``.  Import as ``uav_ac_b200``; the module layout mirrors ``uav_ac``:

    uav_ac_b200.planning.minimum_snap.MinimumSnap      uav_ac_b200.control.controller.CascadedController
    uav_ac_b200.quadrotor.quad.Quad                    uav_ac_b200.main.TrajectoryController
    uav_ac_b200.simulation.batched_sim.BatchedSimulation   uav_ac_b200.planning.rrt.RRTStar

There is no CPU fallback: the kernels are the only implementation, and calls raise when libuavb.so
or a CUDA device is missing.
"""
__version__ = "0.2.1"            # = UAVB_VERSION 210 of include/uavb.h (checked in tests/test_abi_and_host.py)
