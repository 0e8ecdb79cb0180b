// stage_kernels.cu -- one controller / vehicle method for B drones per launch (unit-test granularity).
// The arithmetic is the same flight_core.cuh code the persistent rollout inlines; these kernels only
// move the operands through HBM so that the batched Python classes (CascadedController, Quad,
// TrajectoryController, BatchedSimulation) can expose the reference's method-level API.
#include "flight_core.cuh"
#include "uavb_common.cuh"
#include "veh_setup.cuh"

namespace uavb {

struct StageDev {
  uavb_stage_args a;
  VehU<float> u;
};

__device__ __forceinline__ void load_state(Drone<float>& d, const float* X, long long B, long long i) {
  d.px = (double)X[0 * B + i]; d.py = (double)X[1 * B + i]; d.pz = (double)X[2 * B + i];
  d.dx = d.dy = d.dz = 0.f;
  d.q0 = X[3 * B + i]; d.q1 = X[4 * B + i]; d.q2 = X[5 * B + i]; d.q3 = X[6 * B + i];
  // the reference normalises q before using it (quad.py:141, mujoco_sim.py:36-42); the rollout keeps q unit,
  // the stage entry points accept any q and normalise here
  const float rn = rsqrtf(d.q0 * d.q0 + d.q1 * d.q1 + d.q2 * d.q2 + d.q3 * d.q3);
  d.q0 *= rn; d.q1 *= rn; d.q2 *= rn; d.q3 *= rn;
  d.vx = X[7 * B + i]; d.vy = X[8 * B + i]; d.vz = X[9 * B + i];
  d.wx = X[10 * B + i]; d.wy = X[11 * B + i]; d.wz = X[12 * B + i];
  d.om0 = d.om1 = d.om2 = d.om3 = 0.f;
  d.integral = 0.f; d.thrust_cmd = 0.f; d.coll = 0.f; d.pc = d.qc = d.rc = 0.f;
  d.zbx = d.zby = 0.f; d.zbz = 1.f;
}

__global__ void __launch_bounds__(128) stage_kernel(const StageDev p) {
  const uavb_stage_args& a = p.a;
  const long long B = a.B;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  McValues mc;
  mc_from_vehicle(mc, a.veh);
  if (a.wind) { mc.wind[0] = a.wind[i]; mc.wind[1] = a.wind[B + i]; mc.wind[2] = a.wind[2 * B + i]; }
  VehP<float> v;
  make_vehp<float>(v, a.veh, mc);
  const VehU<float>& u = p.u;
  Drone<float> d;
  load_state(d, a.X, B, i);
  if (a.stage == UAVB_STAGE_OUTER) {
    Target t;
    t.x = a.target[0 * B + i]; t.y = a.target[1 * B + i]; t.z = a.target[2 * B + i];
    t.vx = a.target[3 * B + i]; t.vy = a.target[4 * B + i]; t.vz = a.target[5 * B + i];
    t.ax = a.target[6 * B + i]; t.ay = a.target[7 * B + i]; t.az = a.target[8 * B + i];
    {
      float sy, cy;
      sincosf(a.target[9 * B + i], &sy, &cy);
      t.yc = cy; t.ys = sy;
    }
    d.integral = a.integral[i];
    outer_update<float>(d, u, v, t);
    a.integral[i] = d.integral;
    a.thrust[i] = d.thrust_cmd;
    a.pqr_cmd[0 * B + i] = d.pc; a.pqr_cmd[1 * B + i] = d.qc; a.pqr_cmd[2 * B + i] = d.rc;
  } else if (a.stage == UAVB_STAGE_INNER) {
    set_thrust_cmd<float>(d, u, a.thrust[i]);
    d.pc = a.pqr_cmd[0 * B + i]; d.qc = a.pqr_cmd[1 * B + i]; d.rc = a.pqr_cmd[2 * B + i];
    d.om0 = a.omega[0 * B + i]; d.om1 = a.omega[1 * B + i]; d.om2 = a.omega[2 * B + i]; d.om3 = a.omega[3 * B + i];
    float gx, gy, gz, mom[3], f[4];
    inner_control<float>(d, u, v, &gx, &gy, &gz, mom, f);
    if (a.moment) { a.moment[0 * B + i] = mom[0]; a.moment[1 * B + i] = mom[1]; a.moment[2 * B + i] = mom[2]; }
    if (a.forces) { a.forces[0 * B + i] = f[0]; a.forces[1 * B + i] = f[1]; a.forces[2 * B + i] = f[2]; a.forces[3 * B + i] = f[3]; }
    a.omega[0 * B + i] = d.om0; a.omega[1 * B + i] = d.om1; a.omega[2 * B + i] = d.om2; a.omega[3 * B + i] = d.om3;
  } else {  // UAVB_STAGE_PHYSICS
    d.om0 = a.omega[0 * B + i]; d.om1 = a.omega[1 * B + i]; d.om2 = a.omega[2 * B + i]; d.om3 = a.omega[3 * B + i];
    const float gx = v.dIx * (d.wy * d.wz), gy = v.dIy * (d.wz * d.wx), gz = v.dIz * (d.wx * d.wy);
    float zx, zy, zz;
    if (a.zb) { zx = a.zb[i]; zy = a.zb[B + i]; zz = a.zb[2 * B + i]; }
    else body_z<float>(d, &zx, &zy, &zz);
    physics_step<float>(d, u, v, zx, zy, zz, gx, gy, gz);
    float* X = a.X;
    X[0 * B + i] = (float)(d.px + (double)d.dx); X[1 * B + i] = (float)(d.py + (double)d.dy); X[2 * B + i] = (float)(d.pz + (double)d.dz);
    X[3 * B + i] = d.q0; X[4 * B + i] = d.q1; X[5 * B + i] = d.q2; X[6 * B + i] = d.q3;
    X[7 * B + i] = d.vx; X[8 * B + i] = d.vy; X[9 * B + i] = d.vz;
    X[10 * B + i] = d.wx; X[11 * B + i] = d.wy; X[12 * B + i] = d.wz;
  }
}

}  // namespace uavb

using namespace uavb;

extern "C" int uavb_stage_f32(const uavb_stage_args* args, void* stream) {
  UAVB_REQUIRE(args != nullptr, "stage: args is NULL");
  UAVB_REQUIRE(args->B >= 0, "stage: B must be >= 0");
  UAVB_REQUIRE(args->X != nullptr, "stage: X is required");
  UAVB_REQUIRE(args->veh.dt > 0.0 && args->dt_outer > 0.0 && args->veh.mass > 0.0, "stage: veh.dt, dt_outer, veh.mass must be positive");
  switch (args->stage) {
    case UAVB_STAGE_OUTER:
      UAVB_REQUIRE(args->target && args->integral && args->thrust && args->pqr_cmd, "stage OUTER: target, integral, thrust, pqr_cmd required");
      break;
    case UAVB_STAGE_INNER:
      UAVB_REQUIRE(args->thrust && args->pqr_cmd && args->omega, "stage INNER: thrust, pqr_cmd, omega required");
      break;
    case UAVB_STAGE_PHYSICS:
      UAVB_REQUIRE(args->omega, "stage PHYSICS: omega required");
      break;
    default:
      return set_error(UAVB_EINVAL, "stage: unknown stage %d", args->stage);
  }
  int rc = require_device();
  if (rc) return rc;
  if (args->B == 0) return UAVB_OK;
  StageDev p;
  p.a = *args;
  make_vehu<float>(p.u, args->veh, args->dt_outer);
  stage_kernel<<<div_up(args->B, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
  UAVB_CUDA_OK(cudaGetLastError());
  return UAVB_OK;
}
