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

__device__ __forceinline__ void load_target(Target<float>& t, const float* T, long long B, long long i) {
  t.x = T[0 * B + i]; t.y = T[1 * B + i]; t.z = T[2 * B + i];
  t.vx = T[3 * B + i]; t.vy = T[4 * B + i]; t.vz = T[5 * B + i];
  t.ax = T[6 * B + i]; t.ay = T[7 * B + i]; t.az = T[8 * B + i];
  float sy, cy;
  sincosf(T[9 * B + i], &sy, &cy);
  t.yc = cy; t.ys = sy;
}

// rotation-matrix argument of the reference methods (row-major [9][B]) or the one of the state
__device__ __forceinline__ RotE<float> rot_arg(const float* rot, const Drone<float>& d, long long B, long long i) {
  if (!rot) return rot_entries<float>(d.q0, d.q1, d.q2, d.q3);
  RotE<float> r;
  r.R00 = rot[0 * B + i]; r.R01 = rot[1 * B + i]; r.R02 = rot[2 * B + i];
  r.R10 = rot[3 * B + i]; r.R11 = rot[4 * B + i]; r.R12 = rot[5 * B + i];
  r.R22 = rot[8 * B + i];
  return r;
}

__global__ void __launch_bounds__(128) stage_kernel(const __grid_constant__ StageDev p) {
  const uavb_stage_args& a = p.a;
  const long long B = a.B;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  McValues mc;
  mc_from_vehicle(mc, a.veh);
  if (a.mc_mass) mc.mass = (double)a.mc_mass[i];
  if (a.mc_inertia) { mc.inertia[0] = (double)a.mc_inertia[i]; mc.inertia[1] = (double)a.mc_inertia[B + i]; mc.inertia[2] = (double)a.mc_inertia[2 * B + i]; }
  if (a.mc_gains) {
    for (int k = 0; k < UAVB_N_GAINS; ++k) mc.gains[k] = (double)a.mc_gains[k * B + i];
  }
  if (a.wind) { mc.wind[0] = a.wind[i]; mc.wind[1] = a.wind[B + i]; mc.wind[2] = a.wind[2 * B + i]; }
  VehP<float> v;
  make_vehp<float>(v, a.veh, mc);
  const VehU<float>& u = p.u;
  Drone<float> d;
  if (a.X) load_state(d, a.X, B, i);
  switch (a.stage) {
    case UAVB_STAGE_OUTER: {
      Target<float> t;
      load_target(t, a.target, B, i);
      d.integral = a.integral[i];
      outer_update<float>(d, u, v, t, (float)(t.x - d.px), (float)(t.y - d.py), (float)(t.z - d.pz));
      a.integral[i] = d.integral;
      a.thrust[i] = d.thrust_cmd;
      a.pqr_cmd[0 * B + i] = d.pc; a.pqr_cmd[1 * B + i] = d.qc; a.pqr_cmd[2 * B + i] = d.rc;
      break;
    }
    case UAVB_STAGE_ALTITUDE: {
      const RotE<float> r = rot_arg(a.rot, d, B, i);
      float integ = a.integral[i];
      const float ez = (float)((double)a.target[2 * B + i] - d.pz);
      a.thrust[i] = altitude_cmd<float>(integ, u, v, ez, d.vz, a.target[5 * B + i], a.target[8 * B + i], Math<float>::rcp_fast(r.R22));
      a.integral[i] = integ;
      break;
    }
    case UAVB_STAGE_LATERAL: {
      float bx, by;
      lateral_cmd<float>(u, v, (float)((double)a.target[0 * B + i] - d.px), (float)((double)a.target[1 * B + i] - d.py), d.vx, d.vy,
                         a.target[3 * B + i], a.target[4 * B + i], a.target[6 * B + i], a.target[7 * B + i], a.thrust[i], &bx, &by);
      a.bxy[i] = bx; a.bxy[B + i] = by;
      break;
    }
    case UAVB_STAGE_ROLL_PITCH: {
      const RotE<float> r = rot_arg(a.rot, d, B, i);
      float pc, qc;
      roll_pitch_cmd<float>(v, a.bxy[i], a.bxy[B + i], r, Math<float>::rcp_fast(r.R22), &pc, &qc);
      a.pqr_cmd[0 * B + i] = pc; a.pqr_cmd[1 * B + i] = qc;
      break;
    }
    case UAVB_STAGE_YAW: {
      // Quad.phi/theta/psi use the raw state quaternion (quad.py:189-213); for a unit quaternion that is d.q*
      float sy, cy;
      sincosf(a.target[9 * B + i], &sy, &cy);
      a.pqr_cmd[2 * B + i] = yaw_rate_cmd<float>(v, d.q0, d.q1, d.q2, d.q3, cy, sy, a.pqr_cmd[1 * B + i]);
      break;
    }
    case UAVB_STAGE_BODY_RATE: {
      d.pc = a.pqr_cmd[0 * B + i]; d.qc = a.pqr_cmd[1 * B + i]; d.rc = a.pqr_cmd[2 * B + i];
      float g[3], mom[3];
      body_rate_moment<float>(d, v, g, mom);
      a.moment[0 * B + i] = mom[0]; a.moment[1 * B + i] = mom[1]; a.moment[2 * B + i] = mom[2];
      break;
    }
    case UAVB_STAGE_ALLOCATE:
    case UAVB_STAGE_PROPELLER: {
      const float mom[3] = {a.moment[0 * B + i], a.moment[1 * B + i], a.moment[2 * B + i]};
      const float coll = 0.25f * clampr<float>(a.thrust[i], u.fmin4, u.fmax4);
      float f[4];
      allocate_forces<float>(u, coll, mom, f);
      if (a.forces) { a.forces[0 * B + i] = f[0]; a.forces[1 * B + i] = f[1]; a.forces[2 * B + i] = f[2]; a.forces[3 * B + i] = f[3]; }
      if (a.stage == UAVB_STAGE_PROPELLER) {
        d.om0 = a.omega[0 * B + i]; d.om1 = a.omega[1 * B + i]; d.om2 = a.omega[2 * B + i]; d.om3 = a.omega[3 * B + i];
        float cmd[4];
        motor_lag<float>(d, u, f, cmd);
        a.omega[0 * B + i] = d.om0; a.omega[1 * B + i] = d.om1; a.omega[2 * B + i] = d.om2; a.omega[3 * B + i] = d.om3;
        if (a.omega_cmd) { a.omega_cmd[0 * B + i] = cmd[0]; a.omega_cmd[1 * B + i] = cmd[1]; a.omega_cmd[2 * B + i] = cmd[2]; a.omega_cmd[3 * B + i] = cmd[3]; }
      }
      break;
    }
    case UAVB_STAGE_INNER: {
      d.pc = a.pqr_cmd[0 * B + i]; d.qc = a.pqr_cmd[1 * B + i]; d.rc = a.pqr_cmd[2 * B + i];
      d.om0 = a.omega[0 * B + i]; d.om1 = a.omega[1 * B + i]; d.om2 = a.omega[2 * B + i]; d.om3 = a.omega[3 * B + i];
      float g[3], mom[3], f[4], cmd[4];
      body_rate_moment<float>(d, v, g, mom);
      allocate_forces<float>(u, 0.25f * clampr<float>(a.thrust[i], u.fmin4, u.fmax4), mom, f);
      motor_lag<float>(d, u, f, cmd);
      if (a.moment) { a.moment[0 * B + i] = mom[0]; a.moment[1 * B + i] = mom[1]; a.moment[2 * B + i] = mom[2]; }
      if (a.forces) { a.forces[0 * B + i] = f[0]; a.forces[1 * B + i] = f[1]; a.forces[2 * B + i] = f[2]; a.forces[3 * B + i] = f[3]; }
      if (a.omega_cmd) { a.omega_cmd[0 * B + i] = cmd[0]; a.omega_cmd[1 * B + i] = cmd[1]; a.omega_cmd[2 * B + i] = cmd[2]; a.omega_cmd[3 * B + i] = cmd[3]; }
      a.omega[0 * B + i] = d.om0; a.omega[1 * B + i] = d.om1; a.omega[2 * B + i] = d.om2; a.omega[3 * B + i] = d.om3;
      break;
    }
    case UAVB_STAGE_ATTITUDE: {
      if (a.rot_out) {     // quad.py:129-155 (normalised quaternion)
        const float q0 = d.q0, q1 = d.q1, q2 = d.q2, q3 = d.q3;
        float* o = a.rot_out;
        o[0 * B + i] = 1.f - 2.f * (q2 * q2 + q3 * q3); o[1 * B + i] = 2.f * (q1 * q2 - q0 * q3); o[2 * B + i] = 2.f * (q1 * q3 + q0 * q2);
        o[3 * B + i] = 2.f * (q1 * q2 + q0 * q3); o[4 * B + i] = 1.f - 2.f * (q1 * q1 + q3 * q3); o[5 * B + i] = 2.f * (q2 * q3 - q0 * q1);
        o[6 * B + i] = 2.f * (q1 * q3 - q0 * q2); o[7 * B + i] = 2.f * (q2 * q3 + q0 * q1); o[8 * B + i] = 1.f - 2.f * (q1 * q1 + q2 * q2);
      }
      if (a.euler_out) {   // quad.py:189-213 on the RAW state quaternion, like the reference properties
        const float q0 = a.X[3 * B + i], q1 = a.X[4 * B + i], q2 = a.X[5 * B + i], q3 = a.X[6 * B + i];
        a.euler_out[0 * B + i] = atan2f(2.f * (q0 * q1 + q2 * q3), 1.f - 2.f * (q1 * q1 + q2 * q2));
        a.euler_out[1 * B + i] = asinf(clampr<float>(2.f * (q0 * q2 - q3 * q1), -1.f, 1.f));
        a.euler_out[2 * B + i] = atan2f(2.f * (q0 * q3 + q1 * q2), 1.f - 2.f * (q2 * q2 + q3 * q3));
      }
      break;
    }
    default: {  // UAVB_STAGE_PHYSICS
      d.om0 = a.omega[0 * B + i]; d.om1 = a.omega[1 * B + i]; d.om2 = a.omega[2 * B + i]; d.om3 = a.omega[3 * B + i];
      const float gx = v.dIx * (d.wy * d.wz), gy = v.dIy * (d.wz * d.wx), gz = v.dIz * (d.wx * d.wy);
      float nx, ny, nz;
      body_z<float>(d, &nx, &ny, &nz);                    // what mj_step's forward pass computes from X_k
      float zx = nx, zy = ny, zz = nz;
      if (a.zb) { zx = a.zb[i]; zy = a.zb[B + i]; zz = a.zb[2 * B + i]; }
      physics_step<float>(d, u, v, zx, zy, zz, gx, gy, gz);
      if (a.zb_out) { a.zb_out[i] = nx; a.zb_out[B + i] = ny; a.zb_out[2 * B + i] = nz; }
      float* X = a.X;
      const float fx = (float)(d.px + (double)d.dx), fy = (float)(d.py + (double)d.dy), fz = (float)(d.pz + (double)d.dz);
      X[0 * B + i] = fx; X[1 * B + i] = fy; X[2 * B + i] = fz;
      if (a.collided && a.n_obs > 0) {
        bool h = false;
        for (int k = 0; k < a.n_obs; ++k) {
          const float* q = a.aabbs + 6 * k;
          h |= (q[0] <= fx) & (fx <= q[1]) & (q[2] <= fy) & (fy <= q[3]) & (q[4] <= fz) & (fz <= q[5]);
        }
        if (h) a.collided[i] = 1.f;
      }
      X[3 * B + i] = d.q0; X[4 * B + i] = d.q1; X[5 * B + i] = d.q2; X[6 * B + i] = d.q3;
      X[7 * B + i] = d.vx; X[8 * B + i] = d.vy; X[9 * B + i] = d.vz;
      X[10 * B + i] = d.wx; X[11 * B + i] = d.wy; X[12 * B + i] = d.wz;
    }
  }
}

}  // namespace uavb

using namespace uavb;

extern "C" int uavb_stage_f32(const uavb_stage_args* args, void* stream) {
  UAVB_REQUIRE(args != nullptr, "stage: args is NULL");
  UAVB_REQUIRE(args->B >= 0, "stage: B must be >= 0");
  UAVB_REQUIRE(args->veh.dt > 0.0 && args->dt_outer > 0.0 && args->veh.mass > 0.0, "stage: veh.dt, dt_outer, veh.mass must be positive");
  const uavb_stage_args& a = *args;
  switch (a.stage) {
    case UAVB_STAGE_OUTER:
      UAVB_REQUIRE(a.X && a.target && a.integral && a.thrust && a.pqr_cmd, "stage OUTER: X, target, integral, thrust, pqr_cmd required");
      break;
    case UAVB_STAGE_INNER:
      UAVB_REQUIRE(a.X && a.thrust && a.pqr_cmd && a.omega, "stage INNER: X, thrust, pqr_cmd, omega required");
      break;
    case UAVB_STAGE_PHYSICS:
      UAVB_REQUIRE(a.X && a.omega, "stage PHYSICS: X, omega required");
      UAVB_REQUIRE(a.n_obs >= 0 && (a.n_obs == 0 || a.aabbs), "stage PHYSICS: n_obs > 0 needs aabbs");
      break;
    case UAVB_STAGE_ALTITUDE:
      UAVB_REQUIRE(a.X && a.target && a.integral && a.thrust, "stage ALTITUDE: X, target, integral, thrust required");
      break;
    case UAVB_STAGE_LATERAL:
      UAVB_REQUIRE(a.X && a.target && a.thrust && a.bxy, "stage LATERAL: X, target, thrust, bxy required");
      break;
    case UAVB_STAGE_ROLL_PITCH:
      UAVB_REQUIRE(a.bxy && a.pqr_cmd && (a.rot || a.X), "stage ROLL_PITCH: bxy, pqr_cmd and rot (or X) required");
      break;
    case UAVB_STAGE_YAW:
      UAVB_REQUIRE(a.X && a.target && a.pqr_cmd, "stage YAW: X, target, pqr_cmd required");
      break;
    case UAVB_STAGE_BODY_RATE:
      UAVB_REQUIRE(a.X && a.pqr_cmd && a.moment, "stage BODY_RATE: X, pqr_cmd, moment required");
      break;
    case UAVB_STAGE_ALLOCATE:
      UAVB_REQUIRE(a.thrust && a.moment && a.forces, "stage ALLOCATE: thrust, moment, forces required");
      break;
    case UAVB_STAGE_PROPELLER:
      UAVB_REQUIRE(a.thrust && a.moment && a.omega, "stage PROPELLER: thrust, moment, omega required");
      break;
    case UAVB_STAGE_ATTITUDE:
      UAVB_REQUIRE(a.X && (a.rot_out || a.euler_out), "stage ATTITUDE: X and rot_out or euler_out required");
      break;
    default:
      return set_error(UAVB_EINVAL, "stage: unknown stage %d", a.stage);
  }
  int rc = require_device();
  if (rc) return rc;
  if (a.B == 0) return UAVB_OK;
  StageDev p;
  p.a = a;
  make_vehu<float>(p.u, a.veh, a.dt_outer);
  stage_kernel<<<div_up(a.B, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
  UAVB_CUDA_OK(cudaGetLastError());
  return UAVB_OK;
}
