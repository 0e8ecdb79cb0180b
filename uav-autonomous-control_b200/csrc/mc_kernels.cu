// mc_kernels.cu -- Monte-Carlo inputs from a counter-based generator (Philox4x32-10).
// counter = (global index lo, global index hi, stream id, block), key = seed: the value drawn for
// rollout i does not depend on how rollouts are split over launches or GPUs.
#include "philox.cuh"
#include "uavb_common.cuh"

namespace uavb {

__global__ void __launch_bounds__(256) mc_uniform_kernel(unsigned long long seed, long long index_base, int stream_id, int B, int n_fields,
                                                         const float* __restrict__ lo, const float* __restrict__ hi, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const unsigned long long gi = (unsigned long long)(index_base + i);
  Philox ph{(unsigned)seed, (unsigned)(seed >> 32)};
  for (int k0 = 0; k0 < n_fields; k0 += 4) {
    unsigned r[4];
    ph.block((unsigned)gi, (unsigned)(gi >> 32), (unsigned)stream_id, (unsigned)(k0 >> 2), r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + j;
      if (k < n_fields) {
        // explicit non-fused ops: bit-identical to the fp32 NumPy expression lo + (hi - lo) * u
        const float span = __fsub_rn(hi[k], lo[k]);
        out[(long long)k * B + i] = __fadd_rn(lo[k], __fmul_rn(span, u01f(r[j])));
      }
    }
  }
}

// BASELINE configs[1] missions.
__global__ void __launch_bounds__(128) mc_missions_kernel(unsigned long long seed, long long index_base, int B, int S,
                                                          double* __restrict__ waypoints, double* __restrict__ velocity) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const unsigned long long gi = (unsigned long long)(index_base + i);
  Philox ph{(unsigned)seed, (unsigned)(seed >> 32)};
  unsigned r[4], q[4];
  ph.block((unsigned)gi, (unsigned)(gi >> 32), 0x4d53u, 0u, r);       // first waypoint x, y
  ph.block((unsigned)gi, (unsigned)(gi >> 32), 0x4d53u, 1u, q);       // first waypoint z, velocity
  double x = 2.0 + 20.0 * u01d(r[0], r[1]);
  double y = 2.0 + 10.0 * u01d(r[2], r[3]);
  double z = -5.0 + 4.0 * u01d(q[0], q[1]);
  velocity[i] = 2.0 + u01d(q[2], q[3]);
  double* w = waypoints + (size_t)i * (S + 1) * 3;
  w[0] = x; w[1] = y; w[2] = z;
  for (int s = 0; s < S; ++s) {
    ph.block((unsigned)gi, (unsigned)(gi >> 32), 0x4d53u, (unsigned)(2 + 2 * s), r);
    ph.block((unsigned)gi, (unsigned)(gi >> 32), 0x4d53u, (unsigned)(3 + 2 * s), q);
    const double cz = 2.0 * u01d(r[0], r[1]) - 1.0;                  // uniform on the sphere: z and azimuth
    const double phi = 6.283185307179586476925286766559 * u01d(r[2], r[3]);
    const double step = 2.0 + 3.0 * u01d(q[0], q[1]);
    const double rho = sqrt(fmax(0.0, 1.0 - cz * cz));
    double sp, cp;
    sincos(phi, &sp, &cp);
    x += step * rho * cp; y += step * rho * sp; z += step * 0.4 * cz;
    w[3 * (s + 1) + 0] = x; w[3 * (s + 1) + 1] = y; w[3 * (s + 1) + 2] = z;
  }
}

}  // namespace uavb

using namespace uavb;

extern "C" int uavb_mc_uniform_f32(unsigned long long seed, long long index_base, int stream_id, int B, int n_fields, const float* lo,
                                   const float* hi, float* out, void* stream) {
  UAVB_REQUIRE(lo && hi && out, "mc_uniform: NULL pointer");
  UAVB_REQUIRE(B >= 0 && n_fields >= 1, "mc_uniform: B >= 0 and n_fields >= 1 required");
  int rc = require_device();
  if (rc) return rc;
  if (B == 0) return UAVB_OK;
  mc_uniform_kernel<<<div_up(B, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(seed, index_base, stream_id, B, n_fields, lo, hi, out);
  UAVB_CUDA_OK(cudaGetLastError());
  return UAVB_OK;
}

extern "C" int uavb_mc_missions_f64(unsigned long long seed, long long index_base, int B, int S, double* waypoints_out,
                                    double* velocity_out, void* stream) {
  UAVB_REQUIRE(waypoints_out && velocity_out, "mc_missions: NULL pointer");
  UAVB_REQUIRE(B >= 0 && S >= 1 && S <= UAVB_MAX_SPLINES, "mc_missions: B >= 0 and 1 <= S <= UAVB_MAX_SPLINES required");
  int rc = require_device();
  if (rc) return rc;
  if (B == 0) return UAVB_OK;
  mc_missions_kernel<<<div_up(B, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(seed, index_base, B, S, waypoints_out, velocity_out);
  UAVB_CUDA_OK(cudaGetLastError());
  return UAVB_OK;
}
