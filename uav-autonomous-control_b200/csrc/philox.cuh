// philox.cuh -- Philox4x32-10 counter-based generator shared by the Monte-Carlo input kernels and the RRT* planner.
// counter = (global index lo, global index hi, stream id, block), key = seed: the value drawn for unit i does not depend
// on how units are split over launches or GPUs.  oracle/rrt_np.py restates the same generator in Python.
#pragma once

namespace uavb {

struct Philox {
  unsigned k0, k1;
  __device__ __forceinline__ void block(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned out[4]) const {
    unsigned a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      c0 = hi1 ^ c1 ^ a; c1 = lo1; c2 = hi0 ^ c3 ^ b; c3 = lo0;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};

// (0,1) with 24 random bits, exactly representable in fp32
__device__ __forceinline__ float u01f(unsigned x) { return __fmaf_rn((float)(x >> 8), 5.9604644775390625e-08f, 2.98023223876953125e-08f); }
// [0,1) with 53 random bits
__device__ __forceinline__ double u01d(unsigned hi, unsigned lo) {
  return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * 1.1102230246251565e-16;
}

}  // namespace uavb
