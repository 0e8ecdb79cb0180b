// DEVELOPMENT PROBE (not product): throughput of packed fp32x2 FMA (FFMA2, sm_100) against scalar FFMA, and of a mix of
// FFMA2 with ALU-pipe instructions -- does packing free issue slots on B200?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ffma2_probe tools/ffma2_probe.cu ; build/ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
  float2 x0 = make_float2(threadIdx.x, 1.f), x1 = make_float2(2.f, 3.f), x2 = make_float2(4.f, 5.f), x3 = make_float2(6.f, 7.f);
  float2 x4 = make_float2(8.f, 9.f), x5 = make_float2(1.5f, 2.5f), x6 = make_float2(3.5f, 4.5f), x7 = make_float2(5.5f, 6.5f);
  const float2 A = make_float2(a, a), Bv = make_float2(b, b);
  float m = threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE == 0) {        // 16 scalar FFMA
        x0.x = fmaf(x0.x, a, b); x0.y = fmaf(x0.y, a, b); x1.x = fmaf(x1.x, a, b); x1.y = fmaf(x1.y, a, b);
        x2.x = fmaf(x2.x, a, b); x2.y = fmaf(x2.y, a, b); x3.x = fmaf(x3.x, a, b); x3.y = fmaf(x3.y, a, b);
        x4.x = fmaf(x4.x, a, b); x4.y = fmaf(x4.y, a, b); x5.x = fmaf(x5.x, a, b); x5.y = fmaf(x5.y, a, b);
        x6.x = fmaf(x6.x, a, b); x6.y = fmaf(x6.y, a, b); x7.x = fmaf(x7.x, a, b); x7.y = fmaf(x7.y, a, b);
      } else {                // 8 FFMA2 (same 16 FMAs)
        x0 = __ffma2_rn(x0, A, Bv); x1 = __ffma2_rn(x1, A, Bv); x2 = __ffma2_rn(x2, A, Bv); x3 = __ffma2_rn(x3, A, Bv);
        x4 = __ffma2_rn(x4, A, Bv); x5 = __ffma2_rn(x5, A, Bv); x6 = __ffma2_rn(x6, A, Bv); x7 = __ffma2_rn(x7, A, Bv);
      }
      if (MODE == 2 || MODE == 3) {   // plus 8 ALU-pipe instructions (FMNMX) competing for issue slots
        m = fmaxf(m, x0.x); m = fminf(m, x1.x); m = fmaxf(m, x2.x); m = fminf(m, x3.x);
        m = fmaxf(m, x4.x); m = fminf(m, x5.x); m = fmaxf(m, x6.x); m = fminf(m, x7.x);
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0.x + x0.y + x1.x + x1.y + x2.x + x2.y + x3.x + x3.y + x4.x + x4.y + x5.x + x5.y + x6.x + x6.y + x7.x + x7.y + m;
}

template <int MODE> __global__ void __launch_bounds__(256) kmix(float* out, int iters, float a, float b) {   // MODE 2: scalar + FMNMX, 3: FFMA2 + FMNMX
  k<MODE>(out, iters, a, b);
}

template <int MODE> double run(int sms, const char* name) {
  const int threads = 256, blocks = sms * 8, iters = 4096;
  float* buf;
  cudaMalloc(&buf, sizeof(float) * threads * blocks);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 1e30;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    if (MODE == 0) k<0><<<blocks, threads>>>(buf, iters, 0.999f, 0.001f);
    if (MODE == 1) k<1><<<blocks, threads>>>(buf, iters, 0.999f, 0.001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double fma = 16.0 * 8 * iters * (double)threads * blocks;
  printf("%-28s %.3f ms  %.1f TFLOP/s (2 flop per FMA)\n", name, best, 2 * fma / (best * 1e-3) / 1e12);
  cudaFree(buf);
  return best;
}

__global__ void __launch_bounds__(256) kmix_scalar(float* out, int iters, float a, float b);

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  run<0>(p.multiProcessorCount, "16 FFMA");
  run<1>(p.multiProcessorCount, "8 FFMA2");
  // mixes
  const int threads = 256, blocks = p.multiProcessorCount * 8, iters = 4096;
  float* buf; cudaMalloc(&buf, sizeof(float) * threads * blocks);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 2; mode <= 3; ++mode) {
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      if (mode == 2) k<2><<<blocks, threads>>>(buf, iters, 0.999f, 0.001f); else k<3><<<blocks, threads>>>(buf, iters, 0.999f, 0.001f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    printf("%-28s %.3f ms\n", mode == 2 ? "16 FFMA + 8 FMNMX" : "8 FFMA2 + 8 FMNMX", best);
  }
  return 0;
}
