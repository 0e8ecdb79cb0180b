// DEVELOPMENT PROBE (not product): register-file read bandwidth and banking on sm_100 -- scalar FFMA with three distinct register
// operands whose register numbers differ in a controlled way (see the SASS for the actual allocation), 4 warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/rf_bank_probe tools/rf_bank_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

// VAR 0: x.x = fma(y.x, z.x, x.x)  and  x.y = fma(y.y, z.y, x.y)       (same parity three times)
// VAR 1: x.x = fma(y.x, z.y, x.x)  and  x.y = fma(y.y, z.x, x.y)       (parities e,o,e / o,e,o)
// VAR 2: x.x = fma(y.y, z.y, x.x)  and  x.y = fma(y.x, z.x, x.y)       (parities o,o,e / e,e,o)
// VAR 3: x.x = fma(y.x, s, x.x)  ...                                   (two registers + one register shared by all: reuse)
// VAR 4: x.x = fma(y.x, 0.999f, x.x)                                   (two registers + immediate)
// VAR 5: x.x = x.x + y.x  (FADD, two registers)
// VAR 6: x.x = fma(x.x, x.x, x.x) (one register three times)
template <int VAR> __global__ void __launch_bounds__(128) k(const float2* in, float2* out, int iters, float s) {
  constexpr int N = 8;
  float2 x[N], y[N], z[N];
#pragma unroll
  for (int i = 0; i < N; ++i) { x[i] = in[threadIdx.x + 32 * i]; y[i] = in[threadIdx.x + 32 * (i + N)]; z[i] = in[threadIdx.x + 32 * (i + 2 * N)]; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        if (VAR == 0) { x[i].x = __fmaf_rn(y[i].x, z[i].x, x[i].x); x[i].y = __fmaf_rn(y[i].y, z[i].y, x[i].y); }
        if (VAR == 1) { x[i].x = __fmaf_rn(y[i].x, z[i].y, x[i].x); x[i].y = __fmaf_rn(y[i].y, z[i].x, x[i].y); }
        if (VAR == 2) { x[i].x = __fmaf_rn(y[i].y, z[i].y, x[i].x); x[i].y = __fmaf_rn(y[i].x, z[i].x, x[i].y); }
        if (VAR == 3) { x[i].x = __fmaf_rn(y[i].x, s, x[i].x); x[i].y = __fmaf_rn(y[i].y, s, x[i].y); }
        if (VAR == 4) { x[i].x = __fmaf_rn(y[i].x, 0.999f, x[i].x); x[i].y = __fmaf_rn(y[i].y, 0.999f, x[i].y); }
        if (VAR == 5) { x[i].x = __fadd_rn(x[i].x, y[i].x); x[i].y = __fadd_rn(x[i].y, y[i].y); }
        if (VAR == 6) { x[i].x = __fmaf_rn(x[i].x, x[i].x, x[i].x); x[i].y = __fmaf_rn(x[i].y, x[i].y, x[i].y); }
      }
    }
  }
  float2 r = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < N; ++i) { r.x += x[i].x + y[i].x + z[i].x; r.y += x[i].y + y[i].y + z[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int VAR> void run(int sms, const float2* in, float2* out, double ghz, const char* name) {
  const int iters = 4096, w = 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 1e30;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k<VAR><<<sms * w, 128>>>(in, out, iters, 0.999f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 1 && ms < best) best = ms;
  }
  printf("%-64s %5.2f cycles per instruction per scheduler\n", name, best * 1e-3 * ghz * 1e9 / (4.0 * iters * 16 * w));
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  float2 *in, *out;
  cudaMalloc(&in, sizeof(float2) * 4096); cudaMalloc(&out, sizeof(float2) * 128 * p.multiProcessorCount * 4);
  cudaMemset(in, 0, sizeof(float2) * 4096);
  for (int warm = 0; warm < 200; ++warm) k<0><<<p.multiProcessorCount * 4, 128>>>(in, out, 4096, 0.999f);   // clocks up
  cudaDeviceSynchronize();
  run<0>(p.multiProcessorCount, in, out, ghz, "FFMA d = a*b + d, a b d same register parity");
  run<1>(p.multiProcessorCount, in, out, ghz, "FFMA parities (e,o,e) / (o,e,o)");
  run<2>(p.multiProcessorCount, in, out, ghz, "FFMA parities (o,o,e) / (e,e,o)");
  run<3>(p.multiProcessorCount, in, out, ghz, "FFMA two registers + one shared by all instructions (reuse)");
  run<4>(p.multiProcessorCount, in, out, ghz, "FFMA two registers + immediate");
  run<5>(p.multiProcessorCount, in, out, ghz, "FADD two registers");
  run<6>(p.multiProcessorCount, in, out, ghz, "FFMA one register three times");
  return 0;
}
