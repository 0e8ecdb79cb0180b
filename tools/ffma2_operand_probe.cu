// DEVELOPMENT PROBE (not product): does the throughput of FFMA2 / FADD2 / FFMA depend on how many DISTINCT register operands an
// instruction reads (register-file bandwidth / bank conflicts), as opposed to operands served by the reuse cache or broadcast scalars?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ffma2_operand_probe tools/ffma2_operand_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: x = fma2(x, A, B)        (A, B shared by all chains: reuse cache)
// MODE 1: x = fma2(y_i, z_i, x)    (three distinct register pairs per instruction)
// MODE 2: x = fma2(y_i, s, x)      (two distinct pairs + one broadcast 32-bit register)
// MODE 3: x = add2(x, y_i)         (two distinct pairs)
// MODE 4: scalar: x.x = fma(y_i.x, z_i.x, x.x); x.y = ... (three distinct scalar registers per instruction)
// MODE 5: x = fma2(y_i, z_i, x) with y_i, z_i in the SAME bank parity arrangement but chains paired so consecutive instructions share y
template <int MODE, int N> __global__ void __launch_bounds__(128) k(const float2* in, float2* out, int iters, float s) {
  float2 x[N], y[N], z[N];
#pragma unroll
  for (int i = 0; i < N; ++i) { x[i] = in[threadIdx.x + 32 * i]; y[i] = in[threadIdx.x + 32 * (i + N)]; z[i] = in[threadIdx.x + 32 * (i + 2 * N)]; }
  const float2 A = in[1000], Bv = in[1001];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        if (MODE == 0) x[i] = __ffma2_rn(x[i], A, Bv);
        if (MODE == 1) x[i] = __ffma2_rn(y[i], z[i], x[i]);
        if (MODE == 2) x[i] = __ffma2_rn(y[i], make_float2(s, s), x[i]);
        if (MODE == 3) x[i] = __fadd2_rn(x[i], y[i]);
        if (MODE == 4) { x[i].x = __fmaf_rn(y[i].x, z[i].x, x[i].x); x[i].y = __fmaf_rn(y[i].y, z[i].y, x[i].y); }
        if (MODE == 5) x[i] = __ffma2_rn(y[i / 2], z[i / 2], x[i]);
      }
    }
  }
  float2 r = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < N; ++i) { r.x += x[i].x + y[i].x + z[i].x; r.y += x[i].y + y[i].y + z[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE, int N> void run(int sms, const float2* in, float2* out, double ghz, const char* name) {
  const int iters = 4096;
  for (int w = 2; w <= 4; w += 2) {                      // warps per scheduler (one CTA of 128 threads per warp-per-scheduler)
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      k<MODE, N><<<sms * w, 128>>>(in, out, iters, 0.999f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 1 && ms < best) best = ms;
    }
    const double cycles = best * 1e-3 * ghz * 1e9;
    const double packed = 4.0 * iters * N * w * (MODE == 4 ? 2 : 1);   // instructions per scheduler
    printf("%-58s N=%d, %d warps/scheduler: %5.2f cycles per instruction per scheduler\n", name, N, w, cycles / packed);
  }
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  float2 *in, *out;
  cudaMalloc(&in, sizeof(float2) * 4096); cudaMalloc(&out, sizeof(float2) * 128 * p.multiProcessorCount * 4);
  cudaMemset(in, 0, sizeof(float2) * 4096);
  for (int warm = 0; warm < 200; ++warm) k<0, 8><<<p.multiProcessorCount * 4, 128>>>(in, out, 4096, 0.999f);   // clocks up
  cudaDeviceSynchronize();
  run<0, 8>(p.multiProcessorCount, in, out, ghz, "FFMA2 x = x*A + B (A, B shared: reuse cache)");
  run<1, 8>(p.multiProcessorCount, in, out, ghz, "FFMA2 x = y_i*z_i + x (3 distinct register pairs)");
  run<2, 8>(p.multiProcessorCount, in, out, ghz, "FFMA2 x = y_i*s + x (2 pairs + broadcast scalar)");
  run<3, 8>(p.multiProcessorCount, in, out, ghz, "FADD2 x = x + y_i (2 distinct pairs)");
  run<4, 8>(p.multiProcessorCount, in, out, ghz, "FFMA  x = y_i*z_i + x (3 distinct scalar registers)");
  run<5, 8>(p.multiProcessorCount, in, out, ghz, "FFMA2 x_i = y_(i/2)*z_(i/2) + x_i (pairs shared by 2 instr.)");
  return 0;
}
