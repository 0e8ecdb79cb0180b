// DEVELOPMENT PROBE (not product): throughput and latency of the packed fp32x2 instructions of sm_100 (FFMA2 / FADD2 / FMUL2)
// against their scalar forms, alone and mixed with ALU-pipe instructions (FMNMX) -- does packing free issue slots on B200?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ffma2_probe tools/ffma2_probe.cu ; build/ffma2_probe
// Each mode does the same 16 independent fp32 chains per thread; "ops" counts scalar-equivalent operations.
#include <cstdio>
#include <cuda_runtime.h>

enum { SC_FMA, PK_FMA, SC_FMA_MNMX, PK_FMA_MNMX, SC_ADD, PK_ADD, SC_MUL, PK_MUL, SC_MIX3, PK_MIX3, LAT_SC, LAT_PK, N_MODES };
static const char* kNames[N_MODES] = {"16 FFMA", "8 FFMA2", "16 FFMA + 8 FMNMX", "8 FFMA2 + 8 FMNMX", "16 FADD", "8 FADD2", "16 FMUL", "8 FMUL2",
                                      "8 FFMA+4 FADD+4 FMUL+8 FMNMX", "4 FFMA2+2 FADD2+2 FMUL2+8 FMNMX", "1 dependent FFMA chain", "1 dependent FFMA2 chain"};

template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
  float2 x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x + i, 1.f + 0.5f * i);
  const float2 A = make_float2(a, a), Bv = make_float2(b, b);
  float m = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE == LAT_SC) { x[0].x = fmaf(x[0].x, a, b); continue; }
      if (MODE == LAT_PK) { x[0] = __ffma2_rn(x[0], A, Bv); continue; }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool fma = MODE == SC_FMA || MODE == PK_FMA || MODE == SC_FMA_MNMX || MODE == PK_FMA_MNMX || ((MODE == SC_MIX3 || MODE == PK_MIX3) && i < 4);
        const bool add = MODE == SC_ADD || MODE == PK_ADD || ((MODE == SC_MIX3 || MODE == PK_MIX3) && (i == 4 || i == 5));
        const bool packed = MODE == PK_FMA || MODE == PK_FMA_MNMX || MODE == PK_ADD || MODE == PK_MUL || MODE == PK_MIX3;
        if (fma) { if (packed) x[i] = __ffma2_rn(x[i], A, Bv); else { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); } }
        else if (add) { if (packed) x[i] = __fadd2_rn(x[i], Bv); else { x[i].x = __fadd_rn(x[i].x, b); x[i].y = __fadd_rn(x[i].y, b); } }
        else { if (packed) x[i] = __fmul2_rn(x[i], A); else { x[i].x = __fmul_rn(x[i].x, a); x[i].y = __fmul_rn(x[i].y, a); } }
      }
      if (MODE == SC_FMA_MNMX || MODE == PK_FMA_MNMX || MODE == SC_MIX3 || MODE == PK_MIX3) {
#pragma unroll
        for (int i = 0; i < 8; ++i) m = (i & 1) ? fminf(m, x[i].x) : fmaxf(m, x[i].x);
      }
    }
  }
  float s = m;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(int sms, float* buf, double clock_ghz) {
  const int threads = 256, blocks = sms * 8, iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 1e30;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(buf, iters, 0.999f, 0.001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const bool lat = MODE == LAT_SC || MODE == LAT_PK;
  const double ops = (lat ? (MODE == LAT_PK ? 2.0 : 1.0) : 16.0) * 8 * iters * (double)threads * blocks;
  // warp-instruction slots per SM sub-partition and cycle: 8 CTAs x 8 warps / 4 schedulers = 16 warps per scheduler
  const double cycles = best * 1e-3 * clock_ghz * 1e9;
  const double per_iter = cycles / (8.0 * iters) / 16.0;   // cycles per unrolled body per warp, per scheduler
  printf("%-36s %8.3f ms  %7.2f T scalar-op/s  %6.2f cycles per body per warp (at %.3f GHz)\n", kNames[MODE], best, ops / (best * 1e-3) / 1e12, per_iter, clock_ghz);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
}

template <int M> struct All { static void go(int sms, float* buf, double ghz) { run<M>(sms, buf, ghz); All<M + 1>::go(sms, buf, ghz); } };
template <> struct All<N_MODES> { static void go(int, float*, double) {} };

// FFMA2 latency hiding: ILP independent dependent-chains per thread, W warps per scheduler (one CTA of 128 W threads per SM).
template <int ILP> __global__ void __launch_bounds__(1024) chains(float* out, int iters, float a, float b) {
  float2 x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = make_float2(threadIdx.x + i, 1.f + 0.5f * i);
  const float2 A = make_float2(a, a), Bv = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = __ffma2_rn(x[i], A, Bv);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP> void run_chains(int sms, float* buf, double ghz) {
  for (int w = 1; w <= 4; ++w) {
    const int iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 1e30;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      chains<ILP><<<sms, 128 * w>>>(buf, iters, 0.999f, 0.001f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    const double cycles = best * 1e-3 * ghz * 1e9;
    const double issued = 16.0 * iters * ILP * w;           // FFMA2 per scheduler
    printf("FFMA2 chains: ILP %d, %d warps/scheduler: %6.2f cycles per FFMA2 per scheduler (2.0 = FMA pipe full)\n", ILP, w, cycles / issued);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float* buf; cudaMalloc(&buf, sizeof(float) * 1024 * p.multiProcessorCount * 8);
  printf("%s, %d SMs, clock attr %.3f GHz; 16 warps per scheduler, 16 independent chains per thread\n", p.name, p.multiProcessorCount, khz * 1e-6);
  All<0>::go(p.multiProcessorCount, buf, khz * 1e-6);
  run_chains<1>(p.multiProcessorCount, buf, khz * 1e-6);
  run_chains<2>(p.multiProcessorCount, buf, khz * 1e-6);
  run_chains<3>(p.multiProcessorCount, buf, khz * 1e-6);
  run_chains<4>(p.multiProcessorCount, buf, khz * 1e-6);
  run_chains<8>(p.multiProcessorCount, buf, khz * 1e-6);
  return 0;
}
