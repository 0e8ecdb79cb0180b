// DEVELOPMENT PROBE (not product): how does the hardware place a one-wave grid of small CTAs on the SMs?
// Launches `ctas` CTAs of 64 threads whose residency is capped at `cap` per SM through dynamic shared memory,
// keeps every CTA alive for ~2 ms so all are co-resident, and prints the histogram of CTAs per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o build/smid_probe tools/smid_probe.cu ; build/smid_probe 1563 12
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
__global__ void probe(int* smid_out, long long spin) {
  extern __shared__ int pad[];
  unsigned id;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
  if (threadIdx.x == 0) smid_out[blockIdx.x] = (int)id;
  const long long t0 = clock64();
  while (clock64() - t0 < spin) { }
  if (threadIdx.x == 999) pad[0] = 1;
}
int main(int argc, char** argv) {
  const int ctas = argc > 1 ? atoi(argv[1]) : 1563, cap = argc > 2 ? atoi(argv[2]) : 12;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int per_sm = (int)prop.sharedMemPerMultiprocessor;
  int dyn = per_sm / cap - 1024;
  dyn -= dyn % 128;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
  cudaFuncSetAttribute(probe, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe, 64, dyn);
  int* d;
  cudaMalloc(&d, ctas * sizeof(int));
  probe<<<ctas, 64, dyn>>>(d, 4000000);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<int> h(ctas);
  cudaMemcpy(h.data(), d, ctas * sizeof(int), cudaMemcpyDeviceToHost);
  std::vector<int> cnt(prop.multiProcessorCount, 0);
  for (int x : h) if (x >= 0 && x < (int)cnt.size()) cnt[x]++;
  std::vector<int> hist(64, 0);
  for (int c : cnt) hist[c < 64 ? c : 63]++;
  printf("ctas %d cap %d (dyn smem %d B, occupancy API says %d/SM), SMs %d, status %s\n  CTAs/SM histogram:", ctas, cap, dyn, occ,
         prop.multiProcessorCount, cudaGetErrorString(e));
  for (int i = 0; i < 64; ++i) if (hist[i]) printf(" %dx%d", hist[i], i);
  printf("\n");
  return 0;
}
