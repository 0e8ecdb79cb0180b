// capi.cu -- library-level C-ABI entry points: version, errors, device probing, FMA peak probe.
#include <atomic>
#include <mutex>

#include "tma.cuh"
#include "uavb_common.cuh"

namespace uavb {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return set_error(UAVB_ENODEVICE, "no CUDA device visible (%s); libuavb has no CPU path",
                     e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  return UAVB_OK;
}

// One pool per device, created on first use; std::call_once publishes it to every host thread (the *_host entry points are
// meant to be called from several threads at once).
cudaMemPool_t scratch_pool() {
  static cudaMemPool_t pools[kMaxDevices] = {nullptr};
  static std::once_flag once[kMaxDevices];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  std::call_once(once[dev], [dev]() {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool = nullptr;
    if (cudaMemPoolCreate(&pool, &props) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      pools[dev] = pool;
    } else {
      cudaGetLastError();
    }
  });
  return pools[dev];
}

void* pinned_scratch(int slot, size_t bytes) {
  thread_local void* buf[2] = {nullptr, nullptr};
  thread_local size_t cap[2] = {0, 0};
  if (slot < 0 || slot > 1) return nullptr;
  if (bytes > cap[slot]) {
    if (buf[slot]) cudaFreeHost(buf[slot]);
    buf[slot] = nullptr; cap[slot] = 0;
    const size_t want = bytes < 4096 ? 4096 : bytes;
    if (cudaHostAlloc(&buf[slot], want, cudaHostAllocPortable) != cudaSuccess) {
      cudaGetLastError();
      buf[slot] = nullptr;
      return nullptr;
    }
    cap[slot] = want;
  }
  return buf[slot];
}

TensorMapEncodeTiledFn tensor_map_encoder() {
  static TensorMapEncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TensorMapEncodeTiledFn>(p);
    else
      cudaGetLastError();
  });
  return fn;
}

// SM count of the current device (one query per device and process, thread-safe).
int sm_count_cached(int* sms) {
  static std::atomic<int> cache[kMaxDevices];
  int dev = 0;
  UAVB_CUDA_OK(cudaGetDevice(&dev));
  int n = (dev >= 0 && dev < kMaxDevices) ? cache[dev].load(std::memory_order_relaxed) : 0;
  if (n == 0) {
    UAVB_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    if (dev >= 0 && dev < kMaxDevices) cache[dev].store(n, std::memory_order_relaxed);
  }
  *sms = n;
  return UAVB_OK;
}

// Dependent-chain-free FMA loops: 8 independent accumulators per thread, enough CTAs to fill the chip.
template <class T> __global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters, T a, T b) {
  T x0 = (T)threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
      x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

template <class T> static int measure_one(int sms, double* tflops) {
  const int threads = 256, blocks = sms * 8, iters = sizeof(T) == 4 ? 4096 : 2048;
  T* buf = nullptr;
  UAVB_CUDA_OK(cudaMalloc(&buf, sizeof(T) * threads * blocks));
  cudaEvent_t e0, e1;
  UAVB_CUDA_OK(cudaEventCreate(&e0));
  UAVB_CUDA_OK(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    UAVB_CUDA_OK(cudaEventRecord(e0));
    fma_peak_kernel<T><<<blocks, threads>>>(buf, iters, (T)0.999, (T)0.001);
    UAVB_CUDA_OK(cudaEventRecord(e1));
    UAVB_CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    UAVB_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    const double flop = 2.0 * 64.0 * (double)iters * threads * blocks;
    if (rep > 0 && ms > 0.f) best = fmax(best, flop / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  *tflops = best;
  return UAVB_OK;
}

// The same with THREE distinct register operands per FMA (x_i = y_i * z_i + x_i, nothing for the operand reuse cache): the
// register file delivers two operand words per cycle and scheduler, so these issue every 1.5 cycles (profiles/r02_ffma2_probe.md).
__global__ void __launch_bounds__(256) fma3_peak_kernel(const float* in, float* out, int iters) {
  float x[8], y[8], z[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = in[threadIdx.x + 256 * i]; y[i] = in[threadIdx.x + 256 * (i + 8)]; z[i] = in[threadIdx.x + 256 * (i + 16)]; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = __fmaf_rn(y[i], z[i], x[i]);
    }
  }
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += x[i] + y[i] + z[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

static int measure_three_operand(int sms, double* tflops) {
  const int threads = 256, blocks = sms * 8, iters = 4096;
  float *in = nullptr, *buf = nullptr;
  UAVB_CUDA_OK(cudaMalloc(&in, sizeof(float) * 256 * 24));
  UAVB_CUDA_OK(cudaMemset(in, 0, sizeof(float) * 256 * 24));
  UAVB_CUDA_OK(cudaMalloc(&buf, sizeof(float) * threads * blocks));
  cudaEvent_t e0, e1;
  UAVB_CUDA_OK(cudaEventCreate(&e0));
  UAVB_CUDA_OK(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    UAVB_CUDA_OK(cudaEventRecord(e0));
    fma3_peak_kernel<<<blocks, threads>>>(in, buf, iters);
    UAVB_CUDA_OK(cudaEventRecord(e1));
    UAVB_CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    UAVB_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    const double flop = 2.0 * 64.0 * (double)iters * threads * blocks;
    if (rep > 0 && ms > 0.f) best = fmax(best, flop / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  cudaFree(in);
  *tflops = best;
  return UAVB_OK;
}

}  // namespace uavb

extern "C" int uavb_version(void) { return UAVB_VERSION; }
extern "C" const char* uavb_last_error(void) { return uavb::error_buffer(); }

extern "C" int uavb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" int uavb_device_info(int dev, int* sm_count, int* cc_major, int* cc_minor) {
  int rc = uavb::require_device();
  if (rc) return rc;
  cudaDeviceProp prop;
  UAVB_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return UAVB_OK;
}

extern "C" int uavb_measure_fma_peak(int dev, double* fp32_tflops, double* fp64_tflops) {
  int rc = uavb::require_device();
  if (rc) return rc;
  UAVB_CUDA_OK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  UAVB_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  double a = 0.0, b = 0.0;
  rc = uavb::measure_one<float>(prop.multiProcessorCount, &a);
  if (rc) return rc;
  rc = uavb::measure_one<double>(prop.multiProcessorCount, &b);
  if (rc) return rc;
  if (fp32_tflops) *fp32_tflops = a;
  if (fp64_tflops) *fp64_tflops = b;
  return UAVB_OK;
}

extern "C" int uavb_measure_fma_rates(int dev, double* fp32_tflops, double* fp32_three_operand_tflops, double* fp64_tflops) {
  int rc = uavb_measure_fma_peak(dev, fp32_tflops, fp64_tflops);
  if (rc) return rc;
  cudaDeviceProp prop;
  UAVB_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  double c = 0.0;
  rc = uavb::measure_three_operand(prop.multiProcessorCount, &c);
  if (rc) return rc;
  if (fp32_three_operand_tflops) *fp32_three_operand_tflops = c;
  return UAVB_OK;
}
