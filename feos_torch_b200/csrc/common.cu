// Library-level entry points of include/feos_b200.h: version, error text, launch accounting.
#include <atomic>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/feos_b200.h"
#include "common.h"

namespace feos {
static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
void clear_error() { g_error[0] = 0; }
void count_launch(int64_t k) { g_launches.fetch_add(k, std::memory_order_relaxed); }

// Stream-ordered scratch memory (retry lists, scan tiles).  The default pool gives memory back to the
// OS at every synchronisation (release threshold 0), which made cudaMallocAsync cost milliseconds in
// every step after the Python layer's `.item()`; keep the pool's memory instead.
int scratch_alloc(void** p, size_t bytes, cudaStream_t s) {
  static thread_local int tuned_device = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -2;
  if (tuned_device != dev) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    tuned_device = dev;
  }
  cudaError_t err = cudaMallocAsync(p, bytes ? bytes : 1, s);
  if (err != cudaSuccess) {
    set_error("cudaMallocAsync(%zu) failed: %s", bytes, cudaGetErrorString(err));
    return -2;
  }
  return 0;
}
}  // namespace feos

// Register-resident DFMA chains on every SM: the FP64 roofline denominator (MEASURED_PEAKS.json holds
// HBM and bf16 figures only).  16 independent accumulators per thread hide the pipe latency.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double x, double y) {
  double acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = (double)(threadIdx.x + k);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = fma(acc[k], x, y);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += acc[k];
  if (s == 123.456) out[0] = s;  // keeps the chains alive without a store in practice
}

extern "C" {
int feos_b200_measure_fp64_peak(int iters, double* tflops) {
  feos::clear_error();
  FEOS_ARG(iters > 0 && tflops);
  int dev = 0, sms = 0;
  FEOS_CUDA(cudaGetDevice(&dev));
  FEOS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* d = nullptr;
  FEOS_CUDA(cudaMalloc((void**)&d, 8));
  cudaEvent_t e0, e1;
  FEOS_CUDA(cudaEventCreate(&e0));
  FEOS_CUDA(cudaEventCreate(&e1));
  int blocks = sms * 8;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    FEOS_CUDA(cudaEventRecord(e0, 0));
    fp64_peak_kernel<<<blocks, 256>>>(d, iters, 1.0000001, 1e-9);
    FEOS_LAUNCH_OK();
    FEOS_CUDA(cudaEventRecord(e1, 0));
    FEOS_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    FEOS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double tf = 2.0 * 16.0 * (double)iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return 0;
}

const char* feos_b200_version(void) { return "0.1.0+b200"; }
const char* feos_b200_last_error(void) { return feos::g_error; }
int64_t feos_b200_launch_count(void) { return feos::g_launches.load(std::memory_order_relaxed); }
}
