// Order-preserving stream compaction of per-state results by the failed mask.
//
// Replaces the row removal of the reference's native layer (`filter_binary`, src/pcsaft.rs:216-231,
// and the compaction loops at :94-101, :127-128) and the boolean-mask indexing the Python layer does
// on top of it (pcsaft_pure.py:192-194): three small launches (per-tile counts, single-block scan of
// the tile counts, scatter), HBM-bound and negligible next to the solves.
#include "../../include/feos_b200.h"
#include "common.h"

namespace feos {

constexpr int kTile = 1024;  // elements per block: 256 threads x 4

__global__ void __launch_bounds__(256) compact_count_kernel(const uint8_t* __restrict__ failed, int64_t n,
                                                           int32_t* __restrict__ tile_counts) {
  int64_t base = (int64_t)blockIdx.x * kTile;
  int c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int64_t i = base + threadIdx.x + 256 * k;
    if (i < n && !failed[i]) ++c;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ int ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += ws[w];
    tile_counts[blockIdx.x] = s;
  }
}

// exclusive scan of the tile counts by one block (sequential over chunks of 1024 tiles)
__global__ void __launch_bounds__(1024) compact_scan_kernel(int32_t* __restrict__ tile_counts, int tiles,
                                                           int64_t* __restrict__ n_ok) {
  __shared__ int ws[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < tiles; base += 1024) {
    int t = base + threadIdx.x;
    int v = (t < tiles) ? tile_counts[t] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = ws[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += y;
      }
      ws[threadIdx.x] = w;
    }
    __syncthreads();
    int carry = carry_s;
    int incl = x + ((threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0);
    if (t < tiles) tile_counts[t] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *n_ok = carry_s;  // may live in mapped pinned host memory: the host polls it
    __threadfence_system();
  }
}

__global__ void __launch_bounds__(256)
compact_scatter_kernel(const uint8_t* __restrict__ failed, int64_t n, const int32_t* __restrict__ tile_offsets,
                       int32_t* __restrict__ pos, const double* __restrict__ in, int in_cols,
                       double* __restrict__ out, int out_cols) {
  // thread t owns the 4 consecutive elements base + 4 t .. base + 4 t + 3 (keeps input order)
  int64_t first = (int64_t)blockIdx.x * kTile + 4 * threadIdx.x;
  int ok[4];
  int c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int64_t i = first + k;
    ok[k] = (i < n && !failed[i]) ? 1 : 0;
    c += ok[k];
  }
  int x = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  __shared__ int ws[8];
  if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += ws[w];
  int p = tile_offsets[blockIdx.x] + woff + x - c;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int64_t i = first + k;
    if (i >= n) break;
    if (ok[k]) {
      if (pos) pos[i] = p;
      if (in) {
        for (int col = 0; col < out_cols; ++col) out[(int64_t)p * out_cols + col] = col < in_cols ? in[i * in_cols + col] : 0.0;
      }
      ++p;
    } else if (pos) {
      pos[i] = -1;
    }
  }
}

}  // namespace feos

using namespace feos;

extern "C" int feos_dev_compact(const uint8_t* failed, int64_t n, int32_t* pos, int64_t* n_ok, const double* in,
                                int in_cols, double* out, int out_cols, void* stream) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31 && n_ok);
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) {
    FEOS_CUDA(cudaMemsetAsync(n_ok, 0, sizeof(int64_t), s));
    return 0;
  }
  FEOS_ARG(failed);
  FEOS_ARG(!in || (out && in_cols >= 1 && out_cols >= in_cols));
  int tiles = (int)((n + kTile - 1) / kTile);
  Scratch scratch;
  if (int rc_ = scratch.alloc(sizeof(int32_t) * (size_t)tiles, s)) return rc_;
  int32_t* ws = scratch.as<int32_t>();
  compact_count_kernel<<<tiles, 256, 0, s>>>(failed, n, ws);
  FEOS_LAUNCH_OK();
  compact_scan_kernel<<<1, 1024, 0, s>>>(ws, tiles, n_ok);
  FEOS_LAUNCH_OK();
  compact_scatter_kernel<<<tiles, 256, 0, s>>>(failed, n, ws, pos, in, in_cols, out, out_cols);
  FEOS_LAUNCH_OK();
  return 0;
}

int feos::host_compact_download(HostCall& hc, const uint8_t* d_failed, int64_t n, const double* d_values, int in_cols,
                                int out_cols, double* h_out, uint8_t* h_status, int64_t* n_ok) {
  double* d_out = nullptr;
  int64_t* d_count = nullptr;
  if (int rc = hc.alloc(&d_out, (size_t)n * out_cols)) return rc;
  if (int rc = hc.alloc(&d_count, 1)) return rc;
  if (int rc = feos_dev_compact(d_failed, n, nullptr, d_count, d_values, in_cols, d_out, out_cols, hc.s)) return rc;
  FEOS_CUDA(cudaMemcpyAsync(h_status, d_failed, n, cudaMemcpyDeviceToHost, hc.s));
  int64_t k = 0;
  FEOS_CUDA(cudaMemcpyAsync(&k, d_count, sizeof(k), cudaMemcpyDeviceToHost, hc.s));
  FEOS_CUDA(cudaStreamSynchronize(hc.s));
  if (k > 0) FEOS_CUDA(cudaMemcpyAsync(h_out, d_out, sizeof(double) * (size_t)k * out_cols, cudaMemcpyDeviceToHost, hc.s));
  FEOS_CUDA(cudaStreamSynchronize(hc.s));
  *n_ok = k;
  return 0;
}
