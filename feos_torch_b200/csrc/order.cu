// State order of the bubble / dew point kernels (order.cuh): key kernels + one radix sort of (key, index)
// pairs.  C-ABI entry points feos_dev_mix_point_order / feos_dev_gc_point_order of include/feos_b200.h.
#include <cub/device/device_radix_sort.cuh>

#include "../../include/feos_b200.h"
#include "common.h"
#include "order.cuh"

namespace feos {

constexpr int kOrderBlock = 256;

__global__ void __launch_bounds__(kOrderBlock)
mix_order_key_kernel(const double* __restrict__ params, const double* __restrict__ temperature,
                     const double* __restrict__ molefrac, int64_t n, int flags, uint64_t* __restrict__ keys,
                     int32_t* __restrict__ index) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double prm[16];
  const double2* p = reinterpret_cast<const double2*>(params + ((flags & FEOS_SHARED_PARAMS) ? 0 : 16 * i));
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    double2 t = __ldg(p + k);
    prm[2 * k] = t.x;
    prm[2 * k + 1] = t.y;
  }
  keys[i] = mix_order_key(prm, temperature[i], molefrac[i]);
  index[i] = (int32_t)i;
}

__global__ void __launch_bounds__(kOrderBlock)
pure_order_key_kernel(const double* __restrict__ params, const double* __restrict__ temperature, int64_t n,
                      uint64_t* __restrict__ keys, int32_t* __restrict__ index) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double prm[8];
  const double2* p = reinterpret_cast<const double2*>(params + 8 * i);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double2 t = __ldg(p + k);
    prm[2 * k] = t.x;
    prm[2 * k + 1] = t.y;
  }
  keys[i] = pure_order_key(prm, temperature[i]);
  index[i] = (int32_t)i;
}

__global__ void __launch_bounds__(kOrderBlock)
gc_order_key_kernel(const int32_t* __restrict__ pair_index, const double* __restrict__ temperature,
                    const double* __restrict__ molefrac, int64_t n, uint64_t* __restrict__ keys,
                    int32_t* __restrict__ index) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = gc_order_key(pair_index[i], temperature[i], molefrac[i]);
  index[i] = (int32_t)i;
}

// scratch layout: keys_in [n] u64 | keys_out [n] u64 | index [n] i32 | cub temporary storage
struct OrderScratch {
  Scratch mem;
  void* base = nullptr;
  uint64_t* keys_in = nullptr;
  uint64_t* keys_out = nullptr;
  int32_t* index = nullptr;
  void* temp = nullptr;
  size_t temp_bytes = 0;
};

static int order_scratch(OrderScratch& w, int64_t n, int begin_bit, int end_bit, cudaStream_t s) {
  FEOS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, w.temp_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                            (const int32_t*)nullptr, (int32_t*)nullptr, (int)n, begin_bit, end_bit, s));
  size_t keys = (sizeof(uint64_t) * (size_t)n + 255) / 256 * 256;
  size_t idx = (sizeof(int32_t) * (size_t)n + 255) / 256 * 256;
  if (int rc = w.mem.alloc(2 * keys + idx + w.temp_bytes, s)) return rc;
  w.base = w.mem.p;
  char* p = (char*)w.base;
  w.keys_in = (uint64_t*)p;
  w.keys_out = (uint64_t*)(p + keys);
  w.index = (int32_t*)(p + 2 * keys);
  w.temp = p + 2 * keys + idx;
  return 0;
}

// stable: states with equal key bits in [begin_bit, end_bit) keep their input order
static int order_sort(OrderScratch& w, int64_t n, int begin_bit, int end_bit, int32_t* order, cudaStream_t s) {
  FEOS_CUDA(cub::DeviceRadixSort::SortPairs(w.temp, w.temp_bytes, (const uint64_t*)w.keys_in, w.keys_out,
                                            (const int32_t*)w.index, order, (int)n, begin_bit, end_bit, s));
  count_launch(4);  // histogram + onesweep passes (accounting only)
  return 0;  // w.mem gives the scratch back (stream-ordered) when the caller's OrderScratch goes out of scope
}

}  // namespace feos

using namespace feos;

extern "C" {

int feos_dev_pure_order(const double* params, const double* temperature, int64_t n, int32_t* order, void* stream) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31);
  if (n == 0) return 0;
  FEOS_ARG(params && temperature && order);
  cudaStream_t s = (cudaStream_t)stream;
  OrderScratch w;
  if (int rc = order_scratch(w, n, 0, kPureOrderBits, s)) return rc;
  pure_order_key_kernel<<<grid_for(n, kOrderBlock), kOrderBlock, 0, s>>>(params, temperature, n, w.keys_in, w.index);
  FEOS_LAUNCH_OK();
  return order_sort(w, n, 0, kPureOrderBits, order, s);
}

int feos_dev_mix_point_order(const double* params, const double* temperature, const double* molefracs, int64_t n,
                             int flags, int32_t* order, void* stream) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31);
  if (n == 0) return 0;
  FEOS_ARG(params && temperature && molefracs && order);
  cudaStream_t s = (cudaStream_t)stream;
  OrderScratch w;
  if (int rc = order_scratch(w, n, 0, kMixOrderBits, s)) return rc;
  mix_order_key_kernel<<<grid_for(n, kOrderBlock), kOrderBlock, 0, s>>>(params, temperature, molefracs, n, flags,
                                                                       w.keys_in, w.index);
  FEOS_LAUNCH_OK();
  return order_sort(w, n, 0, kMixOrderBits, order, s);
}

int feos_dev_gc_point_order(const int32_t* pair_index, int32_t n_pairs, const double* temperature,
                            const double* molefracs, int64_t n, int pair_only, int32_t* order, void* stream) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31);
  if (n == 0) return 0;
  FEOS_ARG(pair_index && n_pairs > 0 && temperature && molefracs && order);
  cudaStream_t s = (cudaStream_t)stream;
  int pair_bits = 1;
  while (pair_bits < 31 && ((int64_t)1 << pair_bits) < (int64_t)n_pairs) ++pair_bits;
  const int begin_bit = pair_only ? kOrderStateBits : 0, end_bit = kOrderStateBits + pair_bits;
  OrderScratch w;
  if (int rc = order_scratch(w, n, begin_bit, end_bit, s)) return rc;
  gc_order_key_kernel<<<grid_for(n, kOrderBlock), kOrderBlock, 0, s>>>(pair_index, temperature, molefracs, n, w.keys_in,
                                                                      w.index);
  FEOS_LAUNCH_OK();
  return order_sort(w, n, begin_bit, end_bit, order, s);
}

}  // extern "C"
