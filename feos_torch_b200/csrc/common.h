// Host-side helpers shared by the C-ABI translation units: error text, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace feos {

void set_error(const char* fmt, ...);
void clear_error();
void count_launch(int64_t k = 1);
int scratch_alloc(void** p, size_t bytes, cudaStream_t s);  // cudaMallocAsync from a pool that keeps its memory

// evaluates a CUDA runtime call; on failure records the text and returns -2 from the caller
#define FEOS_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t err__ = (call);                                                           \
    if (err__ != cudaSuccess) {                                                           \
      feos::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
      return -2;                                                                          \
    }                                                                                     \
  } while (0)

#define FEOS_LAUNCH_OK()                                                                  \
  do {                                                                                    \
    cudaError_t err__ = cudaGetLastError();                                               \
    if (err__ != cudaSuccess) {                                                           \
      feos::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(err__), __FILE__, __LINE__); \
      return -3;                                                                          \
    }                                                                                     \
    feos::count_launch();                                                                 \
  } while (0)

#define FEOS_ARG(cond)                                                                    \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      feos::set_error("invalid argument: %s (%s:%d)", #cond, __FILE__, __LINE__);         \
      return -1;                                                                          \
    }                                                                                     \
  } while (0)

inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

// batches below this size are launch-bound: the host entry points skip the state order (order.cuh)
constexpr int64_t kOrderMinStates = 1 << 14;

// bubble / dew forward kernels: 0 = one thread per state (default), 1 = persistent work queue
// (FEOS_B200_POINT_KERNEL=queue).  Measured on B200, 1 M C4 states (profiles/README.md): the queue keeps 32 of
// 32 lanes busy in the evaluation (loop: 16) and matches the loop kernel for bubble points (2.18 vs 2.14 ms),
// but for dew points the three bookkeeping phases of the state machine plus the evaluation no longer fit the
// instruction cache (53 % of the stall samples are "no instruction") and it is 2.5x slower (5.79 vs 2.29 ms).
inline int point_kernel_variant() {
  const char* e = getenv("FEOS_B200_POINT_KERNEL");
  return (e && e[0] == 'q') ? 1 : 0;
}

}  // namespace feos
