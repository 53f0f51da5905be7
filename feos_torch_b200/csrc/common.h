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

// Per-DEVICE one-time values (opt-in shared-memory attribute set, occupancy-derived persistent grid size): function
// attributes and SM counts belong to a device, and one process may drive several.
constexpr int kMaxDevices = 64;
struct PerDevice {
  int value[kMaxDevices] = {};
  bool done[kMaxDevices] = {};
  // returns the cached value of the current device, computing it with `make(dev, &value)` (0 = ok) the first time
  template <class F>
  int get(F make, int* out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -2;
    if (!done[dev]) {
      if (int rc = make(dev, &value[dev])) return rc;
      done[dev] = true;
    }
    *out = value[dev];
    return 0;
  }
};

// Stream-ordered scratch buffer that is given back on EVERY return path of an entry point (error returns included).
struct Scratch {
  void* p = nullptr;
  cudaStream_t s = nullptr;
  Scratch() = default;
  Scratch(const Scratch&) = delete;
  Scratch& operator=(const Scratch&) = delete;
  ~Scratch() {
    if (p) cudaFreeAsync(p, s);
  }
  int alloc(size_t bytes, cudaStream_t stream) {
    s = stream;
    return scratch_alloc(&p, bytes, stream);
  }
  template <class T>
  T* as() const { return (T*)p; }
};

// One call of a host-buffer entry point (feos_host_*): its device buffers come from the stream-ordered pool (no
// cudaMalloc / cudaFree per call) and everything runs on the calling thread's own non-blocking stream instead of
// the legacy default stream, so concurrent callers and the torch layer do not serialise on each other.
struct HostCall {
  static constexpr int kMax = 16;
  Scratch bufs[kMax];
  int used = 0;
  cudaStream_t s = cudaStreamPerThread;
  template <class T>
  int alloc(T** d, size_t count) {
    if (used >= kMax) return -2;
    if (int rc = bufs[used].alloc(sizeof(T) * (count ? count : 1), s)) return rc;
    *d = bufs[used++].as<T>();
    return 0;
  }
  template <class T>
  int upload(T** d, const T* h, size_t count) {
    if (int rc = alloc(d, count)) return rc;
    if (cudaMemcpyAsync(*d, h, sizeof(T) * count, cudaMemcpyHostToDevice, s) != cudaSuccess) {
      set_error("host -> device copy of %zu bytes failed", sizeof(T) * count);
      return -2;
    }
    return 0;
  }
};

// Device-side row removal of the failed states + copy of the converged rows and the mask into the caller's host
// buffers (filter_binary, src/pcsaft.rs:216-231; the compaction loops at :94-101, :127-128).  Synchronises.
int host_compact_download(HostCall& hc, const uint8_t* d_failed, int64_t n, const double* d_values, int in_cols,
                          int out_cols, double* h_out, uint8_t* h_status, int64_t* n_ok);

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
