// sm_100a kernels and C-ABI entry points of the binary PC-SAFT bubble / dew point path
// (include/feos_b200.h: feos_dev_mix_*, feos_host_pcsaft_bubble_point / _dew_point).
//
//   mix_point_fwd_kernel     one thread per state, fast path (mix_solver.cuh); states it cannot finish
//                            go to a retry list handled by
//   mix_point_retry_kernel   the robust path over the compacted retry list;
//   mix_point_bwd_kernel     gradient of the Newton-step pressure w.r.t. 2x8 parameters, k_ij, eps_AiBj
//                            and T: closed-form coefficient adjoints + reverse sweep for non-associating
//                            states, passes of 3 dual directions for associating ones (mix_grad.cuh);
//   mix_state_kernel         a, p, mu_i, v_i at given (T, rho_1, rho_2)  (PcSaftMix.derivatives).
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "../../include/feos_b200.h"
#include "common.h"
#include "mix_grad.cuh"
#include "point_queue.cuh"

namespace feos {

constexpr int kMixBlock = 128;

// The (parameters, k_ij, T)-only coefficients of a state (~70 doubles) live in SHARED memory, one slot
// per thread, so that the registers hold only the hyper-dual working set of the evaluation.  The slot
// stride is an odd number of 8-byte words: thread t touches word (stride * t + field), conflict-free.
template <class Model>
struct ModelSlot {
  static constexpr int kWords = (sizeof(Model) + 7) / 8;
  static constexpr int kStride = kWords | 1;
  static constexpr size_t kSmem = sizeof(double) * kStride * kMixBlock;
  static __device__ __forceinline__ Model& get(double* smem) {
    return *reinterpret_cast<Model*>(smem + (size_t)kStride * threadIdx.x);
  }
};
constexpr size_t kMixSmem = ModelSlot<PcSaftBinary>::kSmem;
constexpr size_t kMixPlainSmem = ModelSlot<PcSaftBinaryPlain>::kSmem;

__device__ __forceinline__ PcSaftBinary& model_slot(double* smem) { return ModelSlot<PcSaftBinary>::get(smem); }

// flags of the bubble / dew entry points: the parameter rows / the k_ij row are ONE row shared by all states
// (a fit of shared parameters: `pair.expand(n, 2, 8)`, `kij.expand(n, 2)` in the reference's API)
constexpr int kSharedParams = FEOS_SHARED_PARAMS, kSharedKij = FEOS_SHARED_KIJ;

__device__ __forceinline__ void load_mix(const double* __restrict__ params, const double* __restrict__ kij, int64_t i,
                                         int flags, double (&prm)[16], double (&kk)[2]) {
  const double2* p = reinterpret_cast<const double2*>(params + ((flags & kSharedParams) ? 0 : 16 * i));
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    double2 t = __ldg(p + k);
    prm[2 * k] = t.x;
    prm[2 * k + 1] = t.y;
  }
  double2 t = __ldg(reinterpret_cast<const double2*>(kij) + ((flags & kSharedKij) ? 0 : i));
  kk[0] = t.x;
  kk[1] = t.y;
}

template <class Model>
__device__ __forceinline__ void store_point(int64_t i, int rc, bool bubble, double z0, double T, const PointState& st,
                                            const PointState& se, const MixEval& es, const MixEval& ei,
                                            double* __restrict__ rho, double* __restrict__ value,
                                            uint8_t* __restrict__ failed) {
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  double r[4] = {nan, nan, nan, nan};
  double val = nan;
  if (rc == kStatusOk) {
    double re[4];
    point_densities(st, bubble, z0, r);
    point_densities(se, bubble, z0, re);
    double spec[2], inc[2];
    if (bubble) { spec[0] = re[2]; spec[1] = re[3]; inc[0] = re[0]; inc[1] = re[1]; }
    else { spec[0] = re[0]; spec[1] = re[1]; inc[0] = re[2]; inc[1] = re[3]; }
    val = point_pressure(es, ei.a, spec, inc) * T * kPaPerRed;  // pcsaft_mix.py:443-444
  }
  double2* o = reinterpret_cast<double2*>(rho + 4 * i);
  o[0] = make_double2(r[0], r[1]);
  o[1] = make_double2(r[2], r[3]);
  if (value) value[i] = val;
  failed[i] = (rc == kStatusOk) ? 0 : 1;
}

#ifndef FEOS_MIX_FWD_MINB
#define FEOS_MIX_FWD_MINB 2
#endif
#ifndef FEOS_MIX_PLAIN_MINB
#define FEOS_MIX_PLAIN_MINB 3
#endif
// Two instantiations share the work by the CLASS of a state's parameter pair (mix_is_plain): PLAIN = hard sphere +
// hard chain + dispersion only, the other one every pair with a dipole or association term.  Each skips the states
// of the other class (the state order puts the classes into separate blocks); the plain one raises `other_class`
// when it meets such a state, and the blocks of the general one leave at once while that flag is down.
template <bool PLAIN>
__global__ void __launch_bounds__(kMixBlock, PLAIN ? FEOS_MIX_PLAIN_MINB : FEOS_MIX_FWD_MINB)
mix_point_fwd_kernel(const double* __restrict__ params, const double* __restrict__ kij,
                     const double* __restrict__ temperature, const double* __restrict__ molefrac, int64_t n, int bubble,
                     int flags, double* __restrict__ rho, double* __restrict__ value, uint8_t* __restrict__ failed,
                     int32_t* __restrict__ evals, int32_t* __restrict__ retry_list, int32_t* __restrict__ retry_count,
                     const int32_t* __restrict__ order, int32_t* __restrict__ other_class) {
  using Model = typename std::conditional<PLAIN, PcSaftBinaryPlain, PcSaftBinary>::type;
  extern __shared__ double smem[];
  if (!PLAIN && *reinterpret_cast<volatile int32_t*>(other_class) == 0) return;
  Model& b = ModelSlot<Model>::get(smem);
  // the plain instantiation is launched with one block per tile of kMixBlock states; the general one with a grid of
  // resident blocks that walk the tiles (so that its launch costs nothing when there is no such state)
#pragma unroll 1
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = order ? (int64_t)order[t] : t;  // states of similar (T, x) and code path share a warp (order.cuh)
    double T, z0;
    {
      double prm[16], kk[2];
      load_mix(params, kij, i, flags, prm, kk);
      if (mix_is_plain(prm) != PLAIN) {
        if (PLAIN) *other_class = 1;
        continue;
      }
      T = temperature[i];
      z0 = molefrac[i];
      pcsaft_binary_init(b, prm, kk, T);
    }
    PointState st, se;
    MixEval es, ei;
    int ne = 0;
    int rc = point_fast(b, bubble != 0, z0, st, es, ei, se, ne);
    if (evals) evals[i] = ne;
    if (rc == kStatusRetry) {
      retry_list[atomicAdd(retry_count, 1)] = (int32_t)i;
      continue;
    }
    store_point<Model>(i, rc, bubble != 0, z0, T, st, se, es, ei, rho, value, failed);
  }
}

// Persistent work-queue variant (point_queue.cuh): same solver as mix_point_fwd_kernel, cut into
// one-evaluation trips so that Newton iterations of divergent lengths do not idle lanes.
struct MixQueueSource {
  const double* params;
  const double* kij;
  const double* temperature;
  const double* molefrac;
  double* rho;
  double* value;
  uint8_t* failed;
  int flags;
  __device__ __forceinline__ void dummy(PcSaftBinary& b) const {
    const double prm[16] = {1.0, 3.5, 150.0, 0, 0, 0, 0, 0, 1.0, 3.5, 150.0, 0, 0, 0, 0, 0};
    const double kk[2] = {0.0, 0.0};
    pcsaft_binary_init(b, prm, kk, 300.0);
  }
  __device__ __forceinline__ bool load(int i, PcSaftBinary& b, double& z0, double& T) const {
    T = temperature[i];
    z0 = molefrac[i];
    double prm[16], kk[2];
    load_mix(params, kij, i, flags, prm, kk);
    pcsaft_binary_init(b, prm, kk, T);
    return z0 > 0.0 && z0 < 1.0;
  }
  __device__ __forceinline__ void store(int i, int rc, bool bubble, double z0, double T,
                                        const PointMachine<PcSaftBinary>& m) const {
    store_point<PcSaftBinary>(i, rc, bubble, z0, T, m.st, m.st_eval, m.es, m.ei, rho, value, failed);
  }
};

__global__ void __launch_bounds__(kMixBlock, 2)
mix_point_queue_kernel(MixQueueSource src, int n, int bubble, int32_t* __restrict__ evals,
                       int32_t* __restrict__ retry_list, int32_t* __restrict__ retry_count,
                       int32_t* __restrict__ queue_head) {
  extern __shared__ double smem[];
  point_queue_body<PcSaftBinary, MixQueueSource>(src, model_slot(smem), n, bubble != 0, evals, retry_list, retry_count,
                                                 queue_head);
}

__global__ void __launch_bounds__(64)
mix_point_retry_kernel(const double* __restrict__ params, const double* __restrict__ kij,
                       const double* __restrict__ temperature, const double* __restrict__ molefrac,
                       const double* __restrict__ start_pressure, int bubble, int flags, double* __restrict__ rho,
                       double* __restrict__ value, uint8_t* __restrict__ failed, int32_t* __restrict__ evals,
                       const int32_t* __restrict__ retry_list, const int32_t* __restrict__ retry_count) {
  int count = *retry_count;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
    int64_t i = retry_list[k];
    double prm[16], kk[2];
    load_mix(params, kij, i, flags, prm, kk);
    double T = temperature[i], z0 = molefrac[i];
    PcSaftBinary b;
    pcsaft_binary_init(b, prm, kk, T);
    PointState st, se;
    MixEval es, ei;
    int ne = 0;
    // the caller's starting pressure (reduced), used where the solver cannot build a start of its own
    double p_start = (start_pressure ? start_pressure[i] : kDefaultStartPressure) / (T * kPaPerRed);
    int rc = point_robust(b, bubble != 0, z0, p_start, st, es, ei, se, ne);
    if (evals) evals[i] += ne;
    store_point<PcSaftBinary>(i, rc, bubble != 0, z0, T, st, se, es, ei, rho, value, failed);
  }
}

// a only (the incipient phase of the Newton-step pressure needs no derivatives)
__device__ __noinline__ double mix_value(const MixCoef<double>& c, double r0, double r1) {
  return mix_helmholtz<double, double, 2, false>(c, r0, r1);
}

// Sum of `v` over the block into dst[k] (k < 19 gradient slots in shared memory `acc`): warp shuffle first --
// shared-memory f64 atomics are compare-and-swap loops -- then one shared atomic per warp.
__device__ __forceinline__ void warp_sum_to(double v, double* acc_k) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(acc_k, v);
}

#ifndef FEOS_MIX_BWD_MINB
#define FEOS_MIX_BWD_MINB 3  // 168 registers: 0.57 -> 0.51 ms per 1 M states (lanes are converged; occupancy pays)
#endif
// flags & kSharedParams: `params` is one [2,8] row and grad_params [2,8] receives the SUM over the states
// (block reduction, one atomic per slot and block -- the fit-loss gradient of shared parameters, SURVEY section
// 2.2 "fused with a per-block gradient reduction where parameters are shared"); likewise kSharedKij / grad_kij [2].
__global__ void __launch_bounds__(kMixBlock, FEOS_MIX_BWD_MINB)
mix_point_bwd_kernel(const double* __restrict__ params, const double* __restrict__ kij,
                     const double* __restrict__ temperature, const double* __restrict__ rho,
                     const uint8_t* __restrict__ failed, const int32_t* __restrict__ pos, const double* __restrict__ g,
                     int g_broadcast, int64_t n, int bubble, int flags, double* __restrict__ grad_params, double* __restrict__ grad_kij,
                     double* __restrict__ grad_t, const int32_t* __restrict__ order) {
  __shared__ double acc[18];
  const bool sh_p = (flags & kSharedParams) && grad_params, sh_k = (flags & kSharedKij) && grad_kij;
  if (sh_p || sh_k) {
    if (threadIdx.x < 18) acc[threadIdx.x] = 0.0;
    __syncthreads();
  }
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  bool active = i < n;
  if (active && order) i = order[i];
  if (active && failed[i]) {
    if (grad_params && !sh_p) {
      double2* gp = reinterpret_cast<double2*>(grad_params + 16 * i);
#pragma unroll
      for (int k = 0; k < 8; ++k) gp[k] = make_double2(0.0, 0.0);
    }
    if (grad_kij && !sh_k) reinterpret_cast<double2*>(grad_kij)[i] = make_double2(0.0, 0.0);
    if (grad_t) grad_t[i] = 0.0;
    active = false;
  }
  double o[19];
#pragma unroll
  for (int k = 0; k < 19; ++k) o[k] = 0.0;
  if (active) {
    double prm[16], kk[2];
    load_mix(params, kij, i, flags, prm, kk);
    double T = temperature[i];
    const double2* rp = reinterpret_cast<const double2*>(rho + 4 * i);
    double2 rv = rp[0], rl = rp[1];
    double spec[2], inc[2];
    if (bubble) { spec[0] = rl.x; spec[1] = rl.y; inc[0] = rv.x; inc[1] = rv.y; }
    else { spec[0] = rv.x; spec[1] = rv.y; inc[0] = rl.x; inc[1] = rl.y; }
    PointWeights q;
    double gi = g[g_broadcast ? 0 : (pos ? (int64_t)pos[i] : i)];
    bool closed;
    {
      PcSaftBinary b;
      pcsaft_binary_init(b, prm, kk, T);
      MixEval es = b.eval(spec[0], spec[1]);
      q = point_weights(es, spec, inc);
      closed = b.c.cat == kAssocNone;
      if (closed) {
        // non-associating state: closed-form coefficient adjoints + one reverse sweep through the coefficient stage
        double p_red = point_pressure(es, mix_value(b.c, inc[0], inc[1]), spec, inc);
        mix_point_grad_closed(prm, kk, T, b.c, spec, inc, q, p_red, gi, o);
      }
    }
    if (!closed) {
      // only the directions whose gradient is requested (a k_ij fit needs 1 pass instead of 7)
      int dirs[24];
      int nd = 0;
      if (grad_params)
        for (int d = 0; d < 16; ++d) dirs[nd++] = d;
      if (grad_kij) {
        dirs[nd++] = 16;
        dirs[nd++] = 17;
      }
      if (grad_t) dirs[nd++] = 18;
      for (int d = 0; d < 3; ++d) dirs[nd + d] = -1;
#pragma unroll 1
      for (int f = 0; f < nd; f += 3) {
        double g3[3];
        mix_point_grad_pass<3>(prm, kk, T, spec, inc, q, dirs + f, gi, g3);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          int d = dirs[f + k];
          if (d < 0) break;
#pragma unroll
          for (int j = 0; j < 19; ++j)
            if (j == d) o[j] = g3[k];
        }
      }
    }
    if (grad_params && !sh_p) {
      double2* gp = reinterpret_cast<double2*>(grad_params + 16 * i);
#pragma unroll
      for (int k = 0; k < 8; ++k) gp[k] = make_double2(o[2 * k], o[2 * k + 1]);
    }
    if (grad_kij && !sh_k) reinterpret_cast<double2*>(grad_kij)[i] = make_double2(o[16], o[17]);
    if (grad_t) grad_t[i] = o[18];
  }
  if (sh_p || sh_k) {
    if (sh_p) {
#pragma unroll
      for (int k = 0; k < 16; ++k) warp_sum_to(o[k], acc + k);
    }
    if (sh_k) {
      warp_sum_to(o[16], acc + 16);
      warp_sum_to(o[17], acc + 17);
    }
    __syncthreads();
    int k = threadIdx.x;
    if (k < 16 && sh_p && acc[k] != 0.0) atomicAdd(grad_params + k, acc[k]);
    if (k >= 16 && k < 18 && sh_k && acc[k] != 0.0) atomicAdd(grad_kij + (k - 16), acc[k]);
  }
}

// d p / d (k_ij, eps_AiBj) alone (a k_ij fit, BASELINE config C4): closed form for k_ij (mix_point_grad_kij: one
// evaluation of the specified phase for the weights + the two dispersion partials), a dual pass for eps_AiBj only
// where it exists (cross-associating pairs with kij[:, 1] != 0).  Same class split as the forward kernel.
#ifndef FEOS_MIX_BWDK_MINB
#define FEOS_MIX_BWDK_MINB 3
#endif
template <bool PLAIN>
__global__ void __launch_bounds__(kMixBlock, PLAIN ? FEOS_MIX_BWDK_MINB : 2)
mix_point_bwd_kij_kernel(const double* __restrict__ params, const double* __restrict__ kij,
                         const double* __restrict__ temperature, const double* __restrict__ rho,
                         const uint8_t* __restrict__ failed, const int32_t* __restrict__ pos, const double* __restrict__ g,
                         int g_broadcast, int64_t n, int bubble, int flags, double* __restrict__ grad_kij,
                         const int32_t* __restrict__ order, int32_t* __restrict__ other_class) {
  using Model = typename std::conditional<PLAIN, PcSaftBinaryPlain, PcSaftBinary>::type;
  __shared__ double acc[2];
  if (!PLAIN && *reinterpret_cast<volatile int32_t*>(other_class) == 0) return;
  const bool sh_k = (flags & kSharedKij) != 0;
  if (sh_k) {
    if (threadIdx.x < 2) acc[threadIdx.x] = 0.0;
    __syncthreads();
  }
  double s16 = 0.0, s17 = 0.0;  // sums over this thread's states (shared k_ij row)
  // plain: one block per tile; general: resident blocks walking the tiles (see mix_point_fwd_kernel)
#pragma unroll 1
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = order ? (int64_t)order[t] : t;
    double prm[16], kk[2];
    load_mix(params, kij, i, flags, prm, kk);
    if (mix_is_plain(prm) != PLAIN) {
      if (PLAIN) *other_class = 1;
      continue;
    }
    double o16 = 0.0, o17 = 0.0;
    if (!failed[i]) {
      double T = temperature[i];
      const double2* rp = reinterpret_cast<const double2*>(rho + 4 * i);
      double2 rv = rp[0], rl = rp[1];
      double spec[2], inc[2];
      if (bubble) { spec[0] = rl.x; spec[1] = rl.y; inc[0] = rv.x; inc[1] = rv.y; }
      else { spec[0] = rv.x; spec[1] = rv.y; inc[0] = rl.x; inc[1] = rl.y; }
      double gi = g[g_broadcast ? 0 : (pos ? (int64_t)pos[i] : i)];
      Model b;
      pcsaft_binary_init(b, prm, kk, T);
      PointWeights q = point_weights(b.eval(spec[0], spec[1]), spec, inc);
      o16 = mix_point_grad_kij(prm, kk[0], T, b.c, spec, inc, q, gi);
      if (!PLAIN && b.c.cat == kAssocCross && kk[1] != 0.0) {
        const int dirs[1] = {17};
        double g1[1];
        mix_point_grad_pass<1>(prm, kk, T, spec, inc, q, dirs, gi, g1);
        o17 = g1[0];
      }
    }
    if (sh_k) {
      s16 += o16;
      s17 += o17;
    } else {
      reinterpret_cast<double2*>(grad_kij)[i] = make_double2(o16, o17);
    }
  }
  if (sh_k) {
    warp_sum_to(s16, acc);
    if (!PLAIN) warp_sum_to(s17, acc + 1);
    __syncthreads();
    if (threadIdx.x < 2 && acc[threadIdx.x] != 0.0) atomicAdd(grad_kij + threadIdx.x, acc[threadIdx.x]);
  }
}

__global__ void __launch_bounds__(kMixBlock)
mix_state_kernel(const double* __restrict__ params, const double* __restrict__ kij,
                 const double* __restrict__ temperature, const double* __restrict__ density, int64_t n,
                 double* __restrict__ a, double* __restrict__ p, double* __restrict__ mu, double* __restrict__ v) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double prm[16], kk[2];
  load_mix(params, kij, i, 0, prm, kk);
  PcSaftBinary b;
  pcsaft_binary_init(b, prm, kk, temperature[i]);
  double2 r = reinterpret_cast<const double2*>(density)[i];
  MixEval e = b.eval(r.x, r.y);
  if (a) a[i] = e.a;
  if (p) p[i] = e.p;
  if (mu) reinterpret_cast<double2*>(mu)[i] = make_double2(e.mu[0], e.mu[1]);
  if (v) reinterpret_cast<double2*>(v)[i] = make_double2(e.v[0], e.v[1]);
}

}  // namespace feos

using namespace feos;

namespace {
// grid of the general-class instantiations: a few blocks per resident slot, each walking its share of the tiles
int general_grid(int64_t n, int resident_blocks) {
  return (int)std::min<int64_t>((n + kMixBlock - 1) / kMixBlock, 4 * (int64_t)std::max(resident_blocks, 1));
}

int host_point(const double* params, const double* kij, const double* temperature, const double* molefracs,
               const double* pressure, int64_t n, int bubble, double* rho_out, uint8_t* status, int64_t* n_ok) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31 && n_ok);
  *n_ok = 0;
  if (n == 0) return 0;
  FEOS_ARG(params && kij && temperature && molefracs && rho_out && status);
  HostCall hc;
  double *dp, *dk, *dt, *dz, *dp0 = nullptr, *drho;
  uint8_t* dfail;
  int32_t* dorder = nullptr;
  if (int rc = hc.upload(&dp, params, 16 * (size_t)n)) return rc;
  if (int rc = hc.upload(&dk, kij, 2 * (size_t)n)) return rc;
  if (int rc = hc.upload(&dt, temperature, (size_t)n)) return rc;
  if (int rc = hc.upload(&dz, molefracs, (size_t)n)) return rc;
  if (pressure)
    if (int rc = hc.upload(&dp0, pressure, (size_t)n)) return rc;
  if (int rc = hc.alloc(&drho, 4 * (size_t)n)) return rc;
  if (int rc = hc.alloc(&dfail, (size_t)n)) return rc;
  if (n >= kOrderMinStates) {
    if (int rc = hc.alloc(&dorder, (size_t)n)) return rc;
    if (int rc = feos_dev_mix_point_order(dp, dt, dz, n, 0, dorder, hc.s)) return rc;
  }
  if (int rc = feos_dev_mix_point_fwd(dp, dk, dt, dz, dp0, n, bubble, 0, drho, nullptr, dfail, nullptr, dorder, hc.s))
    return rc;
  return host_compact_download(hc, dfail, n, drho, 4, 4, rho_out, status, n_ok);  // filter_binary, src/pcsaft.rs:216-231
}
}  // namespace

extern "C" {

int feos_dev_mix_point_fwd(const double* params, const double* kij, const double* temperature,
                           const double* molefracs, const double* start_pressure, int64_t n, int bubble, int flags,
                           double* rho, double* pressure, uint8_t* failed, int32_t* evals, const int32_t* order,
                           void* stream) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31);
  if (n == 0) return 0;
  FEOS_ARG(params && kij && temperature && molefracs && rho && failed);
  cudaStream_t s = (cudaStream_t)stream;
  Scratch scratch;  // [0] = retry counter, [1] = queue head, [2] = "states of the general class present", [4 ..] = retry list
  if (int rc_ = scratch.alloc(sizeof(int32_t) * (size_t)(n + 4), s)) return rc_;
  int32_t* ws = scratch.as<int32_t>();
  FEOS_CUDA(cudaMemsetAsync(ws, 0, 4 * sizeof(int32_t), s));
  // opt-in to > 48 KB of dynamic shared memory and the persistent grid size: once per DEVICE
  static PerDevice cache;
  int blocks = 0;
  if (int rc_ = cache.get([](int dev, int* v) {
        FEOS_CUDA(cudaFuncSetAttribute(mix_point_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMixSmem));
        FEOS_CUDA(cudaFuncSetAttribute(mix_point_queue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMixSmem));
        int sms = 0, per_sm = 0;
        FEOS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        FEOS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mix_point_queue_kernel, kMixBlock, kMixSmem));
        *v = sms * (per_sm > 0 ? per_sm : 1);
        return 0;
      }, &blocks))
    return rc_;
  // default = one thread per state; FEOS_B200_POINT_KERNEL=queue selects the persistent work-queue kernel
  static const int variant = point_kernel_variant();
  if (variant == 0 || n < 65536) {
    mix_point_fwd_kernel<true><<<grid_for(n, kMixBlock), kMixBlock, kMixPlainSmem, s>>>(
        params, kij, temperature, molefracs, n, bubble, flags, rho, pressure, failed, evals, ws + 4, ws, order, ws + 2);
    if (!(flags & FEOS_PLAIN_ONLY)) {
      FEOS_LAUNCH_OK();
      mix_point_fwd_kernel<false><<<general_grid(n, blocks), kMixBlock, kMixSmem, s>>>(
          params, kij, temperature, molefracs, n, bubble, flags, rho, pressure, failed, evals, ws + 4, ws, order, ws + 2);
    }
  } else {
    MixQueueSource src{params, kij, temperature, molefracs, rho, pressure, failed, flags};
    int grid = (int)std::min<int64_t>(blocks, (n + kMixBlock - 1) / kMixBlock);
    mix_point_queue_kernel<<<grid, kMixBlock, kMixSmem, s>>>(src, (int)n, bubble, evals, ws + 4, ws, ws + 1);
  }
  FEOS_LAUNCH_OK();
  int retry_grid = (int)(n < 64 * 592 ? (n + 63) / 64 : 592);
  mix_point_retry_kernel<<<retry_grid, 64, 0, s>>>(params, kij, temperature, molefracs, start_pressure, bubble, flags, rho,
                                                  pressure, failed, evals, ws + 4, ws);
  FEOS_LAUNCH_OK();
  return 0;
}

int feos_dev_mix_point_bwd(const double* params, const double* kij, const double* temperature, const double* rho,
                           const uint8_t* failed, const int32_t* pos, const double* g, int g_broadcast, int64_t n,
                           int bubble, int flags, double* grad_params, double* grad_kij, double* grad_temperature, const int32_t* order,
                           void* stream) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31);
  if (n == 0) return 0;
  FEOS_ARG(params && kij && temperature && rho && failed && g);
  if (grad_kij && !grad_params && !grad_temperature) {
    cudaStream_t s = (cudaStream_t)stream;
    Scratch scratch;  // "states of the general class present"
    if (int rc_ = scratch.alloc(sizeof(int32_t), s)) return rc_;
    FEOS_CUDA(cudaMemsetAsync(scratch.p, 0, sizeof(int32_t), s));
    mix_point_bwd_kij_kernel<true><<<grid_for(n, kMixBlock), kMixBlock, 0, s>>>(
        params, kij, temperature, rho, failed, pos, g, g_broadcast, n, bubble, flags, grad_kij, order, scratch.as<int32_t>());
    FEOS_LAUNCH_OK();
    int sms = 148;
    {
      int dev = 0;
      if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    if (!(flags & FEOS_PLAIN_ONLY)) {
      mix_point_bwd_kij_kernel<false><<<general_grid(n, 2 * sms), kMixBlock, 0, s>>>(
          params, kij, temperature, rho, failed, pos, g, g_broadcast, n, bubble, flags, grad_kij, order, scratch.as<int32_t>());
      FEOS_LAUNCH_OK();
    }
    return 0;
  }
  mix_point_bwd_kernel<<<grid_for(n, kMixBlock), kMixBlock, 0, (cudaStream_t)stream>>>(
      params, kij, temperature, rho, failed, pos, g, g_broadcast, n, bubble, flags, grad_params, grad_kij, grad_temperature,
      order);
  FEOS_LAUNCH_OK();
  return 0;
}

int feos_dev_mix_state(const double* params, const double* kij, const double* temperature, const double* density,
                       int64_t n, double* a, double* p, double* mu, double* v, void* stream) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31);
  if (n == 0) return 0;
  FEOS_ARG(params && kij && temperature && density);
  mix_state_kernel<<<grid_for(n, kMixBlock), kMixBlock, 0, (cudaStream_t)stream>>>(params, kij, temperature, density, n,
                                                                                 a, p, mu, v);
  FEOS_LAUNCH_OK();
  return 0;
}

int feos_host_pcsaft_bubble_point(const double* params, const double* kij, const double* temperature,
                                  const double* liquid_molefracs, const double* pressure, int64_t n, double* rho_out,
                                  uint8_t* status, int64_t* n_ok) {
  return host_point(params, kij, temperature, liquid_molefracs, pressure, n, 1, rho_out, status, n_ok);
}

int feos_host_pcsaft_dew_point(const double* params, const double* kij, const double* temperature,
                               const double* vapor_molefracs, const double* pressure, int64_t n, double* rho_out,
                               uint8_t* status, int64_t* n_ok) {
  return host_point(params, kij, temperature, vapor_molefracs, pressure, n, 0, rho_out, status, n_ok);
}

}  // extern "C"
