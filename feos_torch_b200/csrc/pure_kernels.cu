// sm_100a kernels and C-ABI entry points of the pure-component hot path
// (include/feos_b200.h: feos_dev_pure_*, feos_host_pcsaft_vapor_pressure / _liquid_density).
//
// Roofline: FP64 FMA pipe (no tensor cores: not a contraction).  I/O per state is 72 B in and
// <= 41 B out against ~2e4 FP64 operations, i.e. far right of the HBM ridge.
//
//   pure_vle_fwd_kernel      one thread per state, fast path (pure_solver.cuh); states the fast path
//                            cannot finish are appended to a retry list (atomic counter), so the
//                            expensive robust path never stalls the 31 neighbours of a warp;
//   pure_vle_retry_kernel    robust path over the compacted retry list;
//   pure_vle_bwd_kernel      implicit-function-theorem gradients: closed-form base terms, dual passes only for
//                            dipole / association (pure_grad.cuh);
//   pure_rho_tp_*            liquid density at (T, p), forward and gradient;
//   pure_state_kernel        a, p, dp/drho at given (T, rho).
#include <algorithm>
#include <cstdlib>

#include "../../include/feos_b200.h"
#include "common.h"
#include "pure_grad.cuh"
#include "pure_solver.cuh"

namespace feos {

constexpr int kBlock = 128;
// backward kernels: straight-line code (no loop) of 40 .. 80 KB per state against a 32 KB instruction cache per SM --
// the warps of one block run it nearly in step and share the fetched lines, so the block size sets the hit rate.
// Measured (event-timed, 10 M associating C3 states / 1 M C2 states): pure_vle_bwd_kernel<true> 64 / 128 / 256 / 512
// threads: 2.47 / 2.45 / 2.25 / 2.32 ms; pure_vle_bwd_kernel<false> 0.146 / 0.146 / 0.152 / 0.164 ms;
// pure_rho_tp_bwd_kernel 0.105 / 0.105 / 0.109 / 0.123 ms.
constexpr int kBwdBlock = 128;      // pure_vle_bwd_kernel<false>, pure_rho_tp_bwd_kernel
constexpr int kBwdBlockRleq = 256;  // pure_vle_bwd_kernel<true> (the longest code: vapour pressure and liquid density parts)

struct PureRow {
  double v[8];
};

__device__ __forceinline__ PureRow load_row(const double* __restrict__ params, int64_t i) {
  // one 64 B row = 4 x 16 B vector loads; neighbouring threads read neighbouring rows
  const double2* p = reinterpret_cast<const double2*>(params + 8 * i);
  PureRow r;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double2 t = __ldg(p + k);
    r.v[2 * k] = t.x;
    r.v[2 * k + 1] = t.y;
  }
  return r;
}

__device__ __forceinline__ void make_coef(const PureRow& r, double T, PureCoef<double>& c, double& rho_max) {
  pure_coef<double>(c, r.v[0], r.v[1], r.v[2], r.v[3], r.v[4], r.v[5], r.v[6], r.v[7], T, false);
  rho_max = kMaxEta / (kPi / 6.0 * r.v[0] * r.v[1] * r.v[1] * r.v[1]);
}
// liquid start of the forward solves (liquid_start.cuh)
__device__ __forceinline__ double start_density(const PureRow& r, double T, const PureCoef<double>& c, double rho_max) {
  return pure_start_density(c, rho_max, T / r.v[2]);
}

__device__ __forceinline__ void store_vle(int64_t i, int st, double rho_l, double u, const PureEval& el,
                                          const PureEval& ev, double rle, double rve, double T, double* rho,
                                          double* vp, double* rleq, uint8_t* failed) {
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  double o_v = nan, o_l = nan, o_vp = nan, o_rl = nan;
  if (st == kStatusOk) {
    o_v = exp(u);
    o_l = rho_l;
    pure_vle_properties(el, ev, rle, rve, T, o_vp, o_rl);
  }
  reinterpret_cast<double2*>(rho)[i] = make_double2(o_v, o_l);
  if (vp) vp[i] = o_vp;
  if (rleq) rleq[i] = o_rl;
  failed[i] = (st == kStatusOk) ? 0 : 1;
}

#ifndef FEOS_PURE_FWD_MINB
#define FEOS_PURE_FWD_MINB 4
#endif
__global__ void __launch_bounds__(kBlock, FEOS_PURE_FWD_MINB)
pure_vle_fwd_kernel(const double* __restrict__ params, const double* __restrict__ temperature, int64_t n,
                    double* __restrict__ rho, double* __restrict__ vp, double* __restrict__ rleq,
                    uint8_t* __restrict__ failed, int32_t* __restrict__ evals, int32_t* __restrict__ retry_list,
                    int32_t* __restrict__ retry_count, const int32_t* __restrict__ order) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (order) i = order[i];  // states of similar (m, T / T_c) and code path share a warp (order.cuh)
  PureRow r = load_row(params, i);
  double T = temperature[i];
  PureCoef<double> c;
  double rho_max;
  make_coef(r, T, c, rho_max);
  double rho_l, u, rle, rve;
  PureEval el, ev;
  int ne = 0;
  // associating states call the out-of-line evaluation (instruction cache), the others inline it (pure_solver.cuh)
  int st = c.assoc ? pure_vle_fast<true>(c, rho_max, rho_l, u, el, ev, rle, rve, ne)
                   : pure_vle_fast<false>(c, rho_max, rho_l, u, el, ev, rle, rve, ne, start_density(r, T, c, rho_max));
  if (evals) evals[i] = ne;
  if (st == kStatusRetry) {
    int slot = atomicAdd(retry_count, 1);
    retry_list[slot] = (int32_t)i;
    return;
  }
  store_vle(i, st, rho_l, u, el, ev, rle, rve, T, rho, vp, rleq, failed);
}

// Persistent work-queue variant of the fast path: every lane runs the resumable state machine
// (PureVleMachine) one Helmholtz evaluation per loop trip; lanes whose state has finished (4-11 Newton
// iterations, depending on T/T_c) take the next state from a global queue instead of idling until the
// slowest lane of the warp is done.  Refills are batched (>= kRefillMin idle lanes, warp-aggregated
// atomic) so that the divergent set-up section (parameter load + coefficient stage) is amortised.
constexpr int kRefillMin = 8;

// one copy of the evaluation code that all lanes of the warp enter together (inlined, the compiler
// specialises it per solver phase and the lanes diverge: 12.6 of 32 active lanes per instruction)
__device__ __noinline__ void pure_eval_call(const PureCoef<double>& c, double rho, PureEval& e) { e = pure_eval(c, rho); }

__global__ void __launch_bounds__(kBlock, 5)
pure_vle_queue_kernel(const double* __restrict__ params, const double* __restrict__ temperature, int n,
                      double* __restrict__ rho, double* __restrict__ vp, double* __restrict__ rleq,
                      uint8_t* __restrict__ failed, int32_t* __restrict__ evals, int32_t* __restrict__ retry_list,
                      int32_t* __restrict__ retry_count, int32_t* __restrict__ queue_head,
                      const int32_t* __restrict__ order) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  bool active = false, exhausted = false;
  int done = kRunning;  // status of a finished state whose results are still to be written
  int idx = 0, ne = 0;
  double T = 0.0;
  PureCoef<double> c;
  c.dipolar = c.assoc = false;  // idle lanes evaluate too (converged warp): keep them on the short path
  c.m = c.mm1 = c.c_eta = c.w1 = c.w2 = c.k1 = c.k2 = 1.0;
  PureVleMachine m;
  m.start(1e-3);
  while (true) {
    unsigned idle = __ballot_sync(full, !active);
    if (idle) {
      int nidle = __popc(idle);
      // Divergent section, entered when enough lanes wait: write the results of finished states and hand
      // out new ones.  Batching it amortises the set-up and store code over >= kRefillMin lanes.
      if ((!exhausted && nidle >= kRefillMin) || nidle == 32) {
        if (done != kRunning) {
          if (evals) evals[idx] = ne;
          if (done == kStatusRetry) {
            retry_list[atomicAdd(retry_count, 1)] = idx;
          } else {
            double u = m.u;
            store_vle(idx, done, m.rho_l, u, m.el, m.ev, m.rle, m.rve, T, rho, vp, rleq, failed);
          }
          done = kRunning;
        }
        if (!exhausted) {
          int base = 0;
          if (lane == 0) base = atomicAdd(queue_head, nidle);
          base = __shfl_sync(full, base, 0);
          int mine = base + __popc(idle & ((1u << lane) - 1u));
          if (!active && mine < n) {
            idx = order ? order[mine] : mine;
            PureRow r = load_row(params, idx);
            T = temperature[idx];
            double rho_max;
            make_coef(r, T, c, rho_max);
            m.start(rho_max, start_density(r, T, c, rho_max));
            ne = 0;
            active = true;
          }
          if (base + nidle >= n) exhausted = true;
          idle = __ballot_sync(full, !active);
        }
      }
      if (exhausted && idle == full) {
        // flush what finished since the last refill (reached with every lane idle)
        if (done != kRunning) {
          if (evals) evals[idx] = ne;
          if (done == kStatusRetry) {
            retry_list[atomicAdd(retry_count, 1)] = idx;
          } else {
            double u = m.u;
            store_vle(idx, done, m.rho_l, u, m.el, m.ev, m.rle, m.rve, T, rho, vp, rleq, failed);
          }
        }
        break;
      }
    }
    // converged: evaluation, log(rho_l) and exp(u) are issued by all 32 lanes; only the cheap
    // bookkeeping halves in between diverge by solver phase
    double rho_eval = m.density();
    PureEval e;
    __syncwarp();
    pure_eval_call(c, rho_eval, e);
    double L = log(m.rho_l);
    int st = kRunning;
    if (active) {
      ++ne;
      st = m.advance1(e, L);
    }
    __syncwarp();
    double X = exp(m.u);
    if (active) {
      if (st == kRunning) st = m.advance2(X);
      if (st != kRunning) {
        active = false;
        done = st;
      }
    }
  }
}

__global__ void __launch_bounds__(64)
pure_vle_retry_kernel(const double* __restrict__ params, const double* __restrict__ temperature,
                      double* __restrict__ rho, double* __restrict__ vp, double* __restrict__ rleq,
                      uint8_t* __restrict__ failed, int32_t* __restrict__ evals,
                      const int32_t* __restrict__ retry_list, const int32_t* __restrict__ retry_count) {
  int count = *retry_count;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
    int64_t i = retry_list[k];
    PureRow r = load_row(params, i);
    double T = temperature[i];
    PureCoef<double> c;
    double rho_max;
    make_coef(r, T, c, rho_max);
    double rho_l, u, rle, rve;
    PureEval el, ev;
    int ne = 0;
    int st = pure_vle_robust(c, rho_max, rho_l, u, el, ev, rle, rve, ne);
    if (evals) evals[i] += ne;
    store_vle(i, st, rho_l, u, el, ev, rle, rve, T, rho, vp, rleq, failed);
  }
}

#ifndef FEOS_PURE_BWD_MINB
#define FEOS_PURE_BWD_MINB 4  // 128 registers: 0.179 -> 0.160 ms (vle), 0.132 -> 0.124 ms (density) per 1 M states; 2 and 5 are slower
#endif
template <bool NEED_RLEQ>
__global__ void __launch_bounds__(NEED_RLEQ ? kBwdBlockRleq : kBwdBlock, FEOS_PURE_BWD_MINB * 128 / (NEED_RLEQ ? kBwdBlockRleq : kBwdBlock))
pure_vle_bwd_kernel(const double* __restrict__ params, const double* __restrict__ temperature,
                    const double* __restrict__ rho, const uint8_t* __restrict__ failed,
                    const int32_t* __restrict__ pos, const double* __restrict__ g_vp,
                    const double* __restrict__ g_rho, int g_broadcast, int64_t n, double* __restrict__ grad_params,
                    double* __restrict__ grad_t) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  // every input of the state is requested before the first use: one exposed memory latency per thread instead of
  // one per dependent step (mask -> row -> densities -> row map -> upstream gradient)
  const uint8_t bad = failed[i];
  PureRow r = load_row(params, i);
  double T = temperature[i];
  double2 rr = reinterpret_cast<const double2*>(rho)[i];
  int64_t gi_idx = g_broadcast ? 0 : (pos ? (int64_t)pos[i] : i);
  if (bad) {
    double2* gp = reinterpret_cast<double2*>(grad_params + 8 * i);
#pragma unroll
    for (int k = 0; k < 4; ++k) gp[k] = make_double2(0.0, 0.0);
    if (grad_t) grad_t[i] = 0.0;
    return;
  }
  double gv = g_vp ? g_vp[gi_idx] : 0.0;
  double gr = 0.0, dpl = 1.0;
  if (NEED_RLEQ) {
    gr = g_rho[gi_idx];
    PureCoef<double> c;
    double rho_max;
    make_coef(r, T, c, rho_max);
    dpl = pure_eval(c, rr.y).dp;
  }
  double out[9];
  pure_vle_grad<NEED_RLEQ>(r.v, T, rr.x, rr.y, dpl, gv, gr, grad_t != nullptr, out);
  double2* gp = reinterpret_cast<double2*>(grad_params + 8 * i);
#pragma unroll
  for (int k = 0; k < 4; ++k) gp[k] = make_double2(out[2 * k], out[2 * k + 1]);
  if (grad_t) grad_t[i] = out[8];
}

__global__ void __launch_bounds__(kBlock, 4)
pure_rho_tp_fwd_kernel(const double* __restrict__ params, const double* __restrict__ temperature,
                       const double* __restrict__ pressure, int64_t n, double* __restrict__ rho,
                       double* __restrict__ value, uint8_t* __restrict__ failed, int32_t* __restrict__ evals,
                       const int32_t* __restrict__ order) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (order) i = order[i];
  PureRow r = load_row(params, i);
  double T = temperature[i];
  double p_raw = pressure[i];
  PureCoef<double> c;
  double rho_max;
  make_coef(r, T, c, rho_max);
  double p_t = p_raw / (T * kPaPerRed);  // pcsaft_pure.py:196
  double x;
  PureEval e;
  int ne = 0;
  int st = c.assoc ? pure_liquid_density<true>(c, rho_max, p_t, x, e, ne)
                   : pure_liquid_density<false>(c, rho_max, p_t, x, e, ne, start_density(r, T, c, rho_max));
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  double rn = (st == kStatusOk) ? x - (e.p - p_t) / e.dp : nan;  // pcsaft_pure.py:198
  rho[i] = rn;
  if (value) value[i] = rn * kKmolM3PerRed;
  failed[i] = (st == kStatusOk) ? 0 : 1;
  if (evals) evals[i] = ne;
}

__global__ void __launch_bounds__(kBwdBlock, FEOS_PURE_BWD_MINB)
pure_rho_tp_bwd_kernel(const double* __restrict__ params, const double* __restrict__ temperature,
                       const double* __restrict__ pressure, const double* __restrict__ rho,
                       const uint8_t* __restrict__ failed, const int32_t* __restrict__ pos,
                       const double* __restrict__ g, int g_broadcast, int64_t n, double* __restrict__ grad_params,
                       double* __restrict__ grad_t, double* __restrict__ grad_p) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  // all inputs requested before the first use (see pure_vle_bwd_kernel)
  const uint8_t bad = failed[i];
  PureRow r = load_row(params, i);
  double T = temperature[i];
  double x = rho[i];
  double p_raw = pressure[i];
  const int64_t gi_idx = g_broadcast ? 0 : (pos ? (int64_t)pos[i] : i);
  if (bad) {
    if (grad_params) {
      double2* gp = reinterpret_cast<double2*>(grad_params + 8 * i);
#pragma unroll
      for (int k = 0; k < 4; ++k) gp[k] = make_double2(0.0, 0.0);
    }
    if (grad_t) grad_t[i] = 0.0;
    if (grad_p) grad_p[i] = 0.0;
    return;
  }
  double gi = g[gi_idx];
  PureCoef<double> c;
  double rho_max;
  make_coef(r, T, c, rho_max);
  double dp = pure_eval(c, x).dp;
  double p_t = p_raw / (T * kPaPerRed);
  if (grad_p) grad_p[i] = gi / dp / (T * kPaPerRed) * kKmolM3PerRed;
  if (!grad_params && !grad_t) return;
  double out[9];
  pure_rho_tp_grad(r.v, T, x, p_t, dp, gi, out);
  if (grad_params) {
    double2* gp = reinterpret_cast<double2*>(grad_params + 8 * i);
#pragma unroll
    for (int k = 0; k < 4; ++k) gp[k] = make_double2(out[2 * k], out[2 * k + 1]);
  }
  if (grad_t) grad_t[i] = out[8];
}

__global__ void __launch_bounds__(kBlock)
pure_state_kernel(const double* __restrict__ params, const double* __restrict__ temperature,
                  const double* __restrict__ density, int64_t n, double* __restrict__ a, double* __restrict__ p,
                  double* __restrict__ dp) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  PureRow r = load_row(params, i);
  PureCoef<double> c;
  double rho_max;
  make_coef(r, temperature[i], c, rho_max);
  PureEval e = pure_eval(c, density[i]);
  if (a) a[i] = e.a;
  if (p) p[i] = e.p;
  if (dp) dp[i] = e.dp;
}

}  // namespace feos

using namespace feos;

namespace {
// state order for the host entry points (large batches only; NULL = input order otherwise)
int host_order(HostCall& hc, int32_t** order, const double* params, const double* temperature, int64_t n) {
  *order = nullptr;
  if (n < kOrderMinStates) return 0;
  if (int rc = hc.alloc(order, (size_t)n)) return rc;
  return feos_dev_pure_order(params, temperature, n, *order, hc.s);
}
}  // namespace

extern "C" {

int feos_dev_pure_vle_fwd(const double* params, const double* temperature, int64_t n, double* rho,
                          double* vapor_pressure, double* liquid_density, uint8_t* failed, int32_t* evals,
                          const int32_t* order, void* stream) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31);
  if (n == 0) return 0;
  FEOS_ARG(params && temperature && rho && failed);
  cudaStream_t s = (cudaStream_t)stream;
  Scratch scratch;  // [0] = retry counter, [1] = queue head, [2..n+1] = retry list
  if (int rc_ = scratch.alloc(sizeof(int32_t) * (size_t)(n + 2), s)) return rc_;
  int32_t* ws = scratch.as<int32_t>();
  FEOS_CUDA(cudaMemsetAsync(ws, 0, 2 * sizeof(int32_t), s));
  // FEOS_B200_PURE_VLE_KERNEL = "loop" (one thread per state) / "queue" (persistent work queue) for A/B runs.
  // Default: states in input order -> queue (1 M C2 states, ncu: 0.566 ms, 22.4 of 32 lanes active per
  // instruction; loop 0.572 ms, 16.9 lanes); states in the order of order.cuh -> loop (warps are homogeneous,
  // so the queue's bookkeeping buys nothing: 0.40 ms against 0.49 ms; 10 M associating C3 states: 6.9 ms
  // against 7.0 ms, and 7.6 ms for the queue in input order, profiles/README.md).
  const char* env = getenv("FEOS_B200_PURE_VLE_KERNEL");
  const int forced = !env ? -1 : (env[0] == 'l' ? 0 : (env[0] == 'q' ? 1 : -1));
  const int variant = forced >= 0 ? forced : (order ? 0 : 1);
  if (variant == 0 || n < 65536) {
    pure_vle_fwd_kernel<<<grid_for(n, kBlock), kBlock, 0, s>>>(params, temperature, n, rho, vapor_pressure,
                                                             liquid_density, failed, evals, ws + 2, ws, order);
  } else {
    static PerDevice cache;
    int blocks = 0;
    if (int rc_ = cache.get([](int dev, int* v) {
          int sms = 0, per_sm = 0;
          FEOS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
          FEOS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pure_vle_queue_kernel, kBlock, 0));
          *v = sms * (per_sm > 0 ? per_sm : 1);  // persistent grid: every resident slot of every SM
          return 0;
        }, &blocks))
      return rc_;
    int grid = (int)std::min<int64_t>(blocks, (n + kBlock - 1) / kBlock);
    pure_vle_queue_kernel<<<grid, kBlock, 0, s>>>(params, temperature, (int)n, rho, vapor_pressure, liquid_density,
                                                failed, evals, ws + 2, ws, ws + 1, order);
  }
  FEOS_LAUNCH_OK();
  int retry_grid = (int)(n < 64 * 592 ? (n + 63) / 64 : 592);
  pure_vle_retry_kernel<<<retry_grid, 64, 0, s>>>(params, temperature, rho, vapor_pressure, liquid_density, failed,
                                                 evals, ws + 2, ws);
  FEOS_LAUNCH_OK();
  return 0;
}

int feos_dev_pure_vle_bwd(const double* params, const double* temperature, const double* rho,
                          const uint8_t* failed, const int32_t* pos, const double* g_vp, const double* g_rho,
                          int g_broadcast, int64_t n, double* grad_params, double* grad_temperature, void* stream) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31);
  if (n == 0) return 0;
  FEOS_ARG(params && temperature && rho && failed && grad_params);
  FEOS_ARG(g_vp || g_rho);
  cudaStream_t s = (cudaStream_t)stream;
  if (g_rho)
    pure_vle_bwd_kernel<true><<<grid_for(n, kBwdBlockRleq), kBwdBlockRleq, 0, s>>>(params, temperature, rho, failed, pos, g_vp, g_rho,
                                                                   g_broadcast, n, grad_params, grad_temperature);
  else
    pure_vle_bwd_kernel<false><<<grid_for(n, kBwdBlock), kBwdBlock, 0, s>>>(params, temperature, rho, failed, pos, g_vp, g_rho,
                                                                    g_broadcast, n, grad_params, grad_temperature);
  FEOS_LAUNCH_OK();
  return 0;
}

int feos_dev_pure_rho_tp_fwd(const double* params, const double* temperature, const double* pressure,
                             int64_t n, double* rho, double* liquid_density, uint8_t* failed, int32_t* evals,
                             const int32_t* order, void* stream) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31);
  if (n == 0) return 0;
  FEOS_ARG(params && temperature && pressure && rho && failed);
  pure_rho_tp_fwd_kernel<<<grid_for(n, kBlock), kBlock, 0, (cudaStream_t)stream>>>(params, temperature, pressure, n, rho,
                                                                                 liquid_density, failed, evals, order);
  FEOS_LAUNCH_OK();
  return 0;
}

int feos_dev_pure_rho_tp_bwd(const double* params, const double* temperature, const double* pressure,
                             const double* rho, const uint8_t* failed, const int32_t* pos, const double* g,
                             int g_broadcast, int64_t n, double* grad_params, double* grad_temperature,
                             double* grad_pressure, void* stream) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31);
  if (n == 0) return 0;
  FEOS_ARG(params && temperature && pressure && rho && failed && g);
  pure_rho_tp_bwd_kernel<<<grid_for(n, kBwdBlock), kBwdBlock, 0, (cudaStream_t)stream>>>(
      params, temperature, pressure, rho, failed, pos, g, g_broadcast, n, grad_params, grad_temperature, grad_pressure);
  FEOS_LAUNCH_OK();
  return 0;
}

int feos_dev_pure_state(const double* params, const double* temperature, const double* density, int64_t n,
                        double* a, double* p, double* dp, void* stream) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31);
  if (n == 0) return 0;
  FEOS_ARG(params && temperature && density);
  pure_state_kernel<<<grid_for(n, kBlock), kBlock, 0, (cudaStream_t)stream>>>(params, temperature, density, n, a, p, dp);
  FEOS_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// host-buffer entry points (reference native signatures)
// ---------------------------------------------------------------------------------------------

int feos_host_pcsaft_vapor_pressure(const double* params, const double* temperature, int64_t n, double* rho_out,
                                    uint8_t* status, int64_t* n_ok) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31 && n_ok);
  *n_ok = 0;
  if (n == 0) return 0;
  FEOS_ARG(params && temperature && rho_out && status);
  HostCall hc;
  double *dp, *dt, *drho;
  uint8_t* dfail;
  int32_t* dorder;
  if (int rc = hc.upload(&dp, params, 8 * (size_t)n)) return rc;
  if (int rc = hc.upload(&dt, temperature, (size_t)n)) return rc;
  if (int rc = hc.alloc(&drho, 2 * (size_t)n)) return rc;
  if (int rc = hc.alloc(&dfail, (size_t)n)) return rc;
  if (int rc = host_order(hc, &dorder, dp, dt, n)) return rc;
  if (int rc = feos_dev_pure_vle_fwd(dp, dt, n, drho, nullptr, nullptr, dfail, nullptr, dorder, hc.s)) return rc;
  // rows (rho_V, rho_L, 0, 0) of the converged states: src/pcsaft.rs:95-101 leaves columns 2, 3 at zero
  return host_compact_download(hc, dfail, n, drho, 2, 4, rho_out, status, n_ok);
}

int feos_host_pcsaft_liquid_density(const double* params, const double* temperature, const double* pressure,
                                    int64_t n, double* rho_out, uint8_t* status, int64_t* n_ok) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31 && n_ok);
  *n_ok = 0;
  if (n == 0) return 0;
  FEOS_ARG(params && temperature && pressure && rho_out && status);
  HostCall hc;
  double *dp, *dt, *dpr, *drho;
  uint8_t* dfail;
  int32_t* dorder;
  if (int rc = hc.upload(&dp, params, 8 * (size_t)n)) return rc;
  if (int rc = hc.upload(&dt, temperature, (size_t)n)) return rc;
  if (int rc = hc.upload(&dpr, pressure, (size_t)n)) return rc;
  if (int rc = hc.alloc(&drho, (size_t)n)) return rc;
  if (int rc = hc.alloc(&dfail, (size_t)n)) return rc;
  if (int rc = host_order(hc, &dorder, dp, dt, n)) return rc;
  if (int rc = feos_dev_pure_rho_tp_fwd(dp, dt, dpr, n, drho, nullptr, dfail, nullptr, dorder, hc.s)) return rc;
  return host_compact_download(hc, dfail, n, drho, 1, 1, rho_out, status, n_ok);
}

}  // extern "C"
