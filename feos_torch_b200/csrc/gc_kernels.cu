// sm_100a kernels and C-ABI entry points of the group-contribution PC-SAFT bubble / dew point path
// (include/feos_b200.h: feos_dev_gc_*, feos_host_gc_bubble_point / _dew_point).
//
//   gc_point_fwd_kernel / gc_point_retry_kernel   as the PC-SAFT mixture kernels, with the coefficient
//                            stage of gc_pcsaft.cuh (pair table + segment table, read through L1/L2);
//   gc_point_bwd_kernel      coefficient adjoints (forward-mode passes over the <= 70 coefficients), then
//                            dual passes over the inputs of the coefficient stage only; segment-parameter
//                            and k_ab gradients are SHARED by all states: they are reduced in shared
//                            memory per block and added to the global accumulators with one atomic per
//                            entry and block;
//   gc_state_kernel          a, p, mu_i, v_i at given (T, rho_1, rho_2)  (GcPcSaftMix.derivatives).
#include <algorithm>
#include <cstdlib>

#include "../../include/feos_b200.h"
#include "common.h"
#include "gc_pcsaft.cuh"
#include "point_queue.cuh"

static_assert(sizeof(feos_gc_pair) == sizeof(feos::GcPair), "feos_gc_pair layout");
static_assert(FEOS_GC_MAX_SEG == feos::kGcMaxSeg && FEOS_GC_MAX_BOND == feos::kGcMaxBond, "feos_gc_pair limits");

namespace feos {

constexpr int kGcBlock = 128;
// forward kernels: the ~120-double coefficient structure of a state lives in shared memory (one slot
// per thread, odd stride in 8-byte words: conflict-free), 64 threads per block -> 3 blocks per SM
constexpr int kGcFwdBlock = 64;
constexpr int kGcModelStride = ((sizeof(GcBinary) + 7) / 8) | 1;
constexpr size_t kGcSmem = sizeof(double) * kGcModelStride * kGcFwdBlock;
__device__ __forceinline__ GcBinary& gc_model_slot(double* smem) {
  return *reinterpret_cast<GcBinary*>(smem + (size_t)kGcModelStride * threadIdx.x);
}

struct GcArgs {
  const GcPair* pairs;
  const double* seg_params;
  const double* kab;
  const int32_t* pair_index;
  const double* phi;
  const double* temperature;
  const GcPairCoef* pair_coef = nullptr;  // per-pair part of the slim coefficient stage (gc_pair_coef_kernel), or NULL
};

// one block of two warps per molecule pair: the state-independent part of the slim coefficient stage (gc_pcsaft.cuh).
// It is a chain of a few thousand dependent instructions, so its length is what costs: warp 0 shares the
// O(n_seg^2) dispersion sums among its lanes, lane 0 of warp 1 does the molecule-level and dipole part meanwhile.
__global__ void __launch_bounds__(64)
gc_pair_coef_kernel(const GcPair* __restrict__ pairs, const double* __restrict__ seg_params,
                    const double* __restrict__ kab, GcPairCoef* __restrict__ out) {
  const GcPair& pr = pairs[blockIdx.x];
  GcPairCoef& o = out[blockIdx.x];
  if (threadIdx.x < 32) {
    double S1[3], S2[3];
    gc_pair_dispersion(seg_params, kab, pr, threadIdx.x, 32, S1, S2);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int w = 16; w > 0; w >>= 1) {
        S1[k] += __shfl_xor_sync(0xffffffffu, S1[k], w);
        S2[k] += __shfl_xor_sync(0xffffffffu, S2[k], w);
      }
    }
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        o.S1[k] = S1[k];
        o.S2[k] = S2[k];
      }
    }
  } else if (threadIdx.x == 32) {
    gc_pair_molecules(o, seg_params, kab, pr);
  }
}

__device__ __forceinline__ GcSource make_source(const GcArgs& g, int64_t i) {
  GcSource s;
  const int32_t pi = g.pair_index[i];
  s.pair = g.pairs + pi;
  s.pc = g.pair_coef ? g.pair_coef + pi : nullptr;
  s.seg_params = g.seg_params;
  s.kab = g.kab;
  if (g.phi) {
    double2 p = __ldg(reinterpret_cast<const double2*>(g.phi) + i);
    s.phi_[0] = p.x;
    s.phi_[1] = p.y;
  } else {
    s.phi_[0] = s.phi_[1] = 1.0;
  }
  s.T_ = g.temperature[i];
  return s;
}

__device__ __forceinline__ void store_gc_point(int64_t i, int rc, bool bubble, double z0, double T, const PointState& st,
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
    val = point_pressure(es, ei.a, spec, inc) * T * kPaPerRed;  // gc_pcsaft.py:489-490
  }
  double2* o = reinterpret_cast<double2*>(rho + 4 * i);
  o[0] = make_double2(r[0], r[1]);
  o[1] = make_double2(r[2], r[3]);
  if (value) value[i] = val;
  failed[i] = (rc == kStatusOk) ? 0 : 1;
}

#ifndef FEOS_GC_FWD_MINB
#define FEOS_GC_FWD_MINB 3
#endif
#ifndef FEOS_GC_SLIM_MINB
#define FEOS_GC_SLIM_MINB 3
#endif
constexpr int kGcSlimBlock = 128;
constexpr size_t kGcSlimSmem = sizeof(double) * kGcSlimWords * kGcSlimBlock;  // 45 KB: 3 blocks of 4 warps per SM (168 registers)

// Two instantiations share the work by the CLASS of a state's molecule pair (gc_is_slim): SLIM = no association term
// and at most kGcSlimBond bond types (45 doubles per state, 4-warp blocks), the other one everything else (GcBinary,
// 2-warp blocks).  Each skips the states of the other class -- the state order sorts by pair, so whole warps skip;
// the slim one raises `other_class` when it meets such a state, and the general one, launched as a small grid of
// blocks that walk the tiles, leaves at once while that flag is down.
// BUBBLE = 1 / 0: bubble / dew point compiled in (the slim instantiations: each is a third smaller than one kernel
// with both start procedures, and the kernel is instruction-fetch bound -- 117 KB of code ran at an 82 % hit rate of
// the SM's instruction cache); -1: taken from the argument.
template <bool SLIM, int BUBBLE>
__global__ void __launch_bounds__(SLIM ? kGcSlimBlock : kGcFwdBlock, SLIM ? FEOS_GC_SLIM_MINB : FEOS_GC_FWD_MINB)
gc_point_fwd_kernel(GcArgs g, const double* __restrict__ molefrac, int64_t n, int bubble_arg, double* __restrict__ rho,
                    double* __restrict__ value, uint8_t* __restrict__ failed, int32_t* __restrict__ evals,
                    int32_t* __restrict__ retry_list, int32_t* __restrict__ retry_count,
                    const int32_t* __restrict__ order, int32_t* __restrict__ other_class) {
  extern __shared__ double smem[];
  if (!SLIM && *reinterpret_cast<volatile int32_t*>(other_class) == 0) return;
  const bool bubble = BUBBLE < 0 ? bubble_arg != 0 : BUBBLE != 0;
#pragma unroll 1
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = order ? (int64_t)order[t] : t;  // states of one molecule pair and similar (T, x) share a warp
    GcSource src = make_source(g, i);
    if (gc_is_slim(src, *src.pair) != SLIM) {
      if (SLIM) *other_class = 1;
      continue;
    }
    double z0 = molefrac[i];
    PointState st, se;
    MixEval es, ei;
    int ne = 0;
    int rc;
    if (SLIM) {
      GcSlim b;
      gc_slim_init(b, smem + (size_t)kGcSlimWords * threadIdx.x, src);
      rc = point_fast(b, bubble, z0, st, es, ei, se, ne);
    } else {
      GcBinary& b = gc_model_slot(smem);
      gc_binary_init(b, src);
      rc = point_fast(b, bubble, z0, st, es, ei, se, ne);
    }
    if (evals) evals[i] = ne;
    if (rc == kStatusRetry) {
      retry_list[atomicAdd(retry_count, 1)] = (int32_t)i;
      continue;
    }
    store_gc_point(i, rc, bubble, z0, src.T_, st, se, es, ei, rho, value, failed);
  }
}

// Persistent work-queue variant (point_queue.cuh)
struct GcQueueSource {
  GcArgs g;
  const double* molefrac;
  double* rho;
  double* value;
  uint8_t* failed;
  __device__ __forceinline__ void dummy(GcBinary& b) const { gc_binary_init(b, make_source(g, 0)); }
  __device__ __forceinline__ bool load(int i, GcBinary& b, double& z0, double& T) const {
    GcSource src = make_source(g, i);
    gc_binary_init(b, src);
    T = src.T_;
    z0 = molefrac[i];
    return z0 > 0.0 && z0 < 1.0;
  }
  __device__ __forceinline__ void store(int i, int rc, bool bubble, double z0, double T, const PointMachine<GcBinary>& m) const {
    store_gc_point(i, rc, bubble, z0, T, m.st, m.st_eval, m.es, m.ei, rho, value, failed);
  }
};

__global__ void __launch_bounds__(kGcFwdBlock, 3)
gc_point_queue_kernel(GcQueueSource src, int n, int bubble, int32_t* __restrict__ evals, int32_t* __restrict__ retry_list,
                      int32_t* __restrict__ retry_count, int32_t* __restrict__ queue_head) {
  extern __shared__ double smem[];
  point_queue_body<GcBinary, GcQueueSource>(src, gc_model_slot(smem), n, bubble != 0, evals, retry_list, retry_count,
                                            queue_head);
}

__global__ void __launch_bounds__(64)
gc_point_retry_kernel(GcArgs g, const double* __restrict__ molefrac, const double* __restrict__ start_pressure, int bubble,
                      double* __restrict__ rho, double* __restrict__ value, uint8_t* __restrict__ failed,
                      int32_t* __restrict__ evals, const int32_t* __restrict__ retry_list,
                      const int32_t* __restrict__ retry_count) {
  int count = *retry_count;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
    int64_t i = retry_list[k];
    GcSource src = make_source(g, i);
    GcBinary b;
    gc_binary_init(b, src);
    double z0 = molefrac[i];
    PointState st, se;
    MixEval es, ei;
    int ne = 0;
    double p_start = (start_pressure ? start_pressure[i] : kDefaultStartPressure) / (src.T_ * kPaPerRed);
    int rc = point_robust(b, bubble != 0, z0, p_start, st, es, ei, se, ne);
    if (evals) evals[i] += ne;
    store_gc_point(i, rc, bubble != 0, z0, src.T_, st, se, es, ei, rho, value, failed);
  }
}

#ifndef FEOS_GC_BWD_MINB
#define FEOS_GC_BWD_MINB 4  // registers 255 / 168 / 128: 1.90 / 1.87 / 1.75 ms per 1 M states (converged lanes: occupancy pays)
#endif
__global__ void __launch_bounds__(kGcBlock, FEOS_GC_BWD_MINB)
gc_point_bwd_kernel(GcArgs g, int n_seg, int n_rec, const double* __restrict__ rho, const uint8_t* __restrict__ failed,
                    const int32_t* __restrict__ pos, const double* __restrict__ gout, int g_broadcast, int64_t n,
                    int bubble,
                    double* __restrict__ grad_seg, double* __restrict__ grad_kab, double* __restrict__ grad_phi,
                    double* __restrict__ grad_t, const int32_t* __restrict__ order) {
  extern __shared__ double acc[];  // [8 n_seg + n_rec] block-level sums of the shared-parameter gradients
  const int n_acc = 8 * n_seg + n_rec;
  for (int k = threadIdx.x; k < n_acc; k += blockDim.x) acc[k] = 0.0;
  __syncthreads();
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t i = t < n ? (order ? (int64_t)order[t] : t) : -1;
  const bool valid = i >= 0 && !failed[i];
  // every per-thread array whose address is handed to the (deliberately out-of-line) pass functions lives
  // in ONE stack object: as separate locals nvcc 12.9 -O3 overlapped them on sm_100a (see gc_pcsaft.cuh)
  struct Scratch {
    GcBinary b;
    double cbar[kGcMaxCoef];
    double gseg[8 * kGcMaxSeg];
    double gkab[kGcMaxSeg * kGcMaxSeg];
    int dirs[kGcMaxDirs + 4];
    GcDualCoef cd;
    double* seed[kGcMaxCoef];
  } w;
  const GcPair* my_pair = nullptr;
  if (i >= 0) {
    double gphi[2] = {0.0, 0.0}, gt = 0.0;
    if (valid) {
      GcSource src = make_source(g, i);
      const GcPair& pr = *src.pair;
      my_pair = src.pair;
      GcBinary& b = w.b;
      double* cbar = w.cbar;
      double* gseg = w.gseg;
      double* gkab = w.gkab;
      int* dirs = w.dirs;
      gc_binary_init(b, src);
      const double2* rp = reinterpret_cast<const double2*>(rho + 4 * i);
      double2 rv = rp[0], rl = rp[1];
      double spec[2], inc[2];
      if (bubble) { spec[0] = rl.x; spec[1] = rl.y; inc[0] = rv.x; inc[1] = rv.y; }
      else { spec[0] = rv.x; spec[1] = rv.y; inc[0] = rl.x; inc[1] = rl.y; }
      MixEval es = b.eval(spec[0], spec[1]);
      PointWeights q = point_weights(es, spec, inc);
      double p_red = point_pressure(es, gc_binary_value(b.c, inc[0], inc[1]), spec, inc);
      // A. coefficient adjoints: closed form for the base group, dual passes (dipole / association part of
      // the model only) for pairs that have those terms
      gc_coef_adjoints(b.c, spec, inc, q, w.cd, w.seed, cbar);
      // B. inputs of the coefficient stage
      double gi = gout[g_broadcast ? 0 : (pos ? (int64_t)pos[i] : i)];
      double scale = gi * src.T_ * kPaPerRed;
      // base group (z, m, cd, q1, q2): one reverse-mode sweep; gradients in units of p_red
      for (int k = 0; k < 8 * pr.nseg; ++k) gseg[k] = 0.0;
      double gp2[2] = {0.0, 0.0}, gtr = 0.0;
      for (int a = 0; a < pr.nseg; ++a)
        for (int c2 = 0; c2 < pr.nseg; ++c2) gkab[kGcMaxSeg * a + c2] = 0.0;
      gc_base_adjoint(src, b.c, cbar, gseg, gp2, &gtr, gkab);
      if (b.c.dipolar) gc_dipole_adjoint(src, b.c, cbar + 16 + b.c.nbond, gseg, &gtr);
      // association group (only for pairs that associate): dual passes, 4 directions each
      int n_dir = gc_active_directions(src, b.c, n_rec, dirs);
      for (int k = 0; k < 4; ++k) dirs[n_dir + k] = -1;
#pragma unroll 1
      for (int f = 0; f < n_dir; f += 4) {
        double g4[4];
        gc_input_grad_pass<4, false>(src, n_rec, dirs + f, cbar, g4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          int d = dirs[f + k];
          if (d < 0) break;
          if (d < 8 * pr.nseg) gseg[d] += g4[k];
          else gtr += g4[k];
        }
      }
      for (int k = 0; k < 8 * pr.nseg; ++k) gseg[k] *= scale;
      for (int a = 0; a < pr.nseg; ++a)
        for (int c2 = 0; c2 < pr.nseg; ++c2) gkab[kGcMaxSeg * a + c2] *= scale;
      gphi[0] = gp2[0] * scale;
      gphi[1] = gp2[1] * scale;
      gt = gtr * scale + gi * p_red * kPaPerRed;  // explicit factor T of p = p_red T kB/A^3
    }
    if (grad_phi) reinterpret_cast<double2*>(grad_phi)[i] = make_double2(gphi[0], gphi[1]);
    if (grad_t) grad_t[i] = gt;
  }
  // C. shared-parameter gradients -> block accumulators.  Shared-memory f64 atomics are compare-and-swap loops;
  // in the state order of order.cuh a warp holds ONE molecule pair, i.e. 32 lanes adding to the same addresses.
  // Such a warp sums each entry with shuffles and issues one atomic; mixed warps fall back to per-lane atomics.
  const unsigned vm = __ballot_sync(full, valid);
  if (vm) {
    const unsigned long long mine = (unsigned long long)my_pair;
    const unsigned long long lead = __shfl_sync(full, mine, __ffs(vm) - 1);
    if (__all_sync(full, !valid || mine == lead)) {
      const GcPair& pr = *reinterpret_cast<const GcPair*>(lead);
      const int ns = pr.nseg;
      if (grad_seg) {
        for (int k = 0; k < 8 * ns; ++k) {
          double v = valid ? w.gseg[k] : 0.0;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
          if (lane == 0 && v != 0.0) atomicAdd(&acc[8 * pr.seg[k >> 3] + (k & 7)], v);
        }
      }
      if (grad_kab) {
        for (int a = 0; a < ns; ++a)
          for (int c2 = 0; c2 < ns; ++c2) {
            const int r = pr.krec[a][c2];
            if (r < 0) continue;  // uniform over the warp: same pair
            double v = valid ? w.gkab[kGcMaxSeg * a + c2] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
            if (lane == 0 && v != 0.0) atomicAdd(&acc[8 * n_seg + r], v);
          }
      }
    } else if (valid) {
      const GcPair& pr = *my_pair;
      if (grad_seg) {
        for (int k = 0; k < 8 * pr.nseg; ++k) {
          double v = w.gseg[k];
          if (v != 0.0) atomicAdd(&acc[8 * pr.seg[k >> 3] + (k & 7)], v);
        }
      }
      if (grad_kab) {
        for (int a = 0; a < pr.nseg; ++a)
          for (int c2 = 0; c2 < pr.nseg; ++c2) {
            int r = pr.krec[a][c2];
            double v = w.gkab[kGcMaxSeg * a + c2];
            if (r >= 0 && v != 0.0) atomicAdd(&acc[8 * n_seg + r], v);
          }
      }
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n_acc; k += blockDim.x) {
    double v = acc[k];
    if (v != 0.0) {
      if (k < 8 * n_seg) {
        if (grad_seg) atomicAdd(grad_seg + k, v);
      } else if (grad_kab) {
        atomicAdd(grad_kab + (k - 8 * n_seg), v);
      }
    }
  }
}

// Gradient with respect to the k_ab of the binary segment records ONLY (a k_ab fit: segment parameters, phi and T
// carry no gradient).  k_ab enters through the cross dispersion coefficients q1[01], q2[01] alone
// (gc_pcsaft.py:177-194), so their two closed-form adjoints and one pass over the (a, b) segment pairs that have
// a record replace the reverse sweep through the whole coefficient stage -- and the per-thread scratch of the
// general kernel (dense segment / k_ab gradient arrays, dual coefficient copy) is not needed at all.
__global__ void __launch_bounds__(kGcBlock, 3)
gc_point_bwd_kab_kernel(GcArgs g, int n_rec, const double* __restrict__ rho, const uint8_t* __restrict__ failed,
                        const int32_t* __restrict__ pos, const double* __restrict__ gout, int g_broadcast, int64_t n,
                        int bubble, double* __restrict__ grad_kab, const int32_t* __restrict__ order) {
  extern __shared__ double acc[];  // [n_rec]
  for (int k = threadIdx.x; k < n_rec; k += blockDim.x) acc[k] = 0.0;
  __syncthreads();
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t i = t < n ? (order ? (int64_t)order[t] : t) : -1;
  const bool valid = i >= 0 && !failed[i];
  const GcPair* my_pair = nullptr;
  GcSource src;
  double cb1 = 0.0, cb2 = 0.0, scale = 0.0;  // adjoints of q1[01], q2[01]; g * T * kB / A^3
  if (valid) {
    src = make_source(g, i);
    my_pair = src.pair;
    const double2* rp = reinterpret_cast<const double2*>(rho + 4 * i);
    double2 rv = rp[0], rl = rp[1];
    double spec[2], inc[2];
    if (bubble) { spec[0] = rl.x; spec[1] = rl.y; inc[0] = rv.x; inc[1] = rv.y; }
    else { spec[0] = rv.x; spec[1] = rv.y; inc[0] = rl.x; inc[1] = rl.y; }
    // the adjoints of q1[01], q2[01] are the two dispersion partials (disp_q_partials, mix_grad.cuh) combined over the
    // phases; the weights come from one evaluation of the specified phase -- through the slim model where
    // the pair has no association term (the usual case), through GcBinary otherwise
    struct {
      GcBinary b;
      double slot[kGcSlimWords];
    } w;  // one stack object: see the compiler note in gc_pcsaft.cuh
    if (gc_is_slim(src, *src.pair)) {
      GcSlim b;
      gc_slim_init(b, w.slot, src);
      PointWeights q = point_weights(b.eval(spec[0], spec[1]), spec, inc);
      point_q_adjoints(GcSlimView(w.slot, src.pair), spec, inc, q, cb1, cb2);
    } else {
      gc_binary_init(w.b, src);
      PointWeights q = point_weights(w.b.eval(spec[0], spec[1]), spec, inc);
      point_q_adjoints(w.b.c, spec, inc, q, cb1, cb2);
    }
    scale = gout[g_broadcast ? 0 : (pos ? (int64_t)pos[i] : i)] * src.T_ * kPaPerRed;
  }
  const unsigned vm = __ballot_sync(full, valid);
  if (vm) {
    const unsigned long long mine = (unsigned long long)my_pair;
    const unsigned long long lead = __shfl_sync(full, mine, __ffs(vm) - 1);
    if (__all_sync(full, !valid || mine == lead)) {
      // one molecule pair per warp (state order): sum over the lanes with shuffles, one shared atomic per record
      const GcPair& pr = *reinterpret_cast<const GcPair*>(lead);
      for (int a = 0; a < pr.nseg; ++a)
        for (int b = 0; b < pr.nseg; ++b) {
          const int r = pr.krec[a][b];
          if (r < 0 || pr.cnt[0][a] * pr.cnt[1][b] == 0.0) continue;  // uniform over the warp
          double v = valid ? gc_kab_term(src, pr, a, b, cb1, cb2) * scale : 0.0;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
          if (lane == 0 && v != 0.0) atomicAdd(&acc[r], v);
        }
    } else if (valid) {
      const GcPair& pr = *my_pair;
      for (int a = 0; a < pr.nseg; ++a)
        for (int b = 0; b < pr.nseg; ++b) {
          const int r = pr.krec[a][b];
          if (r < 0 || pr.cnt[0][a] * pr.cnt[1][b] == 0.0) continue;
          double v = gc_kab_term(src, pr, a, b, cb1, cb2) * scale;
          if (v != 0.0) atomicAdd(&acc[r], v);
        }
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n_rec; k += blockDim.x)
    if (acc[k] != 0.0) atomicAdd(grad_kab + k, acc[k]);
}

__global__ void __launch_bounds__(kGcBlock)
gc_state_kernel(GcArgs g, const double* __restrict__ density, int64_t n, double* __restrict__ a, double* __restrict__ p,
                double* __restrict__ mu, double* __restrict__ v) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  GcSource src = make_source(g, i);
  GcBinary b;
  gc_binary_init(b, src);
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
// per-pair table of the slim coefficient stage, computed once per call into stream-ordered scratch
int make_pair_coef(GcArgs& g, int32_t n_pairs, Scratch& mem, cudaStream_t s) {
  if (int rc = mem.alloc(sizeof(GcPairCoef) * (size_t)n_pairs, s)) return rc;
  gc_pair_coef_kernel<<<n_pairs, 64, 0, s>>>(g.pairs, g.seg_params, g.kab, mem.as<GcPairCoef>());
  FEOS_LAUNCH_OK();
  g.pair_coef = mem.as<GcPairCoef>();
  return 0;
}

int check_tables(const feos_gc_pair* pairs, int32_t n_pairs, const double* seg_params, int32_t n_seg, int32_t n_rec,
                 const double* kab, const int32_t* pair_index) {
  FEOS_ARG(pairs && n_pairs > 0 && seg_params && n_seg > 0 && n_rec >= 0 && pair_index);
  FEOS_ARG(n_rec == 0 || kab);
  return 0;
}
}  // namespace

extern "C" {

int feos_dev_gc_point_fwd(const feos_gc_pair* pairs, int32_t n_pairs, const double* seg_params, int32_t n_seg,
                          const double* kab, int32_t n_rec, const int32_t* pair_index, const double* phi,
                          const double* temperature, const double* molefracs, const double* start_pressure, int64_t n,
                          int bubble, int flags, double* rho, double* pressure, uint8_t* failed, int32_t* evals,
                          const int32_t* order, void* stream) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31);
  if (n == 0) return 0;
  if (int rc = check_tables(pairs, n_pairs, seg_params, n_seg, n_rec, kab, pair_index)) return rc;
  FEOS_ARG(temperature && molefracs && rho && failed);
  cudaStream_t s = (cudaStream_t)stream;
  GcArgs g{reinterpret_cast<const GcPair*>(pairs), seg_params, kab, pair_index, phi, temperature};
  Scratch pair_coef;
  if (int rc_ = make_pair_coef(g, n_pairs, pair_coef, s)) return rc_;
  Scratch scratch;  // [0] = retry counter, [1] = queue head, [2] = "states of the general class present", [4 ..] = retry list
  if (int rc_ = scratch.alloc(sizeof(int32_t) * (size_t)(n + 4), s)) return rc_;
  int32_t* ws = scratch.as<int32_t>();
  FEOS_CUDA(cudaMemsetAsync(ws, 0, 4 * sizeof(int32_t), s));
  static PerDevice cache;  // shared-memory opt-in and persistent grid size, once per DEVICE
  int blocks = 0;
  if (int rc_ = cache.get([](int dev, int* v) {
        FEOS_CUDA(cudaFuncSetAttribute(gc_point_fwd_kernel<false, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGcSmem));
        FEOS_CUDA(cudaFuncSetAttribute(gc_point_fwd_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGcSlimSmem));
        FEOS_CUDA(cudaFuncSetAttribute(gc_point_fwd_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGcSlimSmem));
        FEOS_CUDA(cudaFuncSetAttribute(gc_point_queue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGcSmem));
        int sms = 0, per_sm = 0;
        FEOS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        FEOS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gc_point_queue_kernel, kGcFwdBlock, kGcSmem));
        *v = sms * (per_sm > 0 ? per_sm : 1);
        return 0;
      }, &blocks))
    return rc_;
  static const int variant = point_kernel_variant();
  if (variant == 0 || n < 65536) {
    if (bubble)
      gc_point_fwd_kernel<true, 1><<<grid_for(n, kGcSlimBlock), kGcSlimBlock, kGcSlimSmem, s>>>(
          g, molefracs, n, bubble, rho, pressure, failed, evals, ws + 4, ws, order, ws + 2);
    else
      gc_point_fwd_kernel<true, 0><<<grid_for(n, kGcSlimBlock), kGcSlimBlock, kGcSlimSmem, s>>>(
          g, molefracs, n, bubble, rho, pressure, failed, evals, ws + 4, ws, order, ws + 2);
    if (!(flags & FEOS_PLAIN_ONLY)) {
      FEOS_LAUNCH_OK();
      int general = (int)std::min<int64_t>((n + kGcFwdBlock - 1) / kGcFwdBlock, 4 * (int64_t)std::max(blocks, 1));
      gc_point_fwd_kernel<false, -1><<<general, kGcFwdBlock, kGcSmem, s>>>(g, molefracs, n, bubble, rho, pressure, failed,
                                                                      evals, ws + 4, ws, order, ws + 2);
    }
  } else {
    GcQueueSource src{g, molefracs, rho, pressure, failed};
    int grid = (int)std::min<int64_t>(blocks, (n + kGcFwdBlock - 1) / kGcFwdBlock);
    gc_point_queue_kernel<<<grid, kGcFwdBlock, kGcSmem, s>>>(src, (int)n, bubble, evals, ws + 4, ws, ws + 1);
  }
  FEOS_LAUNCH_OK();
  int retry_grid = (int)(n < 64 * 592 ? (n + 63) / 64 : 592);
  gc_point_retry_kernel<<<retry_grid, 64, 0, s>>>(g, molefracs, start_pressure, bubble, rho, pressure, failed, evals, ws + 4, ws);
  FEOS_LAUNCH_OK();
  return 0;
}

int feos_dev_gc_point_bwd(const feos_gc_pair* pairs, int32_t n_pairs, const double* seg_params, int32_t n_seg,
                          const double* kab, int32_t n_rec, const int32_t* pair_index, const double* phi,
                          const double* temperature, const double* rho, const uint8_t* failed, const int32_t* pos,
                          const double* g, int g_broadcast, int64_t n, int bubble, double* grad_seg_params,
                          double* grad_kab,
                          double* grad_phi, double* grad_temperature, const int32_t* order, void* stream) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31);
  if (n == 0) return 0;
  if (int rc = check_tables(pairs, n_pairs, seg_params, n_seg, n_rec, kab, pair_index)) return rc;
  FEOS_ARG(temperature && rho && failed && g);
  GcArgs a{reinterpret_cast<const GcPair*>(pairs), seg_params, kab, pair_index, phi, temperature};
  Scratch pair_coef;
  if (int rc_ = make_pair_coef(a, n_pairs, pair_coef, (cudaStream_t)stream)) return rc_;
  size_t shmem = sizeof(double) * (size_t)(8 * n_seg + n_rec);
  FEOS_ARG(shmem <= 48 * 1024);
  // experiment knob: pad the dynamic shared memory to cap the number of resident blocks per SM (the
  // kernel's per-thread local-memory working set is several KB: fewer resident threads thrash L1/L2 less)
  static const int pad = [] {
    const char* e = getenv("FEOS_B200_GC_BWD_SMEM_PAD");
    int v = e ? atoi(e) : 0;
    if (v > 0) cudaFuncSetAttribute(gc_point_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v + 48 * 1024);
    return v;
  }();
  shmem += (size_t)pad;
  if (!grad_seg_params && !grad_phi && !grad_temperature) {
    if (!grad_kab || n_rec == 0) return 0;
    gc_point_bwd_kab_kernel<<<grid_for(n, kGcBlock), kGcBlock, sizeof(double) * (size_t)n_rec, (cudaStream_t)stream>>>(
        a, n_rec, rho, failed, pos, g, g_broadcast, n, bubble, grad_kab, order);
    FEOS_LAUNCH_OK();
    return 0;
  }
  gc_point_bwd_kernel<<<grid_for(n, kGcBlock), kGcBlock, shmem, (cudaStream_t)stream>>>(
      a, n_seg, n_rec, rho, failed, pos, g, g_broadcast, n, bubble, grad_seg_params, grad_kab, grad_phi, grad_temperature,
      order);
  FEOS_LAUNCH_OK();
  return 0;
}

int feos_dev_gc_state(const feos_gc_pair* pairs, int32_t n_pairs, const double* seg_params, int32_t n_seg,
                      const double* kab, int32_t n_rec, const int32_t* pair_index, const double* phi,
                      const double* temperature, const double* density, int64_t n, double* a, double* p, double* mu,
                      double* v, void* stream) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31);
  if (n == 0) return 0;
  if (int rc = check_tables(pairs, n_pairs, seg_params, n_seg, n_rec, kab, pair_index)) return rc;
  FEOS_ARG(temperature && density);
  GcArgs g{reinterpret_cast<const GcPair*>(pairs), seg_params, kab, pair_index, phi, temperature};
  gc_state_kernel<<<grid_for(n, kGcBlock), kGcBlock, 0, (cudaStream_t)stream>>>(g, density, n, a, p, mu, v);
  FEOS_LAUNCH_OK();
  return 0;
}

int feos_host_gc_point(const feos_gc_pair* pairs, int32_t n_pairs, const double* seg_params, int32_t n_seg,
                       const double* kab, int32_t n_rec, const int32_t* pair_index, const double* phi,
                       const double* temperature, const double* molefracs, const double* pressure, int64_t n,
                       int bubble, double* rho_out, uint8_t* status, int64_t* n_ok) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31 && n_ok);
  *n_ok = 0;
  if (n == 0) return 0;
  if (int rc = check_tables(pairs, n_pairs, seg_params, n_seg, n_rec, kab, pair_index)) return rc;
  FEOS_ARG(temperature && molefracs && rho_out && status);
  HostCall hc;
  feos_gc_pair* dpairs;
  double *dseg, *dkab = nullptr, *dphi = nullptr, *dt, *dz, *dp0 = nullptr, *drho;
  int32_t *didx, *dorder = nullptr;
  uint8_t* dfail;
  if (int rc = hc.upload(&dpairs, pairs, (size_t)n_pairs)) return rc;
  if (int rc = hc.upload(&dseg, seg_params, 8 * (size_t)n_seg)) return rc;
  if (int rc = n_rec ? hc.upload(&dkab, kab, (size_t)n_rec) : hc.alloc(&dkab, 1)) return rc;
  if (int rc = hc.upload(&didx, pair_index, (size_t)n)) return rc;
  if (phi)
    if (int rc = hc.upload(&dphi, phi, 2 * (size_t)n)) return rc;
  if (int rc = hc.upload(&dt, temperature, (size_t)n)) return rc;
  if (int rc = hc.upload(&dz, molefracs, (size_t)n)) return rc;
  if (pressure)
    if (int rc = hc.upload(&dp0, pressure, (size_t)n)) return rc;
  if (int rc = hc.alloc(&drho, 4 * (size_t)n)) return rc;
  if (int rc = hc.alloc(&dfail, (size_t)n)) return rc;
  if (n >= kOrderMinStates) {
    if (int rc = hc.alloc(&dorder, (size_t)n)) return rc;
    if (int rc = feos_dev_gc_point_order(didx, n_pairs, dt, dz, n, 0, dorder, hc.s)) return rc;
  }
  if (int rc = feos_dev_gc_point_fwd(dpairs, n_pairs, dseg, n_seg, dkab, n_rec, didx, dphi, dt, dz, dp0, n, bubble, 0, drho,
                                     nullptr, dfail, nullptr, dorder, hc.s))
    return rc;
  return host_compact_download(hc, dfail, n, drho, 4, 4, rho_out, status, n_ok);
}

}  // extern "C"
