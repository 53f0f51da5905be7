// sm_100a kernels and C-ABI entry points for the GRADIENTS of the public state functions
// (include/feos_b200.h: feos_dev_pure_state_bwd, feos_dev_mix_state_bwd, feos_dev_gc_state_bwd).
//
// Replaces torch autograd through the reference's dual-number `derivatives` (pcsaft_pure.py:180-182,
// pcsaft_mix.py:395-420, gc_pcsaft.py:443-468); see state_grad.cuh.  One thread per state, K input
// directions per pass.  Not part of the solver hot path: FP64-pipe bound
// like everything else here, but written for clarity; the forward values come from the *_state kernels.
#include "../../include/feos_b200.h"
#include "common.h"
#include "state_grad.cuh"

namespace feos {

constexpr int kStateBlock = 128;

__global__ void __launch_bounds__(kStateBlock)
pure_state_bwd_kernel(const double* __restrict__ params, const double* __restrict__ temperature,
                      const double* __restrict__ density, int64_t n, const double* __restrict__ g_a,
                      const double* __restrict__ g_p, const double* __restrict__ g_dp, double* __restrict__ grad_params,
                      double* __restrict__ grad_t, double* __restrict__ grad_rho) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double prm[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) prm[k] = params[8 * i + k];
  double ga = g_a ? g_a[i] : 0.0, gp = g_p ? g_p[i] : 0.0, gdp = g_dp ? g_dp[i] : 0.0;
  double T = temperature[i], rho = density[i];
  const int first = grad_params ? 0 : 8;
#pragma unroll 1
  for (int f = first; f < 10; f += 2) {
    double o[2];
    pure_state_grad_pass<2>(prm, T, rho, f, ga, gp, gdp, o);
    if (f < 8) {
      reinterpret_cast<double2*>(grad_params + 8 * i)[f >> 1] = make_double2(o[0], o[1]);
    } else {
      if (grad_t) grad_t[i] = o[0];
      if (grad_rho) grad_rho[i] = o[1];
    }
  }
}

__device__ __forceinline__ void load_weights(int64_t i, const double* g_a, const double* g_p, const double* g_mu,
                                             const double* g_v, double (&gw)[6]) {
  gw[0] = g_a ? g_a[i] : 0.0;
  gw[1] = g_p ? g_p[i] : 0.0;
  gw[2] = g_mu ? g_mu[2 * i] : 0.0;
  gw[3] = g_mu ? g_mu[2 * i + 1] : 0.0;
  gw[4] = g_v ? g_v[2 * i] : 0.0;
  gw[5] = g_v ? g_v[2 * i + 1] : 0.0;
}

__global__ void __launch_bounds__(kStateBlock)
mix_state_bwd_kernel(const double* __restrict__ params, const double* __restrict__ kij,
                     const double* __restrict__ temperature, const double* __restrict__ density, int64_t n,
                     const double* __restrict__ g_a, const double* __restrict__ g_p, const double* __restrict__ g_mu,
                     const double* __restrict__ g_v, double* __restrict__ grad_params, double* __restrict__ grad_kij,
                     double* __restrict__ grad_t, double* __restrict__ grad_rho) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double prm[16], kk[2], gw[6];
#pragma unroll
  for (int k = 0; k < 16; ++k) prm[k] = params[16 * i + k];
  kk[0] = kij[2 * i];
  kk[1] = kij[2 * i + 1];
  load_weights(i, g_a, g_p, g_mu, g_v, gw);
  double T = temperature[i], r0 = density[2 * i], r1 = density[2 * i + 1];
  int dirs[24];
  int nd = 0;
  if (grad_params)
    for (int d = 0; d < 16; ++d) dirs[nd++] = d;
  if (grad_kij) {
    dirs[nd++] = 16;
    dirs[nd++] = 17;
  }
  if (grad_t) dirs[nd++] = 18;
  if (grad_rho) {
    dirs[nd++] = 19;
    dirs[nd++] = 20;
  }
  for (int d = 0; d < 3; ++d) dirs[nd + d] = -1;
#pragma unroll 1
  for (int f = 0; f < nd; f += 3) {
    double o[3];
    mix_state_grad_pass<3>(prm, kk, T, r0, r1, dirs + f, gw, o);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      int d = dirs[f + k];
      if (d < 0) break;
      if (d < 16) grad_params[16 * i + d] = o[k];
      else if (d < 18) grad_kij[2 * i + (d - 16)] = o[k];
      else if (d == 18) grad_t[i] = o[k];
      else grad_rho[2 * i + (d - 19)] = o[k];
    }
  }
}

__global__ void __launch_bounds__(kStateBlock)
gc_state_bwd_kernel(const GcPair* __restrict__ pairs, const double* __restrict__ seg_params, const double* __restrict__ kab,
                    int n_rec, const int32_t* __restrict__ pair_index, const double* __restrict__ phi,
                    const double* __restrict__ temperature, const double* __restrict__ density, int64_t n,
                    const double* __restrict__ g_a, const double* __restrict__ g_p, const double* __restrict__ g_mu,
                    const double* __restrict__ g_v, double* __restrict__ grad_seg, double* __restrict__ grad_kab,
                    double* __restrict__ grad_phi, double* __restrict__ grad_t, double* __restrict__ grad_rho) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  GcSource src;
  src.pair = pairs + pair_index[i];
  src.seg_params = seg_params;
  src.kab = kab;
  src.phi_[0] = phi ? phi[2 * i] : 1.0;
  src.phi_[1] = phi ? phi[2 * i + 1] : 1.0;
  src.T_ = temperature[i];
  const GcPair& pr = *src.pair;
  double gw[6];
  load_weights(i, g_a, g_p, g_mu, g_v, gw);
  double r0 = density[2 * i], r1 = density[2 * i + 1];
  const int ns8 = 8 * pr.nseg;
  const int n_all = ns8 + n_rec + 5;
#pragma unroll 1
  for (int f = 0; f < n_all; f += 3) {
    // skip passes whose outputs are all unrequested
    int last = f + 2 < n_all ? f + 2 : n_all - 1;
    if (last < ns8 && !grad_seg) continue;
    if (f >= ns8 && last < ns8 + n_rec && !grad_kab) continue;
    int ids[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) ids[k] = (f + k < n_all) ? f + k : -1;
    double o[3];
    gc_state_grad_pass<3>(src, n_rec, r0, r1, ids, gw, o);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      int d = ids[k];
      if (d < 0) break;
      if (d < ns8) {
        if (grad_seg && o[k] != 0.0) atomicAdd(grad_seg + 8 * pr.seg[d >> 3] + (d & 7), o[k]);
      } else if (d < ns8 + n_rec) {
        if (grad_kab && o[k] != 0.0) atomicAdd(grad_kab + (d - ns8), o[k]);
      } else if (d < ns8 + n_rec + 2) {
        if (grad_phi) grad_phi[2 * i + (d - ns8 - n_rec)] = o[k];
      } else if (d == ns8 + n_rec + 2) {
        if (grad_t) grad_t[i] = o[k];
      } else if (grad_rho) {
        grad_rho[2 * i + (d - ns8 - n_rec - 3)] = o[k];
      }
    }
  }
}

}  // namespace feos

using namespace feos;

extern "C" {

int feos_dev_pure_state_bwd(const double* params, const double* temperature, const double* density, int64_t n,
                            const double* g_a, const double* g_p, const double* g_dp, double* grad_params,
                            double* grad_temperature, double* grad_density, void* stream) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31);
  if (n == 0) return 0;
  FEOS_ARG(params && temperature && density);
  FEOS_ARG(g_a || g_p || g_dp);
  pure_state_bwd_kernel<<<grid_for(n, kStateBlock), kStateBlock, 0, (cudaStream_t)stream>>>(
      params, temperature, density, n, g_a, g_p, g_dp, grad_params, grad_temperature, grad_density);
  FEOS_LAUNCH_OK();
  return 0;
}

int feos_dev_mix_state_bwd(const double* params, const double* kij, const double* temperature, const double* density,
                           int64_t n, const double* g_a, const double* g_p, const double* g_mu, const double* g_v,
                           double* grad_params, double* grad_kij, double* grad_temperature, double* grad_density,
                           void* stream) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31);
  if (n == 0) return 0;
  FEOS_ARG(params && kij && temperature && density);
  FEOS_ARG(g_a || g_p || g_mu || g_v);
  mix_state_bwd_kernel<<<grid_for(n, kStateBlock), kStateBlock, 0, (cudaStream_t)stream>>>(
      params, kij, temperature, density, n, g_a, g_p, g_mu, g_v, grad_params, grad_kij, grad_temperature, grad_density);
  FEOS_LAUNCH_OK();
  return 0;
}

int feos_dev_gc_state_bwd(const feos_gc_pair* pairs, int32_t n_pairs, const double* seg_params, int32_t n_seg,
                          const double* kab, int32_t n_rec, const int32_t* pair_index, const double* phi,
                          const double* temperature, const double* density, int64_t n, const double* g_a,
                          const double* g_p, const double* g_mu, const double* g_v, double* grad_seg_params,
                          double* grad_kab, double* grad_phi, double* grad_temperature, double* grad_density,
                          void* stream) {
  clear_error();
  FEOS_ARG(n >= 0 && n < (int64_t)1 << 31);
  if (n == 0) return 0;
  FEOS_ARG(pairs && n_pairs > 0 && seg_params && n_seg > 0 && n_rec >= 0 && pair_index && (n_rec == 0 || kab));
  FEOS_ARG(temperature && density);
  FEOS_ARG(g_a || g_p || g_mu || g_v);
  gc_state_bwd_kernel<<<grid_for(n, kStateBlock), kStateBlock, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const GcPair*>(pairs), seg_params, kab, n_rec, pair_index, phi, temperature, density, n, g_a, g_p,
      g_mu, g_v, grad_seg_params, grad_kab, grad_phi, grad_temperature, grad_density);
  FEOS_LAUNCH_OK();
  return 0;
}

}  // extern "C"
