// Gradients of the public state functions  helmholtz_energy / derivatives  at given (T, rho).
//
// The reference obtains them by torch autograd through its dual-number evaluation
// (/root/reference/feos_torch/pcsaft_pure.py:106-182, pcsaft_mix.py:31-154 and :395-420,
// gc_pcsaft.py:116-253 and :443-468).  Here the same outputs are formed from one evaluation in nested
// dual numbers: the density derivatives of `derivatives` (D2T / H2) over a field DN<K> that carries K
// input directions -- parameters, k_ij / k_ab, phi, T and the densities themselves -- so that one pass
// yields d(sum_j g_j out_j)/d(direction) for K directions.  Not on the solver hot path: clarity over speed.
#pragma once
#include "gc_pcsaft.cuh"
#include "pure_grad.cuh"

namespace feos {

// ---- pure component: out = (a, p, dp) of pcsaft_pure.py:180-182; directions 0..7 parameters, 8 T, 9 rho ----
template <int K>
FD_NOINLINE void pure_state_grad_pass(const double* prm, double T, double rho, int first, double ga, double gp, double gdp,
                                      double (&out)[K]) {
  DN<K> x[9];
  pure_seed<K>(prm, T, first, x);
  PureCoef<DN<K>> c;
  pure_coef(c, x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7], x[8], true);
  DN<K> r0(rho);
#pragma unroll
  for (int k = 0; k < K; ++k) r0.e[k] = (first + k == 9) ? 1.0 : 0.0;
  D2T<DN<K>> r(r0, DN<K>(1.0), DN<K>(0.0));
  D2T<DN<K>> a = pure_helmholtz<DN<K>, D2T<DN<K>>>(c, r);
  DN<K> p = r0 - a.re + r0 * a.d1;  // rho - a + rho a'
  DN<K> dp = r0 * a.d2 + 1.0;       // 1 + rho a''
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = ga * a.re.e[k] + gp * p.e[k] + gdp * dp.e[k];
}

// ---- mixtures: out = (a, p, mu_0, mu_1, v_0, v_1) of pcsaft_mix.py:413-418 with weights gw[6] -----------------
template <class P>
FD P state_loss(const H2<P>& a, const P& rho0, const P& rho1, const double* gw) {
  P p = rho0 + rho1 - a.re + rho0 * a.g0 + rho1 * a.g1;
  P dp0 = rho0 * a.h00 + rho1 * a.h01 + 1.0;
  P dp1 = rho0 * a.h01 + rho1 * a.h11 + 1.0;
  P iden = ad_recip(rho0 * dp0 + rho1 * dp1);
  return a.re * gw[0] + p * gw[1] + a.g0 * gw[2] + a.g1 * gw[3] + (dp0 * iden) * gw[4] + (dp1 * iden) * gw[5];
}

template <int K>
FD void seed_densities(double r0, double r1, const int* ids, int id0, H2<DN<K>>& rho0, H2<DN<K>>& rho1) {
  rho0 = H2<DN<K>>(0.0);
  rho1 = H2<DN<K>>(0.0);
  rho0.re = DN<K>(r0);
  rho1.re = DN<K>(r1);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    rho0.re.e[k] = (ids[k] == id0) ? 1.0 : 0.0;
    rho1.re.e[k] = (ids[k] == id0 + 1) ? 1.0 : 0.0;
  }
  rho0.g0 = DN<K>(1.0);
  rho1.g1 = DN<K>(1.0);
}

// binary PC-SAFT: directions 0..15 parameters, 16/17 kij, 18 T, 19/20 rho
template <int K>
FD_NOINLINE void mix_state_grad_pass(const double (&prm)[16], const double (&kij)[2], double T, double r0, double r1,
                                     const int* ids, const double* gw, double (&out)[K]) {
  DN<K> p[16], k[2], t;
  mix_seed<K>(prm, kij, T, ids, p, k, t);
  MixCoef<DN<K>> c;
  mix_coef<DN<K>>(c, p, k, t, true);
  H2<DN<K>> rho0, rho1;
  seed_densities<K>(r0, r1, ids, 19, rho0, rho1);
  H2<DN<K>> a = mix_helmholtz<DN<K>, H2<DN<K>>, 2, false>(c, rho0, rho1);
  DN<K> L = state_loss<DN<K>>(a, rho0.re, rho1.re, gw);
#pragma unroll
  for (int d = 0; d < K; ++d) out[d] = L.e[d];
}

// group contribution: directions as GcDualSource (8 a + j | 8 nseg + r | + nrec + c (phi) | + nrec + 2 (T)),
// then 8 nseg + nrec + 3 / + 4 for the two densities
template <int K>
FD_NOINLINE void gc_state_grad_pass(const GcSource& src, int nrec, double r0, double r1, const int* ids, const double* gw,
                                    double (&out)[K]) {
  GcDualSource<K> ds;
  ds.s = src;
#pragma unroll
  for (int k = 0; k < K; ++k) ds.ids[k] = ids[k];
  ds.nrec = nrec;
  MixCoef<DN<K>, kGcMaxBond, true> c;
  gc_coef<DN<K>, GcDualSource<K>, true>(c, ds, *src.pair);
  H2<DN<K>> rho0, rho1;
  seed_densities<K>(r0, r1, ids, 8 * src.pair->nseg + nrec + 3, rho0, rho1);
  H2<DN<K>> a = mix_helmholtz<DN<K>, H2<DN<K>>, kGcMaxBond, true>(c, rho0, rho1);
  DN<K> L = state_loss<DN<K>>(a, rho0.re, rho1.re, gw);
#pragma unroll
  for (int d = 0; d < K; ++d) out[d] = L.e[d];
}

}  // namespace feos
