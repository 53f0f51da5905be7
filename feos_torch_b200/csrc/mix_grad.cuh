// Parameter gradient of the Newton-step bubble / dew pressure of a binary PC-SAFT mixture.
//
// Reference: torch autograd through pcsaft_mix.py:435-444 / :459-468 at fixed solver densities.
// With  N(theta) = a_inc/rho_inc + p_spec v + sum_i w_i (ln(rho_inc,i / rho_spec,i) - mu_spec,i) - 1
// and   p = -N / (1/rho_inc - v),  w = incipient-phase mole fractions, v = sum_i w_i v_spec,i,
// the theta-derivative of v multiplies (p_spec - p), which vanishes at the converged state, so only
//   a_inc(theta)  and  p_spec(theta) v - sum_i w_i mu_spec,i(theta)
//     = v rho_spec - v a_spec(theta) + sum_i (v rho_spec,i - w_i) d a_spec / d rho_i (theta)
// are needed: one plain evaluation of the incipient phase and ONE directional density derivative of
// the specified phase (D1T), both over K parameter directions (DN<K>) per pass.  19 directions:
// 2 x 8 parameters, k_ij, eps_AiBj, T.
#pragma once
#include "mix_solver.cuh"

namespace feos {

constexpr int kMixDirs = 19;

// ids[d] = direction carried by dual slot d: 0..15 parameters, 16/17 kij, 18 T, -1 unused
template <int K>
FD void mix_seed(const double (&prm)[16], const double (&kij)[2], double T, const int* ids, DN<K> (&p)[16],
                 DN<K> (&k)[2], DN<K>& t) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    p[i] = DN<K>(prm[i]);
#pragma unroll
    for (int d = 0; d < K; ++d) p[i].e[d] = (i == ids[d]) ? 1.0 : 0.0;
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    k[i] = DN<K>(kij[i]);
#pragma unroll
    for (int d = 0; d < K; ++d) k[i].e[d] = (16 + i == ids[d]) ? 1.0 : 0.0;
  }
  t = DN<K>(T);
#pragma unroll
  for (int d = 0; d < K; ++d) t.e[d] = (18 == ids[d]) ? 1.0 : 0.0;
}

// weights of the Newton-step formula from the primal evaluation of the specified phase
struct PointWeights {
  double v, w[2], dir[2], cst, denom, rho_inc, rho_spec;
};
FD PointWeights point_weights(const MixEval& es, const double (&spec)[2], const double (&inc)[2]) {
  PointWeights q;
  q.rho_inc = inc[0] + inc[1];
  q.rho_spec = spec[0] + spec[1];
  q.w[0] = inc[0] / q.rho_inc;
  q.w[1] = inc[1] / q.rho_inc;
  q.v = q.w[0] * es.v[0] + q.w[1] * es.v[1];
  q.dir[0] = q.v * spec[0] - q.w[0];
  q.dir[1] = q.v * spec[1] - q.w[1];
  q.cst = q.w[0] * log(inc[0] / spec[0]) + q.w[1] * log(inc[1] / spec[1]) - 1.0 + q.v * q.rho_spec;
  q.denom = 1.0 / q.rho_inc - q.v;
  return q;
}

// out[k] = g * d(p_bubble/dew [Pa]) / d(direction ids[k])
template <int K>
FD_NOINLINE void mix_point_grad_pass(const double (&prm)[16], const double (&kij)[2], double T, const double (&spec)[2],
                            const double (&inc)[2], const PointWeights& q, const int* ids, double g, double (&out)[K]) {
  DN<K> p[16], k[2], t;
  mix_seed<K>(prm, kij, T, ids, p, k, t);
  MixCoef<DN<K>> c;
  mix_coef<DN<K>>(c, p, k, t, true);
  DN<K> a_inc = mix_helmholtz<DN<K>, DN<K>, 2, false>(c, DN<K>(inc[0]), DN<K>(inc[1]));
  D1T<DN<K>> r0(DN<K>(spec[0]), DN<K>(q.dir[0])), r1(DN<K>(spec[1]), DN<K>(q.dir[1]));
  D1T<DN<K>> A = mix_helmholtz<DN<K>, D1T<DN<K>>, 2, false>(c, r0, r1);
  DN<K> N = a_inc * (1.0 / q.rho_inc) - A.re * q.v + A.d1 + q.cst;
  DN<K> val = (N * (-1.0 / q.denom)) * t * kPaPerRed;
#pragma unroll
  for (int d = 0; d < K; ++d) out[d] = g * val.e[d];
}

}  // namespace feos
