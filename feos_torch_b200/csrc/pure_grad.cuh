// Closed-form property formulas of the pure-component path and their parameter gradients.
//
// The reference obtains gradients by PyTorch autograd through ONE Newton step of the target property
// evaluated at the (constant) solver densities (pcsaft_pure.py:196-198, 210-215, 226-233).  The
// kernels differentiate the same formulas at the same fixed densities with forward-mode duals over
// the nine inputs (m, sigma, eps, mu, kappa, eps_ab, na, nb, T), K directions per pass to bound the
// register footprint -- which is the implicit-function-theorem gradient at the converged state.
#pragma once
#include "pcsaft_pure.cuh"

namespace feos {

// vapor pressure [Pa] and saturated liquid density [kmol/m^3] from the last evaluation point
// (pcsaft_pure.py:210-215 and :226-233)
FD void pure_vle_properties(const PureEval& el, const PureEval& ev, double rho_l, double rho_v, double T,
                            double& vp_pa, double& rho_l_eq) {
  double dv = 1.0 / rho_v - 1.0 / rho_l;
  double pm = -(ev.a / rho_v - el.a / rho_l + log(rho_v / rho_l)) / dv;
  vp_pa = pm * T * kPaPerRed;
  rho_l_eq = (rho_l - (el.p - pm) / el.dp) * kKmolM3PerRed;
}

template <int K>
FD void pure_seed(const double* prm, double T, int first, DN<K> (&x)[9]) {
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    x[i] = DN<K>(i < 8 ? prm[i] : T);
#pragma unroll
    for (int k = 0; k < K; ++k) x[i].e[k] = (i == first + k) ? 1.0 : 0.0;
  }
}

// One pass of K directions [first, first+K) of  gv * d(vp)/d(theta) + gr * d(rho_l_eq)/d(theta);
// out[k] belongs to direction first + k.
template <int K, bool NEED_RLEQ>
FD void pure_vle_grad_pass(const double* prm, double T, double rho_v, double rho_l, double dp_l, int first,
                           double gv, double gr, double (&out)[K]) {
  DN<K> x[9];
  pure_seed<K>(prm, T, first, x);
  PureCoef<DN<K>> c;
  pure_coef(c, x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7], x[8], true);
  DN<K> av = pure_helmholtz<DN<K>, DN<K>>(c, DN<K>(rho_v));
  DN<K> al, pl;
  if (NEED_RLEQ) {
    D1T<DN<K>> r(DN<K>(rho_l), DN<K>(1.0));
    D1T<DN<K>> a = pure_helmholtz<DN<K>, D1T<DN<K>>>(c, r);
    al = a.re;
    pl = rho_l - a.re + rho_l * a.d1;
  } else {
    al = pure_helmholtz<DN<K>, DN<K>>(c, DN<K>(rho_l));
  }
  double dv = 1.0 / rho_v - 1.0 / rho_l;
  DN<K> pm = (av * (1.0 / rho_v) - al * (1.0 / rho_l) + log(rho_v / rho_l)) * (-1.0 / dv);
  DN<K> vp = pm * x[8] * kPaPerRed;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double g = gv * vp.e[k];
    if (NEED_RLEQ) g += gr * (-(pl.e[k] - pm.e[k]) / dp_l) * kKmolM3PerRed;
    out[k] = g;
  }
}

// liquid density at (T, p): value = (rho - (p(rho) - p_t)/dp) * unit, p_t = P / (T kPaPerRed)
// out[k] = g * d(value)/d(direction first + k) over (params, T)
template <int K>
FD void pure_rho_tp_grad_pass(const double* prm, double T, double rho, double p_t, double dp, int first, double g,
                              double (&out)[K]) {
  DN<K> x[9];
  pure_seed<K>(prm, T, first, x);
  PureCoef<DN<K>> c;
  pure_coef(c, x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7], x[8], true);
  D1T<DN<K>> r(DN<K>(rho), DN<K>(1.0));
  D1T<DN<K>> a = pure_helmholtz<DN<K>, D1T<DN<K>>>(c, r);
  DN<K> p = rho - a.re + rho * a.d1;
  DN<K> pt = p_t * x[8].re * ad_recip(x[8]);  // P / (T kPa): only T-dependent
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = g * (-(p.e[k] - pt.e[k]) / dp) * kKmolM3PerRed;
}

}  // namespace feos
