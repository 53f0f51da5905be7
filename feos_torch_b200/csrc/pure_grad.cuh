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

// ---------------------------------------------------------------------------------------------
// Closed-form parameter gradient of the hard-sphere + hard-chain + dispersion part.
//
// At fixed density the base part of a depends on (m, sigma, eps, T) only through the six coefficients
// c_eta (via eta = c_eta rho), m, w1, w2, k1, k2 of PureCoef, and it is linear in w1, w2, k1, k2:
//   a  = rho f + rho^2 g,  f = m F_hs - (m-1) F_hc,  g = k1 I1 + k2 C1 I2,  I = I^0 + w1 I^1 + w2 I^2.
// So the partial sums of the series are the coefficient gradients, and the chain rule through the
// coefficient stage is a handful of multiplications -- instead of three passes of dual arithmetic
// through the whole model (the generic passes above remain for the dipole / association terms and as
// the cross-check in tests/test_hostsim_pure.py).
// Output order: d/dm, d/dsigma, d/deps, d/dT of a (da) and, if NEED_AP, of a' = da/drho (dap).
// ---------------------------------------------------------------------------------------------
struct Poly3 {
  double v, d1, d2;
};
FD Poly3 horner3(const double* k, double x) {
  double p = k[6], p1 = 0.0, p2 = 0.0;
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    p2 = p2 * x + p1;
    p1 = p1 * x + p;
    p = p * x + k[i];
  }
  Poly3 r;
  r.v = p;
  r.d1 = p1;
  r.d2 = 2.0 * p2;
  return r;
}

template <bool NEED_AP>
FD void pure_base_grad(double m, double sigma, double eps, double T, double rho, double (&da)[4], double (&dap)[4]) {
  const double ti = 1.0 / T;
  const double e = eps * ti;
  const double ex = exp(-3.0 * e);
  const double d = sigma * (1.0 - 0.12 * ex);
  const double c_eta = (kPi / 6.0) * m * (d * d * d);
  const double eta = rho * c_eta;
  const double mm1 = m - 1.0, mi = 1.0 / m;
  const double w1 = mm1 * mi, w2 = w1 * (m - 2.0) * mi;
  const double s3g = sigma * sigma * sigma;
  const double mes = m * m * e * s3g;
  const double k1 = (-2.0 * kPi) * mes, k2 = (-kPi) * mes * m * e;
  // derivatives of the coefficients
  const double dd_eps = sigma * 0.36 * ex * ti;  // d d / d eps ; d d / d T = -dd_eps eps / T
  const double ce_m = c_eta * mi, ce_s = 3.0 * c_eta / sigma, ce_e = 3.0 * c_eta * dd_eps / d, ce_t = -ce_e * e;
  const double w1_m = mi * mi, w2_m = (3.0 - 4.0 * mi) * mi * mi;

  const double s = 1.0 / (1.0 - eta), t = 1.0 / (2.0 - eta);
  const double s2 = s * s, s3 = s2 * s, s4 = s2 * s2;
  const double fhs = (4.0 - 3.0 * eta) * eta * s2, fhs1 = (4.0 - 2.0 * eta) * s3, fhs2 = (10.0 - 4.0 * eta) * s4;
  const double fhc = log((1.0 - 0.5 * eta) * s3), fhc1 = 3.0 * s - t, fhc2 = 3.0 * s2 - t * t;
  const double f1 = m * fhs1 - mm1 * fhc1, f2 = m * fhs2 - mm1 * fhc2;
  const Poly3 a0 = horner3(cA0, eta), a1 = horner3(cA1, eta), a2 = horner3(cA2, eta);
  const Poly3 b0 = horner3(cB0, eta), b1 = horner3(cB1, eta), b2 = horner3(cB2, eta);
  const double i1 = a0.v + w1 * a1.v + w2 * a2.v, i1d = a0.d1 + w1 * a1.d1 + w2 * a2.d1, i1dd = a0.d2 + w1 * a1.d2 + w2 * a2.d2;
  const double i2 = b0.v + w1 * b1.v + w2 * b2.v, i2d = b0.d1 + w1 * b1.d1 + w2 * b2.d1, i2dd = b0.d2 + w1 * b1.d2 + w2 * b2.d2;
  const double x0 = (8.0 - 2.0 * eta) * eta * s4;
  const double x1 = (8.0 + (20.0 - 4.0 * eta) * eta) * s4 * s;
  const double x2 = (60.0 + (72.0 - 12.0 * eta) * eta) * s4 * s2;
  const double n0 = (((-2.0 * eta + 12.0) * eta - 27.0) * eta + 20.0) * eta;
  const double n1 = ((-8.0 * eta + 36.0) * eta - 54.0) * eta + 20.0;
  const double n2 = (-24.0 * eta + 72.0) * eta - 54.0;
  const double w = s2 * t * t, st = s + t;
  const double y0 = n0 * w;
  const double y1 = w * (n1 + 2.0 * n0 * st);
  const double y2 = w * (n2 + 4.0 * n1 * st + n0 * (4.0 * st * st + 2.0 * (s2 + t * t)));
  const double d0 = 1.0 + m * x0 - mm1 * y0, d1 = m * x1 - mm1 * y1, d2 = m * x2 - mm1 * y2;
  const double c1 = 1.0 / d0;
  const double c1d = -d1 * c1 * c1;
  const double c1dd = (2.0 * d1 * d1 * c1 - d2) * c1 * c1;
  const double z0 = x0 - y0, z1 = x1 - y1;  // d D / d m and its eta-derivative
  const double c1m = -c1 * c1 * z0;         // d C1 / d m
  const double c1md = -2.0 * c1 * c1d * z0 - c1 * c1 * z1;
  const double g1 = k1 * i1d + k2 * (c1d * i2 + c1 * i2d);
  const double rho2 = rho * rho;
  // gradient of a with respect to (eta, m explicit, w1, w2, k1, k2)
  const double A_eta = rho * f1 + rho2 * g1;
  const double A_m = rho * (fhs - fhc) + rho2 * k2 * i2 * c1m;
  const double A_w1 = rho2 * (k1 * a1.v + k2 * c1 * b1.v), A_w2 = rho2 * (k1 * a2.v + k2 * c1 * b2.v);
  const double A_k1 = rho2 * i1, A_k2 = rho2 * c1 * i2;
  const double kk_m = (2.0 * A_k1 * k1 + 3.0 * A_k2 * k2) * mi;
  const double kk_s = 3.0 * (A_k1 * k1 + A_k2 * k2) / sigma;
  const double kk_e = (A_k1 * k1 + 2.0 * A_k2 * k2) / eps;
  da[0] = A_eta * rho * ce_m + A_m + A_w1 * w1_m + A_w2 * w2_m + kk_m;
  da[1] = A_eta * rho * ce_s + kk_s;
  da[2] = A_eta * rho * ce_e + kk_e;
  da[3] = A_eta * rho * ce_t - kk_e * e;
  if (NEED_AP) {
    const double g2 = k1 * i1dd + k2 * (c1dd * i2 + 2.0 * c1d * i2d + c1 * i2dd);
    const double B_eta = 2.0 * f1 + eta * f2 + rho * (3.0 * g1 + eta * g2);
    const double gm = k2 * i2 * c1m, gmd = k2 * (c1md * i2 + c1m * i2d);
    const double B_m = (fhs - fhc) + eta * (fhs1 - fhc1) + rho * (2.0 * gm + eta * gmd);
    const double gw1 = k1 * a1.v + k2 * c1 * b1.v, gw1d = k1 * a1.d1 + k2 * (c1d * b1.v + c1 * b1.d1);
    const double gw2 = k1 * a2.v + k2 * c1 * b2.v, gw2d = k1 * a2.d1 + k2 * (c1d * b2.v + c1 * b2.d1);
    const double B_w1 = rho * (2.0 * gw1 + eta * gw1d), B_w2 = rho * (2.0 * gw2 + eta * gw2d);
    const double B_k1 = rho * (2.0 * i1 + eta * i1d), B_k2 = rho * (2.0 * c1 * i2 + eta * (c1d * i2 + c1 * i2d));
    const double bk_m = (2.0 * B_k1 * k1 + 3.0 * B_k2 * k2) * mi;
    const double bk_s = 3.0 * (B_k1 * k1 + B_k2 * k2) / sigma;
    const double bk_e = (B_k1 * k1 + 2.0 * B_k2 * k2) / eps;
    dap[0] = B_eta * rho * ce_m + B_m + B_w1 * w1_m + B_w2 * w2_m + bk_m;
    dap[1] = B_eta * rho * ce_s + bk_s;
    dap[2] = B_eta * rho * ce_e + bk_e;
    dap[3] = B_eta * rho * ce_t - bk_e * e;
  }
}

// K directions [first, first + K) of d(extras)/d(theta) and, if NEED_AP, d(extras')/d(theta) for the
// dipole + association terms (generic dual arithmetic; only called for states that have them).
template <int K, bool NEED_AP>
FD void pure_extras_grad_pass(const double* prm, double T, double rho, int first, double (&da)[K], double (&dap)[K]) {
  DN<K> x[9];
  pure_seed<K>(prm, T, first, x);
  PureCoef<DN<K>> c;
  pure_coef(c, x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7], x[8], true);
  if (NEED_AP) {
    typedef D1T<DN<K>> R;
    R r(DN<K>(rho), DN<K>(1.0));
    R eta = r * c.c_eta;
    R e2 = eta * eta;
    R a = pure_helmholtz_extras<DN<K>, R>(c, r, eta, e2, e2 * eta, e2 * e2, ad_recip(1.0 - eta));
#pragma unroll
    for (int k = 0; k < K; ++k) {
      da[k] = a.re.e[k];
      dap[k] = a.d1.e[k];
    }
  } else {
    typedef DN<K> R;
    R r(rho);
    R eta = r * c.c_eta;
    R e2 = eta * eta;
    R a = pure_helmholtz_extras<DN<K>, R>(c, r, eta, e2, e2 * eta, e2 * e2, ad_recip(1.0 - eta));
#pragma unroll
    for (int k = 0; k < K; ++k) da[k] = a.e[k];
  }
}

// Association term of a non-polar state: it depends on the nine inputs only through FOUR coefficients
// (c_eta, delta0, na, nb), so one pass of four dual directions over the term + the hand-written derivatives of
// the coefficient stage replaces three passes of three directions over coefficient stage and term
// (C3: pure_vle_bwd_kernel 4.6 -> see profiles/README.md).  Adds to da / dap.
template <bool NEED_AP>
FD void pure_assoc_grad_coef(const double* prm, double T, double rho, double (&da)[9], double (&dap)[9]) {
  const double m = prm[0], sigma = prm[1], eps = prm[2], kappa = prm[4], eps_ab = prm[5];
  const double ti = 1.0 / T;
  const double ex = exp(-3.0 * eps * ti);
  const double f = 1.0 - 0.12 * ex;
  const double d = sigma * f, d2 = d * d;
  const double s2 = sigma * sigma, s3 = s2 * sigma;
  const double E = exp(eps_ab * ti);
  PureCoef<DN<4>> c;
  c.dipolar = false;
  c.assoc = true;
  c.c_eta = DN<4>((kPi / 6.0) * m * (d2 * d));
  c.c_eta.e[0] = 1.0;
  c.delta0 = DN<4>((E - 1.0) * s3 * kappa);
  c.delta0.e[1] = 1.0;
  c.na = DN<4>(prm[6]);
  c.na.e[2] = 1.0;
  c.nb = DN<4>(prm[7]);
  c.nb.e[3] = 1.0;
  double A[4], Ap[4];
  if (NEED_AP) {
    typedef D1T<DN<4>> R;
    R r(DN<4>(rho), DN<4>(1.0));
    R eta = r * c.c_eta;
    R e2 = eta * eta;
    R a = pure_helmholtz_extras<DN<4>, R>(c, r, eta, e2, e2 * eta, e2 * e2, ad_recip(1.0 - eta));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      A[k] = a.re.e[k];
      Ap[k] = a.d1.e[k];
    }
  } else {
    typedef DN<4> R;
    R r(rho);
    R eta = r * c.c_eta;
    R e2 = eta * eta;
    R a = pure_helmholtz_extras<DN<4>, R>(c, r, eta, e2, e2 * eta, e2 * e2, ad_recip(1.0 - eta));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      A[k] = a.e[k];
      Ap[k] = 0.0;
    }
  }
  // coefficient stage (pure_coef): c_eta = pi/6 m d^3, d = sigma (1 - 0.12 exp(-3 eps / T));
  // delta0 = (exp(eps_ab / T) - 1) sigma^3 kappa
  const double ce_d = (kPi / 6.0) * m * 3.0 * d2;      // d c_eta / d d
  const double d_eps = sigma * 0.36 * ex * ti;         // d d / d eps
  const double d_T = -sigma * 0.36 * ex * eps * ti * ti;
  const double ce[4] = {(kPi / 6.0) * (d2 * d), ce_d * f, ce_d * d_eps, ce_d * d_T};  // m, sigma, eps, T
  const double de[4] = {(E - 1.0) * 3.0 * s2 * kappa, (E - 1.0) * s3, E * ti * s3 * kappa,
                        -E * eps_ab * ti * ti * s3 * kappa};                          // sigma, kappa, eps_ab, T
  da[0] += A[0] * ce[0];
  da[1] += A[0] * ce[1] + A[1] * de[0];
  da[2] += A[0] * ce[2];
  da[4] += A[1] * de[1];
  da[5] += A[1] * de[2];
  da[6] += A[2];
  da[7] += A[3];
  da[8] += A[0] * ce[3] + A[1] * de[3];
  if (NEED_AP) {
    dap[0] += Ap[0] * ce[0];
    dap[1] += Ap[0] * ce[1] + Ap[1] * de[0];
    dap[2] += Ap[0] * ce[2];
    dap[4] += Ap[1] * de[1];
    dap[5] += Ap[1] * de[2];
    dap[6] += Ap[2];
    dap[7] += Ap[3];
    dap[8] += Ap[0] * ce[3] + Ap[1] * de[3];
  }
}

// d a / d(m, sigma, eps, mu, kappa, eps_ab, na, nb, T) at fixed rho (da[9]) and the same for a' (dap[9])
template <bool NEED_AP>
FD void pure_state_grad(const double* prm, double T, double rho, double (&da)[9], double (&dap)[9]) {
  double b[4], bp[4];
  pure_base_grad<NEED_AP>(prm[0], prm[1], prm[2], T, rho, b, bp);
#pragma unroll
  for (int k = 0; k < 9; ++k) da[k] = dap[k] = 0.0;
  da[0] = b[0]; da[1] = b[1]; da[2] = b[2]; da[8] = b[3];
  if (NEED_AP) { dap[0] = bp[0]; dap[1] = bp[1]; dap[2] = bp[2]; dap[8] = bp[3]; }
  bool dipolar = prm[3] != 0.0;
  bool assoc = (prm[6] != 0.0) || (prm[7] != 0.0);  // as pure_coef(..., grad_mode = true)
  if (assoc && !dipolar) {
    pure_assoc_grad_coef<NEED_AP>(prm, T, rho, da, dap);
  } else if (dipolar) {
#pragma unroll
    for (int f = 0; f < 9; f += 3) {
      double x[3], xp[3];
      pure_extras_grad_pass<3, NEED_AP>(prm, T, rho, f, x, xp);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        da[f + k] += x[k];
        if (NEED_AP) dap[f + k] += xp[k];
      }
    }
  }
}

// gradients of the pure-component properties from the state gradients
//   out[0..8] = gv d(vp)/d(theta) + gr d(rho_l_eq)/d(theta)
template <bool NEED_RLEQ>
FD void pure_vle_grad(const double* prm, double T, double rho_v, double rho_l, double dp_l, double gv, double gr,
                      bool need_t, double (&out)[9]) {
  double av[9], avp[9], al[9], alp[9];
  pure_state_grad<false>(prm, T, rho_v, av, avp);
  pure_state_grad<NEED_RLEQ>(prm, T, rho_l, al, alp);
  const double dv = 1.0 / rho_v - 1.0 / rho_l;
  // the value of p_m is needed for the explicit factor T of vp = p_m T kB/A^3
  double pm = 0.0;
  if (need_t) {
    PureCoef<double> c;
    pure_coef<double>(c, prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], prm[6], prm[7], T, false);
    PureEval ev = pure_eval(c, rho_v), el = pure_eval(c, rho_l);
    pm = -(ev.a / rho_v - el.a / rho_l + log(rho_v / rho_l)) / dv;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    double pm_k = -(av[k] / rho_v - al[k] / rho_l) / dv;
    double g = gv * (pm_k * T + (k == 8 ? pm : 0.0)) * kPaPerRed;
    if (NEED_RLEQ) {
      double pl_k = -al[k] + rho_l * alp[k];
      g += gr * (-(pl_k - pm_k) / dp_l) * kKmolM3PerRed;
    }
    out[k] = g;
  }
}

//   out[0..8] = g d(rho(T,p))/d(theta)  (liquid density at given pressure)
FD void pure_rho_tp_grad(const double* prm, double T, double rho, double p_t, double dp, double g, double (&out)[9]) {
  double a[9], ap[9];
  pure_state_grad<true>(prm, T, rho, a, ap);
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    double p_k = -a[k] + rho * ap[k];
    double pt_k = (k == 8) ? -p_t / T : 0.0;
    out[k] = g * (-(p_k - pt_k) / dp) * kKmolM3PerRed;
  }
}

}  // namespace feos
