// Binary-mixture PC-SAFT reduced residual Helmholtz energy density a(T, rho_1, rho_2) [1/A^3].
//
// Device-side, in-register replacement for PcSaftMix.helmholtz_energy_density and its helpers
// (/root/reference/feos_torch/pcsaft_mix.py:31-154 HS/HC/dispersion, :156-208 dipole, :210-239 self-,
// :241-321 cross-, :323-393 induced association, :482-522 integrals and association strength) for the
// binary case the native boundary supports (src/pcsaft.rs:150-214).
//   mix_coef<P>()        everything that depends on (parameters, k_ij, T) only;
//   mix_helmholtz<P,R>() the density dependent part; R is the number type of the partial densities:
//                        H2<double> (value, gradient, Hessian -> a, mu_i, p, v_i; replaces the
//                        hyper-dual DualTensor evaluation of :395-420), DN<K> (parameter directions at
//                        fixed densities) or D1T<DN<K>> (one density direction x parameter directions).
// The site-fraction systems of cross/induced association are solved in plain doubles (Newton from
// 0.2 with the reference's negative-iterate rule) and then differentiated by `NSTEP` Newton steps
// carried out in R arithmetic (each step from a converged root doubles the exact derivative order),
// which is implicit differentiation of the converged site fractions.
#pragma once
#include "pcsaft_pure.cuh"

namespace feos {

enum : int { kAssocNone = 0, kAssocSelf = 1, kAssocCross = 2, kAssocInduced = 3 };

// NB = capacity of the bond-type list of the hard-chain term, GC = group-contribution association
// formulas.  A homosegmented PC-SAFT binary is the special case NB = 2 with one "bond type" per
// component: contact distance d_i / 2 and m_i - 1 bonds (pcsaft_mix.py:63-65 == gc_pcsaft.py:156-165).
template <class P, int NB = 2, bool GC = false>
struct MixCoef {
  // hard sphere: zeta_k = sum_i z[k][i] rho_i
  P z[4][2];
  P m[2];  // segments per molecule
  // hard chain: a_hc = - sum_b (rho_0 bn[b][0] + rho_1 bn[b][1]) ln g(cd[b])
  int nbond;
  P cd[NB];     // d_a d_b / (d_a + d_b)
  P bn[NB][2];  // number of bonds of this type per molecule
  // dispersion: rho1mix = sum q1[ij] rho_i rho_j (ij = 00, 01 (both orders), 11), rho2mix with q2
  P q1[3], q2[3];
  // dipole: phi2 = sum pair terms, phi3 = sum triplet terms
  bool dipolar;
  P mu2t[2];                   // sigma^3 eps mu2 / T per component (:163)
  P pw1[3], pw2[3], pe[3], ps[3];  // pairs 00, 01, 11: m_ij weights, eps_ij/T, multiplicity / sigma_ij^3
  P tw1[4], tw2[4], ts[4];         // triplets 000, 001, 011, 111: weights, multiplicity / (s_ij s_ik s_jk)
  // association
  int cat;
  P dlt0[4];   // (sigma_i sigma_j)^1.5 sqrt(kappa_i kappa_j) (exp(eps_AiBj / T) - 1), ij = 00, 01, 10, 11
  P dd[4];     // d_i d_j / (d_i + d_j)
  P na[2], nb[2];
  double sgn[2];  // GC: 1 if the component carries the associating segment (gc_pcsaft.py:317-320)
};

template <class P>
FD P ad_cbrt_pos(const P& x);
template <>
FD double ad_cbrt_pos<double>(const double& x) { return cbrt(x); }
template <int N>
FD DN<N> ad_cbrt_pos(const DN<N>& x) {
  double c = cbrt(x.re);
  return x.chain(c, c / (3.0 * x.re));
}

// Gross-Vrabec dipole coefficients from molecule-level m, sigma, eps and mu2_term = (reduced mu^2)/T
// (pcsaft_mix.py:156-208, gc_pcsaft.py:255-307)
template <class C, class P>
FD void dipole_coef(C& c, const P (&m)[2], const P (&sig)[2], const P (&eps)[2], const P (&mu2t)[2], const P& ti,
                    bool dipolar) {
  c.dipolar = dipolar;
  if (!dipolar) return;
  P mc[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    c.mu2t[i] = mu2t[i];
    mc[i] = ad_clamp_max(m[i], 2.0);
  }
  const int pi_[3] = {0, 0, 1}, pj_[3] = {0, 1, 1};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int i = pi_[k], j = pj_[k];
    P sij = 0.5 * (sig[i] + sig[j]);
    P mij = ad_sqrt(mc[i] * mc[j]);
    P mi = ad_recip(mij);
    c.pw1[k] = (mij - 1.0) * mi;
    c.pw2[k] = c.pw1[k] * (mij - 2.0) * mi;
    c.pe[k] = ad_sqrt(eps[i] * eps[j]) * ti;
    c.ps[k] = ad_recip(sij * sij * sij) * (i == j ? 1.0 : 2.0);
  }
  const int ti_[4] = {0, 0, 0, 1}, tj_[4] = {0, 0, 1, 1}, tk_[4] = {0, 1, 1, 1};
  const double mult[4] = {1.0, 3.0, 3.0, 1.0};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    int i = ti_[t], j = tj_[t], k = tk_[t];
    P sij = 0.5 * (sig[i] + sig[j]), sik = 0.5 * (sig[i] + sig[k]), sjk = 0.5 * (sig[j] + sig[k]);
    P mijk = ad_cbrt_pos(mc[i] * mc[j] * mc[k]);
    P mi = ad_recip(mijk);
    c.tw1[t] = (mijk - 1.0) * mi;
    c.tw2[t] = c.tw1[t] * (mijk - 2.0) * mi;
    c.ts[t] = ad_recip(sij * sik * sjk) * mult[t];
  }
}

// Association category of a parameter pair (pcsaft_mix.py:118-152), from the site counts alone
FD int mix_category(double na0, double nb0, double na1, double nb1) {
  int n_assoc = ((na0 + nb0) != 0.0) + ((na1 + nb1) != 0.0);
  int n_self = ((na0 * nb0) != 0.0) + ((na1 * nb1) != 0.0);
  if (n_assoc == 1 && n_self == 1) return kAssocSelf;
  if (n_assoc == 2 && n_self == 2) return kAssocCross;
  if (n_assoc == 2 && n_self == 1) return kAssocInduced;
  return kAssocNone;
}

// "plain" pair: hard sphere + hard chain + dispersion only (no dipole term, no association term)
FD bool mix_is_plain(const double (&prm)[16]) {
  return prm[3] == 0.0 && prm[11] == 0.0 && mix_category(prm[6], prm[7], prm[14], prm[15]) == kAssocNone;
}

// The coefficients of the base part alone (24 doubles): what the kernels specialised for plain pairs keep per state.
// Same member names as MixCoef, so that mix_base_eval and the solvers take either.
struct MixCoefBase {
  double z[4][2];
  double m[2];
  double cd[2];
  double bn[2][2];
  double q1[3], q2[3];
  static constexpr int nbond = 2;
  static constexpr int cat = kAssocNone;
  static constexpr bool dipolar = false;
};

// hard-sphere, hard-chain and dispersion coefficients (pcsaft_mix.py:31-33, :56-90); returns sigma, eps, d per component
template <class C, class P>
FD void mix_coef_base(C& c, const P (&prm)[16], const P (&kij)[2], const P& ti, P (&s3)[2], P (&eps)[2], P (&sig)[2],
                      P (&dseg)[2]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    P m = prm[8 * i], s = prm[8 * i + 1], e = prm[8 * i + 2];
    sig[i] = s;
    eps[i] = e;
    s3[i] = s * s * s;
    P d = s * (1.0 - 0.12 * ad_exp(-3.0 * (e * ti)));  // pcsaft_mix.py:33
    c.m[i] = m;
    dseg[i] = d;
    c.cd[i] = 0.5 * d;
    c.bn[i][i] = m - 1.0;
    c.bn[i][1 - i] = P(0.0);
    P md = (kPi / 6.0) * m;
    c.z[0][i] = md;
    md = md * d;
    c.z[1][i] = md;
    md = md * d;
    c.z[2][i] = md;
    c.z[3][i] = md * d;
  }
  // dispersion (:69-90): eps_ij = sqrt(eps_i eps_j)/T (1 - k_ij [i != j]), sigma_ij^3, m_i m_j
  {
    P e00 = eps[0] * ti, e11 = eps[1] * ti;
    P e01 = ad_sqrt(eps[0] * eps[1]) * ti * (1.0 - kij[0]);
    P s01 = 0.5 * (sig[0] + sig[1]);
    c.q1[0] = c.m[0] * c.m[0] * e00 * s3[0];
    c.q1[1] = c.m[0] * c.m[1] * e01 * (s01 * s01 * s01);
    c.q1[2] = c.m[1] * c.m[1] * e11 * s3[1];
    c.q2[0] = c.q1[0] * e00;
    c.q2[1] = c.q1[1] * e01;
    c.q2[2] = c.q1[2] * e11;
  }
}

// prm: 16 values = 2 rows of [m, sigma, eps, mu, kappa_ab, eps_ab, na, nb]; kij = (k_ij, eps_AiBj)
template <class P>
FD void mix_coef(MixCoef<P>& c, const P (&prm)[16], const P (&kij)[2], const P& T, bool grad_mode) {
  P ti = ad_recip(T);
  P s3[2], eps[2], sig[2], dseg[2];
  c.nbond = 2;
  mix_coef_base(c, prm, kij, ti, s3, eps, sig, dseg);
  // dipole (:111-115, :156-208)
  {
    P mu2t[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      P mu = prm[8 * i + 3];
      // mu2 = mu^2 / (m sigma^3 eps) * 1e-19 J/K/kB (:17-22); mu2_term = sigma^3 eps mu2 / T (:163)
      P mu2 = mu * mu / (c.m[i] * s3[i] * eps[i]) * kDipoleScale;
      mu2t[i] = s3[i] * eps[i] * mu2 * ti;
    }
    dipole_coef(c, c.m, sig, eps, mu2t, ti, ad_primal(prm[3]) != 0.0 || ad_primal(prm[11]) != 0.0);
  }
  // association categories (:118-152)
  {
    c.cat = mix_category(ad_primal(prm[6]), ad_primal(prm[7]), ad_primal(prm[14]), ad_primal(prm[15]));
    c.na[0] = prm[6];
    c.nb[0] = prm[7];
    c.na[1] = prm[14];
    c.nb[1] = prm[15];
    c.sgn[0] = c.sgn[1] = 0.0;
    (void)grad_mode;
    if (c.cat == kAssocSelf) {
      // :210-239 -- kappa, eps_ab summed over components, sigma and d weighted with na
      P kap = prm[4] + prm[12];
      P eab = prm[5] + prm[13];
      P nas = c.na[0] + c.na[1];
      P nai = ad_recip(nas);
      P sg = (c.na[0] * sig[0] + c.na[1] * sig[1]) * nai;
      P dg = (c.na[0] * dseg[0] + c.na[1] * dseg[1]) * nai;
      // (sigma sigma)^1.5 sqrt(kappa kappa) = sigma^3 kappa
      c.dlt0[0] = ad_sqrt(sg * sg) * (sg * sg) * ad_sqrt(kap * kap) * (ad_exp(eab * ti) - 1.0);
      c.dd[0] = dg * dg / (dg + dg);
    } else if (c.cat != kAssocNone) {
      const int di_[4] = {0, 0, 1, 1}, dj_[4] = {0, 1, 0, 1};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int i = di_[k], j = dj_[k];
        P e = 0.5 * (prm[8 * i + 5] + prm[8 * j + 5]);
        // eps_AiBj from kij[:, 1] for i != j when non-zero, cross association only (:148, :509-516)
        if (c.cat == kAssocCross && i != j && ad_primal(kij[1]) != 0.0) e = kij[1];
        P ss = sig[i] * sig[j];
        c.dlt0[k] = ad_sqrt(ss) * ss * ad_sqrt(prm[8 * i + 4] * prm[8 * j + 4]) * (ad_exp(e * ti) - 1.0);
        c.dd[k] = dseg[i] * dseg[j] / (dseg[i] + dseg[j]);
      }
    }
  }
}

// Delta^{AiBj} (:500-522): zeta3_m1 (k (2k + 3) + 1) dlt0, k = dd zeta2 zeta3_m1
template <class C, class R>
FD R mix_strength(const C& c, int k, const R& z2om, const R& omi) {
  R kk = z2om * c.dd[k];
  return (omi * (kk * (2.0 * kk + 3.0) + 1.0)) * c.dlt0[k];
}

// ---- cross association (:241-321): unknowns X_A0, X_A1 ------------------------------------------
template <class R>
FD void cross_residual(const R& x0, const R& x1, const R (&e)[4], const R (&cc)[4], R& f0, R& f1, R& j00, R& j01,
                       R& j10, R& j11) {
  // b_i = 1 / X_Bi;  e = rho_A Delta (columns), cc = rho_B Delta
  R ib0 = ad_recip(1.0 + x0 * e[0] + x1 * e[1]);
  R ib1 = ad_recip(1.0 + x0 * e[2] + x1 * e[3]);
  R s0 = cc[0] * ib0 + cc[1] * ib1;
  R s1 = cc[2] * ib0 + cc[3] * ib1;
  f0 = x0 - 1.0 + x0 * s0;
  f1 = x1 - 1.0 + x1 * s1;
  R q0 = ib0 * ib0, q1 = ib1 * ib1;
  R a00 = cc[0] * q0, a01 = cc[1] * q1, a10 = cc[2] * q0, a11 = cc[3] * q1;
  j00 = 1.0 + s0 - x0 * (a00 * e[0] + a01 * e[2]);
  j01 = -(x0 * (a00 * e[1] + a01 * e[3]));
  j10 = -(x1 * (a10 * e[0] + a11 * e[2]));
  j11 = 1.0 + s1 - x1 * (a10 * e[1] + a11 * e[3]);
}

template <class R, int NSTEP>
FD R cross_assoc(const R (&ra)[2], const R (&rb)[2], const R (&dl)[4]) {
  // e[2 i + j] = d b_i / d x_j = rho_Aj Delta_ij ; cc[2 i + j] = rho_Bj Delta_ij
  R e[4] = {ra[0] * dl[0], ra[1] * dl[1], ra[0] * dl[2], ra[1] * dl[3]};
  R cc[4] = {rb[0] * dl[0], rb[1] * dl[1], rb[0] * dl[2], rb[1] * dl[3]};
  double ep[4], cp[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    ep[k] = ad_primal(e[k]);
    cp[k] = ad_primal(cc[k]);
  }
  double x0 = 0.2, x1 = 0.2;
  for (int it = 0; it < 50; ++it) {
    double f0, f1, j00, j01, j10, j11;
    cross_residual<double>(x0, x1, ep, cp, f0, f1, j00, j01, j10, j11);
    double det = j00 * j11 - j01 * j10;
    double n0 = x0 - (j11 * f0 - j01 * f1) / det;
    double n1 = x1 - (-j10 * f0 + j00 * f1) / det;
    if (n0 < 0.0) n0 = 0.2 * x0;  // :304-308
    if (n1 < 0.0) n1 = 0.2 * x1;
    bool done = fabs(n0 - x0) <= 1e-15 && fabs(n1 - x1) <= 1e-15;
    x0 = n0;
    x1 = n1;
    if (done) break;
  }
  R X0(x0), X1(x1);
#pragma unroll
  for (int s = 0; s < NSTEP; ++s) {
    R f0, f1, j00, j01, j10, j11;
    cross_residual<R>(X0, X1, e, cc, f0, f1, j00, j01, j10, j11);
    R idet = ad_recip(j00 * j11 - j01 * j10);
    R n0 = X0 - (j11 * f0 - j01 * f1) * idet;
    R n1 = X1 - (j00 * f1 - j10 * f0) * idet;
    X0 = n0;
    X1 = n1;
  }
  R xb0 = ad_recip(1.0 + X0 * e[0] + X1 * e[1]);
  R xb1 = ad_recip(1.0 + X0 * e[2] + X1 * e[3]);
  return ra[0] * site_term(X0) + ra[1] * site_term(X1) + rb[0] * site_term(xb0) + rb[1] * site_term(xb1);
}

// ---- induced association (:323-393): one unknown X_A -------------------------------------------
template <class R>
FD void induced_residual(const R& x, const R& s0, const R& s1, const R (&t)[4], const R& na0, const R& na1, R& f,
                         R& df) {
  // b_i = 1 + x s_i ; f_i = x (b0 b1 + t[2i] b1 + t[2i+1] b0) - b0 b1 ; f = na0 f0 + na1 f1
  R b0 = 1.0 + x * s0, b1 = 1.0 + x * s1;
  R bb = b0 * b1;
  R dbb = s0 * b1 + b0 * s1;
  R g0 = bb + t[0] * b1 + t[1] * b0;
  R g1 = bb + t[2] * b1 + t[3] * b0;
  R dg0 = dbb + t[0] * s1 + t[1] * s0;
  R dg1 = dbb + t[2] * s1 + t[3] * s0;
  f = na0 * (x * g0 - bb) + na1 * (x * g1 - bb);
  df = na0 * (g0 + x * dg0 - dbb) + na1 * (g1 + x * dg1 - dbb);
}

template <class R, int NSTEP, bool GC>
FD R induced_assoc(const R (&rho)[2], const R (&dl)[4], const R (&na)[2], const R (&nb)[2]) {
  // d_ij = Delta_ij rho_j (:353-366)
  R d00 = dl[0] * rho[0], d01 = dl[1] * rho[1], d10 = dl[2] * rho[0], d11 = dl[3] * rho[1];
  R s0 = na[0] * d00 + na[1] * d01;
  R s1 = na[0] * d10 + na[1] * d11;
  R t[4] = {nb[0] * d00, nb[1] * d01, nb[0] * d10, nb[1] * d11};
  double tp[4] = {ad_primal(t[0]), ad_primal(t[1]), ad_primal(t[2]), ad_primal(t[3])};
  double s0p = ad_primal(s0), s1p = ad_primal(s1), na0p = ad_primal(na[0]), na1p = ad_primal(na[1]);
  double x = 0.2;
  for (int it = 0; it < 50; ++it) {
    double f, df;
    induced_residual<double>(x, s0p, s1p, tp, na0p, na1p, f, df);
    double nx = x - f / df;
    if (!GC && nx < 0.0) nx = 0.2 * x;  // :381-383 (not in gc_pcsaft.py:414-433)
    bool done = fabs(nx - x) <= 1e-15;
    x = nx;
    if (done) break;
  }
  R X(x);
#pragma unroll
  for (int s = 0; s < NSTEP; ++s) {
    R f, df;
    induced_residual<R>(X, s0, s1, t, na[0], na[1], f, df);
    R nx = X - f * ad_recip(df);
    X = nx;
  }
  R xb0 = ad_recip(1.0 + X * s0);
  R xb1 = ad_recip(1.0 + X * s1);
  return rho[0] * (site_term(X) * na[0] + site_term(xb0) * nb[0]) + rho[1] * (site_term(X) * na[1] + site_term(xb1) * nb[1]);
}

// number of differentiable Newton steps needed for exact derivatives of the site fractions
template <class R> struct AssocSteps { static constexpr int value = 1; };                 // first order
template <class T> struct AssocSteps<H2<T>> { static constexpr int value = 2; };          // second order
template <class T> struct AssocSteps<D2T<T>> { static constexpr int value = 2; };
template <int N> struct AssocSteps<D1T<DN<N>>> { static constexpr int value = 2; };       // mixed second order
template <> struct AssocSteps<double> { static constexpr int value = 0; };

// GC cross association (gc_pcsaft.py:332-380, nA = nB = 1): f_i = x_i + x_i sum_j x_j d_ij - 1
template <class R>
FD void gc_cross_residual(const R& x0, const R& x1, const R (&d)[4], R& f0, R& f1, R& j00, R& j01, R& j10, R& j11) {
  f0 = x0 + x0 * (x0 * d[0] + x1 * d[1]) - 1.0;
  f1 = x1 + x1 * (x0 * d[2] + x1 * d[3]) - 1.0;
  j00 = 1.0 + 2.0 * (x0 * d[0]) + x1 * d[1];
  j01 = x0 * d[1];
  j10 = x1 * d[2];
  j11 = 1.0 + x0 * d[2] + 2.0 * (x1 * d[3]);
}

template <class R, int NSTEP>
FD R gc_cross_assoc(const R (&rho)[2], const R (&dl)[4]) {
  R d[4] = {dl[0] * rho[0], dl[1] * rho[1], dl[2] * rho[0], dl[3] * rho[1]};  // Delta_ij rho_j
  double dp[4] = {ad_primal(d[0]), ad_primal(d[1]), ad_primal(d[2]), ad_primal(d[3])};
  double x0 = 0.2, x1 = 0.2;
  for (int it = 0; it < 50; ++it) {
    double f0, f1, j00, j01, j10, j11;
    gc_cross_residual<double>(x0, x1, dp, f0, f1, j00, j01, j10, j11);
    double det = j00 * j11 - j01 * j10;
    double n0 = x0 - (j11 * f0 - j01 * f1) / det;
    double n1 = x1 - (-j10 * f0 + j00 * f1) / det;
    bool done = fabs(n0 - x0) <= 1e-15 && fabs(n1 - x1) <= 1e-15;
    x0 = n0;
    x1 = n1;
    if (done) break;
  }
  R X0(x0), X1(x1);
#pragma unroll
  for (int s = 0; s < NSTEP; ++s) {
    R f0, f1, j00, j01, j10, j11;
    gc_cross_residual<R>(X0, X1, d, f0, f1, j00, j01, j10, j11);
    R idet = ad_recip(j00 * j11 - j01 * j10);
    R n0 = X0 - (j11 * f0 - j01 * f1) * idet;
    R n1 = X1 - (j00 * f1 - j10 * f0) * idet;
    X0 = n0;
    X1 = n1;
  }
  return rho[0] * (2.0 * ad_log(X0) - X0 + 1.0) + rho[1] * (2.0 * ad_log(X1) - X1 + 1.0);
}

// PARTS: bit 0 = hard sphere + hard chain + dispersion (the "base" group), bit 1 = dipole term, bit 2 =
// association.  The gradient kernels evaluate the groups separately (closed-form coefficient adjoints for
// the base and dipole groups, dual passes for association).
template <class P, class R, int NB, bool GC, int PARTS = 7>
FD R mix_helmholtz(const MixCoef<P, NB, GC>& c, const R& rho0, const R& rho1) {
  R z0 = rho0 * c.z[0][0] + rho1 * c.z[0][1];
  R z1 = rho0 * c.z[1][0] + rho1 * c.z[1][1];
  R z2 = rho0 * c.z[2][0] + rho1 * c.z[2][1];
  R z3 = rho0 * c.z[3][0] + rho1 * c.z[3][1];
  R om = 1.0 - z3;
  R omi = ad_recip(om);
  R omi2 = omi * omi;
  R z23 = z2 * ad_recip(z3);

  R a(0.0);
  R e2 = z3 * z3, e3 = e2 * z3, e4 = e2 * e2;
  if (PARTS & 1) {
  // hard sphere (:56-60)
  a = (6.0 / kPi) * ((3.0 * (z1 * z2)) * omi + ((z2 * z2) * omi2) * z23 + ((z2 * z23) * z23 - z0) * ad_log(om));

  // hard chain (:63-65; gc_pcsaft.py:156-165): g = 1/(1-z3) + 3 c cd + 2 (c cd)^2 (1 - z3)
  R cz = z2 * omi2;
  if (NB == 2) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      R x = cz * c.cd[b];
      R g = omi + 3.0 * x + (2.0 * (x * x)) * om;
      a = a - (rho0 * c.bn[b][0] + rho1 * c.bn[b][1]) * ad_log(g);
    }
  } else {
    for (int b = 0; b < c.nbond; ++b) {
      R x = cz * c.cd[b];
      R g = omi + 3.0 * x + (2.0 * (x * x)) * om;
      a = a - (rho0 * c.bn[b][0] + rho1 * c.bn[b][1]) * ad_log(g);
    }
  }

  // dispersion (:67-106)
  R rt = rho0 + rho1;
  R mbar = (rho0 * c.m[0] + rho1 * c.m[1]) * ad_recip(rt);
  R r00 = rho0 * rho0, r01 = rho0 * rho1, r11 = rho1 * rho1;
  R rho1mix = r00 * c.q1[0] + (2.0 * r01) * c.q1[1] + r11 * c.q1[2];
  R rho2mix = r00 * c.q2[0] + (2.0 * r01) * c.q2[1] + r11 * c.q2[2];
  R e5 = e2 * e3, e6 = e3 * e3;
  {
    R mi = ad_recip(mbar);
    R w1 = (mbar - 1.0) * mi;
    R w2 = (w1 * (mbar - 2.0)) * mi;
    R sa0 = cA0[0] + z3 * cA0[1] + e2 * cA0[2] + e3 * cA0[3] + e4 * cA0[4] + e5 * cA0[5] + e6 * cA0[6];
    R sa1 = cA1[0] + z3 * cA1[1] + e2 * cA1[2] + e3 * cA1[3] + e4 * cA1[4] + e5 * cA1[5] + e6 * cA1[6];
    R sa2 = cA2[0] + z3 * cA2[1] + e2 * cA2[2] + e3 * cA2[3] + e4 * cA2[4] + e5 * cA2[5] + e6 * cA2[6];
    R sb0 = cB0[0] + z3 * cB0[1] + e2 * cB0[2] + e3 * cB0[3] + e4 * cB0[4] + e5 * cB0[5] + e6 * cB0[6];
    R sb1 = cB1[0] + z3 * cB1[1] + e2 * cB1[2] + e3 * cB1[3] + e4 * cB1[4] + e5 * cB1[5] + e6 * cB1[6];
    R sb2 = cB2[0] + z3 * cB2[1] + e2 * cB2[2] + e3 * cB2[3] + e4 * cB2[4] + e5 * cB2[5] + e6 * cB2[6];
    R i1 = sa0 + w1 * sa1 + w2 * sa2;
    R i2 = sb0 + w1 * sb1 + w2 * sb2;
    R tmi = ad_recip(2.0 - z3);
    R c1 = ad_recip(1.0 + mbar * ((8.0 * z3 - 2.0 * e2) * (omi2 * omi2)) +
                    (1.0 - mbar) * ((20.0 * z3 - 27.0 * e2 + 12.0 * e3 - 2.0 * e4) * (omi2 * (tmi * tmi))));
    a = a + (-kPi) * (2.0 * (rho1mix * i1) + (rho2mix * c1) * (i2 * mbar));
  }
  }  // PARTS & 1

  // dipoles (:156-208)
  if ((PARTS & 2) && c.dipolar) {
    R t0 = rho0 * c.mu2t[0], t1 = rho1 * c.mu2t[1];
    R tt[3] = {t0 * t0, t0 * t1, t1 * t1};
    R phi2(0.0), phi3(0.0);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      // pair integral: sum_n eta^n ((a_n0 + w1 a_n1 + w2 a_n2) + eps_ij/T (b_n0 + w1 b_n1 + w2 b_n2))
      R pint = (cAD[0][0] + c.pw1[k] * cAD[0][1] + c.pw2[k] * cAD[0][2] + c.pe[k] * (cBD[0][0] + c.pw1[k] * cBD[0][1] + c.pw2[k] * cBD[0][2])) +
               z3 * (cAD[1][0] + c.pw1[k] * cAD[1][1] + c.pw2[k] * cAD[1][2] + c.pe[k] * (cBD[1][0] + c.pw1[k] * cBD[1][1] + c.pw2[k] * cBD[1][2])) +
               e2 * (cAD[2][0] + c.pw1[k] * cAD[2][1] + c.pw2[k] * cAD[2][2] + c.pe[k] * (cBD[2][0] + c.pw1[k] * cBD[2][1] + c.pw2[k] * cBD[2][2])) +
               e3 * (cAD[3][0] + c.pw1[k] * cAD[3][1] + c.pw2[k] * cAD[3][2]) + e4 * (cAD[4][0] + c.pw1[k] * cAD[4][1] + c.pw2[k] * cAD[4][2]);
      phi2 = phi2 - (tt[k] * pint) * c.ps[k];
    }
    R ttt[4] = {tt[0] * t0, tt[0] * t1, tt[2] * t0, tt[2] * t1};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      R tint = (cCD[0][0] + c.tw1[t] * cCD[0][1] + c.tw2[t] * cCD[0][2]) + z3 * (cCD[1][0] + c.tw1[t] * cCD[1][1] + c.tw2[t] * cCD[1][2]) +
               e2 * (cCD[2][0] + c.tw1[t] * cCD[2][1] + c.tw2[t] * cCD[2][2]) + e3 * (cCD[3][0] + c.tw1[t] * cCD[3][1] + c.tw2[t] * cCD[3][2]);
      phi3 = phi3 - (ttt[t] * tint) * c.ts[t];
    }
    phi2 = phi2 * kPi;
    phi3 = phi3 * (4.0 / 3.0 * kPi * kPi);
    a = a + (phi2 * phi2) * ad_recip(phi2 - phi3);
  }

  // association (:118-152)
  if ((PARTS & 4) && c.cat != kAssocNone) {
    R z2om = z2 * omi;
    if (c.cat == kAssocSelf) {
      R delta = mix_strength(c, 0, z2om, omi);
      if (GC) {
        // gc_pcsaft.py:309-330 (nA = nB = 1): X = 2 / (sqrt(1 + 4 Delta rho) + 1)
        R rs = rho0 * c.sgn[0] + rho1 * c.sgn[1];
        R xa = 2.0 * ad_recip(ad_sqrt(1.0 + 4.0 * (delta * rs)) + 1.0);
        a = a + rs * (2.0 * ad_log(xa) - xa + 1.0);
      } else {
        R ra = rho0 * c.na[0] + rho1 * c.na[1];
        R rb = rho0 * c.nb[0] + rho1 * c.nb[1];
        R dd = (rb - ra) * delta;
        R aux = 1.0 - dd;
        R sq = ad_sqrt(aux * aux + 4.0 * (rb * delta));
        R xa = 2.0 * ad_recip(sq + 1.0 + dd);
        R xb = 2.0 * ad_recip(sq + 1.0 - dd);
        a = a + ra * site_term(xa) + rb * site_term(xb);
      }
    } else {
      R dl[4] = {mix_strength(c, 0, z2om, omi), mix_strength(c, 1, z2om, omi), mix_strength(c, 2, z2om, omi),
                 mix_strength(c, 3, z2om, omi)};
      if (c.cat == kAssocCross) {
        if (GC) {
          R rr[2] = {rho0, rho1};
          a = a + gc_cross_assoc<R, AssocSteps<R>::value>(rr, dl);
        } else {
          R ra[2] = {rho0 * c.na[0], rho1 * c.na[1]};
          R rb[2] = {rho0 * c.nb[0], rho1 * c.nb[1]};
          a = a + cross_assoc<R, AssocSteps<R>::value>(ra, rb, dl);
        }
      } else {
        R rr[2] = {rho0, rho1};
        R na[2] = {R(0.0) + c.na[0], R(0.0) + c.na[1]};
        R nb[2] = {R(0.0) + c.nb[0], R(0.0) + c.nb[1]};
        a = a + induced_assoc<R, AssocSteps<R>::value, GC>(rr, dl, na, nb);
      }
    }
  }
  return a;
}

// ---- thermodynamic properties from one H2 evaluation (pcsaft_mix.py:413-418) --------------------
struct MixEval {
  double a, mu[2], h00, h01, h11;  // a, a_i, a_ij
  double p;                        // sum rho - a + sum rho_i a_i
  double dp[2];                    // dp/drho_j = 1 + sum_i rho_i a_ij
  double v[2];                     // partial molar volumes (reduced)
};

FD MixEval mix_eval_from(const H2<double>& a, double rho0, double rho1);

// every term in hyper-dual arithmetic (the reference's way; mix_eval in mix_eval_fast.cuh is checked against it)
template <int NB, bool GC>
FD MixEval mix_eval_generic(const MixCoef<double, NB, GC>& c, double rho0, double rho1) {
  H2<double> r0(rho0), r1(rho1);
  r0.g0 = 1.0;
  r1.g1 = 1.0;
  return mix_eval_from(mix_helmholtz<double, H2<double>, NB, GC>(c, r0, r1), rho0, rho1);
}

FD MixEval mix_eval_from(const H2<double>& a, double rho0, double rho1) {
  MixEval e;
  e.a = a.re;
  e.mu[0] = a.g0;
  e.mu[1] = a.g1;
  e.h00 = a.h00;
  e.h01 = a.h01;
  e.h11 = a.h11;
  e.p = rho0 + rho1 - a.re + rho0 * a.g0 + rho1 * a.g1;
  e.dp[0] = 1.0 + rho0 * a.h00 + rho1 * a.h01;
  e.dp[1] = 1.0 + rho0 * a.h01 + rho1 * a.h11;
  double den = rho0 * e.dp[0] + rho1 * e.dp[1];  // = rho + sum_ij rho_i rho_j a_ij
  e.v[0] = e.dp[0] / den;
  e.v[1] = e.dp[1] / den;
  return e;
}

}  // namespace feos
