// Pure-component PC-SAFT reduced residual Helmholtz energy density a = A/(V kT) [1/A^3].
//
// Device-side, in-register replacement for PcSaftPure.helmholtz_energy
// (/root/reference/feos_torch/pcsaft_pure.py:106-178) and the parameter preprocessing of
// PcSaftPure.__init__ (:90-103).  The evaluation is split in two stages:
//   pure_coef<P>()      everything that depends on (parameters, T) only -- evaluated once per
//                       state and kept in registers across all Newton iterations;
//   pure_helmholtz<P,R> the density dependent part, templated on the number type R of the density
//                       (double, D2 = value+1st+2nd density derivative, DN<K> = parameter
//                       directions, D1T<DN<K>> = both).
#pragma once
#include "ad.cuh"

namespace feos {

constexpr double kPi = 3.14159265358979323846;
constexpr double kKB = 1.380649e-23;          // J/K (exact, CODATA 2018; si_units.KB)
constexpr double kNAV = 6.02214076e23;        // 1/mol
constexpr double kDipoleScale = 1e-19 / kKB;  // pcsaft_pure.py:95-100
constexpr double kPaPerRed = kKB / 1e-30;     // p[Pa] = p_red * T * kPaPerRed  (pcsaft_pure.py:215)
constexpr double kKmolM3PerRed = 1.0 / (1.0e3 * kNAV * 1e-30);  // pcsaft_pure.py:199
constexpr double kMaxEta = 0.5;               // rho_max = 0.5 / (pi/6 m sigma^3), as feos' PC-SAFT

// universal constants, pcsaft_pure.py:10-86 (Gross & Sadowski 2001; Gross & Vrabec 2006)
FEOS_CONSTANT double cA0[7] = {0.91056314451539, 0.63612814494991, 2.68613478913903, -26.5473624914884,
                               97.7592087835073, -159.591540865600, 91.2977740839123};
FEOS_CONSTANT double cA1[7] = {-0.30840169182720, 0.18605311591713, -2.50300472586548, 21.4197936296668,
                               -65.2558853303492, 83.3186804808856, -33.7469229297323};
FEOS_CONSTANT double cA2[7] = {-0.09061483509767, 0.45278428063920, 0.59627007280101, -1.72418291311787,
                               -4.13021125311661, 13.7766318697211, -8.67284703679646};
FEOS_CONSTANT double cB0[7] = {0.72409469413165, 2.23827918609380, -4.00258494846342, -21.00357681484648,
                               26.8556413626615, 206.5513384066188, -355.60235612207947};
FEOS_CONSTANT double cB1[7] = {-0.57554980753450, 0.69950955214436, 3.89256733895307, -17.21547164777212,
                               192.6722644652495, -161.8264616487648, -165.2076934555607};
FEOS_CONSTANT double cB2[7] = {0.09768831158356, -0.25575749816100, -9.15585615297321, 20.64207597439724,
                               -38.80443005206285, 93.6267740770146, -29.66690558514725};
FEOS_CONSTANT double cAD[5][3] = {{0.30435038064, 0.95346405973, -1.16100802773},
                                  {-0.13585877707, -1.83963831920, 4.52586067320},
                                  {1.44933285154, 2.01311801180, 0.97512223853},
                                  {0.35569769252, -7.37249576667, -12.2810377713},
                                  {-2.06533084541, 8.23741345333, 5.93975747420}};
FEOS_CONSTANT double cBD[3][3] = {{0.21879385627, -0.58731641193, 3.48695755800},
                                  {-1.18964307357, 1.24891317047, -14.9159739347},
                                  {1.16268885692, -0.50852797392, 15.3720218600}};  // rows 3,4 are zero
FEOS_CONSTANT double cCD[4][3] = {{-0.06467735252, -0.95208758351, -0.62609792333},
                                  {0.19758818347, 2.99242575222, 1.29246858189},
                                  {-0.80875619458, -2.38026356489, 1.65427830900},
                                  {0.69028490492, -0.27012609786, -3.43967436378}};

template <class P>
struct PureCoef {
  P m, mm1, c_eta, w1, w2, k1, k2;  // hard sphere / chain / dispersion
  P mu2t, v1, v2, e, inv_s3;        // dipole
  P delta0, na, nb;                 // association
  bool dipolar, assoc;
};

// Stage 1: (parameters, T) -> coefficients.  `grad_mode` keeps association terms whose value
// vanishes but whose parameter derivative does not (kappa = 0 or eps_ab = 0 with sites present).
template <class P>
FD void pure_coef(PureCoef<P>& c, const P& m, const P& sigma, const P& eps, const P& mu, const P& kappa,
                  const P& eps_ab, const P& na, const P& nb, const P& T, bool grad_mode) {
  P s3 = sigma * sigma * sigma;
  P ti = ad_recip(T);
  P e = eps * ti;
  P d = sigma * (1.0 - 0.12 * ad_exp(-3.0 * e));  // pcsaft_pure.py:108
  c.m = m;
  c.mm1 = m - 1.0;
  c.c_eta = (kPi / 6.0) * m * (d * d * d);  // eta = c_eta * rho   (:110)
  P mi = ad_recip(m);
  P m1 = (m - 1.0) * mi;
  P m2 = (m - 2.0) * mi;
  c.w1 = m1;
  c.w2 = m1 * m2;  // (m1*(m2*A2+A1)+A0)  (:129-133)
  P mes = m * m * e * s3;
  c.k1 = (-2.0 * kPi) * mes;  // disp = rho^2 (k1 I1 + k2 C1 I2)   (:141-142)
  c.k2 = (-kPi) * mes * m * e;
  c.e = e;
  // dipole (:145-160); identically zero with zero gradient for mu = 0 (README.md:28)
  c.dipolar = (ad_primal(mu) != 0.0);
  if (c.dipolar) {
    P mu2 = mu * mu / (m * s3 * eps) * kDipoleScale;  // :94-99
    c.mu2t = mu2 * e * s3;
    P mc = ad_clamp_max(m, 2.0);
    P mci = ad_recip(mc);
    c.v1 = (mc - 1.0) * mci;
    c.v2 = c.v1 * (mc - 2.0) * mci;
    c.inv_s3 = ad_recip(s3);
  }
  // association (:163-176).  Value and density derivatives vanish identically when kappa or
  // eps_ab is zero (Delta = 0) or when a site count is zero; all parameter derivatives vanish only
  // when na = nb = 0.
  double na0 = ad_primal(na), nb0 = ad_primal(nb);
  bool sites = (na0 != 0.0) || (nb0 != 0.0);
  bool strength = (ad_primal(kappa) != 0.0) && (ad_primal(eps_ab) != 0.0);
  c.assoc = grad_mode ? sites : (strength && na0 != 0.0 && nb0 != 0.0);
  if (c.assoc) {
    c.delta0 = (ad_exp(eps_ab * ti) - 1.0) * s3 * kappa;
    c.na = na;
    c.nb = nb;
  }
}

template <class R>
FD R site_term(const R& x) { return ad_log(x) - 0.5 * x + 0.5; }

template <class P, class R>
FD R pure_helmholtz_extras(const PureCoef<P>& c, const R& rho, const R& eta, const R& e2, const R& e3, const R& e4,
                           const R& omi);

// Stage 2: a(rho).  pcsaft_pure.py:110-178.
template <class P, class R>
FD R pure_helmholtz(const PureCoef<P>& c, const R& rho) {
  R eta = rho * c.c_eta;
  R e2 = eta * eta;
  R e3 = e2 * eta;
  R e4 = e2 * e2;
  R e5 = e2 * e3;
  R e6 = e3 * e3;
  R om = 1.0 - eta;
  R omi = ad_recip(om);
  R omi2 = omi * omi;

  // hard sphere (:118) and hard chain (:121-122)
  R a = (rho * c.m) * ((4.0 * eta - 3.0 * e2) * omi2);
  R g = (1.0 - 0.5 * eta) * (omi * omi2);
  a = a - (rho * c.mm1) * ad_log(g);

  // dispersion (:125-142)
  R i1 = eta * (cA0[1] + c.w1 * cA1[1] + c.w2 * cA2[1]) + (cA0[0] + c.w1 * cA1[0] + c.w2 * cA2[0]);
  R i2 = eta * (cB0[1] + c.w1 * cB1[1] + c.w2 * cB2[1]) + (cB0[0] + c.w1 * cB1[0] + c.w2 * cB2[0]);
  i1 = i1 + e2 * (cA0[2] + c.w1 * cA1[2] + c.w2 * cA2[2]);
  i2 = i2 + e2 * (cB0[2] + c.w1 * cB1[2] + c.w2 * cB2[2]);
  i1 = i1 + e3 * (cA0[3] + c.w1 * cA1[3] + c.w2 * cA2[3]);
  i2 = i2 + e3 * (cB0[3] + c.w1 * cB1[3] + c.w2 * cB2[3]);
  i1 = i1 + e4 * (cA0[4] + c.w1 * cA1[4] + c.w2 * cA2[4]);
  i2 = i2 + e4 * (cB0[4] + c.w1 * cB1[4] + c.w2 * cB2[4]);
  i1 = i1 + e5 * (cA0[5] + c.w1 * cA1[5] + c.w2 * cA2[5]);
  i2 = i2 + e5 * (cB0[5] + c.w1 * cB1[5] + c.w2 * cB2[5]);
  i1 = i1 + e6 * (cA0[6] + c.w1 * cA1[6] + c.w2 * cA2[6]);
  i2 = i2 + e6 * (cB0[6] + c.w1 * cB1[6] + c.w2 * cB2[6]);
  R tmi = ad_recip(2.0 - eta);
  R c1 = ad_recip(1.0 + ((8.0 * eta - 2.0 * e2) * (omi2 * omi2)) * c.m -
                  ((20.0 * eta - 27.0 * e2 + 12.0 * e3 - 2.0 * e4) * (omi2 * (tmi * tmi))) * c.mm1);
  R rho2 = rho * rho;
  a = a + rho2 * (i1 * c.k1 + (c1 * i2) * c.k2);

  if (c.dipolar || c.assoc) a = a + pure_helmholtz_extras<P, R>(c, rho, eta, e2, e3, e4, omi);
  return a;
}

// dipole and association contributions (zero for the non-polar, non-associating majority)
template <class P, class R>
FD R pure_helmholtz_extras(const PureCoef<P>& c, const R& rho, const R& eta, const R& e2, const R& e3, const R& e4,
                           const R& omi) {
  R a(0.0);
  R rho2 = rho * rho;
  // dipole (:145-160)
  if (c.dipolar) {
    R j1 = ((cAD[0][0] + c.v1 * cAD[0][1] + c.v2 * cAD[0][2]) + c.e * (cBD[0][0] + c.v1 * cBD[0][1] + c.v2 * cBD[0][2])) +
           eta * ((cAD[1][0] + c.v1 * cAD[1][1] + c.v2 * cAD[1][2]) + c.e * (cBD[1][0] + c.v1 * cBD[1][1] + c.v2 * cBD[1][2])) +
           e2 * ((cAD[2][0] + c.v1 * cAD[2][1] + c.v2 * cAD[2][2]) + c.e * (cBD[2][0] + c.v1 * cBD[2][1] + c.v2 * cBD[2][2])) +
           e3 * (cAD[3][0] + c.v1 * cAD[3][1] + c.v2 * cAD[3][2]) + e4 * (cAD[4][0] + c.v1 * cAD[4][1] + c.v2 * cAD[4][2]);
    R j2 = eta * (cCD[1][0] + c.v1 * cCD[1][1] + c.v2 * cCD[1][2]) + (cCD[0][0] + c.v1 * cCD[0][1] + c.v2 * cCD[0][2]) +
           e2 * (cCD[2][0] + c.v1 * cCD[2][1] + c.v2 * cCD[2][2]) + e3 * (cCD[3][0] + c.v1 * cCD[3][1] + c.v2 * cCD[3][2]);
    R phi2 = (rho2 * j1) * (c.inv_s3 * (-kPi));
    R phi3 = ((rho2 * rho) * j2) * (c.inv_s3 * (-4.0 / 3.0 * kPi * kPi));
    R num = (phi2 * phi2) * (c.mu2t * c.mu2t);
    a = a + num / (phi2 - phi3 * c.mu2t);
  }

  // association, closed-form two-site solution (:163-176)
  if (c.assoc) {
    R k = eta * omi;
    R delta = ((1.0 + k * (1.5 + 0.5 * k)) * omi) * c.delta0;
    R ra = rho * c.na;
    R rb = rho * c.nb;
    R dd = (rb - ra) * delta;
    R aux = 1.0 - dd;
    R sq = ad_sqrt(aux * aux + 4.0 * (rb * delta));
    R xa = 2.0 / (sq + 1.0 + dd);
    R xb = 2.0 / (sq + 1.0 - dd);
    a = a + ra * site_term(xa) + rb * site_term(xb);
  }
  return a;
}

// (p/kT, d(p/kT)/drho, a, a') from one second-order evaluation (pcsaft_pure.py:180-182)
struct PureEval {
  double a, a1, p, dp;
};
FD PureEval pure_eval_generic(const PureCoef<double>& c, double rho) {
  D2 r(rho, 1.0, 0.0);
  D2 a = pure_helmholtz<double, D2>(c, r);
  PureEval o;
  o.a = a.re;
  o.a1 = a.d1;
  o.p = rho - a.re + rho * a.d1;
  o.dp = 1.0 + rho * a.d2;
  return o;
}

// The same quantities with the hard-sphere, hard-chain and dispersion terms differentiated in closed
// form: they depend on the density only through eta = c_eta rho, so value, first and second derivative
// follow from scalar recurrences (Horner with derivatives for the I1/I2 series) instead of generic
// second-order dual arithmetic -- about 2.5x fewer FP64 instructions per evaluation.  With
//   a = rho f(eta) + rho^2 g(eta),  f = m F_hs - (m-1) F_hc,  g = k1 I1 + k2 C1 I2:
//   a'  = f + eta f' + rho (2 g + eta g'),   a'' = c_eta (2 f' + eta f'') + 2 g + 4 eta g' + eta^2 g''.
// Dipole and association terms (when present) still go through the dual-number path.
FD PureEval pure_eval(const PureCoef<double>& c, double rho) {
  const double eta = rho * c.c_eta;
  const double s = 1.0 / (1.0 - eta);  // d s / d eta = s^2
  const double t = 1.0 / (2.0 - eta);  // d t / d eta = t^2
  const double s2 = s * s, s3 = s2 * s, s4 = s2 * s2;
  // hard sphere F_hs = (4 eta - 3 eta^2) s^2 (:118) and hard chain F_hc = ln((1 - eta/2) s^3) (:121-122)
  const double fhs = (4.0 - 3.0 * eta) * eta * s2, fhs1 = (4.0 - 2.0 * eta) * s3, fhs2 = (10.0 - 4.0 * eta) * s4;
  const double fhc = log((1.0 - 0.5 * eta) * s3), fhc1 = 3.0 * s - t, fhc2 = 3.0 * s2 - t * t;
  const double f = c.m * fhs - c.mm1 * fhc, f1 = c.m * fhs1 - c.mm1 * fhc1, f2 = c.m * fhs2 - c.mm1 * fhc2;
  // dispersion series (:125-133): Horner with first and second derivative
  double i1 = cA0[6] + c.w1 * cA1[6] + c.w2 * cA2[6], i1d = 0.0, i1dd = 0.0;
  double i2 = cB0[6] + c.w1 * cB1[6] + c.w2 * cB2[6], i2d = 0.0, i2dd = 0.0;
#pragma unroll
  for (int k = 5; k >= 0; --k) {
    i1dd = i1dd * eta + i1d;
    i1d = i1d * eta + i1;
    i1 = i1 * eta + (cA0[k] + c.w1 * cA1[k] + c.w2 * cA2[k]);
    i2dd = i2dd * eta + i2d;
    i2d = i2d * eta + i2;
    i2 = i2 * eta + (cB0[k] + c.w1 * cB1[k] + c.w2 * cB2[k]);
  }
  i1dd *= 2.0;
  i2dd *= 2.0;
  // C1 = 1 / D, D = 1 + m X + (1 - m) Y (:134-140), X = (8 eta - 2 eta^2) s^4, Y = N s^2 t^2
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
  const double d0 = 1.0 + c.m * x0 - c.mm1 * y0, d1 = c.m * x1 - c.mm1 * y1, d2 = c.m * x2 - c.mm1 * y2;
  const double c1 = 1.0 / d0;
  const double c1d = -d1 * c1 * c1;
  const double c1dd = (2.0 * d1 * d1 * c1 - d2) * c1 * c1;
  const double g = c.k1 * i1 + c.k2 * (c1 * i2);
  const double g1 = c.k1 * i1d + c.k2 * (c1d * i2 + c1 * i2d);
  const double g2 = c.k1 * i1dd + c.k2 * (c1dd * i2 + 2.0 * c1d * i2d + c1 * i2dd);
  double a = rho * (f + rho * g);
  double a1 = f + eta * f1 + rho * (2.0 * g + eta * g1);
  double a2 = c.c_eta * (2.0 * f1 + eta * f2) + 2.0 * g + eta * (4.0 * g1 + eta * g2);
  if (c.assoc && !c.dipolar) {
    // association (:163-176) in closed form.  Delta = delta0 g(eta), g = (1 + 1.5 k + 0.5 k^2) s with k = eta s,
    // s = 1 / (1 - eta) (d k / d eta = d s / d eta = s^2).  With the site fractions at their mass-action solution,
    //   d a / d rho_S = ln X_S (fixed Delta),  d a / d Delta = - rho_A rho_B X_A X_B,
    // so a' needs no derivative of X; a'' takes X_A', X_B' from the linearised mass-action equations
    // (2 x 2 system, determinant X_A + X_B - X_A X_B > 0).
    const double k = eta * s;
    const double q = 1.0 + k * (1.5 + 0.5 * k), q1 = (1.5 + k) * s2, q2 = s4 + (1.5 + k) * 2.0 * s3;
    const double gq = q * s, gq1 = q1 * s + q * s2, gq2 = q2 * s + 2.0 * q1 * s2 + 2.0 * q * s3;
    const double delta = gq * c.delta0;  // same operations as the generic path: a is bit-identical
    const double dl1 = c.delta0 * gq1 * c.c_eta, dl2 = c.delta0 * gq2 * c.c_eta * c.c_eta;
    const double ra = rho * c.na, rb = rho * c.nb;
    const double dd = (rb - ra) * delta;
    const double aux = 1.0 - dd;
    const double sq = sqrt(aux * aux + 4.0 * (rb * delta));
    const double xa = 2.0 / (sq + 1.0 + dd), xb = 2.0 / (sq + 1.0 - dd);
    const double la = log(xa), lb = log(xb);
    a += ra * (la - 0.5 * xa + 0.5) + rb * (lb - 0.5 * xb + 0.5);
    const double w = ra * rb * xa * xb;
    a1 += c.na * la + c.nb * lb - w * dl1;
    // X_A' = u, X_B' = v:  u + X_A^2 rb Delta v = -p1,  X_B^2 ra Delta u + v = -p2
    const double ca = xa * xa * rb * delta, cb = xb * xb * ra * delta;
    const double p1 = xa * xa * xb * (c.nb * delta + rb * dl1), p2 = xb * xb * xa * (c.na * delta + ra * dl1);
    const double idet = 1.0 / (xa + xb - xa * xb);
    const double u = (ca * p2 - p1) * idet, v = (cb * p1 - p2) * idet;
    a2 += c.na * u / xa + c.nb * v / xb - (2.0 * c.na * c.nb * rho * xa * xb + ra * rb * (u * xb + xa * v)) * dl1 - w * dl2;
  } else if (c.dipolar || c.assoc) {
    D2 r(rho, 1.0, 0.0);
    D2 et = r * c.c_eta;
    D2 e2 = et * et;
    D2 x = pure_helmholtz_extras<double, D2>(c, r, et, e2, e2 * et, e2 * e2, ad_recip(1.0 - et));
    a += x.re;
    a1 += x.d1;
    a2 += x.d2;
  }
  PureEval o;
  o.a = a;
  o.a1 = a1;
  o.p = rho - a + rho * a1;
  o.dp = 1.0 + rho * a2;
  return o;
}

}  // namespace feos
