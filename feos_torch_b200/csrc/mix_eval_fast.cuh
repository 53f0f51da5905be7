// Closed-form value, gradient and Hessian in the partial densities of the BASE part (hard sphere + hard
// chain + dispersion) of the binary-mixture Helmholtz energy density -- the evaluation the bubble / dew
// solvers call ~10 times per state.
//
// The generic path evaluates mix_helmholtz in H2<double> arithmetic (value, 2 first, 3 second derivatives
// carried through every operation: ~1700 FP64 instructions, 37 KB of straight-line code).  Here the
// structure of the model is used instead: all density dependence enters through
//   zeta_0..zeta_3 (linear in rho), the bond weights B_b (linear), mbar (rational), Q1, Q2 (quadratic),
// and the only non-linear variables are (zeta_2, eta = zeta_3) for the hard-sphere / hard-chain terms and
// (eta, mbar) for dispersion.  The partial derivatives with respect to those few variables are written
// out by hand and contracted with the (constant) gradients of the intermediates.  Same formulas as
// mix_helmholtz (pcsaft_mix.py:56-106, gc_pcsaft.py:149-194), checked against it to round-off in
// tests/test_hostsim_mix.py / test_hostsim_gc.py.
#pragma once
#include "pcsaft_mix.cuh"

namespace feos {

struct BaseH2 {
  double a, g0, g1, h00, h01, h11;
};

// value, first and second derivative of a degree-6 polynomial
FD void horner7_2(const double (&k)[7], double x, double& v, double& d1, double& d2) {
  v = k[6];
  d1 = 0.0;
  d2 = 0.0;
#pragma unroll
  for (int n = 5; n >= 0; --n) {
    d2 = d2 * x + d1;
    d1 = d1 * x + v;
    v = v * x + k[n];
  }
  d2 *= 2.0;
}

template <int NB, class Coef>
FD BaseH2 mix_base_eval_impl(const Coef& c, double r0, double r1) {
  const double z00 = c.z[0][0], z01 = c.z[0][1], z10 = c.z[1][0], z11 = c.z[1][1];
  const double z20 = c.z[2][0], z21 = c.z[2][1], z30 = c.z[3][0], z31 = c.z[3][1];
  const double zt0 = r0 * z00 + r1 * z01, zt1 = r0 * z10 + r1 * z11, zt2 = r0 * z20 + r1 * z21;
  const double eta = r0 * z30 + r1 * z31;
  const double om = 1.0 - eta, omi = 1.0 / om, omi2 = omi * omi, omi3 = omi2 * omi, omi4 = omi2 * omi2;
  const double ie = 1.0 / eta, ie2 = ie * ie, ie3 = ie2 * ie;
  const double L = log(om);
  const double t2 = zt2 * zt2, t3 = t2 * zt2;
  const double cs = 6.0 / kPi;

  // ---- hard sphere: G = 3 z1 z2 / om + z2^3 / (eta om^2) + (z2^3 / eta^2 - z0) L --------------------
  const double u = 2.0 * ie * omi3 - ie2 * omi2;  // d/d eta of 1 / (eta om^2)
  double a = cs * (3.0 * zt1 * zt2 * omi + t3 * ie * omi2 + (t3 * ie2 - zt0) * L);
  double A0 = -cs * L;
  double A1 = cs * 3.0 * zt2 * omi;
  double A2 = cs * 3.0 * (zt1 * omi + t2 * ie * omi2 + t2 * ie2 * L);
  double A3 = cs * (3.0 * zt1 * zt2 * omi2 + t3 * u - 2.0 * t3 * ie3 * L - (t3 * ie2 - zt0) * omi);
  const double H03 = cs * omi;
  const double H12 = cs * 3.0 * omi;
  const double H13 = cs * 3.0 * zt2 * omi2;
  double H22 = cs * 6.0 * zt2 * (ie * omi2 + ie2 * L);
  double H23 = cs * 3.0 * (zt1 * omi2 + t2 * u - 2.0 * t2 * ie3 * L - t2 * ie2 * omi);
  double H33 = cs * (6.0 * zt1 * zt2 * omi3 + t3 * (6.0 * ie * omi4 - 4.0 * ie2 * omi3 + 2.0 * ie3 * omi2) +
                     6.0 * t3 * ie2 * ie2 * L + 4.0 * t3 * ie3 * omi - (t3 * ie2 - zt0) * omi2);
  double g0 = 0.0, g1 = 0.0, h00 = 0.0, h01 = 0.0, h11 = 0.0;  // parts that do not go through zeta_k

  // ---- hard chain: - sum_b B_b ln g_b,  g = 1/om + 3 x + 2 x^2 om,  x = zeta_2 cd / om^2 -----------------
  {
    const int nb = (NB == 2) ? 2 : c.nbond;
    for (int b = 0; b < nb; ++b) {
      const double cd = c.cd[b], bn0 = c.bn[b][0], bn1 = c.bn[b][1];
      const double B = r0 * bn0 + r1 * bn1;
      const double x2 = omi2 * cd;         // dx/dzeta2
      const double x = zt2 * x2;
      const double x3 = 2.0 * x * omi;     // dx/deta
      const double x23 = 2.0 * x2 * omi;
      const double x33 = 3.0 * x3 * omi;
      const double gx = 3.0 + 4.0 * x * om;  // dg/dx at fixed om
      const double gg = omi + 3.0 * x + 2.0 * x * x * om;
      const double gi = 1.0 / gg;
      const double gd2 = gx * x2;
      const double gd3 = omi2 + gx * x3 - 2.0 * x * x;
      const double gxe = 4.0 * (x3 * om - x);  // d gx / d eta
      const double gd22 = 4.0 * om * x2 * x2;
      const double gd23 = gxe * x2 + gx * x23;
      const double gd33 = 2.0 * omi3 + gxe * x3 + gx * x33 - 4.0 * x * x3;
      const double f = log(gg);
      const double f2 = gd2 * gi, f3 = gd3 * gi;
      const double f22 = gd22 * gi - f2 * f2, f23 = gd23 * gi - f2 * f3, f33 = gd33 * gi - f3 * f3;
      a -= B * f;
      A2 -= B * f2;
      A3 -= B * f3;
      H22 -= B * f22;
      H23 -= B * f23;
      H33 -= B * f33;
      g0 -= bn0 * f;
      g1 -= bn1 * f;
      const double t0 = f2 * z20 + f3 * z30, t1 = f2 * z21 + f3 * z31;
      h00 -= 2.0 * bn0 * t0;
      h01 -= bn0 * t1 + bn1 * t0;
      h11 -= 2.0 * bn1 * t1;
    }
  }

  // ---- dispersion: -pi [2 Q1 I1(eta, mbar) + Q2 K(eta, mbar)],  K = mbar C1 I2 ---------------------------
  {
    const double rti = 1.0 / (r0 + r1);
    const double mb = (r0 * c.m[0] + r1 * c.m[1]) * rti;
    const double m0 = (c.m[0] - mb) * rti, m1 = (c.m[1] - mb) * rti;  // d mbar / d rho_i
    const double m00 = -2.0 * m0 * rti, m01 = -(m0 + m1) * rti, m11 = -2.0 * m1 * rti;
    const double Q1 = r0 * r0 * c.q1[0] + 2.0 * r0 * r1 * c.q1[1] + r1 * r1 * c.q1[2];
    const double Q2 = r0 * r0 * c.q2[0] + 2.0 * r0 * r1 * c.q2[1] + r1 * r1 * c.q2[2];
    const double Q1_0 = 2.0 * (c.q1[0] * r0 + c.q1[1] * r1), Q1_1 = 2.0 * (c.q1[1] * r0 + c.q1[2] * r1);
    const double Q2_0 = 2.0 * (c.q2[0] * r0 + c.q2[1] * r1), Q2_1 = 2.0 * (c.q2[1] * r0 + c.q2[2] * r1);
    const double mi = 1.0 / mb, mi2 = mi * mi, mi3 = mi2 * mi;
    const double w1 = 1.0 - mi, w2 = w1 * (1.0 - 2.0 * mi);
    const double w1m = mi2, w1mm = -2.0 * mi3;
    const double w2m = mi2 * (3.0 - 4.0 * mi), w2mm = mi3 * (12.0 * mi - 6.0);
    double p0, p0e, p0ee, p1, p1e, p1ee, p2, p2e, p2ee;
    horner7_2(cA0, eta, p0, p0e, p0ee);
    horner7_2(cA1, eta, p1, p1e, p1ee);
    horner7_2(cA2, eta, p2, p2e, p2ee);
    const double I1 = p0 + w1 * p1 + w2 * p2;
    const double I1e = p0e + w1 * p1e + w2 * p2e, I1ee = p0ee + w1 * p1ee + w2 * p2ee;
    const double I1m = w1m * p1 + w2m * p2, I1em = w1m * p1e + w2m * p2e, I1mm = w1mm * p1 + w2mm * p2;
    horner7_2(cB0, eta, p0, p0e, p0ee);
    horner7_2(cB1, eta, p1, p1e, p1ee);
    horner7_2(cB2, eta, p2, p2e, p2ee);
    const double I2 = p0 + w1 * p1 + w2 * p2;
    const double I2e = p0e + w1 * p1e + w2 * p2e, I2ee = p0ee + w1 * p1ee + w2 * p2ee;
    const double I2m = w1m * p1 + w2m * p2, I2em = w1m * p1e + w2m * p2e, I2mm = w1mm * p1 + w2mm * p2;
    // C1 = 1 / D,  D = 1 + mbar X(eta) + (1 - mbar) Y(eta)
    const double e2 = eta * eta, e3 = e2 * eta;
    const double xn = 8.0 * eta - 2.0 * e2, xn1 = 8.0 - 4.0 * eta;
    const double omi5 = omi4 * omi;
    const double X = xn * omi4;
    const double Xe = xn1 * omi4 + 4.0 * xn * omi5;
    const double Xee = -4.0 * omi4 + 8.0 * xn1 * omi5 + 20.0 * xn * omi5 * omi;
    const double tmi = 1.0 / (2.0 - eta);
    const double yn = 20.0 * eta - 27.0 * e2 + 12.0 * e3 - 2.0 * e2 * e2;
    const double yn1 = 20.0 - 54.0 * eta + 36.0 * e2 - 8.0 * e3, yn2 = -54.0 + 72.0 * eta - 24.0 * e2;
    const double s = omi2 * tmi * tmi, so = omi + tmi;
    const double s1 = 2.0 * s * so;
    const double s2 = 2.0 * s1 * so + 2.0 * s * (omi2 + tmi * tmi);
    const double Y = yn * s, Ye = yn1 * s + yn * s1, Yee = yn2 * s + 2.0 * yn1 * s1 + yn * s2;
    const double D = 1.0 + mb * X + (1.0 - mb) * Y;
    const double De = mb * Xe + (1.0 - mb) * Ye, Dm = X - Y, Dee = mb * Xee + (1.0 - mb) * Yee, Dem = Xe - Ye;
    const double C = 1.0 / D, C2 = C * C, C3 = C2 * C;
    const double Ce = -C2 * De, Cm = -C2 * Dm;
    const double Cee = 2.0 * C3 * De * De - C2 * Dee, Cem = 2.0 * C3 * De * Dm - C2 * Dem, Cmm = 2.0 * C3 * Dm * Dm;
    // J = C1 I2 (bivariate product), K = mbar J
    const double J = C * I2;
    const double Je = Ce * I2 + C * I2e, Jm = Cm * I2 + C * I2m;
    const double Jee = Cee * I2 + 2.0 * Ce * I2e + C * I2ee;
    const double Jem = Cem * I2 + Ce * I2m + Cm * I2e + C * I2em;
    const double Jmm = Cmm * I2 + 2.0 * Cm * I2m + C * I2mm;
    const double K = mb * J, Ke = mb * Je, Km = J + mb * Jm;
    const double Kee = mb * Jee, Kem = Je + mb * Jem, Kmm = 2.0 * Jm + mb * Jmm;
    const double np = -kPi;
    a += np * (2.0 * Q1 * I1 + Q2 * K);
    A3 += np * (2.0 * Q1 * I1e + Q2 * Ke);
    H33 += np * (2.0 * Q1 * I1ee + Q2 * Kee);
    const double Sm = np * (2.0 * Q1 * I1m + Q2 * Km);
    const double Sem = np * (2.0 * Q1 * I1em + Q2 * Kem);
    const double Smm = np * (2.0 * Q1 * I1mm + Q2 * Kmm);
    g0 += Sm * m0 + np * (2.0 * Q1_0 * I1 + Q2_0 * K);
    g1 += Sm * m1 + np * (2.0 * Q1_1 * I1 + Q2_1 * K);
    // e1_j = d I1 / d rho_j, ek_j = d K / d rho_j
    const double e1_0 = I1e * z30 + I1m * m0, e1_1 = I1e * z31 + I1m * m1;
    const double ek_0 = Ke * z30 + Km * m0, ek_1 = Ke * z31 + Km * m1;
    h00 += Sem * 2.0 * z30 * m0 + Smm * m0 * m0 + Sm * m00 +
           np * (4.0 * Q1_0 * e1_0 + 2.0 * Q2_0 * ek_0 + 4.0 * c.q1[0] * I1 + 2.0 * c.q2[0] * K);
    h01 += Sem * (z30 * m1 + m0 * z31) + Smm * m0 * m1 + Sm * m01 +
           np * (2.0 * (Q1_0 * e1_1 + Q1_1 * e1_0) + (Q2_0 * ek_1 + Q2_1 * ek_0) + 4.0 * c.q1[1] * I1 + 2.0 * c.q2[1] * K);
    h11 += Sem * 2.0 * z31 * m1 + Smm * m1 * m1 + Sm * m11 +
           np * (4.0 * Q1_1 * e1_1 + 2.0 * Q2_1 * ek_1 + 4.0 * c.q1[2] * I1 + 2.0 * c.q2[2] * K);
  }

  // ---- contraction with the gradients of zeta_k ----------------------------------------------------------
  BaseH2 o;
  o.a = a;
  o.g0 = g0 + A0 * z00 + A1 * z10 + A2 * z20 + A3 * z30;
  o.g1 = g1 + A0 * z01 + A1 * z11 + A2 * z21 + A3 * z31;
  o.h00 = h00 + 2.0 * (H03 * z00 * z30 + H12 * z10 * z20 + H13 * z10 * z30 + H23 * z20 * z30) + H22 * z20 * z20 + H33 * z30 * z30;
  o.h01 = h01 + H03 * (z00 * z31 + z30 * z01) + H12 * (z10 * z21 + z20 * z11) + H13 * (z10 * z31 + z30 * z11) +
          H23 * (z20 * z31 + z30 * z21) + H22 * z20 * z21 + H33 * z30 * z31;
  o.h11 = h11 + 2.0 * (H03 * z01 * z31 + H12 * z11 * z21 + H13 * z11 * z31 + H23 * z21 * z31) + H22 * z21 * z21 + H33 * z31 * z31;
  return o;
}

template <int NB, bool GC>
FD BaseH2 mix_base_eval(const MixCoef<double, NB, GC>& c, double r0, double r1) {
  return mix_base_eval_impl<NB>(c, r0, r1);
}

// Gross-Vrabec dipole term a = phi2^2 / (phi2 - phi3) in closed form (value, gradient, Hessian), added to `o`:
//   phi2 = -pi sum_k tau_k(rho) ps_k J2_k(eta),  phi3 = -4/3 pi^2 sum_t tautau_t(rho) ts_t J3_t(eta)
// with tau = (t0^2, t0 t1, t1^2), tautau = (t0^3, t0^2 t1, t1^2 t0, t1^3), t_i = rho_i mu2t_i and the pair /
// triplet integrals J polynomial in eta (pcsaft_mix.py:156-208).
template <class C>
FD void mix_dipole_eval(const C& c, double r0, double r1, BaseH2& o) {
  const double z30 = c.z[3][0], z31 = c.z[3][1];
  const double eta = r0 * z30 + r1 * z31;
  // polynomials in eta with two derivatives: PA_j (degree 4), PB_j (degree 2), PC_j (degree 3)
  double pa[3], pa1[3], pa2[3], pb[3], pb1[3], pb2[3], pc[3], pc1[3], pc2[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    double v = cAD[4][j], d1 = 0.0, d2 = 0.0;
#pragma unroll
    for (int n = 3; n >= 0; --n) {
      d2 = d2 * eta + d1;
      d1 = d1 * eta + v;
      v = v * eta + cAD[n][j];
    }
    pa[j] = v; pa1[j] = d1; pa2[j] = 2.0 * d2;
    v = cBD[2][j]; d1 = 0.0; d2 = 0.0;
#pragma unroll
    for (int n = 1; n >= 0; --n) {
      d2 = d2 * eta + d1;
      d1 = d1 * eta + v;
      v = v * eta + cBD[n][j];
    }
    pb[j] = v; pb1[j] = d1; pb2[j] = 2.0 * d2;
    v = cCD[3][j]; d1 = 0.0; d2 = 0.0;
#pragma unroll
    for (int n = 2; n >= 0; --n) {
      d2 = d2 * eta + d1;
      d1 = d1 * eta + v;
      v = v * eta + cCD[n][j];
    }
    pc[j] = v; pc1[j] = d1; pc2[j] = 2.0 * d2;
  }
  const double u0 = c.mu2t[0], u1 = c.mu2t[1];
  const double t0 = r0 * u0, t1 = r1 * u1;
  // tau_k and tautau_t with their density derivatives: value, d/d rho0, d/d rho1, d00, d01, d11
  const double tau[3][6] = {{t0 * t0, 2.0 * u0 * t0, 0.0, 2.0 * u0 * u0, 0.0, 0.0},
                            {t0 * t1, u0 * t1, u1 * t0, 0.0, u0 * u1, 0.0},
                            {t1 * t1, 0.0, 2.0 * u1 * t1, 0.0, 0.0, 2.0 * u1 * u1}};
  const double ttt[4][6] = {{t0 * t0 * t0, 3.0 * u0 * t0 * t0, 0.0, 6.0 * u0 * u0 * t0, 0.0, 0.0},
                            {t0 * t0 * t1, 2.0 * u0 * t0 * t1, u1 * t0 * t0, 2.0 * u0 * u0 * t1, 2.0 * u0 * u1 * t0, 0.0},
                            {t1 * t1 * t0, u0 * t1 * t1, 2.0 * u1 * t1 * t0, 0.0, 2.0 * u0 * u1 * t1, 2.0 * u1 * u1 * t0},
                            {t1 * t1 * t1, 0.0, 3.0 * u1 * t1 * t1, 0.0, 0.0, 6.0 * u1 * u1 * t1}};
  // phi[6] = value, d0, d1, d00, d01, d11 (before the constant factors)
  double p2[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, p3[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double w1 = c.pw1[k], w2 = c.pw2[k], pe = c.pe[k], ps = c.ps[k];
    const double j = ps * ((pa[0] + w1 * pa[1] + w2 * pa[2]) + pe * (pb[0] + w1 * pb[1] + w2 * pb[2]));
    const double j1 = ps * ((pa1[0] + w1 * pa1[1] + w2 * pa1[2]) + pe * (pb1[0] + w1 * pb1[1] + w2 * pb1[2]));
    const double j2 = ps * ((pa2[0] + w1 * pa2[1] + w2 * pa2[2]) + pe * (pb2[0] + w1 * pb2[1] + w2 * pb2[2]));
    const double* t = tau[k];
    p2[0] += t[0] * j;
    p2[1] += t[1] * j + t[0] * j1 * z30;
    p2[2] += t[2] * j + t[0] * j1 * z31;
    p2[3] += t[3] * j + 2.0 * t[1] * j1 * z30 + t[0] * j2 * z30 * z30;
    p2[4] += t[4] * j + (t[1] * z31 + t[2] * z30) * j1 + t[0] * j2 * z30 * z31;
    p2[5] += t[5] * j + 2.0 * t[2] * j1 * z31 + t[0] * j2 * z31 * z31;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double w1 = c.tw1[k], w2 = c.tw2[k], ts = c.ts[k];
    const double j = ts * (pc[0] + w1 * pc[1] + w2 * pc[2]);
    const double j1 = ts * (pc1[0] + w1 * pc1[1] + w2 * pc1[2]);
    const double j2 = ts * (pc2[0] + w1 * pc2[1] + w2 * pc2[2]);
    const double* t = ttt[k];
    p3[0] += t[0] * j;
    p3[1] += t[1] * j + t[0] * j1 * z30;
    p3[2] += t[2] * j + t[0] * j1 * z31;
    p3[3] += t[3] * j + 2.0 * t[1] * j1 * z30 + t[0] * j2 * z30 * z30;
    p3[4] += t[4] * j + (t[1] * z31 + t[2] * z30) * j1 + t[0] * j2 * z30 * z31;
    p3[5] += t[5] * j + 2.0 * t[2] * j1 * z31 + t[0] * j2 * z31 * z31;
  }
  const double k2 = -kPi, k3 = -4.0 / 3.0 * kPi * kPi;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    p2[q] *= k2;
    p3[q] *= k3;
  }
  if (p2[0] == 0.0) return;  // no dipole-dipole pair at this composition: the term and its derivatives vanish
  const double D = p2[0] - p3[0], Di = 1.0 / D, Di2 = Di * Di, Di3 = Di2 * Di;
  const double f2 = p2[0] * (p2[0] - 2.0 * p3[0]) * Di2, f3 = p2[0] * p2[0] * Di2;
  const double f22 = 2.0 * p3[0] * p3[0] * Di3, f23 = -2.0 * p2[0] * p3[0] * Di3, f33 = 2.0 * p2[0] * p2[0] * Di3;
  o.a += p2[0] * p2[0] * Di;
  o.g0 += f2 * p2[1] + f3 * p3[1];
  o.g1 += f2 * p2[2] + f3 * p3[2];
  o.h00 += f22 * p2[1] * p2[1] + 2.0 * f23 * p2[1] * p3[1] + f33 * p3[1] * p3[1] + f2 * p2[3] + f3 * p3[3];
  o.h01 += f22 * p2[1] * p2[2] + f23 * (p2[1] * p3[2] + p3[1] * p2[2]) + f33 * p3[1] * p3[2] + f2 * p2[4] + f3 * p3[4];
  o.h11 += f22 * p2[2] * p2[2] + 2.0 * f23 * p2[2] * p3[2] + f33 * p3[2] * p3[2] + f2 * p2[5] + f3 * p3[5];
}

// ---- the dipole term with per-STATE polynomial coefficients ------------------------------------------------------
// J2_k(eta) and J3_t(eta) are polynomials in eta whose coefficients depend on the state only through the pair /
// triplet weights (pw1, pw2, pe, ps; tw1, tw2, ts), so they are combined ONCE per state:
//   phi2 = t0^2 A_0(eta) + t0 t1 A_1(eta) + t1^2 A_2(eta),      A_k = sum_n c2[k][n] eta^n   (degree 4, -pi ps_k folded in)
//   phi3 = t0^3 B_0 + t0^2 t1 B_1 + t1^2 t0 B_2 + t1^3 B_3,      B_t = sum_n c3[t][n] eta^n   (degree 3, -4/3 pi^2 ts_t folded in)
// and the evaluation differentiates a = phi2^2 / (phi2 - phi3) in the three variables (t0, t1, eta), which are LINEAR in
// the densities (t_i = rho_i mu2t_i, eta = rho . z3), and maps gradient and Hessian to (rho_0, rho_1) at the end.
// 33 doubles per state instead of 26, but ~215 FP64 instructions per evaluation instead of ~600 and a third of the
// code: the bubble / dew loop of a dipolar GC pair then fits the SM's 32 KB instruction cache.
struct DipolePoly {
  double mu2t[2];
  double c2[3][5];
  double c3[4][4];
};

template <class C>
FD void dipole_poly_from(DipolePoly& d, const C& c) {
  d.mu2t[0] = c.mu2t[0];
  d.mu2t[1] = c.mu2t[1];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double w1 = c.pw1[k], w2 = c.pw2[k], pe = c.pe[k], f = -kPi * c.ps[k];
#pragma unroll
    for (int n = 0; n < 5; ++n) {
      double v = cAD[n][0] + w1 * cAD[n][1] + w2 * cAD[n][2];
      if (n < 3) v += pe * (cBD[n][0] + w1 * cBD[n][1] + w2 * cBD[n][2]);
      d.c2[k][n] = f * v;
    }
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const double w1 = c.tw1[t], w2 = c.tw2[t], f = (-4.0 / 3.0 * kPi * kPi) * c.ts[t];
#pragma unroll
    for (int n = 0; n < 4; ++n) d.c3[t][n] = f * (cCD[n][0] + w1 * cCD[n][1] + w2 * cCD[n][2]);
  }
}

// value, first and second derivative of sum_n k[n] x^n, n <= DEG
template <int DEG>
FD void horner_2(const double* k, double x, double& v, double& d1, double& d2) {
  v = k[DEG];
  d1 = 0.0;
  d2 = 0.0;
#pragma unroll
  for (int n = DEG - 1; n >= 0; --n) {
    d2 = d2 * x + d1;
    d1 = d1 * x + v;
    v = v * x + k[n];
  }
  d2 *= 2.0;
}

// adds the dipole term (value, gradient, Hessian in the partial densities) to `o`; z30, z31 = d eta / d rho_i.
// The three coefficient groups may live in different places (GC: the triplet polynomials c3 do not depend on the
// state at all and are read from the per-pair table).
FD void dipole_poly_eval(const double* mu2t, const double (*c2)[5], const double (*c3)[4], double z30, double z31,
                         double r0, double r1, BaseH2& o);
FD void dipole_poly_eval(const DipolePoly& d, double z30, double z31, double r0, double r1, BaseH2& o) {
  dipole_poly_eval(d.mu2t, d.c2, d.c3, z30, z31, r0, r1, o);
}
FD void dipole_poly_eval(const double* mu2t, const double (*c2)[5], const double (*c3)[4], double z30, double z31,
                         double r0, double r1, BaseH2& o) {
  const double u0 = mu2t[0], u1 = mu2t[1];
  const double x = r0 * u0, y = r1 * u1, eta = r0 * z30 + r1 * z31;
  const double xx = x * x, xy = x * y, yy = y * y;
  double A0, A0e, A0ee, A1, A1e, A1ee, A2, A2e, A2ee;
  horner_2<4>(c2[0], eta, A0, A0e, A0ee);
  horner_2<4>(c2[1], eta, A1, A1e, A1ee);
  horner_2<4>(c2[2], eta, A2, A2e, A2ee);
  const double p2 = xx * A0 + xy * A1 + yy * A2;
  if (p2 == 0.0) return;  // no dipole-dipole pair at this composition: the term and its derivatives vanish
  const double p2x = 2.0 * x * A0 + y * A1, p2y = x * A1 + 2.0 * y * A2, p2e = xx * A0e + xy * A1e + yy * A2e;
  const double p2xx = 2.0 * A0, p2xy = A1, p2yy = 2.0 * A2;
  const double p2xe = 2.0 * x * A0e + y * A1e, p2ye = x * A1e + 2.0 * y * A2e, p2ee = xx * A0ee + xy * A1ee + yy * A2ee;
  double B0, B0e, B0ee, B1, B1e, B1ee, B2, B2e, B2ee, B3, B3e, B3ee;
  horner_2<3>(c3[0], eta, B0, B0e, B0ee);
  horner_2<3>(c3[1], eta, B1, B1e, B1ee);
  horner_2<3>(c3[2], eta, B2, B2e, B2ee);
  horner_2<3>(c3[3], eta, B3, B3e, B3ee);
  const double xxx = xx * x, xxy = xx * y, yyx = yy * x, yyy = yy * y;
  const double p3 = xxx * B0 + xxy * B1 + yyx * B2 + yyy * B3;
  const double p3x = 3.0 * xx * B0 + 2.0 * xy * B1 + yy * B2, p3y = xx * B1 + 2.0 * xy * B2 + 3.0 * yy * B3;
  const double p3e = xxx * B0e + xxy * B1e + yyx * B2e + yyy * B3e;
  const double p3xx = 6.0 * x * B0 + 2.0 * y * B1, p3xy = 2.0 * (x * B1 + y * B2), p3yy = 2.0 * x * B2 + 6.0 * y * B3;
  const double p3xe = 3.0 * xx * B0e + 2.0 * xy * B1e + yy * B2e, p3ye = xx * B1e + 2.0 * xy * B2e + 3.0 * yy * B3e;
  const double p3ee = xxx * B0ee + xxy * B1ee + yyx * B2ee + yyy * B3ee;
  // f = p2^2 / (p2 - p3)
  const double Di = 1.0 / (p2 - p3), Di2 = Di * Di, Di3 = Di2 * Di;
  const double f2 = p2 * (p2 - 2.0 * p3) * Di2, f3 = p2 * p2 * Di2;
  const double f22 = 2.0 * p3 * p3 * Di3, f23 = -2.0 * p2 * p3 * Di3, f33 = 2.0 * p2 * p2 * Di3;
  const double gx = f2 * p2x + f3 * p3x, gy = f2 * p2y + f3 * p3y, ge = f2 * p2e + f3 * p3e;
  const double ax = f22 * p2x + f23 * p3x, ay = f22 * p2y + f23 * p3y, ae = f22 * p2e + f23 * p3e;
  const double bx = f23 * p2x + f33 * p3x, by = f23 * p2y + f33 * p3y, be = f23 * p2e + f33 * p3e;
  const double Hxx = ax * p2x + bx * p3x + f2 * p2xx + f3 * p3xx;
  const double Hxy = ax * p2y + bx * p3y + f2 * p2xy + f3 * p3xy;
  const double Hyy = ay * p2y + by * p3y + f2 * p2yy + f3 * p3yy;
  const double Hxe = ax * p2e + bx * p3e + f2 * p2xe + f3 * p3xe;
  const double Hye = ay * p2e + by * p3e + f2 * p2ye + f3 * p3ye;
  const double Hee = ae * p2e + be * p3e + f2 * p2ee + f3 * p3ee;
  o.a += p2 * p2 * Di;
  o.g0 += u0 * gx + z30 * ge;
  o.g1 += u1 * gy + z31 * ge;
  o.h00 += u0 * (u0 * Hxx + 2.0 * z30 * Hxe) + z30 * z30 * Hee;
  o.h01 += u0 * (u1 * Hxy + z31 * Hxe) + z30 * (u1 * Hye + z31 * Hee);
  o.h11 += u1 * (u1 * Hyy + 2.0 * z31 * Hye) + z31 * z31 * Hee;
}

// The evaluation the solvers and state kernels use: closed-form base and dipole parts; the association
// term (only where present) in hyper-dual arithmetic on top.
template <int NB, bool GC>
FD MixEval mix_eval(const MixCoef<double, NB, GC>& c, double rho0, double rho1) {
  BaseH2 b = mix_base_eval<NB, GC>(c, rho0, rho1);
  if (c.dipolar) mix_dipole_eval(c, rho0, rho1, b);
  H2<double> a;
  a.re = b.a; a.g0 = b.g0; a.g1 = b.g1; a.h00 = b.h00; a.h01 = b.h01; a.h11 = b.h11;
  if (c.cat != kAssocNone) {
    H2<double> r0(rho0), r1(rho1);
    r0.g0 = 1.0;
    r1.g1 = 1.0;
    a = a + mix_helmholtz<double, H2<double>, NB, GC, 4>(c, r0, r1);
  }
  return mix_eval_from(a, rho0, rho1);
}

// plain pairs (MixCoefBase): the base part is the whole evaluation
FD MixEval mix_eval(const MixCoefBase& c, double rho0, double rho1) {
  BaseH2 b = mix_base_eval_impl<2>(c, rho0, rho1);
  H2<double> a;
  a.re = b.a; a.g0 = b.g0; a.g1 = b.g1; a.h00 = b.h00; a.h01 = b.h01; a.h11 = b.h11;
  return mix_eval_from(a, rho0, rho1);
}

}  // namespace feos
