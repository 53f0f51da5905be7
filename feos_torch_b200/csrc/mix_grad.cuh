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

// ---- closed-form coefficient adjoints of the base group (hard sphere + hard chain + dispersion) ----
// a_base depends on the coefficients only through
//   zeta_k = sum_i z[k][i] rho_i,  mbar = sum_i m[i] rho_i / rho,  Q1 = sum q1 rho rho,  Q2,  cd[b],
// so  d a / d coef  is 4 + 1 + 2 partial derivatives times the (known) densities plus one partial per bond
// type.  BasePhase<R> holds those partials at one density point in the number type R of the densities:
// R = double for the incipient phase, R = D1T<double> (density + direction) for the specified phase, whose
// .d1 part is then the directional density derivative of the coefficient gradient.  The bubble / dew
// pressure adjoint of a coefficient combines the two phases (point_comb).  Replaces 16 + nbond forward-mode
// passes over the whole Helmholtz energy (one per coefficient).
// Everything is scalar state in registers (no locally indexed arrays): the per-bond terms are consumed
// as they are produced.
template <class R>
struct BasePhase {
  R rho0, rho1;
  R z0, z1, z2, z3, om, omi, omi2, omi3, cz, dcz3;
  R a0, a1, a2, a3;  // d a / d zeta_k
  R mb;              // d a / d mbar / rho
  R q1, q2;          // d a / d Q1, d a / d Q2
};

template <class R>
FD void horner7(const double (&k)[7], const R& x, R& v, R& dv) {
  v = R(k[6]);
  dv = R(0.0);
#pragma unroll
  for (int n = 5; n >= 0; --n) {
    dv = dv * x + v;
    v = v * x + k[n];
  }
}

// hard-sphere partials; sets up the packing-fraction helpers the other terms use
template <class R, class C>
FD void base_phase_begin(BasePhase<R>& s, const C& c, const R& rho0, const R& rho1) {
  s.rho0 = rho0;
  s.rho1 = rho1;
  s.z0 = rho0 * c.z[0][0] + rho1 * c.z[0][1];
  s.z1 = rho0 * c.z[1][0] + rho1 * c.z[1][1];
  s.z2 = rho0 * c.z[2][0] + rho1 * c.z[2][1];
  s.z3 = rho0 * c.z[3][0] + rho1 * c.z[3][1];
  s.om = 1.0 - s.z3;
  s.omi = ad_recip(s.om);
  s.omi2 = s.omi * s.omi;
  s.omi3 = s.omi2 * s.omi;
  R iz3 = ad_recip(s.z3);
  R L = ad_log(s.om);
  // a_hs = 6/pi [3 z1 z2 / om + z2^3 / (z3 om^2) + (z2^3 / z3^2 - z0) ln om]
  R z2sq = s.z2 * s.z2;
  R t = (z2sq * s.z2) * (iz3 * iz3);  // z2^3 / z3^2
  s.a0 = (-6.0 / kPi) * L;
  s.a1 = (18.0 / kPi) * (s.z2 * s.omi);
  s.a2 = (18.0 / kPi) * (s.z1 * s.omi + (z2sq * iz3) * s.omi2 + (z2sq * (iz3 * iz3)) * L);
  s.a3 = (6.0 / kPi) * ((3.0 * (s.z1 * s.z2)) * s.omi2 + (t * s.z3) * (2.0 * s.omi3 - iz3 * s.omi2) - (2.0 * (t * iz3)) * L -
                        (t - s.z0) * s.omi);
  s.cz = s.z2 * s.omi2;
  s.dcz3 = 2.0 * (s.z2 * s.omi3);  // d cz / d zeta3
}

// hard chain, one bond type: a = - B ln g,  g = 1/om + 3 x + 2 x^2 om,  x = z2 cd / om^2.
// Returns d a / d cd and adds the zeta_2 / zeta_3 partials.
template <class R>
FD R base_phase_bond(BasePhase<R>& s, double cd, double bn0, double bn1, R* lng = nullptr) {
  R x = s.cz * cd;
  R x2 = x * x;
  R gg = s.omi + 3.0 * x + (2.0 * x2) * s.om;
  if (lng) *lng = ad_log(gg);
  R gx = 3.0 + 4.0 * (x * s.om);
  R w = (s.rho0 * bn0 + s.rho1 * bn1) * ad_recip(gg);  // B / g
  R wgx = w * gx;
  s.a2 = s.a2 - wgx * (s.omi2 * cd);
  s.a3 = s.a3 - w * (s.omi2 + gx * (s.dcz3 * cd) - 2.0 * x2);
  return -(wgx * s.cz);
}

// dispersion: a = -pi [2 Q1 I1(eta, mbar) + Q2 C1(eta, mbar) I2(eta, mbar) mbar]
template <class R, class C>
FD void base_phase_end(BasePhase<R>& s, const C& c) {
  const R& z3 = s.z3;
  R rti = ad_recip(s.rho0 + s.rho1);
  R mbar = (s.rho0 * c.m[0] + s.rho1 * c.m[1]) * rti;
  R r00 = s.rho0 * s.rho0, r01 = s.rho0 * s.rho1, r11 = s.rho1 * s.rho1;
  R Q1 = r00 * c.q1[0] + (2.0 * r01) * c.q1[1] + r11 * c.q1[2];
  R Q2 = r00 * c.q2[0] + (2.0 * r01) * c.q2[1] + r11 * c.q2[2];
  R mi = ad_recip(mbar);
  R w1 = 1.0 - mi;
  R w2 = w1 * (1.0 - 2.0 * mi);
  R dw1 = mi * mi;
  R dw2 = dw1 * (3.0 - 4.0 * mi);
  R s0, s1, s2, d0, d1, d2;
  horner7(cA0, z3, s0, d0);
  horner7(cA1, z3, s1, d1);
  horner7(cA2, z3, s2, d2);
  R i1 = s0 + w1 * s1 + w2 * s2;
  R i1e = d0 + w1 * d1 + w2 * d2;
  R i1m = dw1 * s1 + dw2 * s2;
  horner7(cB0, z3, s0, d0);
  horner7(cB1, z3, s1, d1);
  horner7(cB2, z3, s2, d2);
  R i2 = s0 + w1 * s1 + w2 * s2;
  R i2e = d0 + w1 * d1 + w2 * d2;
  R i2m = dw1 * s1 + dw2 * s2;
  R e2 = z3 * z3, e3 = e2 * z3;
  R tmi = ad_recip(2.0 - z3);
  R omi4 = s.omi2 * s.omi2;
  R xn = 8.0 * z3 - 2.0 * e2;
  R X = xn * omi4;
  R dX = (8.0 - 4.0 * z3) * omi4 + (4.0 * xn) * (omi4 * s.omi);
  R yn = 20.0 * z3 - 27.0 * e2 + 12.0 * e3 - 2.0 * (e2 * e2);
  R ow = s.omi2 * (tmi * tmi);
  R Y = yn * ow;
  R dY = (20.0 - 54.0 * z3 + 36.0 * e2 - 8.0 * e3) * ow + (2.0 * yn) * (ow * (s.omi + tmi));
  R c1 = ad_recip(1.0 + mbar * X + (1.0 - mbar) * Y);
  R c1sq = c1 * c1;
  R c1e = -(c1sq * (mbar * dX + (1.0 - mbar) * dY));
  R c1m = -(c1sq * (X - Y));
  s.q1 = (-2.0 * kPi) * i1;
  s.q2 = (-kPi) * ((c1 * i2) * mbar);
  s.a3 = s.a3 + (-kPi) * (2.0 * (Q1 * i1e) + (Q2 * mbar) * (c1e * i2 + c1 * i2e));
  s.mb = ((-kPi) * (2.0 * (Q1 * i1m) + Q2 * ((c1m * i2 + c1 * i2m) * mbar + c1 * i2))) * rti;
}

FD double point_comb(double gi, const D1& gs, const PointWeights& q) {
  // - d N / d coef / denom with N = a_inc / rho_inc - v a_spec + D_dir a_spec + cst
  return -(gi / q.rho_inc - gs.re * q.v + gs.d1) / q.denom;
}

// cbar[idx] = d p_red / d coef[idx] for the base coefficients in the flat order of FEOS_VISIT_COEF:
// z[k][i] -> 2 k + i, m[i] -> 8 + i, q1[k] -> 10 + k, q2[k] -> 13 + k, cd[b] -> 16 + b.
// bnbar (optional, [2 NB]): adjoints of the bond numbers bn[b][i] (a_hc = - sum_b (rho . bn_b) ln g_b); they are
// parameters (m_i - 1) for the homosegmented model and constants (counts) for the group-contribution model.
template <int NB, bool GC>
FD_NOINLINE void mix_base_coef_adjoint(const MixCoef<double, NB, GC>& c, const double (&spec)[2], const double (&inc)[2],
                                       const PointWeights& q, double* cbar, double* bnbar = nullptr) {
  BasePhase<double> pi;
  BasePhase<D1> ps;
  const double i0 = inc[0], i1 = inc[1];
  const D1 s0(spec[0], q.dir[0]), s1(spec[1], q.dir[1]);
  base_phase_begin(pi, c, i0, i1);
  base_phase_begin(ps, c, s0, s1);
  const int nb = (NB == 2) ? 2 : c.nbond;
#pragma unroll 1
  for (int b = 0; b < nb; ++b) {
    double cd = c.cd[b], bn0 = c.bn[b][0], bn1 = c.bn[b][1];
    if (bnbar) {
      double li;
      D1 ls;
      double gi = base_phase_bond(pi, cd, bn0, bn1, &li);
      D1 gs = base_phase_bond(ps, cd, bn0, bn1, &ls);
      cbar[16 + b] = point_comb(gi, gs, q);
      bnbar[2 * b] = point_comb(-(li * i0), -(ls * s0), q);
      bnbar[2 * b + 1] = point_comb(-(li * i1), -(ls * s1), q);
    } else {
      double gi = base_phase_bond(pi, cd, bn0, bn1);
      D1 gs = base_phase_bond(ps, cd, bn0, bn1);
      cbar[16 + b] = point_comb(gi, gs, q);
    }
  }
  base_phase_end(pi, c);
  base_phase_end(ps, c);
  cbar[0] = point_comb(pi.a0 * i0, ps.a0 * s0, q);
  cbar[1] = point_comb(pi.a0 * i1, ps.a0 * s1, q);
  cbar[2] = point_comb(pi.a1 * i0, ps.a1 * s0, q);
  cbar[3] = point_comb(pi.a1 * i1, ps.a1 * s1, q);
  cbar[4] = point_comb(pi.a2 * i0, ps.a2 * s0, q);
  cbar[5] = point_comb(pi.a2 * i1, ps.a2 * s1, q);
  cbar[6] = point_comb(pi.a3 * i0, ps.a3 * s0, q);
  cbar[7] = point_comb(pi.a3 * i1, ps.a3 * s1, q);
  cbar[8] = point_comb(pi.mb * i0, ps.mb * s0, q);
  cbar[9] = point_comb(pi.mb * i1, ps.mb * s1, q);
  const D1 r00 = s0 * s0, r01 = 2.0 * (s0 * s1), r11 = s1 * s1;
  cbar[10] = point_comb(pi.q1 * (i0 * i0), ps.q1 * r00, q);
  cbar[11] = point_comb(pi.q1 * (2.0 * i0 * i1), ps.q1 * r01, q);
  cbar[12] = point_comb(pi.q1 * (i1 * i1), ps.q1 * r11, q);
  cbar[13] = point_comb(pi.q2 * (i0 * i0), ps.q2 * r00, q);
  cbar[14] = point_comb(pi.q2 * (2.0 * i0 * i1), ps.q2 * r01, q);
  cbar[15] = point_comb(pi.q2 * (i1 * i1), ps.q2 * r11, q);
}

// ---- k_ij alone --------------------------------------------------------------------------------------------
// A k_ij fit (BASELINE config C4) asks for d p / d k_ij only.  k_ij enters the Helmholtz energy through the cross
// dispersion coefficients q1[01], q2[01] alone (pcsaft_mix.py:69-90: eps_ij (1 - k_ij)) -- for EVERY pair, polar and
// associating ones included -- so the gradient needs the two partials
//     d a / d Q1 = -2 pi I1(eta, mbar),   d a / d Q2 = -pi mbar C1(eta, mbar) I2(eta, mbar)
// at the incipient densities (double) and with the density direction of the specified phase (D1), instead of the
// whole coefficient adjoint + reverse sweep.
template <class R, class C>
FD void disp_q_partials(const C& c, const R& rho0, const R& rho1, R& pq1, R& pq2) {
  R z3 = rho0 * c.z[3][0] + rho1 * c.z[3][1];
  R mbar = (rho0 * c.m[0] + rho1 * c.m[1]) * ad_recip(rho0 + rho1);
  R mi = ad_recip(mbar);
  R w1 = 1.0 - mi;
  R w2 = w1 * (1.0 - 2.0 * mi);
  R s0, s1, s2, d0, d1, d2;
  horner7(cA0, z3, s0, d0);
  horner7(cA1, z3, s1, d1);
  horner7(cA2, z3, s2, d2);
  pq1 = (-2.0 * kPi) * (s0 + w1 * s1 + w2 * s2);
  horner7(cB0, z3, s0, d0);
  horner7(cB1, z3, s1, d1);
  horner7(cB2, z3, s2, d2);
  R i2 = s0 + w1 * s1 + w2 * s2;
  R e2 = z3 * z3, e3 = e2 * z3;
  R omi = ad_recip(1.0 - z3), tmi = ad_recip(2.0 - z3);
  R omi2 = omi * omi;
  R X = (8.0 * z3 - 2.0 * e2) * (omi2 * omi2);
  R Y = (20.0 * z3 - 27.0 * e2 + 12.0 * e3 - 2.0 * (e2 * e2)) * (omi2 * (tmi * tmi));
  R c1 = ad_recip(1.0 + mbar * X + (1.0 - mbar) * Y);
  pq2 = (-kPi) * ((c1 * i2) * mbar);
}

// cb1 = d p_red / d q1[01], cb2 = d p_red / d q2[01] of the Newton-step pressure (the two coefficient adjoints a
// k_ij or k_ab gradient needs); `q` = point_weights of the specified phase's full evaluation
template <class C>
FD void point_q_adjoints(const C& c, const double (&spec)[2], const double (&inc)[2], const PointWeights& q, double& cb1,
                         double& cb2) {
  double iq1, iq2;
  D1 sq1, sq2;
  disp_q_partials<double>(c, inc[0], inc[1], iq1, iq2);
  const D1 s0(spec[0], q.dir[0]), s1(spec[1], q.dir[1]);
  disp_q_partials<D1>(c, s0, s1, sq1, sq2);
  const D1 r01 = 2.0 * (s0 * s1);
  const double i01 = 2.0 * inc[0] * inc[1];
  cb1 = point_comb(iq1 * i01, sq1 * r01, q);
  cb2 = point_comb(iq2 * i01, sq2 * r01, q);
}

// g * d(p_bubble/dew [Pa]) / d k_ij of any binary
template <class C>
FD double mix_point_grad_kij(const double (&prm)[16], double kij0, double T, const C& c, const double (&spec)[2],
                             const double (&inc)[2], const PointWeights& q, double g) {
  double cb1, cb2;
  point_q_adjoints(c, spec, inc, q, cb1, cb2);
  // q1[01] = m0 m1 sqrt(eps0 eps1)/T (1 - k) s01^3, q2[01] = q1[01] sqrt(eps0 eps1)/T (1 - k)
  const double ti = 1.0 / T;
  const double s01 = 0.5 * (prm[1] + prm[9]);
  const double e = sqrt(prm[2] * prm[10]) * ti;
  const double base = prm[0] * prm[8] * e * (s01 * s01 * s01);  // - d q1[01] / d k
  return -(base * (cb1 + cb2 * (2.0 * e * (1.0 - kij0)))) * (g * T * kPaPerRed);
}

// ---- closed-form coefficient adjoints of the Gross-Vrabec dipole term -------------------------------
//   a_dip = phi2^2 / (phi2 - phi3),
//   phi2 = -pi sum_k tt_k J2_k ps_k,            J2_k = PA0 + pw1_k PA1 + pw2_k PA2 + pe_k (PB0 + pw1_k PB1 + pw2_k PB2)
//   phi3 = -4/3 pi^2 sum_t ttt_t J3_t ts_t,     J3_t = PC0 + tw1_t PC1 + tw2_t PC2
// with tt = (t0^2, t0 t1, t1^2), ttt = (t0^3, t0^2 t1, t1^2 t0, t1^3), t_i = rho_i mu2t_i and the
// polynomials PA_j, PB_j, PC_j in eta = zeta_3.  Every one of the 26 coefficients enters linearly or
// through these few products, so the gradient costs about as much as the term itself.
template <class R>
struct DipolePhase {
  R pa[3], pb[3], pc[3], dpa[3], dpb[3], dpc[3];  // polynomials in eta and their eta-derivatives
  R t0, t1, tt[3], ttt[4];
  R j2[3], j3[4];
  R w2, w3;  // d a / d phi2 * (-pi),  d a / d phi3 * (-4/3 pi^2)
  R deta;    // d a_dip / d eta
};

template <class R, class C>
FD void dipole_phase(DipolePhase<R>& d, const C& c, const R& rho0, const R& rho1, const R& eta) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    // Horner with derivative; orders 4 (AD), 2 (BD), 3 (CD)
    R v(cAD[4][j]), dv(0.0);
#pragma unroll
    for (int n = 3; n >= 0; --n) {
      dv = dv * eta + v;
      v = v * eta + cAD[n][j];
    }
    d.pa[j] = v;
    d.dpa[j] = dv;
    v = R(cBD[2][j]);
    dv = R(0.0);
#pragma unroll
    for (int n = 1; n >= 0; --n) {
      dv = dv * eta + v;
      v = v * eta + cBD[n][j];
    }
    d.pb[j] = v;
    d.dpb[j] = dv;
    v = R(cCD[3][j]);
    dv = R(0.0);
#pragma unroll
    for (int n = 2; n >= 0; --n) {
      dv = dv * eta + v;
      v = v * eta + cCD[n][j];
    }
    d.pc[j] = v;
    d.dpc[j] = dv;
  }
  d.t0 = rho0 * c.mu2t[0];
  d.t1 = rho1 * c.mu2t[1];
  d.tt[0] = d.t0 * d.t0;
  d.tt[1] = d.t0 * d.t1;
  d.tt[2] = d.t1 * d.t1;
  d.ttt[0] = d.tt[0] * d.t0;
  d.ttt[1] = d.tt[0] * d.t1;
  d.ttt[2] = d.tt[2] * d.t0;
  d.ttt[3] = d.tt[2] * d.t1;
  R phi2(0.0), phi3(0.0), e2(0.0), e3(0.0);  // e2, e3: d phi / d eta (before the constant factors)
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    d.j2[k] = (d.pa[0] + d.pa[1] * c.pw1[k] + d.pa[2] * c.pw2[k]) + (d.pb[0] + d.pb[1] * c.pw1[k] + d.pb[2] * c.pw2[k]) * c.pe[k];
    R dj = (d.dpa[0] + d.dpa[1] * c.pw1[k] + d.dpa[2] * c.pw2[k]) + (d.dpb[0] + d.dpb[1] * c.pw1[k] + d.dpb[2] * c.pw2[k]) * c.pe[k];
    R w = d.tt[k] * c.ps[k];
    phi2 = phi2 + w * d.j2[k];
    e2 = e2 + w * dj;
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    d.j3[t] = d.pc[0] + d.pc[1] * c.tw1[t] + d.pc[2] * c.tw2[t];
    R dj = d.dpc[0] + d.dpc[1] * c.tw1[t] + d.dpc[2] * c.tw2[t];
    R w = d.ttt[t] * c.ts[t];
    phi3 = phi3 + w * d.j3[t];
    e3 = e3 + w * dj;
  }
  phi2 = phi2 * (-kPi);
  phi3 = phi3 * (-4.0 / 3.0 * kPi * kPi);
  R idf = ad_recip(phi2 - phi3);
  R r = phi2 * idf;
  d.w2 = (r * (phi2 - 2.0 * phi3) * idf) * (-kPi);
  d.w3 = (r * r) * (-4.0 / 3.0 * kPi * kPi);
  d.deta = d.w2 * e2 + d.w3 * e3;
}

// cbar[first .. first + 25] = d p_red / d (dipole coefficients) in the order of FEOS_VISIT_COEF (mu2t[2]; pw1, pw2,
// pe, ps per pair; tw1, tw2, ts per triplet); ADDS the dipole term's share to the z[3][i] adjoints cbar[6], cbar[7].
template <int NB, bool GC>
FD_NOINLINE void mix_dipole_coef_adjoint(const MixCoef<double, NB, GC>& c, const double (&spec)[2], const double (&inc)[2],
                                         const PointWeights& q, double* cbar, int first) {
  const double i0 = inc[0], i1 = inc[1];
  const D1 s0(spec[0], q.dir[0]), s1(spec[1], q.dir[1]);
  DipolePhase<double> di;
  DipolePhase<D1> ds;
  dipole_phase(di, c, i0, i1, i0 * c.z[3][0] + i1 * c.z[3][1]);
  dipole_phase(ds, c, s0, s1, s0 * c.z[3][0] + s1 * c.z[3][1]);
  cbar[6] += point_comb(di.deta * i0, ds.deta * s0, q);
  cbar[7] += point_comb(di.deta * i1, ds.deta * s1, q);
  // mu2t_i through t_i = rho_i mu2t_i
  {
    double a2 = (2.0 * di.t0) * (c.ps[0] * di.j2[0]) + di.t1 * (c.ps[1] * di.j2[1]);
    double a3 = (3.0 * di.tt[0]) * (c.ts[0] * di.j3[0]) + (2.0 * di.tt[1]) * (c.ts[1] * di.j3[1]) + di.tt[2] * (c.ts[2] * di.j3[2]);
    D1 b2 = (2.0 * ds.t0) * (ds.j2[0] * c.ps[0]) + ds.t1 * (ds.j2[1] * c.ps[1]);
    D1 b3 = (3.0 * ds.tt[0]) * (ds.j3[0] * c.ts[0]) + (2.0 * ds.tt[1]) * (ds.j3[1] * c.ts[1]) + ds.tt[2] * (ds.j3[2] * c.ts[2]);
    cbar[first] = point_comb((di.w2 * a2 + di.w3 * a3) * i0, (ds.w2 * b2 + ds.w3 * b3) * s0, q);
    a2 = di.t0 * (c.ps[1] * di.j2[1]) + (2.0 * di.t1) * (c.ps[2] * di.j2[2]);
    a3 = di.tt[0] * (c.ts[1] * di.j3[1]) + (2.0 * di.tt[1]) * (c.ts[2] * di.j3[2]) + (3.0 * di.tt[2]) * (c.ts[3] * di.j3[3]);
    b2 = ds.t0 * (ds.j2[1] * c.ps[1]) + (2.0 * ds.t1) * (ds.j2[2] * c.ps[2]);
    b3 = ds.tt[0] * (ds.j3[1] * c.ts[1]) + (2.0 * ds.tt[1]) * (ds.j3[2] * c.ts[2]) + (3.0 * ds.tt[2]) * (ds.j3[3] * c.ts[3]);
    cbar[first + 1] = point_comb((di.w2 * a2 + di.w3 * a3) * i1, (ds.w2 * b2 + ds.w3 * b3) * s1, q);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double wi = di.w2 * di.tt[k];
    D1 ws = ds.w2 * ds.tt[k];
    double wpi = wi * c.ps[k];
    D1 wps = ws * c.ps[k];
    double* o = cbar + first + 2 + 4 * k;
    o[0] = point_comb(wpi * (di.pa[1] + di.pb[1] * c.pe[k]), wps * (ds.pa[1] + ds.pb[1] * c.pe[k]), q);
    o[1] = point_comb(wpi * (di.pa[2] + di.pb[2] * c.pe[k]), wps * (ds.pa[2] + ds.pb[2] * c.pe[k]), q);
    o[2] = point_comb(wpi * (di.pb[0] + di.pb[1] * c.pw1[k] + di.pb[2] * c.pw2[k]),
                      wps * (ds.pb[0] + ds.pb[1] * c.pw1[k] + ds.pb[2] * c.pw2[k]), q);
    o[3] = point_comb(wi * di.j2[k], ws * ds.j2[k], q);
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    double wi = di.w3 * di.ttt[t];
    D1 ws = ds.w3 * ds.ttt[t];
    double* o = cbar + first + 14 + 3 * t;
    o[0] = point_comb((wi * c.ts[t]) * di.pc[1], (ws * c.ts[t]) * ds.pc[1], q);
    o[1] = point_comb((wi * c.ts[t]) * di.pc[2], (ws * c.ts[t]) * ds.pc[2], q);
    o[2] = point_comb(wi * di.j3[t], ws * ds.j3[t], q);
  }
}

// ---- reverse sweeps through the coefficient stage -------------------------------------------------------
// Adjoint of dipole_coef (pcsaft_mix.cuh): cb = adjoints of the 26 dipole coefficients in the flat order
// (mu2t[2]; pw1, pw2, pe, ps per pair; tw1, tw2, ts per triplet).  ADDS the adjoints of the (clamped) segment
// numbers mc_i, the molecule diameters sg_i and energies ep_i, and of T.
template <class C>
FD void dipole_coef_adjoint(const C& c, const double (&mc)[2], const double (&sg)[2], const double (&ep)[2], double ti,
                            const double* cb, double (&mcb)[2], double (&sgb)[2], double (&epb)[2], double& gT) {
  const int pi_[3] = {0, 0, 1}, pj_[3] = {0, 1, 1};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int i = pi_[k], j = pj_[k];
    const double* o = cb + 2 + 4 * k;  // pw1, pw2, pe, ps
    double x = sqrt(mc[i] * mc[j]), xi = 1.0 / x;
    double xb = (o[0] + o[1] * (3.0 - 4.0 * xi)) * xi * xi;
    mcb[i] += xb * x / (2.0 * mc[i]);
    mcb[j] += xb * x / (2.0 * mc[j]);
    double pe = c.pe[k];
    epb[i] += o[2] * pe / (2.0 * ep[i]);
    epb[j] += o[2] * pe / (2.0 * ep[j]);
    gT -= o[2] * pe * ti;
    double sij = 0.5 * (sg[i] + sg[j]);
    double sb = -1.5 * o[3] * c.ps[k] / sij;  // (d ps / d sij) / 2
    sgb[i] += sb;
    sgb[j] += sb;
  }
  const int ti_[4] = {0, 0, 0, 1}, tj_[4] = {0, 0, 1, 1}, tk_[4] = {0, 1, 1, 1};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int i = ti_[t], j = tj_[t], k = tk_[t];
    const double* o = cb + 14 + 3 * t;  // tw1, tw2, ts
    double y = cbrt(mc[i] * mc[j] * mc[k]), yi = 1.0 / y;
    double yb = (o[0] + o[1] * (3.0 - 4.0 * yi)) * yi * yi * y / 3.0;
    mcb[i] += yb / mc[i];
    mcb[j] += yb / mc[j];
    mcb[k] += yb / mc[k];
    double tb = -0.5 * o[2] * c.ts[t];
    double sij = 0.5 * (sg[i] + sg[j]), sik = 0.5 * (sg[i] + sg[k]), sjk = 0.5 * (sg[j] + sg[k]);
    sgb[i] += tb / sij + tb / sik;
    sgb[j] += tb / sij + tb / sjk;
    sgb[k] += tb / sik + tb / sjk;
  }
}

// Reverse sweep through mix_coef for a NON-ASSOCIATING binary: given the coefficient adjoints (cbar: flat order of
// the base group 0 .. 17 then, if dipolar, the 26 dipole coefficients; bnbar: bond numbers) it returns
//   out[0..15] = d p_red / d(parameters of the two components), out[16] = d/d k_ij, out[17] = 0, out[18] = d/d T.
FD_NOINLINE void mix_coef_sweep(const double (&prm)[16], const double (&kij)[2], double T, const MixCoef<double>& c,
                                const double* cbar, const double* bnbar, double (&out)[19]) {
  const double ti = 1.0 / T;
  double mb[2], sb[2], eb[2], mub[2] = {0.0, 0.0}, kb = 0.0, Tb = 0.0;
  double d[2], ex[2], dbar[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double m = prm[8 * i], s = prm[8 * i + 1], e = prm[8 * i + 2];
    ex[i] = exp(-3.0 * e * ti);
    d[i] = s * (1.0 - 0.12 * ex[i]);
    const double dd = d[i];
    const double z0 = cbar[i], z1 = cbar[2 + i], z2 = cbar[4 + i], z3 = cbar[6 + i];
    mb[i] = (kPi / 6.0) * (z0 + dd * (z1 + dd * (z2 + dd * z3))) + cbar[8 + i] + bnbar[3 * i];  // bn[i][i]: index 2 i + i
    dbar[i] = (kPi / 6.0) * m * (z1 + dd * (2.0 * z2 + 3.0 * dd * z3)) + 0.5 * cbar[16 + i];
    sb[i] = eb[i] = 0.0;
  }
  // dispersion: q1[k] (10 + k), q2[k] = q1[k] e_k (13 + k)
  {
    const double e00 = prm[2] * ti, e11 = prm[10] * ti;
    const double e01 = sqrt(prm[2] * prm[10]) * ti * (1.0 - kij[0]);
    const double s01 = 0.5 * (prm[1] + prm[9]);
    double tb = cbar[10] + cbar[13] * e00;  // adjoint of q1[0]
    double ebar = (cbar[13] + tb / e00) * c.q1[0];
    mb[0] += 2.0 * tb * c.q1[0] / prm[0];
    sb[0] += 3.0 * tb * c.q1[0] / prm[1];
    eb[0] += ebar * ti;
    Tb -= ebar * e00 * ti;
    tb = cbar[12] + cbar[15] * e11;
    ebar = (cbar[15] + tb / e11) * c.q1[2];
    mb[1] += 2.0 * tb * c.q1[2] / prm[8];
    sb[1] += 3.0 * tb * c.q1[2] / prm[9];
    eb[1] += ebar * ti;
    Tb -= ebar * e11 * ti;
    tb = cbar[11] + cbar[14] * e01;
    ebar = (cbar[14] + tb / e01) * c.q1[1] * e01;  // e01 times the adjoint of e01
    mb[0] += tb * c.q1[1] / prm[0];
    mb[1] += tb * c.q1[1] / prm[8];
    const double s01b = 1.5 * tb * c.q1[1] / s01;  // (d / d s01) / 2
    sb[0] += s01b;
    sb[1] += s01b;
    eb[0] += 0.5 * ebar / prm[2];
    eb[1] += 0.5 * ebar / prm[10];
    Tb -= ebar * ti;
    kb -= ebar / (1.0 - kij[0]);
  }
  // dipole group
  if (c.dipolar) {
    const double* cb = cbar + 18;
    double mc[2], sg[2], ep[2], mcb[2] = {0.0, 0.0}, sgb[2] = {0.0, 0.0}, epb[2] = {0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mc[i] = prm[8 * i] <= 2.0 ? prm[8 * i] : 2.0;
      sg[i] = prm[8 * i + 1];
      ep[i] = prm[8 * i + 2];
    }
    dipole_coef_adjoint(c, mc, sg, ep, ti, cb, mcb, sgb, epb, Tb);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const double m = prm[8 * i], mu = prm[8 * i + 3];
      // mu2t_i = mu_i^2 kDipoleScale / (m_i T)
      mub[i] += cb[i] * 2.0 * mu * kDipoleScale * ti / m;
      mb[i] -= cb[i] * c.mu2t[i] / m;
      Tb -= cb[i] * c.mu2t[i] * ti;
      if (m <= 2.0) mb[i] += mcb[i];
      sb[i] += sgb[i];
      eb[i] += epb[i];
    }
  }
  // d_i = sigma_i (1 - 0.12 exp(-3 eps_i / T))
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double s = prm[8 * i + 1], e = prm[8 * i + 2];
    sb[i] += dbar[i] * d[i] / s;
    const double de = dbar[i] * s * 0.36 * ex[i] * ti;
    eb[i] += de;
    Tb -= de * e * ti;
  }
#pragma unroll
  for (int k = 0; k < 19; ++k) out[k] = 0.0;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    out[8 * i] = mb[i];
    out[8 * i + 1] = sb[i];
    out[8 * i + 2] = eb[i];
    out[8 * i + 3] = mub[i];
  }
  out[16] = kb;
  out[18] = Tb;
}

// Closed-form gradient of the bubble / dew pressure [Pa] of a NON-ASSOCIATING binary with respect to all 19
// inputs (2 x 8 parameters, k_ij, eps_AiBj (= 0 here), T), times g: coefficient adjoints (closed form) and one
// reverse sweep through the coefficient stage.  Replaces 7 forward-mode passes of 3 directions each.
// (cbar / bnbar live in one struct: see the compiler note in gc_pcsaft.cuh.)
FD void mix_point_grad_closed(const double (&prm)[16], const double (&kij)[2], double T, const MixCoef<double>& c,
                              const double (&spec)[2], const double (&inc)[2], const PointWeights& q, double p_red,
                              double g, double (&out)[19]) {
  struct {
    double cbar[18 + 26];
    double bnbar[4];
  } w;
  mix_base_coef_adjoint<2, false>(c, spec, inc, q, w.cbar, w.bnbar);
  if (c.dipolar) mix_dipole_coef_adjoint<2, false>(c, spec, inc, q, w.cbar, 18);
  mix_coef_sweep(prm, kij, T, c, w.cbar, w.bnbar, out);
  const double scale = g * T * kPaPerRed;
#pragma unroll
  for (int k = 0; k < 19; ++k) out[k] *= scale;
  out[18] += g * p_red * kPaPerRed;  // explicit factor T of p = p_red T kB / A^3
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
