// Per-state solvers for the pure-component hot path.
//
// Replaces what the reference's native layer delegates to the un-vendored feos crate:
//   PhaseEquilibrium::pure(eos, T)              /root/reference/src/pcsaft.rs:82-103
//   State::new_npt(eos, T, p, Liquid)           /root/reference/src/pcsaft.rs:105-129
// The returned properties are second-order insensitive to the solver path,
// so the algorithms below are designed for the GPU instead of restating feos' iteration:
//
//  pure_vle_fast    1. liquid density iteration from rho_max towards the ideal-vapour estimate of the
//                      saturation pressure p_est = rho_L exp(mu_L^res - p_L / rho_L); p(rho) is convex
//                      on the liquid branch, so Newton from the dense side converges monotonically;
//                   2. Newton on (p_L = p_V, mu_L = mu_V) in the variables (rho_L, ln rho_V) with
//                      step limits and back-tracking out of the mechanically unstable region.
//                   Anything that does not converge is handed to
//  pure_vle_robust  scan for the mechanically unstable density range (none -> super-critical ->
//                   failed), bisect both spinodals, then a bracketed Newton iteration in ln p on
//                   mu_L(p) - mu_V(p) with bracketed density solves, polished by step 2 above.
//  pure_liquid_density  Newton from rho_max at fixed (T, p).
#pragma once
#include "pcsaft_pure.cuh"

namespace feos {

constexpr double kVleTol = 1e-8;  // accepted Newton step; the applied step squares the error
constexpr int kStatusOk = 0, kStatusFailed = 1, kStatusRetry = 2;

FD double clampd(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

// Coupled Newton iteration.  On return 0: (rho_l, u = ln rho_v) hold the accepted step and
// (el, ev, rho_l_eval, rho_v_eval) the last evaluation point, one Newton step before it.
FD int pure_vle_newton(const PureCoef<double>& c, double& rho_l, double& u, PureEval& el, PureEval& ev,
                       double& rho_l_eval, double& rho_v_eval, int max_iter, int& evals) {
  double prev_l = rho_l, prev_u = u, dl = 0.0, du = 0.0;
  bool have_prev = false;
  int halvings = 0;
  for (int it = 0; it < max_iter; ++it) {
    double rho_v = exp(u);
    el = pure_eval(c, rho_l);
    ev = pure_eval(c, rho_v);
    evals += 2;
    bool bad = !(el.dp > 0.0) || !(ev.dp > 0.0) || !(rho_l > rho_v);
    if (bad) {
      if (!have_prev || ++halvings > 10) return kStatusRetry;
      dl *= 0.5;
      du *= 0.5;
      rho_l = prev_l + dl;
      u = prev_u + du;
      continue;
    }
    halvings = 0;
    double D = 1.0 - rho_v / rho_l;
    if (D < 1e-5) return kStatusRetry;  // (near-)trivial solution
    double F1 = el.p - ev.p;
    double F2 = (el.a1 - ev.a1) + (log(rho_l) - u);
    dl = -(F1 - rho_v * F2) / (el.dp * D);
    du = (F2 - F1 / rho_l) / (ev.dp * D);
    if (!(fabs(dl) < 1e300) || !(fabs(du) < 1e300)) return kStatusRetry;
    dl = clampd(dl, -0.15 * rho_l, 0.15 * rho_l);
    du = clampd(du, -2.0, 2.0);
    prev_l = rho_l;
    prev_u = u;
    have_prev = true;
    rho_l_eval = rho_l;
    rho_v_eval = rho_v;
    rho_l += dl;
    u += du;
    if (fabs(dl) <= kVleTol * prev_l && fabs(du) <= kVleTol) return kStatusOk;
  }
  return kStatusRetry;
}

FD int pure_vle_fast(const PureCoef<double>& c, double rho_max, double& rho_l, double& u, PureEval& el,
                     PureEval& ev, double& rho_l_eval, double& rho_v_eval, int& evals) {
  rho_l = rho_max;
  u = 0.0;
  bool ready = false;
  for (int it = 0; it < 24; ++it) {
    el = pure_eval(c, rho_l);
    evals += 1;
    if (!(el.dp > 0.0)) return kStatusRetry;
    u = log(rho_l) + el.a1 - el.p / rho_l;  // ideal vapour in equilibrium with this liquid
    double delta = (exp(u) - el.p) / el.dp;
    delta = fmax(delta, -0.3 * rho_l);
    rho_l += delta;
    if (fabs(delta) < 0.02 * rho_l) {
      ready = true;
      break;
    }
  }
  if (!ready || !(u == u)) return kStatusRetry;
  return pure_vle_newton(c, rho_l, u, el, ev, rho_l_eval, rho_v_eval, 30, evals);
}

// Root of p(rho) = p_t inside [lo, hi] on a mechanically stable branch (p monotone increasing).
FD double pure_bracketed_density(const PureCoef<double>& c, double p_t, double lo, double hi, double rho, int& evals) {
  for (int it = 0; it < 60; ++it) {
    PureEval e = pure_eval(c, rho);
    evals += 1;
    if (e.p < p_t) lo = rho; else hi = rho;
    double nw = rho - (e.p - p_t) / e.dp;
    if (!(nw > lo) || !(nw < hi) || !(e.dp > 0.0)) nw = 0.5 * (lo + hi);
    bool done = fabs(nw - rho) <= 1e-13 * rho;
    rho = nw;
    if (done) break;
  }
  return rho;
}

FD int pure_vle_robust(const PureCoef<double>& c, double rho_max, double& rho_l, double& u, PureEval& el,
                       PureEval& ev, double& rho_l_eval, double& rho_v_eval, int& evals) {
  // 1. scan d(p)/d(rho) on a grid
  const int NG = 40;
  const double top = 1.1 * rho_max;
  double a_lo = 0.0, a_hi = -1.0;  // [stable, unstable] bracket of the vapour spinodal
  double b_lo = -1.0, b_hi = -1.0;  // [unstable, stable] bracket of the liquid spinodal
  double min_dp = 1e300, rho_min = 0.0;
  double last_stable = 0.0;
  for (int i = 1; i <= NG; ++i) {
    double x = (double)i / NG;
    double rho = top * x * (0.25 + 0.75 * x);  // denser sampling at low density
    PureEval e = pure_eval(c, rho);
    evals += 1;
    bool unstable = !(e.dp > 0.0);
    if (e.dp < min_dp) { min_dp = e.dp; rho_min = rho; }
    if (unstable) {
      if (a_hi < 0.0) { a_lo = last_stable; a_hi = rho; }
      b_lo = rho;
      b_hi = -1.0;
    } else {
      if (a_hi < 0.0) last_stable = rho;
      if (b_lo > 0.0 && b_hi < 0.0) { b_hi = rho; break; }
    }
  }
  if (a_hi < 0.0) {
    // no unstable grid point: refine the minimum of dp by golden section inside its grid cell(s)
    double w = top / NG * 2.0;
    double lo = fmax(rho_min - w, 1e-3 * rho_min), hi = rho_min + w;
    const double gr = 0.6180339887498949;
    double x1 = hi - gr * (hi - lo), x2 = lo + gr * (hi - lo);
    double f1 = pure_eval(c, x1).dp, f2 = pure_eval(c, x2).dp;
    evals += 2;
    for (int it = 0; it < 24 && f1 > 0.0 && f2 > 0.0; ++it) {
      if (f1 < f2) { hi = x2; x2 = x1; f2 = f1; x1 = hi - gr * (hi - lo); f1 = pure_eval(c, x1).dp; }
      else { lo = x1; x1 = x2; f1 = f2; x2 = lo + gr * (hi - lo); f2 = pure_eval(c, x2).dp; }
      evals += 1;
    }
    double xu = (f1 < f2) ? x1 : x2;
    if (!(fmin(f1, f2) <= 0.0)) return kStatusFailed;  // monotonic isotherm: super-critical
    // bracket outwards from xu with the grid spacing
    a_hi = xu; b_lo = xu;
    a_lo = fmax(xu - 2.0 * w, 0.0);
    b_hi = xu + 2.0 * w;
    if (!(pure_eval(c, b_hi).dp > 0.0)) return kStatusFailed;
    if (a_lo > 0.0 && !(pure_eval(c, a_lo).dp > 0.0)) a_lo = 0.0;
    evals += 2;
  }
  if (b_hi < 0.0) return kStatusFailed;  // unstable up to 1.1 rho_max: no liquid branch
  // 2. bisect both spinodals; keep the stable-side end points
  for (int it = 0; it < 40; ++it) {
    double mid = 0.5 * (a_lo + a_hi);
    if (pure_eval(c, mid).dp > 0.0) a_lo = mid; else a_hi = mid;
    mid = 0.5 * (b_lo + b_hi);
    if (pure_eval(c, mid).dp > 0.0) b_hi = mid; else b_lo = mid;
    evals += 2;
  }
  double rho_sv = a_lo, rho_sl = b_hi;
  if (!(rho_sv > 0.0)) return kStatusFailed;
  double p_sv = pure_eval(c, rho_sv).p, p_sl = pure_eval(c, rho_sl).p;
  evals += 2;
  if (!(p_sv > 0.0) || !(p_sv > p_sl)) return kStatusFailed;
  // 3. bracketed Newton in q = ln p on G(p) = mu_L(p) - mu_V(p), dG/dp = 1/rho_L - 1/rho_V < 0
  double q_hi = log(p_sv);
  double q_lo = (p_sl > 0.0) ? log(p_sl) : q_hi - 60.0;
  double q = (p_sl > 0.0) ? log(0.5 * (p_sl + p_sv)) : q_hi - 0.7;
  double rv = 0.5 * rho_sv, rl = 1.05 * rho_sl;
  // upper end of the liquid bracket: a density whose pressure exceeds p_sv
  double l_top = fmax(1.2 * rho_max, 1.3 * rho_sl);
  for (int k = 0; k < 8 && !(pure_eval(c, l_top).p > p_sv); ++k) { l_top *= 1.15; evals += 1; }
  bool conv = false;
  for (int it = 0; it < 80; ++it) {
    double p = exp(q);
    rv = pure_bracketed_density(c, p, 0.0, rho_sv, fmin(rv, rho_sv), evals);
    rl = pure_bracketed_density(c, p, rho_sl, l_top, fmax(rl, rho_sl), evals);
    PureEval gl = pure_eval(c, rl), gv = pure_eval(c, rv);
    evals += 2;
    double G = (gl.a1 - gv.a1) + log(rl / rv);
    if (G > 0.0) q_lo = q; else q_hi = q;
    double dq = -G / (p * (1.0 / rl - 1.0 / rv));
    double qn = q + dq;
    if (!(qn > q_lo) || !(qn < q_hi)) qn = 0.5 * (q_lo + q_hi);
    bool done = fabs(qn - q) < 1e-9;
    q = qn;
    if (done) { conv = true; break; }
  }
  if (!conv) return kStatusFailed;
  if (fabs(rl / rv - 1.0) < 1e-5) return kStatusFailed;
  rho_l = rl;
  u = log(rv);
  int st = pure_vle_newton(c, rho_l, u, el, ev, rho_l_eval, rho_v_eval, 20, evals);
  return st == kStatusOk ? kStatusOk : kStatusFailed;
}

// Liquid-like density at (T, p): Newton from rho_max (State::new_npt(.., Liquid), pcsaft.rs:116-122).
// p(rho) is convex on the liquid branch: Newton from the dense side never overshoots the root.  If the
// iteration runs into the mechanically unstable region with p still above the target, no liquid root
// exists at this pressure and -- like feos' density iteration, which then hops branches -- the vapour
// root is returned instead (Newton from the ideal gas; p(rho) is concave there).
// On success `rho` is the evaluation point and `e` its evaluation; the caller applies the Newton step
// (pcsaft_pure.py:196-198).
FD int pure_liquid_density(const PureCoef<double>& c, double rho_max, double p_t, double& rho, PureEval& e, int& evals) {
  rho = rho_max;
  bool vapour_side = false;
  for (int it = 0; it < 60; ++it) {
    e = pure_eval(c, rho);
    evals += 1;
    if (!(e.dp > 0.0)) {
      if (vapour_side || !(e.p > p_t) || !(p_t > 0.0)) return kStatusFailed;
      vapour_side = true;
      rho = fmin(p_t, 0.5 * rho);
      continue;
    }
    double step = (p_t - e.p) / e.dp;
    if (!(fabs(step) < 1e300)) return kStatusFailed;
    if (fabs(step) <= kVleTol * rho) return kStatusOk;
    rho += fmax(step, -0.5 * rho);
  }
  return kStatusFailed;
}

}  // namespace feos
