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
#include "liquid_start.cuh"
#include "pcsaft_pure.cuh"

namespace feos {

// Evaluation call of the fast VLE path and the density iteration.  SHARED = true calls ONE out-of-line copy of
// pure_eval instead of inlining it at each call site: with the association term in it the inlined loop bodies
// no longer fit the instruction cache (ncu: `no_instruction` 1.9 per issue-active cycle; C3 forward 5.98 -> 5.43 ms
// per 10 M states), while for the short non-associating evaluation inlining wins (C2 forward 0.42 vs 0.46 ms).
// The kernels pick per state; in the state order of order.cuh a warp holds one kind.
#if defined(__CUDACC__)
static __device__ __noinline__ PureEval pure_eval_shared(const PureCoef<double>& c, double rho) { return pure_eval(c, rho); }
#endif
template <bool SHARED>
FD PureEval pure_eval_sel(const PureCoef<double>& c, double rho) {
#if defined(__CUDACC__)
  if (SHARED) return pure_eval_shared(c, rho);
#endif
  return pure_eval(c, rho);
}

// liquid density iterations start at packing fraction 0.45: right of the root for all but the coldest
// states (from the left Newton overshoots once and then descends monotonically as well); measured on
// the C2 grid: 12.2 instead of 12.9 evaluations per state, same retry fraction
constexpr double kLiquidStart = 0.9;
constexpr double kVleTol = 1e-8;  // accepted Newton step; the applied step squares the error
constexpr int kStatusOk = 0, kStatusFailed = 1, kStatusRetry = 2;

FD double clampd(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

// Coupled Newton iteration.  On return 0: (rho_l, u = ln rho_v) hold the accepted step and
// (el, ev, rho_l_eval, rho_v_eval) the last evaluation point, one Newton step before it.
template <bool SHARED = false>
FD int pure_vle_newton(const PureCoef<double>& c, double& rho_l, double& u, PureEval& el, PureEval& ev,
                       double& rho_l_eval, double& rho_v_eval, int max_iter, int& evals) {
  double prev_l = rho_l, prev_u = u, dl = 0.0, du = 0.0;
  bool have_prev = false;
  int halvings = 0;
  for (int it = 0; it < max_iter; ++it) {
    double rho_v = exp(u);
    el = pure_eval_sel<SHARED>(c, rho_l);
    ev = pure_eval_sel<SHARED>(c, rho_v);
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

// Liquid start of a pure-component solve: the correlation of liquid_start.cuh for non-associating, non-polar states
// inside its range (`tstar` = T / eps), the fixed packing fraction otherwise.
FD double pure_start_density(const PureCoef<double>& c, double rho_max, double tstar) {
  double f = (c.assoc || c.dipolar) ? 0.0 : liquid_start_fraction(tstar, c.m);
  return (f > 0.0 ? f : kLiquidStart) * rho_max;
}

// rho_start: from pure_start_density (0 = the fixed packing fraction).  SHARED (= associating states, which never
// use the correlation) compiles exactly the loop it had before: that variant's loop body sits at the edge of the
// instruction cache (10 M C3 states: 5.3 ms, and 7.6 ms with the three extra instructions of the restart test).
template <bool SHARED = false>
FD int pure_vle_fast(const PureCoef<double>& c, double rho_max, double& rho_l, double& u, PureEval& el,
                     PureEval& ev, double& rho_l_eval, double& rho_v_eval, int& evals, double rho_start = 0.0) {
  const double rho_default = kLiquidStart * rho_max;
  rho_l = (!SHARED && rho_start > 0.0) ? rho_start : rho_default;
  u = 0.0;
  bool ready = false;
  for (int it = 0; it < 24; ++it) {
    el = pure_eval_sel<SHARED>(c, rho_l);
    evals += 1;
    if (!(el.dp > 0.0)) {
      if (!SHARED && it == 0 && rho_l != rho_default) {  // the correlated start is mechanically unstable: take the fixed one
        rho_l = rho_default;
        continue;
      }
      return kStatusRetry;
    }
    u = log(rho_l) + el.a1 - el.p / rho_l;  // ideal vapour in equilibrium with this liquid
    double delta = (exp(u) - el.p) / el.dp;
    delta = fmax(delta, -0.3 * rho_l);
    if (!SHARED) delta = fmin(delta, 1.05 * rho_max - rho_l);  // (a start left of the root overshoots once)
    rho_l += delta;
    if (fabs(delta) < 0.02 * rho_l) {
      ready = true;
      break;
    }
  }
  if (!ready || !(u == u)) return kStatusRetry;
  return pure_vle_newton<SHARED>(c, rho_l, u, el, ev, rho_l_eval, rho_v_eval, 30, evals);
}

// The fast path as a resumable state machine: every call of advance() consumes ONE Helmholtz evaluation
// at density(), so that the lanes of a warp can run states of different progress in lock step and pick
// up new states from a work queue as they finish (pure_vle_queue_kernel).  Same arithmetic, iteration
// limits and results as pure_vle_fast().
constexpr int kRunning = -1;
// Every trip costs one evaluation at density(), one log and one exp, issued by ALL lanes of the warp
// between the two (cheap, divergent) bookkeeping halves:
//   e = eval(density());  st = advance1(e, log(rho_l));  if (st == kRunning) st = advance2(exp(u));
struct PureVleMachine {
  double rho_l, u;                 // current iterate (u = ln rho_v)
  double prev_l, prev_u, dl, du;   // back-tracking
  double rle, rve;                 // last evaluation point, one Newton step before (rho_l, u)
  double rho_v;                    // exp(u) of the pending vapour evaluation
  PureEval el, ev;
  int phase, it, halvings;
  bool have_prev;

  FD void start(double rho_max, double rho_start = 0.0) {
    rho_l = rho_start > 0.0 ? rho_start : kLiquidStart * rho_max;
    u = 0.0;
    rho_v = 1.0;
    phase = 0;
    it = 0;
  }
  FD double density() const { return phase == 2 ? rho_v : rho_l; }

  // first half: consumes the evaluation and L = log(rho_l) (of the density before any update)
  FD int advance1(const PureEval& e, double L) {
    if (phase == 0) {  // liquid density iteration towards the ideal-vapour pressure estimate
      if (!(e.dp > 0.0)) return kStatusRetry;
      el = e;
      u = L + e.a1 - e.p / rho_l;
      return kRunning;
    }
    if (phase == 1) {  // liquid evaluation of a Newton iteration: wait for the vapour one
      el = e;
      return kRunning;
    }
    ev = e;
    ++it;
    bool bad = !(el.dp > 0.0) || !(ev.dp > 0.0) || !(rho_l > rho_v);
    if (bad) {
      if (!have_prev || ++halvings > 10) return kStatusRetry;
      dl *= 0.5;
      du *= 0.5;
      rho_l = prev_l + dl;
      u = prev_u + du;
      return it >= 30 ? kStatusRetry : kRunning;
    }
    halvings = 0;
    double D = 1.0 - rho_v / rho_l;
    if (D < 1e-5) return kStatusRetry;
    double F1 = el.p - ev.p;
    double F2 = (el.a1 - ev.a1) + (L - u);
    dl = -(F1 - rho_v * F2) / (el.dp * D);
    du = (F2 - F1 / rho_l) / (ev.dp * D);
    if (!(fabs(dl) < 1e300) || !(fabs(du) < 1e300)) return kStatusRetry;
    dl = clampd(dl, -0.15 * rho_l, 0.15 * rho_l);
    du = clampd(du, -2.0, 2.0);
    prev_l = rho_l;
    prev_u = u;
    have_prev = true;
    rle = rho_l;
    rve = rho_v;
    rho_l += dl;
    u += du;
    if (fabs(dl) <= kVleTol * prev_l && fabs(du) <= kVleTol) return kStatusOk;
    return it >= 30 ? kStatusRetry : kRunning;
  }

  // second half: X = exp(u)
  FD int advance2(double X) {
    if (phase == 0) {
      double delta = (X - el.p) / el.dp;
      delta = fmax(delta, -0.3 * rho_l);
      rho_l += delta;
      ++it;
      if (fabs(delta) < 0.02 * rho_l) {
        if (!(u == u)) return kStatusRetry;
        phase = 1;
        it = 0;
        rho_v = X;
        prev_l = rho_l;
        prev_u = u;
        dl = du = 0.0;
        have_prev = false;
        halvings = 0;
      } else if (it >= 24) {
        return kStatusRetry;
      }
      return kRunning;
    }
    if (phase == 1) {
      phase = 2;
      return kRunning;
    }
    rho_v = X;
    phase = 1;
    return kRunning;
  }
};

// Root of p(rho) = p_t inside [lo, hi] on a mechanically stable branch (p monotone increasing).
FD double pure_bracketed_density(const PureCoef<double>& c, double p_t, double lo, double hi, double rho, int& evals) {
  for (int it = 0; it < 60; ++it) {
    PureEval e = pure_eval(c, rho);
    evals += 1;
    if (e.p < p_t) lo = rho; else hi = rho;
    double nw = rho - (e.p - p_t) / e.dp;
    if (!(nw > lo) || !(nw < hi) || !(e.dp > 0.0)) nw = 0.5 * (lo + hi);
    bool done = fabs(nw - rho) <= 1e-9 * rho;
    rho = nw;
    if (done) break;
  }
  return rho;
}

// Root of dp/drho between a (dp > 0) and b (dp <= 0) by the Illinois variant of regula falsi; returns the
// end point on the mechanically stable side.
FD double pure_spinodal(const PureCoef<double>& c, double a, double fa, double b, double fb, int& evals) {
  for (int it = 0; it < 12; ++it) {
    double x = (a * fb - b * fa) / (fb - fa);
    if (!(x > fmin(a, b)) || !(x < fmax(a, b))) x = 0.5 * (a + b);
    double fx = pure_eval(c, x).dp;
    evals += 1;
    if (fx > 0.0) {
      a = x;
      fa = fx;
      fb *= 0.5;
    } else {
      b = x;
      fb = fx;
      fa *= 0.5;
    }
    if (fabs(b - a) < 1e-4 * fabs(a) || fa < 1e-6) break;
  }
  return a;
}

FD int pure_vle_robust(const PureCoef<double>& c, double rho_max, double& rho_l, double& u, PureEval& el,
                       PureEval& ev, double& rho_l_eval, double& rho_v_eval, int& evals) {
  // 1. coarse scan of dp/drho: brackets of the mechanically unstable range or the location of min dp
  const int NG = 14;
  const double top = 1.1 * rho_max;
  double a_lo = 0.0, fa_lo = 1.0, a_hi = -1.0, fa_hi = 0.0;  // vapour spinodal in [a_lo (stable), a_hi (unstable)]
  double b_lo = -1.0, fb_lo = 0.0, b_hi = -1.0, fb_hi = 0.0;  // liquid spinodal in [b_lo (unstable), b_hi (stable)]
  double min_dp = 1e300;
  int i_min = 1;
  double prev_rho = 0.0, prev_dp = 1.0;
  double dp_before = 1.0, dp_after = -1.0;  // samples next to the smallest one (dp_after < 0: not taken yet)
  for (int i = 1; i <= NG; ++i) {
    double x = (double)i / NG;
    double rho = top * x * (0.3 + 0.7 * x);  // denser sampling at low density
    double dp = pure_eval(c, rho).dp;
    evals += 1;
    if (dp < min_dp) { min_dp = dp; i_min = i; dp_before = prev_dp; dp_after = -1.0; }
    else if (i == i_min + 1) dp_after = dp;
    if (!(dp > 0.0)) {
      if (a_hi < 0.0) { a_lo = prev_rho; fa_lo = prev_dp; a_hi = rho; fa_hi = dp; }
      b_lo = rho; fb_lo = dp; b_hi = -1.0;
    } else if (b_lo > 0.0 && b_hi < 0.0) {
      b_hi = rho; fb_hi = dp;
      break;
    }
    prev_rho = rho; prev_dp = dp;
  }
  if (a_hi < 0.0) {
    auto grid = [&](int i) { double x = (double)i / NG; return top * x * (0.3 + 0.7 * x); };
    // no unstable grid point.  Clearly super-critical states (half of the retry list of a fit with a
    // super-critical slice) are recognised from the scan alone: the parabola through the smallest sample and
    // its two neighbours bounds the minimum of dp/drho (ideal gas: 1) well above zero.
    if (i_min > 1 && i_min < NG && dp_after > 0.0) {
      double x0 = grid(i_min - 1), x1 = grid(i_min), x2 = grid(i_min + 1);
      double d01 = (min_dp - dp_before) / (x1 - x0), d12 = (dp_after - min_dp) / (x2 - x1);
      double curv = (d12 - d01) / (x2 - x0);
      if (curv > 0.0) {
        double xv = 0.5 * (x0 + x1) - 0.5 * d01 / curv;
        double fv = dp_before + d01 * (xv - x0) + curv * (xv - x0) * (xv - x1);
        if (fv > 0.05 && min_dp > 0.05) return kStatusFailed;
      }
    }
    // otherwise: golden-section search for the minimum of dp around the smallest sample
    double lo = grid(i_min - 1), hi = grid(i_min < NG ? i_min + 1 : NG);
    if (i_min == 1) lo = 0.05 * grid(1);
    const double gr = 0.6180339887498949;
    double x1 = hi - gr * (hi - lo), x2 = lo + gr * (hi - lo);
    double f1 = pure_eval(c, x1).dp, f2 = pure_eval(c, x2).dp;
    evals += 2;
    for (int it = 0; it < 14 && f1 > 0.0 && f2 > 0.0 && (hi - lo) > 2e-3 * rho_max; ++it) {
      if (f1 < f2) { hi = x2; x2 = x1; f2 = f1; x1 = hi - gr * (hi - lo); f1 = pure_eval(c, x1).dp; }
      else { lo = x1; x1 = x2; f1 = f2; x2 = lo + gr * (hi - lo); f2 = pure_eval(c, x2).dp; }
      evals += 1;
    }
    if (f1 > 0.0 && f2 > 0.0) return kStatusFailed;  // monotonic isotherm: super-critical
    double xu = (f1 < f2) ? x1 : x2, fu = fmin(f1, f2);
    // stable neighbours: the enclosing grid points
    a_hi = xu; fa_hi = fu; b_lo = xu; fb_lo = fu;
    a_lo = grid(i_min - 1); b_hi = grid(i_min < NG ? i_min + 1 : NG);
    if (i_min == 1) a_lo = 0.05 * grid(1);
    fa_lo = pure_eval(c, a_lo).dp;
    fb_hi = pure_eval(c, b_hi).dp;
    evals += 2;
    if (!(fa_lo > 0.0) || !(fb_hi > 0.0)) return kStatusFailed;
  }
  if (b_hi < 0.0) return kStatusFailed;  // unstable up to 1.1 rho_max: no liquid branch
  if (!(a_lo > 0.0)) {  // first grid point already unstable: a tiny density is always stable
    a_lo = 1e-3 * a_hi;
    fa_lo = pure_eval(c, a_lo).dp;
    evals += 1;
    if (!(fa_lo > 0.0)) return kStatusFailed;
  }
  // 2. spinodals (stable-side end points)
  double rho_sv = pure_spinodal(c, a_lo, fa_lo, a_hi, fa_hi, evals);
  double rho_sl = pure_spinodal(c, b_hi, fb_hi, b_lo, fb_lo, evals);
  double p_sv = pure_eval(c, rho_sv).p, p_sl = pure_eval(c, rho_sl).p;
  evals += 2;
  if (!(p_sv > 0.0) || !(p_sv > p_sl)) return kStatusFailed;
  // 3a. near the critical point the binodal is the spinodal scaled by sqrt(3) about its mid density (Landau
  //     expansion): from there the coupled Newton iteration usually converges in a few steps, instead of the
  //     ~60 evaluations of the bracketed pressure iteration below (which remains the fallback)
  {
    double mid = 0.5 * (rho_sv + rho_sl), half = 0.5 * (rho_sl - rho_sv);
    double rv0 = mid - 1.7320508075688772 * half;
    if (rv0 > 0.25 * rho_sv) {
      rho_l = mid + 1.7320508075688772 * half;
      u = log(rv0);
      if (pure_vle_newton(c, rho_l, u, el, ev, rho_l_eval, rho_v_eval, 30, evals) == kStatusOk) return kStatusOk;
    }
  }
  // 3. bracketed Newton in q = ln p on G(p) = mu_L(p) - mu_V(p), dG/dp = 1/rho_L - 1/rho_V < 0; coarse:
  //    the coupled Newton iteration polishes the result
  double q_hi = log(p_sv);
  double q_lo = (p_sl > 0.0) ? log(p_sl) : q_hi - 60.0;
  double q = (p_sl > 0.0) ? log(0.5 * (p_sl + p_sv)) : q_hi - 0.7;
  double rv = 0.5 * rho_sv, rl = 1.05 * rho_sl;
  double l_top = fmax(1.2 * rho_max, 1.3 * rho_sl);
  for (int k = 0; k < 8 && !(pure_eval(c, l_top).p > p_sv); ++k) { l_top *= 1.15; evals += 1; }
  evals += 1;
  bool conv = false;
  for (int it = 0; it < 60; ++it) {
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
    bool done = fabs(qn - q) < 1e-3;
    q = qn;
    if (done) { conv = true; break; }
  }
  if (!conv) return kStatusFailed;
  if (fabs(rl / rv - 1.0) < 1e-5) return kStatusFailed;
  rho_l = rl;
  u = log(rv);
  int st = pure_vle_newton(c, rho_l, u, el, ev, rho_l_eval, rho_v_eval, 30, evals);
  return st == kStatusOk ? kStatusOk : kStatusFailed;
}

// Liquid-like density at (T, p): Newton from rho_max (State::new_npt(.., Liquid), pcsaft.rs:116-122).
// p(rho) is convex on the liquid branch: Newton from the dense side never overshoots the root.  If the
// iteration runs into the mechanically unstable region with p still above the target, no liquid root
// exists at this pressure and -- like feos' density iteration, which then hops branches -- the vapour
// root is returned instead (Newton from the ideal gas; p(rho) is concave there).
// On success `rho` is the evaluation point and `e` its evaluation; the caller applies the Newton step
// (pcsaft_pure.py:196-198).
// rho_start (optional, from pure_start_density): a start near the saturated liquid density instead of rho_max; it is
// only kept if it is mechanically stable, so the branch logic below sees the same situations as from rho_max.
template <bool SHARED = false>
FD int pure_liquid_density(const PureCoef<double>& c, double rho_max, double p_t, double& rho, PureEval& e, int& evals,
                           double rho_start = 0.0) {
  rho = (!SHARED && rho_start > 0.0) ? rho_start : rho_max;
  bool vapour_side = false;
  for (int it = 0; it < 60; ++it) {
    e = pure_eval_sel<SHARED>(c, rho);
    evals += 1;
    if (!(e.dp > 0.0)) {
      if (!SHARED && it == 0 && rho != rho_max) {
        rho = rho_max;
        continue;
      }
      if (vapour_side || !(e.p > p_t) || !(p_t > 0.0)) return kStatusFailed;
      vapour_side = true;
      rho = fmin(p_t, 0.5 * rho);
      continue;
    }
    double step = (p_t - e.p) / e.dp;
    if (!(fabs(step) < 1e300)) return kStatusFailed;
    if (fabs(step) <= kVleTol * rho) return kStatusOk;
    step = fmax(step, -0.5 * rho);
    // a correlated start left of the root overshoots once: not beyond rho_max, from where the iteration is the old one
    if (!SHARED && rho_start > 0.0 && rho < rho_max && rho + step > rho_max) step = rho_max - rho;
    rho += step;
  }
  return kStatusFailed;
}

}  // namespace feos
