// Per-state binary bubble / dew point solver, shared by the PC-SAFT mixture kernels (and, through the
// Model template parameter, by the group-contribution kernels).
//
// Replaces what the reference's native layer delegates to the un-vendored feos crate:
//   PhaseEquilibrium::bubble_point / dew_point(eos, T, [z, 1-z], Some(p0), None, default)
//   (/root/reference/src/pcsaft.rs:150-214, src/gc_pcsaft.rs:103-171).
// The returned pressure is second-order insensitive to the solver path, so the
// algorithm is designed for the GPU:
//   "specified" phase = the phase whose composition is given (liquid for bubble, vapour for dew):
//                       one unknown, its total density;
//   "incipient" phase = the other one: two unknown partial densities.
//   1. start: density iteration of a liquid at (nearly) zero pressure + ideal-vapour partial pressures
//      with Poynting correction; for dew points a few successive-substitution rounds on x;
//   2. Newton on (mu_1, mu_2, p) equalities in the logarithms of the three unknown densities, step
//      limited to |d ln rho| <= 1, with back-tracking out of mechanically unstable states.
// Model must provide:  Eval eval(rho0, rho1) (a, mu[2], h00, h01, h11, p, dp[2], v[2]) and
//                      double max_density(x0) (total density at packing fraction 0.5).
#pragma once
#include "liquid_start.cuh"
#include "mix_eval_fast.cuh"
#include "pure_solver.cuh"

namespace feos {

#ifndef FEOS_DEW_SIM
#define FEOS_DEW_SIM 1
#endif
#ifndef FEOS_DEW_ROUNDS
#define FEOS_DEW_ROUNDS 4
#endif
#ifndef FEOS_DEW_XTOL
#define FEOS_DEW_XTOL 1e-2
#endif
constexpr bool kDewSimultaneous = FEOS_DEW_SIM;
constexpr int kDewRounds = FEOS_DEW_ROUNDS;
constexpr double kDewXTol = FEOS_DEW_XTOL;
constexpr double kDefaultStartPressure = 1e5;  // Pa: start pressure when the caller gives none (the value of the reference's tests)

struct PointState {
  double s;       // ln(total density of the specified phase)
  double u[2];    // ln(partial densities of the incipient phase)
};

// densities in the order of the native API: (rho_V1, rho_V2, rho_L1, rho_L2), src/pcsaft.rs:225-228
FD void point_densities(const PointState& st, bool bubble, double z0, double (&rho)[4]) {
  double rs = exp(st.s);
  double spec[2] = {rs * z0, rs * (1.0 - z0)};
  double inc[2] = {exp(st.u[0]), exp(st.u[1])};
  if (bubble) {
    rho[0] = inc[0]; rho[1] = inc[1]; rho[2] = spec[0]; rho[3] = spec[1];
  } else {
    rho[0] = spec[0]; rho[1] = spec[1]; rho[2] = inc[0]; rho[3] = inc[1];
  }
}

// Newton-step bubble/dew pressure (reduced) at fixed densities (pcsaft_mix.py:435-443, :459-467):
// `es` = evaluation of the specified phase (with derivatives), `a_inc` = a of the incipient phase.
FD double point_pressure(const MixEval& es, double a_inc, const double (&spec)[2], const double (&inc)[2]) {
  double ri = inc[0] + inc[1];
  double w0 = inc[0] / ri, w1 = inc[1] / ri;
  double v = w0 * es.v[0] + w1 * es.v[1];
  double g = w0 * (log(inc[0] / spec[0]) - es.mu[0]) + w1 * (log(inc[1] / spec[1]) - es.mu[1]);
  return -(a_inc / ri + es.p * v + g - 1.0) / (1.0 / ri - v);
}

FD bool solve3(const double (&J)[3][3], const double (&F)[3], double (&dx)[3]) {
  double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
  double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
  double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
  if (!(fabs(det) > 0.0)) return false;
  double id = 1.0 / det;
  double r0 = -F[0], r1 = -F[1], r2 = -F[2];
  dx[0] = (r0 * c00 + r1 * (J[0][2] * J[2][1] - J[0][1] * J[2][2]) + r2 * (J[0][1] * J[1][2] - J[0][2] * J[1][1])) * id;
  dx[1] = (r0 * c01 + r1 * (J[0][0] * J[2][2] - J[0][2] * J[2][0]) + r2 * (J[0][2] * J[1][0] - J[0][0] * J[1][2])) * id;
  dx[2] = (r0 * c02 + r1 * (J[0][1] * J[2][0] - J[0][0] * J[2][1]) + r2 * (J[0][0] * J[1][1] - J[0][1] * J[1][0])) * id;
  return true;
}

// On return kStatusOk: `st` holds the accepted step; (es, ei, st_eval) the last evaluation point.
template <class Model>
FD int point_newton(const Model& mdl, bool bubble, double z0, PointState& st, MixEval& es, MixEval& ei,
                    PointState& st_eval, int max_iter, double max_step, int& evals) {
  const double z[2] = {z0, 1.0 - z0};
  PointState prev = st;
  double dx[3] = {0.0, 0.0, 0.0};
  bool have_prev = false;
  int halvings = 0;
  for (int it = 0; it < max_iter; ++it) {
    double rs = exp(st.s);
    double spec[2] = {rs * z[0], rs * z[1]};
    double inc[2] = {exp(st.u[0]), exp(st.u[1])};
    es = mdl.eval(spec[0], spec[1]);
    ei = mdl.eval(inc[0], inc[1]);
    evals += 2;
    double ri = inc[0] + inc[1];
    double dps = z[0] * es.dp[0] + z[1] * es.dp[1];
    double dpi = (inc[0] * ei.dp[0] + inc[1] * ei.dp[1]) / ri;
    double rho_l = bubble ? rs : ri, rho_v = bubble ? ri : rs;
    bool bad = !(dps > 0.0) || !(dpi > 0.0) || !(rho_l > rho_v);
    if (bad) {
      if (!have_prev || ++halvings > 10) return kStatusRetry;
      st.s = prev.s + (dx[0] *= 0.5);
      st.u[0] = prev.u[0] + (dx[1] *= 0.5);
      st.u[1] = prev.u[1] + (dx[2] *= 0.5);
      continue;
    }
    halvings = 0;
    if (fabs(rho_l / rho_v - 1.0) < 1e-5 && fabs(inc[0] / ri - z[0]) < 1e-5) return kStatusRetry;  // trivial
    double F[3] = {es.mu[0] + log(spec[0]) - ei.mu[0] - st.u[0], es.mu[1] + log(spec[1]) - ei.mu[1] - st.u[1],
                   es.p - ei.p};
    double J[3][3];
    J[0][0] = rs * (z[0] * es.h00 + z[1] * es.h01) + 1.0;
    J[1][0] = rs * (z[0] * es.h01 + z[1] * es.h11) + 1.0;
    J[2][0] = rs * dps;
    J[0][1] = -(ei.h00 * inc[0] + 1.0);
    J[0][2] = -(ei.h01 * inc[1]);
    J[1][1] = -(ei.h01 * inc[0]);
    J[1][2] = -(ei.h11 * inc[1] + 1.0);
    J[2][1] = -(ei.dp[0] * inc[0]);
    J[2][2] = -(ei.dp[1] * inc[1]);
    if (!solve3(J, F, dx)) return kStatusRetry;
    double mx = fmax(fabs(dx[0]), fmax(fabs(dx[1]), fabs(dx[2])));
    if (!(mx < 1e300)) return kStatusRetry;
    if (mx > max_step) {
      double sc = max_step / mx;
      dx[0] *= sc; dx[1] *= sc; dx[2] *= sc;
    }
    prev = st;
    st_eval = st;
    have_prev = true;
    st.s += dx[0];
    st.u[0] += dx[1];
    st.u[1] += dx[2];
    if (mx <= kVleTol) return kStatusOk;
  }
  return kStatusRetry;
}

// Liquid start of composition x (liquid_start.cuh) for non-associating, non-polar mixtures: the pure-fluid correlation
// with one-fluid values  m = sum x_i m_i,  m sigma^3 = sum x_i m_i sigma_i^3  (= 6/pi sum x_i msig3_i)  and
// eps / T = sum x_i x_j q1_ij / (m^2 sigma^3)  (q1_ij = m_i m_j eps_ij sigma_ij^3 / T: the van der Waals one-fluid
// rule on the model's own dispersion coefficients, k_ij included).  Falls back to the packing-fraction-0.5 density.
template <class Coef>
FD double one_fluid_start_density(const Coef& c, const double (&msig3)[2], double x0, double rho_max) {
  if (c.cat != kAssocNone || c.dipolar) return rho_max;
  const double x1 = 1.0 - x0;
  const double m = x0 * c.m[0] + x1 * c.m[1];
  const double ms3 = (6.0 / kPi) * (x0 * msig3[0] + x1 * msig3[1]);
  const double q = x0 * x0 * c.q1[0] + 2.0 * x0 * x1 * c.q1[1] + x1 * x1 * c.q1[2];
  if (!(q > 0.0)) return rho_max;
  const double f = liquid_start_fraction(m * ms3 / q, m);
  return f > 0.0 ? f * rho_max : rho_max;
}

// Density iteration of a liquid of composition x towards the pressure of the ideal vapour it would be
// in equilibrium with; returns ln of the ideal-vapour partial densities (Poynting corrected) in u.
template <class Model>
FD int liquid_start(const Model& mdl, double x0, double& rho, double (&u)[2], int max_iter, double tol, int& evals) {
  const double x[2] = {x0, 1.0 - x0};
  const double rho_max = mdl.max_density(x0);
  for (int it = 0; it < max_iter; ++it) {
    MixEval e = mdl.eval(rho * x[0], rho * x[1]);
    evals += 1;
    double dps = x[0] * e.dp[0] + x[1] * e.dp[1];
    if (!(dps > 0.0)) {
      // a warm start that landed in the mechanically unstable range: restart from the dense side
      if (rho < 0.999 * rho_max) {
        rho = rho_max;
        continue;
      }
      return kStatusRetry;
    }
    // ideal-vapour partial densities rho x_i exp(mu_i - v_i p); their logarithms are only needed at the end
    double w0 = e.mu[0] - e.v[0] * e.p, w1 = e.mu[1] - e.v[1] * e.p;
    double r0 = rho * x[0], r1 = rho * x[1];
    double delta = (r0 * exp(w0) + r1 * exp(w1) - e.p) / dps;
    delta = fmin(fmax(delta, -0.3 * rho), 1.05 * rho_max - rho);  // (a start left of the root overshoots once)
    rho += delta;
    if (fabs(delta) < tol * rho) {
      u[0] = log(r0) + w0;
      u[1] = log(r1) + w1;
      return kStatusOk;
    }
  }
  return kStatusRetry;
}

// rounds = number of successive-substitution rounds on the liquid composition (dew point only)
template <class Model>
FD int point_solve(const Model& mdl, bool bubble, double z0, int ss_rounds, double start_tol, int newton_iter,
                   double max_step, PointState& st, MixEval& es, MixEval& ei, PointState& st_eval, int& evals,
                   bool correlated_start = false) {
  if (!(z0 > 0.0) || !(z0 < 1.0)) return kStatusFailed;
  double u[2];
  if (bubble) {
    double rho = correlated_start ? mdl.start_density(z0) : mdl.max_density(z0);
    int rc = liquid_start(mdl, z0, rho, u, 30, start_tol, evals);
    if (rc != kStatusOk) return rc;
    st.s = log(rho);
    st.u[0] = u[0];
    st.u[1] = u[1];
  } else if (correlated_start && kDewSimultaneous) {
    // fast path: liquid composition and density move TOGETHER, one evaluation per round.  w_i = mu_i - v_i p is
    // first-order insensitive to the density error (d mu_i / d rho = v_i d p / d rho along the composition line), so the
    // K-factors of a liquid that is a Newton step or two from its density are as good as those of the converged one.
    double x0 = z0;
    double rho = mdl.start_density(x0);
    double p = 0.0;
    bool ok = false;
    for (int r = 0; r < kDewRounds; ++r) {
      const double rho_max = mdl.max_density(x0);
      MixEval e = mdl.eval(rho * x0, rho * (1.0 - x0));
      evals += 1;
      double dps = x0 * e.dp[0] + (1.0 - x0) * e.dp[1];
      if (!(dps > 0.0)) {
        if (rho < 0.999 * rho_max) { rho = rho_max; continue; }
        return kStatusRetry;
      }
      double k0 = rho * exp(e.mu[0] - e.v[0] * e.p), k1 = rho * exp(e.mu[1] - e.v[1] * e.p);  // kappa_i = p y_i / x_i
      double delta = (x0 * k0 + (1.0 - x0) * k1 - e.p) / dps;
      delta = fmin(fmax(delta, -0.3 * rho), 1.05 * rho_max - rho);
      double t0 = z0 / k0, t1 = (1.0 - z0) / k1;
      p = 1.0 / (t0 + t1);
      double xn = t0 * p;
      if (!(xn > 0.0) || !(xn < 1.0)) return kStatusRetry;
      bool done = fabs(delta) < start_tol * rho && fabs(xn - x0) < kDewXTol * fmin(xn, 1.0 - xn);
      // next point: the Newton-updated packing fraction at the new composition
      rho = (rho + delta) * mdl.max_density(xn) / rho_max;
      x0 = xn;
      if (done) { ok = true; break; }
    }
    (void)ok;
    if (!(p > 0.0)) return kStatusRetry;
    st.s = log(p);
    st.u[0] = log(rho * x0);
    st.u[1] = log(rho * (1.0 - x0));
  } else {
    double x0 = z0;
    double rho = correlated_start ? mdl.start_density(x0) : mdl.max_density(x0);
    double p = 0.0;
    for (int r = 0; r <= ss_rounds; ++r) {
      int rc = liquid_start(mdl, x0, rho, u, 30, start_tol, evals);
      if (rc != kStatusOk) return rc;
      // ideal vapour: y_i p = x_i kappa_i with kappa_i = exp(u_i) / x_i  ->  x_i = y_i p / kappa_i
      double t0 = z0 * x0 / exp(u[0]), t1 = (1.0 - z0) * (1.0 - x0) / exp(u[1]);
      p = 1.0 / (t0 + t1);
      if (r == ss_rounds) break;
      double xn = t0 * p;
      if (!(xn > 0.0) || !(xn < 1.0)) return kStatusRetry;
      bool done = fabs(xn - x0) < 1e-3 * fmin(xn, 1.0 - xn);
      // warm start of the next density iteration at (slightly above) the same packing fraction
      rho = fmin(1.05 * rho * mdl.max_density(xn) / mdl.max_density(x0), mdl.max_density(xn));
      x0 = xn;
      if (done) ss_rounds = r + 1;  // one more density iteration at the final composition
    }
    if (!(p > 0.0)) return kStatusRetry;
    st.s = log(p);
    st.u[0] = log(rho * x0);
    st.u[1] = log(rho * (1.0 - x0));
  }
  return point_newton(mdl, bubble, z0, st, es, ei, st_eval, newton_iter, max_step, evals);
}

template <class Model>
FD int point_fast(const Model& mdl, bool bubble, double z0, PointState& st, MixEval& es, MixEval& ei,
                  PointState& st_eval, int& evals) {
  return point_solve(mdl, bubble, z0, 2, 0.02, 30, 1.0, st, es, ei, st_eval, evals, true);
}

// ---- retry path ------------------------------------------------------------------------------------
// Stage A: the fast path's algorithm with fully converged starting states, many successive-substitution
//          rounds and short Newton steps.
// Stage B (dew points): the starting liquid composition is BRACKETED instead of iterated -- x <- y p / kappa(x)
//          from x = y fails when no liquid of the vapour's composition exists (light component above its
//          critical temperature, x_1 << y_1) and oscillates for strongly associating pairs.
// Stage C: successive substitution on the K-factors with a REAL vapour phase (density iterations of both
//          phases at the current pressure estimate, the structure of feos' bubble / dew point iteration,
//          SURVEY.md Appendix B), then Newton -- for states whose vapour is far from ideal (near-critical).

// f(s) = ln(y_1 / y_2) of the ideal vapour over the liquid of composition x_1 = 1 / (1 + e^-s), minus the target
template <class Model>
FD bool dew_residual(const Model& mdl, double s, double s_target, double& f, double& x0, double& rho, double (&u)[2],
                     int& evals) {
  x0 = 1.0 / (1.0 + exp(-s));
  rho = mdl.max_density(x0);
  if (liquid_start(mdl, x0, rho, u, 40, 1e-8, evals) != kStatusOk) return false;
  f = (u[0] - u[1]) - s_target;
  return f == f;
}

template <class Model>
FD int dew_start_bracketed(const Model& mdl, double z0, PointState& st, int& evals) {
  const double s_t = log(z0 / (1.0 - z0));
  double lo = -25.0, hi = 25.0, f_lo = 0.0, f_hi = 0.0, f, x0, rho, u[2];
  // the ends: an almost pure liquid of component 2 (s = lo; its vapour holds almost no component 1: f < 0) and of
  // component 1 (s = hi: f > 0).  An end without a liquid (component above its critical temperature) keeps the
  // sign it would have; states between it and the root without a liquid are moved to that end's side.
  bool ok_lo = dew_residual(mdl, lo, s_t, f_lo, x0, rho, u, evals);
  bool ok_hi = dew_residual(mdl, hi, s_t, f_hi, x0, rho, u, evals);
  if (!ok_lo && !ok_hi) return kStatusFailed;
  if ((ok_lo && !(f_lo < 0.0)) || (ok_hi && !(f_hi > 0.0))) return kStatusFailed;
  bool have = false;
  for (int it = 0; it < 60; ++it) {
    double s = 0.5 * (lo + hi);
    if (ok_lo && ok_hi && (it & 1)) {  // regula falsi on alternate steps once both ends carry values
      double t = lo - f_lo * (hi - lo) / (f_hi - f_lo);
      if (t > lo + 0.01 * (hi - lo) && t < hi - 0.01 * (hi - lo)) s = t;
    }
    if (dew_residual(mdl, s, s_t, f, x0, rho, u, evals)) {
      have = true;
      if (f < 0.0) { lo = s; f_lo = f; ok_lo = true; } else { hi = s; f_hi = f; ok_hi = true; }
      if (fabs(f) < 1e-6 || hi - lo < 1e-9) break;
    } else {
      have = false;
      if (!ok_lo) lo = s; else if (!ok_hi) hi = s; else return kStatusFailed;
      if (hi - lo < 1e-9) return kStatusFailed;
    }
  }
  if (!have) return kStatusFailed;
  double p = exp(u[0]) + exp(u[1]);  // total density of the ideal vapour = its reduced pressure
  if (!(p > 0.0)) return kStatusFailed;
  st.s = log(p);
  st.u[0] = log(rho * x0);
  st.u[1] = log(rho * (1.0 - x0));
  return kStatusOk;
}

// Total density of a phase of composition x at reduced pressure p: Newton along the composition line from `rho`.
template <class Model>
FD MixEval density_at_pressure(const Model& mdl, double x0, double p, bool liquid, int iters, double& rho, int& evals) {
  const double rho_max = mdl.max_density(x0);
  MixEval e;
  for (int it = 0;; ++it) {
    e = mdl.eval(rho * x0, rho * (1.0 - x0));
    evals += 1;
    if (it == iters) break;
    double dp = x0 * e.dp[0] + (1.0 - x0) * e.dp[1];
    double next;
    if (liquid) {
      double step = fmin(fmax((p - e.p) / dp, -0.075 * rho_max), 0.075 * rho_max);
      next = dp > 0.0 ? rho + step : fmin(1.1 * rho, 1.2 * rho_max);
    } else {
      next = dp > 0.0 ? fmax(rho + (p - e.p) / dp, 0.2 * rho) : 0.5 * rho;
    }
    if (!(next == next) || !(next > 0.0)) break;
    if (fabs(next - rho) <= 1e-13 * rho) { rho = next; it = iters - 1; continue; }
    rho = next;
  }
  return e;
}

template <class Model>
FD int point_substitution(const Model& mdl, bool bubble, double z0, double p_start, PointState& st, int& evals) {
  // start: the specified-composition liquid (bubble) / the bracketed liquid (dew) with its ideal vapour
  double x0, y0, p, rho_l, rho_v;
  if (bubble) {
    double u[2];
    rho_l = mdl.max_density(z0);
    if (liquid_start(mdl, z0, rho_l, u, 40, 1e-8, evals) != kStatusOk) {
      // no liquid at its ideal vapour's pressure (near-critical): come from the caller's starting pressure
      // (the reference hands it to feos as Some(p0), src/pcsaft.rs:170-177), low enough that both density
      // roots exist, and let the sweeps below climb
      rho_l = mdl.max_density(z0);
      MixEval e = density_at_pressure(mdl, z0, p_start, true, 40, rho_l, evals);
      u[0] = log(rho_l * z0) + e.mu[0];
      u[1] = log(rho_l * (1.0 - z0)) + e.mu[1];
      if (!(u[0] == u[0]) || !(u[1] == u[1])) return kStatusFailed;
    }
    p = exp(u[0]) + exp(u[1]);
    x0 = z0;
    y0 = exp(u[0]) / p;
  } else {
    PointState t;
    if (dew_start_bracketed(mdl, z0, t, evals) != kStatusOk) return kStatusFailed;
    p = exp(t.s);
    double r0 = exp(t.u[0]), r1 = exp(t.u[1]);
    rho_l = r0 + r1;
    x0 = r0 / rho_l;
    y0 = z0;
  }
  if (!(p > 0.0) || !(p < 1e300)) return kStatusFailed;
  rho_v = p;
  for (int sweep = 0; sweep < 60; ++sweep) {
    MixEval el = density_at_pressure(mdl, x0, p, true, 6, rho_l, evals);
    MixEval ev = density_at_pressure(mdl, y0, p, false, 6, rho_v, evals);
    // ln K_i = ln phi_i^L - ln phi_i^V,  ln phi_i = mu_i^res - ln(p / rho)
    double lz = log(p / rho_l) - log(p / rho_v);
    double k0 = exp(el.mu[0] - ev.mu[0] - lz), k1 = exp(el.mu[1] - ev.mu[1] - lz);
    double n0, n1, w0;
    if (bubble) { n0 = x0 * k0; n1 = (1.0 - x0) * k1; w0 = y0; } else { n0 = y0 / k0; n1 = (1.0 - y0) / k1; w0 = x0; }
    double sum = n0 + n1;
    if (!(sum > 0.0) || !(sum < 1e300)) return kStatusFailed;
    // damped (geometric mean of old and new composition): plain substitution oscillates with growing amplitude
    // for strongly non-ideal pairs
    double g0 = sqrt(w0 * n0 / sum), g1 = sqrt((1.0 - w0) * n1 / sum);
    double upd = g0 / (g0 + g1);
    double change = fabs(upd - w0);
    double fac = sqrt(fmin(fmax(sum, 0.5), 2.0));
    if (bubble) { y0 = upd; p *= fac; } else { x0 = upd; p /= fac; }
    if (change < 1e-7 && fabs(sum - 1.0) < 1e-7) break;
  }
  if (!(rho_l > 0.0) || !(rho_v > 0.0)) return kStatusFailed;
  if (bubble) {
    st.s = log(rho_l);
    st.u[0] = log(rho_v * y0);
    st.u[1] = log(rho_v * (1.0 - y0));
  } else {
    st.s = log(rho_v);
    st.u[0] = log(rho_l * x0);
    st.u[1] = log(rho_l * (1.0 - x0));
  }
  return kStatusOk;
}

template <class Model>
FD int point_robust(const Model& mdl, bool bubble, double z0, double p_start, PointState& st, MixEval& es, MixEval& ei,
                    PointState& st_eval, int& evals) {
  if (!(z0 > 0.0) || !(z0 < 1.0)) return kStatusFailed;
  if (point_solve(mdl, bubble, z0, 40, 1e-8, 100, 0.3, st, es, ei, st_eval, evals) == kStatusOk) return kStatusOk;
  if (!bubble && dew_start_bracketed(mdl, z0, st, evals) == kStatusOk &&
      point_newton(mdl, bubble, z0, st, es, ei, st_eval, 100, 0.3, evals) == kStatusOk)
    return kStatusOk;
  if (point_substitution(mdl, bubble, z0, p_start, st, evals) == kStatusOk &&
      point_newton(mdl, bubble, z0, st, es, ei, st_eval, 100, 1.0, evals) == kStatusOk)
    return kStatusOk;
  return kStatusFailed;
}

// PC-SAFT binary with k_ij as a Model
// One out-of-line copy of the hyper-dual evaluation per kernel: inlined at its three call sites the
// solver needs 255 registers and still spills ~1.5 KB per thread.
FD_NOINLINE void pcsaft_binary_eval(const MixCoef<double>& c, double r0, double r1, MixEval& e) { e = mix_eval(c, r0, r1); }

struct PcSaftBinary {
  MixCoef<double> c;
  double msig3[2];  // pi/6 m_i sigma_i^3
  FD MixEval eval(double r0, double r1) const {
    MixEval e;
    pcsaft_binary_eval(c, r0, r1, e);
    return e;
  }
  FD double max_density(double x0) const { return kMaxEta / (x0 * msig3[0] + (1.0 - x0) * msig3[1]); }
  FD double start_density(double x0) const { return one_fluid_start_density(c, msig3, x0, max_density(x0)); }
};

FD void pcsaft_binary_init(PcSaftBinary& b, const double (&prm)[16], const double (&kij)[2], double T) {
  mix_coef<double>(b.c, prm, kij, T, false);
  b.msig3[0] = kPi / 6.0 * prm[0] * prm[1] * prm[1] * prm[1];
  b.msig3[1] = kPi / 6.0 * prm[8] * prm[9] * prm[9] * prm[9];
}

// The same Model for PLAIN pairs (mix_is_plain: no dipole, no association term): 26 doubles per state instead of 73
// and an evaluation without the hyper-dual association code -- the kernels instantiated with it need a third fewer
// registers (3 instead of 2 resident blocks per SM) and a third of the instruction-cache footprint.
FD_NOINLINE void pcsaft_plain_eval(const MixCoefBase& c, double r0, double r1, MixEval& e) { e = mix_eval(c, r0, r1); }

struct PcSaftBinaryPlain {
  MixCoefBase c;
  double msig3[2];
  FD MixEval eval(double r0, double r1) const {
    MixEval e;
    pcsaft_plain_eval(c, r0, r1, e);
    return e;
  }
  FD double max_density(double x0) const { return kMaxEta / (x0 * msig3[0] + (1.0 - x0) * msig3[1]); }
  FD double start_density(double x0) const { return one_fluid_start_density(c, msig3, x0, max_density(x0)); }
};

FD void pcsaft_binary_init(PcSaftBinaryPlain& b, const double (&prm)[16], const double (&kij)[2], double T) {
  double s3[2], eps[2], sig[2], dseg[2];
  mix_coef_base(b.c, prm, kij, 1.0 / T, s3, eps, sig, dseg);
  b.msig3[0] = kPi / 6.0 * prm[0] * prm[1] * prm[1] * prm[1];
  b.msig3[1] = kPi / 6.0 * prm[8] * prm[9] * prm[9] * prm[9];
}

}  // namespace feos
