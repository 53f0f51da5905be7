// Resumable form of the fast bubble / dew point path (point_fast in mix_solver.cuh) for the persistent
// work-queue kernels: exactly the same arithmetic, cut into trips of ONE Helmholtz evaluation each, so that
// all lanes of a warp enter the (expensive, ~1000 instruction) evaluation together whatever solver phase
// their state is in, and a lane whose state has finished takes the next state from the queue instead of
// idling until the slowest lane of the warp is done (4 .. 22 evaluations per state, mean ~10).
//   m.start(mdl, bubble, z0);  loop { m.density(r0, r1); e = mdl.eval(r0, r1); rc = m.advance(mdl, e); }
// Results on kStatusOk are those of point_fast: st (accepted step), st_eval, es, ei.
#pragma once
#include "mix_solver.cuh"

namespace feos {

template <class Model>
struct PointMachine {
  static constexpr int kSsRounds = 2, kStartIter = 30, kNewtonIter = 30;
  static constexpr double kStartTol = 0.02, kMaxStep = 1.0;
  PointState st, prev, st_eval;
  MixEval es, ei;
  double dx[3];
  double spec[2], inc[2], rs;  // densities of the Newton iterate being evaluated
  double x0, rho, rho_max, z0, p_start;
  int phase;  // 0 liquid start, 1 specified-phase evaluation, 2 incipient-phase evaluation
  int it, r, ss_rounds, halvings;
  bool bubble, have_prev;

  FD void start(const Model& mdl, bool bubble_, double z0_) {
    bubble = bubble_;
    z0 = z0_;
    x0 = z0_;
    rho_max = mdl.max_density(x0);
    rho = mdl.start_density(x0);
    phase = 0;
    it = 0;
    r = 0;
    p_start = 0.0;
    ss_rounds = bubble_ ? 0 : kSsRounds;
  }
  FD void density(double& r0, double& r1) const {
    if (phase == 0) {
      r0 = rho * x0;
      r1 = rho * (1.0 - x0);
    } else if (phase == 1) {
      r0 = spec[0];
      r1 = spec[1];
    } else {
      r0 = inc[0];
      r1 = inc[1];
    }
  }
  FD void newton_point() {
    rs = exp(st.s);
    spec[0] = rs * z0;
    spec[1] = rs * (1.0 - z0);
    inc[0] = exp(st.u[0]);
    inc[1] = exp(st.u[1]);
    phase = 1;
  }
  FD void begin_newton() {
    prev = st;
    dx[0] = dx[1] = dx[2] = 0.0;
    have_prev = false;
    halvings = 0;
    it = 0;
    newton_point();
  }

  FD int advance(const Model& mdl, const MixEval& e) {
    if (phase == 0 && !bubble && kDewSimultaneous) {  // one round of the simultaneous dew start of point_solve()
      ++r;
      bool last = r >= kDewRounds;
      bool done = false;
      double dps = x0 * e.dp[0] + (1.0 - x0) * e.dp[1];
      if (!(dps > 0.0)) {
        if (!(rho < 0.999 * rho_max)) return kStatusRetry;
        rho = rho_max;
      } else {
        double k0 = rho * exp(e.mu[0] - e.v[0] * e.p), k1 = rho * exp(e.mu[1] - e.v[1] * e.p);
        double delta = (x0 * k0 + (1.0 - x0) * k1 - e.p) / dps;
        delta = fmin(fmax(delta, -0.3 * rho), 1.05 * rho_max - rho);
        double t0 = z0 / k0, t1 = (1.0 - z0) / k1;
        p_start = 1.0 / (t0 + t1);
        double xn = t0 * p_start;
        if (!(xn > 0.0) || !(xn < 1.0)) return kStatusRetry;
        done = fabs(delta) < kStartTol * rho && fabs(xn - x0) < kDewXTol * fmin(xn, 1.0 - xn);
        double mx_n = mdl.max_density(xn);
        rho = (rho + delta) * mx_n / rho_max;
        x0 = xn;
        rho_max = mx_n;
      }
      if (!done && !last) return kRunning;
      if (!(p_start > 0.0)) return kStatusRetry;
      st.s = log(p_start);
      st.u[0] = log(rho * x0);
      st.u[1] = log(rho * (1.0 - x0));
      begin_newton();
      return kRunning;
    }
    if (phase == 0) {  // liquid_start(), one iteration
      const double x[2] = {x0, 1.0 - x0};
      ++it;
      double dps = x[0] * e.dp[0] + x[1] * e.dp[1];
      if (!(dps > 0.0)) {
        if (rho < 0.999 * rho_max) {
          rho = rho_max;
          return it >= kStartIter ? kStatusRetry : kRunning;
        }
        return kStatusRetry;
      }
      double w0 = e.mu[0] - e.v[0] * e.p, w1 = e.mu[1] - e.v[1] * e.p;
      double r0 = rho * x[0], r1 = rho * x[1];
      double delta = (r0 * exp(w0) + r1 * exp(w1) - e.p) / dps;
      delta = fmin(fmax(delta, -0.3 * rho), 1.05 * rho_max - rho);
      rho += delta;
      if (!(fabs(delta) < kStartTol * rho)) return it >= kStartIter ? kStatusRetry : kRunning;
      // the density iteration has converged: point_solve() continues
      double u0 = log(r0) + w0, u1 = log(r1) + w1;
      double eu0 = exp(u0), eu1 = exp(u1);
      if (bubble) {
        st.s = log(rho);
        st.u[0] = u0;
        st.u[1] = u1;
        begin_newton();
        return kRunning;
      }
      double t0 = z0 * x0 / eu0, t1 = (1.0 - z0) * (1.0 - x0) / eu1;
      double p = 1.0 / (t0 + t1);
      if (r == ss_rounds) {
        if (!(p > 0.0)) return kStatusRetry;
        st.s = log(p);
        st.u[0] = log(rho * x0);
        st.u[1] = log(rho * (1.0 - x0));
        begin_newton();
        return kRunning;
      }
      double xn = t0 * p;
      if (!(xn > 0.0) || !(xn < 1.0)) return kStatusRetry;
      bool done = fabs(xn - x0) < 1e-3 * fmin(xn, 1.0 - xn);
      double mx_n = mdl.max_density(xn);
      rho = fmin(1.05 * rho * mx_n / mdl.max_density(x0), mx_n);
      x0 = xn;
      if (done) ss_rounds = r + 1;
      ++r;
      it = 0;
      rho_max = mx_n;
      return kRunning;
    }
    if (phase == 1) {
      es = e;
      phase = 2;
      return kRunning;
    }
    // phase 2: both evaluations of a Newton iteration are in (point_newton(), one iteration)
    ei = e;
    const double z[2] = {z0, 1.0 - z0};
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
      if (++it >= kNewtonIter) return kStatusRetry;
      newton_point();
      return kRunning;
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
    if (mx > kMaxStep) {
      double sc = kMaxStep / mx;
      dx[0] *= sc; dx[1] *= sc; dx[2] *= sc;
    }
    prev = st;
    st_eval = st;
    have_prev = true;
    st.s += dx[0];
    st.u[0] += dx[1];
    st.u[1] += dx[2];
    if (mx <= kVleTol) return kStatusOk;
    if (++it >= kNewtonIter) return kStatusRetry;
    newton_point();
    return kRunning;
  }
};

// the machine driven sequentially (CPU tests; also documents the calling protocol)
template <class Model>
FD int point_fast_machine(const Model& mdl, bool bubble, double z0, PointState& st, MixEval& es, MixEval& ei,
                          PointState& st_eval, int& evals) {
  if (!(z0 > 0.0) || !(z0 < 1.0)) return kStatusFailed;
  PointMachine<Model> m;
  m.start(mdl, bubble, z0);
  int rc = kRunning;
  while (rc == kRunning) {
    double r0, r1;
    m.density(r0, r1);
    MixEval e = mdl.eval(r0, r1);
    ++evals;
    rc = m.advance(mdl, e);
  }
  st = m.st;
  es = m.es;
  ei = m.ei;
  st_eval = m.st_eval;
  return rc;
}

}  // namespace feos
