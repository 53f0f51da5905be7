// CPU build of the kernels' per-state device functions (same headers, -DFEOS_HOST_SIM).
// Debugging / CPU-test aid only: lets the solver logic that runs inside the CUDA kernels be
// exercised in the GPU-less build container.  Not part of the product; nothing in feos_torch_b200
// loads it.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../feos_torch_b200/csrc/pure_solver.cuh"
#include "../../feos_torch_b200/csrc/pure_grad.cuh"

using namespace feos;

static void load_coef(const double* p, double T, PureCoef<double>& c, double& rho_max) {
  pure_coef<double>(c, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], T, false);
  rho_max = kMaxEta / (kPi / 6.0 * p[0] * p[1] * p[1] * p[1]);
}
static int g_start_corr = 1;  // liquid start correlation (liquid_start.cuh) on / off for A/B tests
static double start_of(const double* p, double T, const PureCoef<double>& c, double rho_max) {
  return g_start_corr ? pure_start_density(c, rho_max, T / p[2]) : 0.0;
}

extern "C" {

void sim_set_start_correlation(int on) { g_start_corr = on; }

void sim_pure_vle(const double* params, const double* T, int64_t n, int use_robust, double* rho, double* vp,
                  double* rleq, uint8_t* status, int32_t* evals) {
  for (int64_t i = 0; i < n; ++i) {
    PureCoef<double> c;
    double rho_max;
    load_coef(params + 8 * i, T[i], c, rho_max);
    double rl, u, rle, rve;
    PureEval el, ev;
    int ne = 0;
    int st = pure_vle_fast(c, rho_max, rl, u, el, ev, rle, rve, ne, start_of(params + 8 * i, T[i], c, rho_max));
    if (st == kStatusRetry && use_robust) st = pure_vle_robust(c, rho_max, rl, u, el, ev, rle, rve, ne);
    status[i] = (uint8_t)st;
    evals[i] = ne;
    if (st == kStatusOk) {
      rho[2 * i] = exp(u);
      rho[2 * i + 1] = rl;
      pure_vle_properties(el, ev, rle, rve, T[i], vp[i], rleq[i]);
    } else {
      rho[2 * i] = rho[2 * i + 1] = vp[i] = rleq[i] = NAN;
    }
  }
}

// the resumable state machine of the work-queue kernel, driven sequentially
void sim_pure_vle_machine(const double* params, const double* T, int64_t n, double* rho, double* vp, double* rleq,
                          uint8_t* status, int32_t* evals) {
  for (int64_t i = 0; i < n; ++i) {
    PureCoef<double> c;
    double rho_max;
    load_coef(params + 8 * i, T[i], c, rho_max);
    PureVleMachine m;
    m.start(rho_max, start_of(params + 8 * i, T[i], c, rho_max));
    int ne = 0, st = kRunning;
    while (st == kRunning) {
      PureEval e = pure_eval(c, m.density());
      ++ne;
      st = m.advance1(e, log(m.rho_l));
      if (st == kRunning) st = m.advance2(exp(m.u));
    }
    status[i] = (uint8_t)st;
    evals[i] = ne;
    if (st == kStatusOk) {
      rho[2 * i] = exp(m.u);
      rho[2 * i + 1] = m.rho_l;
      pure_vle_properties(m.el, m.ev, m.rle, m.rve, T[i], vp[i], rleq[i]);
    } else {
      rho[2 * i] = rho[2 * i + 1] = vp[i] = rleq[i] = NAN;
    }
  }
}

void sim_pure_vle_bwd(const double* params, const double* T, const double* rho, int64_t n, const double* gv,
                      const double* gr, double* grad /*[n,9]*/) {
  for (int64_t i = 0; i < n; ++i) {
    double rv = rho[2 * i], rl = rho[2 * i + 1];
    PureCoef<double> c;
    double rho_max;
    load_coef(params + 8 * i, T[i], c, rho_max);
    double dpl = pure_eval(c, rl).dp;
    for (int f = 0; f < 9; f += 3) {
      double g3[3];
      if (gr) pure_vle_grad_pass<3, true>(params + 8 * i, T[i], rv, rl, dpl, f, gv ? gv[i] : 0.0, gr[i], g3);
      else pure_vle_grad_pass<3, false>(params + 8 * i, T[i], rv, rl, dpl, f, gv[i], 0.0, g3);
      for (int k = 0; k < 3; ++k) grad[9 * i + f + k] = g3[k];
    }
  }
}

// closed-form gradients (the kernels' path); the pass-based functions above are the cross-check
void sim_pure_vle_bwd_fast(const double* params, const double* T, const double* rho, int64_t n, const double* gv,
                           const double* gr, double* grad /*[n,9]*/) {
  for (int64_t i = 0; i < n; ++i) {
    double rv = rho[2 * i], rl = rho[2 * i + 1];
    PureCoef<double> c;
    double rho_max;
    load_coef(params + 8 * i, T[i], c, rho_max);
    double dpl = pure_eval(c, rl).dp;
    double out[9];
    if (gr) pure_vle_grad<true>(params + 8 * i, T[i], rv, rl, dpl, gv ? gv[i] : 0.0, gr[i], true, out);
    else pure_vle_grad<false>(params + 8 * i, T[i], rv, rl, dpl, gv[i], 0.0, true, out);
    for (int k = 0; k < 9; ++k) grad[9 * i + k] = out[k];
  }
}

void sim_pure_rho_tp_bwd_fast(const double* params, const double* T, const double* P, const double* rho, int64_t n,
                              const double* g, double* grad /*[n,9]*/) {
  for (int64_t i = 0; i < n; ++i) {
    PureCoef<double> c;
    double rho_max;
    load_coef(params + 8 * i, T[i], c, rho_max);
    double p_t = P[i] / (T[i] * kPaPerRed);
    double dp = pure_eval(c, rho[i]).dp;
    double out[9];
    pure_rho_tp_grad(params + 8 * i, T[i], rho[i], p_t, dp, g[i], out);
    for (int k = 0; k < 9; ++k) grad[9 * i + k] = out[k];
  }
}

void sim_pure_rho_tp(const double* params, const double* T, const double* P, int64_t n, double* rho_out,
                     double* value, uint8_t* status, int32_t* evals) {
  for (int64_t i = 0; i < n; ++i) {
    PureCoef<double> c;
    double rho_max;
    load_coef(params + 8 * i, T[i], c, rho_max);
    double p_t = P[i] / (T[i] * kPaPerRed);
    double rho;
    PureEval e;
    int ne = 0;
    int st = pure_liquid_density(c, rho_max, p_t, rho, e, ne, start_of(params + 8 * i, T[i], c, rho_max));
    status[i] = (uint8_t)st;
    evals[i] = ne;
    double rn = rho - (e.p - p_t) / e.dp;
    rho_out[i] = st == kStatusOk ? rn : NAN;
    value[i] = st == kStatusOk ? rn * kKmolM3PerRed : NAN;
  }
}

void sim_pure_rho_tp_bwd(const double* params, const double* T, const double* P, const double* rho, int64_t n,
                         const double* g, double* grad /*[n,9]*/, double* grad_p) {
  for (int64_t i = 0; i < n; ++i) {
    PureCoef<double> c;
    double rho_max;
    load_coef(params + 8 * i, T[i], c, rho_max);
    double p_t = P[i] / (T[i] * kPaPerRed);
    double dp = pure_eval(c, rho[i]).dp;
    for (int f = 0; f < 9; f += 3) {
      double g3[3];
      pure_rho_tp_grad_pass<3>(params + 8 * i, T[i], rho[i], p_t, dp, f, g[i], g3);
      for (int k = 0; k < 3; ++k) grad[9 * i + f + k] = g3[k];
    }
    grad_p[i] = g[i] / dp / (T[i] * kPaPerRed) * kKmolM3PerRed;
  }
}

// generic second-order dual evaluation (the closed-form pure_eval is checked against it)
void sim_pure_state_generic(const double* params, const double* T, const double* rho, int64_t n, double* a, double* p,
                            double* dp) {
  for (int64_t i = 0; i < n; ++i) {
    PureCoef<double> c;
    double rho_max;
    load_coef(params + 8 * i, T[i], c, rho_max);
    PureEval e = pure_eval_generic(c, rho[i]);
    a[i] = e.a; p[i] = e.p; dp[i] = e.dp;
  }
}

void sim_pure_state(const double* params, const double* T, const double* rho, int64_t n, double* a, double* p, double* dp) {
  for (int64_t i = 0; i < n; ++i) {
    PureCoef<double> c;
    double rho_max;
    load_coef(params + 8 * i, T[i], c, rho_max);
    PureEval e = pure_eval(c, rho[i]);
    a[i] = e.a; p[i] = e.p; dp[i] = e.dp;
  }
}
}

// ---------------------------------------------------------------------------------------------
// binary mixtures
// ---------------------------------------------------------------------------------------------
#include "../../feos_torch_b200/csrc/mix_grad.cuh"
#include "../../feos_torch_b200/csrc/mix_machine.cuh"
static int g_use_machine = 0;
static int g_mix_bwd_closed = 1;
extern "C" void sim_mix_bwd_closed(int on) { g_mix_bwd_closed = on; }
extern "C" void sim_use_point_machine(int on) { g_use_machine = on; }

extern "C" {

void sim_mix_state(const double* params, const double* kij, const double* T, const double* rho, int64_t n, double* a,
                   double* p, double* mu, double* v) {
  for (int64_t i = 0; i < n; ++i) {
    double prm[16], kk[2] = {kij[2 * i], kij[2 * i + 1]};
    for (int k = 0; k < 16; ++k) prm[k] = params[16 * i + k];
    PcSaftBinary b;
    pcsaft_binary_init(b, prm, kk, T[i]);
    MixEval e = b.eval(rho[2 * i], rho[2 * i + 1]);
    a[i] = e.a; p[i] = e.p; mu[2 * i] = e.mu[0]; mu[2 * i + 1] = e.mu[1]; v[2 * i] = e.v[0]; v[2 * i + 1] = e.v[1];
  }
}

void sim_mix_point(const double* params, const double* kij, const double* T, const double* z, int64_t n, int bubble,
                   int use_robust, double* rho /*[n,4]*/, double* value, uint8_t* status, int32_t* evals) {
  for (int64_t i = 0; i < n; ++i) {
    double prm[16], kk[2] = {kij[2 * i], kij[2 * i + 1]};
    for (int k = 0; k < 16; ++k) prm[k] = params[16 * i + k];
    PcSaftBinary b;
    pcsaft_binary_init(b, prm, kk, T[i]);
    PointState st, se;
    MixEval es, ei;
    int ne = 0;
    int rc = g_use_machine ? point_fast_machine(b, bubble != 0, z[i], st, es, ei, se, ne)
                           : point_fast(b, bubble != 0, z[i], st, es, ei, se, ne);
    if (rc == kStatusRetry && use_robust) rc = point_robust(b, bubble != 0, z[i], kDefaultStartPressure / (T[i] * kPaPerRed), st, es, ei, se, ne);
    status[i] = (uint8_t)rc;
    evals[i] = ne;
    if (rc == kStatusOk) {
      double r[4], re[4];
      point_densities(st, bubble != 0, z[i], r);
      point_densities(se, bubble != 0, z[i], re);
      for (int k = 0; k < 4; ++k) rho[4 * i + k] = r[k];
      double spec[2], inc[2];
      if (bubble) { spec[0] = re[2]; spec[1] = re[3]; inc[0] = re[0]; inc[1] = re[1]; }
      else { spec[0] = re[0]; spec[1] = re[1]; inc[0] = re[2]; inc[1] = re[3]; }
      value[i] = point_pressure(es, ei.a, spec, inc) * T[i] * kPaPerRed;
    } else {
      for (int k = 0; k < 4; ++k) rho[4 * i + k] = NAN;
      value[i] = NAN;
    }
  }
}

void sim_mix_point_bwd(const double* params, const double* kij, const double* T, const double* rho, int64_t n,
                       int bubble, const double* g, double* grad /*[n,19]*/) {
  for (int64_t i = 0; i < n; ++i) {
    double prm[16], kk[2] = {kij[2 * i], kij[2 * i + 1]};
    for (int k = 0; k < 16; ++k) prm[k] = params[16 * i + k];
    PcSaftBinary b;
    pcsaft_binary_init(b, prm, kk, T[i]);
    const double* r = rho + 4 * i;
    double spec[2], inc[2];
    if (bubble) { spec[0] = r[2]; spec[1] = r[3]; inc[0] = r[0]; inc[1] = r[1]; }
    else { spec[0] = r[0]; spec[1] = r[1]; inc[0] = r[2]; inc[1] = r[3]; }
    MixEval es = b.eval(spec[0], spec[1]);
    PointWeights q = point_weights(es, spec, inc);
    if (g_mix_bwd_closed && b.c.cat == kAssocNone) {
      double p_red = point_pressure(es, mix_helmholtz<double, double, 2, false>(b.c, inc[0], inc[1]), spec, inc);
      double o[19];
      mix_point_grad_closed(prm, kk, T[i], b.c, spec, inc, q, p_red, g[i], o);
      for (int k = 0; k < 19; ++k) grad[19 * i + k] = o[k];
      continue;
    }
    for (int f = 0; f < 21; f += 3) {
      double g3[3];
      int ids[3] = {f < 19 ? f : -1, f + 1 < 19 ? f + 1 : -1, f + 2 < 19 ? f + 2 : -1};
      mix_point_grad_pass<3>(prm, kk, T[i], spec, inc, q, ids, g[i], g3);
      for (int k = 0; k < 3; ++k) if (f + k < 19) grad[19 * i + f + k] = g3[k];
    }
  }
}

// k_ij-only gradient (mix_point_bwd_kij_kernel): closed form for k_ij; plain pairs through the plain model
void sim_mix_point_bwd_kij(const double* params, const double* kij, const double* T, const double* rho, int64_t n,
                           int bubble, const double* g, double* grad /*[n]*/) {
  for (int64_t i = 0; i < n; ++i) {
    double prm[16], kk[2] = {kij[2 * i], kij[2 * i + 1]};
    for (int k = 0; k < 16; ++k) prm[k] = params[16 * i + k];
    const double* r = rho + 4 * i;
    double spec[2], inc[2];
    if (bubble) { spec[0] = r[2]; spec[1] = r[3]; inc[0] = r[0]; inc[1] = r[1]; }
    else { spec[0] = r[0]; spec[1] = r[1]; inc[0] = r[2]; inc[1] = r[3]; }
    if (mix_is_plain(prm)) {
      PcSaftBinaryPlain b;
      pcsaft_binary_init(b, prm, kk, T[i]);
      PointWeights q = point_weights(b.eval(spec[0], spec[1]), spec, inc);
      grad[i] = mix_point_grad_kij(prm, kk[0], T[i], b.c, spec, inc, q, g[i]);
    } else {
      PcSaftBinary b;
      pcsaft_binary_init(b, prm, kk, T[i]);
      PointWeights q = point_weights(b.eval(spec[0], spec[1]), spec, inc);
      grad[i] = mix_point_grad_kij(prm, kk[0], T[i], b.c, spec, inc, q, g[i]);
    }
  }
}

// forward solve of plain pairs through the plain model (mix_point_fwd_kernel<true>); status 3 = not a plain pair
void sim_mix_point_plain(const double* params, const double* kij, const double* T, const double* z, int64_t n, int bubble,
                         double* rho, double* value, uint8_t* status, int32_t* evals) {
  for (int64_t i = 0; i < n; ++i) {
    double prm[16], kk[2] = {kij[2 * i], kij[2 * i + 1]};
    for (int k = 0; k < 16; ++k) prm[k] = params[16 * i + k];
    if (!mix_is_plain(prm)) { status[i] = 3; continue; }
    PcSaftBinaryPlain b;
    pcsaft_binary_init(b, prm, kk, T[i]);
    PointState st, se;
    MixEval es, ei;
    int ne = 0;
    int rc = point_fast(b, bubble != 0, z[i], st, es, ei, se, ne);
    status[i] = (uint8_t)rc;
    evals[i] = ne;
    if (rc == kStatusOk) {
      double re[4];
      point_densities(st, bubble != 0, z[i], *reinterpret_cast<double(*)[4]>(rho + 4 * i));
      point_densities(se, bubble != 0, z[i], re);
      double spec[2], inc[2];
      if (bubble) { spec[0] = re[2]; spec[1] = re[3]; inc[0] = re[0]; inc[1] = re[1]; }
      else { spec[0] = re[0]; spec[1] = re[1]; inc[0] = re[2]; inc[1] = re[3]; }
      value[i] = point_pressure(es, ei.a, spec, inc) * T[i] * kPaPerRed;
    }
  }
}
}

// ---------------------------------------------------------------------------------------------
// group-contribution PC-SAFT
// ---------------------------------------------------------------------------------------------
#include "../../feos_torch_b200/csrc/gc_pcsaft.cuh"

static GcSource gc_source(const GcPair* pairs, const double* seg, const double* kab, const int32_t* idx,
                          const double* phi, const double* T, int64_t i) {
  GcSource s;
  s.pair = pairs + idx[i];
  s.seg_params = seg;
  s.kab = kab;
  s.phi_[0] = phi ? phi[2 * i] : 1.0;
  s.phi_[1] = phi ? phi[2 * i + 1] : 1.0;
  s.T_ = T[i];
  return s;
}

extern "C" {

int sim_gc_pair_size() { return (int)sizeof(GcPair); }

void sim_gc_state(const GcPair* pairs, const double* seg, const double* kab, const int32_t* idx, const double* phi,
                  const double* T, const double* rho, int64_t n, double* a, double* p, double* mu, double* v) {
  for (int64_t i = 0; i < n; ++i) {
    GcSource s = gc_source(pairs, seg, kab, idx, phi, T, i);
    GcBinary b;
    gc_binary_init(b, s);
    MixEval e = b.eval(rho[2 * i], rho[2 * i + 1]);
    a[i] = e.a; p[i] = e.p; mu[2 * i] = e.mu[0]; mu[2 * i + 1] = e.mu[1]; v[2 * i] = e.v[0]; v[2 * i + 1] = e.v[1];
  }
}

void sim_gc_point(const GcPair* pairs, const double* seg, const double* kab, const int32_t* idx, const double* phi,
                  const double* T, const double* z, int64_t n, int bubble, double* rho, double* value, uint8_t* status,
                  int32_t* evals) {
  for (int64_t i = 0; i < n; ++i) {
    GcSource s = gc_source(pairs, seg, kab, idx, phi, T, i);
    GcBinary b;
    gc_binary_init(b, s);
    PointState st, se;
    MixEval es, ei;
    int ne = 0;
    int rc = g_use_machine ? point_fast_machine(b, bubble != 0, z[i], st, es, ei, se, ne)
                           : point_fast(b, bubble != 0, z[i], st, es, ei, se, ne);
    if (rc == kStatusRetry) rc = point_robust(b, bubble != 0, z[i], kDefaultStartPressure / (T[i] * kPaPerRed), st, es, ei, se, ne);
    status[i] = (uint8_t)rc;
    evals[i] = ne;
    if (rc == kStatusOk) {
      double r[4], re[4];
      point_densities(st, bubble != 0, z[i], r);
      point_densities(se, bubble != 0, z[i], re);
      for (int k = 0; k < 4; ++k) rho[4 * i + k] = r[k];
      double spec[2], inc[2];
      if (bubble) { spec[0] = re[2]; spec[1] = re[3]; inc[0] = re[0]; inc[1] = re[1]; }
      else { spec[0] = re[0]; spec[1] = re[1]; inc[0] = re[2]; inc[1] = re[3]; }
      value[i] = point_pressure(es, ei.a, spec, inc) * T[i] * kPaPerRed;
    } else {
      for (int k = 0; k < 4; ++k) rho[4 * i + k] = NAN;
      value[i] = NAN;
    }
  }
}

// state functions through the slim model (53 + 8 doubles per state, DipolePoly dipole term); ok[i] = 0 for pairs of the general class
void sim_gc_state_slim(const GcPair* pairs, const double* seg, const double* kab, const int32_t* idx, const double* phi,
                       const double* T, const double* rho, int64_t n, double* a, double* p, double* mu, double* v,
                       double* h, uint8_t* ok) {
  for (int64_t i = 0; i < n; ++i) {
    GcSource s = gc_source(pairs, seg, kab, idx, phi, T, i);
    GcPairCoef pc;  // (the kernels compute this once per pair and call: gc_pair_coef_kernel)
    gc_pair_coef(pc, seg, kab, *s.pair);
    s.pc = &pc;
    ok[i] = gc_is_slim(s, *s.pair);
    MixEval e;
    if (ok[i]) {
      double slot[kGcSlimWords];
      GcSlim b;
      gc_slim_init(b, slot, s);
      e = b.eval(rho[2 * i], rho[2 * i + 1]);
    } else {
      GcBinary b;
      gc_binary_init(b, s);
      e = b.eval(rho[2 * i], rho[2 * i + 1]);
    }
    a[i] = e.a; p[i] = e.p; mu[2 * i] = e.mu[0]; mu[2 * i + 1] = e.mu[1]; v[2 * i] = e.v[0]; v[2 * i + 1] = e.v[1];
    h[3 * i] = e.h00; h[3 * i + 1] = e.h01; h[3 * i + 2] = e.h11;
  }
}

// fast path of the slim-class forward kernel (gc_point_fwd_kernel<true>); status 3 = not a slim pair, 2 = retry list
void sim_gc_point_slim(const GcPair* pairs, const double* seg, const double* kab, const int32_t* idx, const double* phi,
                       const double* T, const double* z, int64_t n, int bubble, double* rho, double* value,
                       uint8_t* status, int32_t* evals) {
  for (int64_t i = 0; i < n; ++i) {
    GcSource s = gc_source(pairs, seg, kab, idx, phi, T, i);
    GcPairCoef pc;  // (the kernels compute this once per pair and call: gc_pair_coef_kernel)
    gc_pair_coef(pc, seg, kab, *s.pair);
    s.pc = &pc;
    if (!gc_is_slim(s, *s.pair)) { status[i] = 3; continue; }
    double slot[kGcSlimWords];
    GcSlim b;
    gc_slim_init(b, slot, s);
    PointState st, se;
    MixEval es, ei;
    int ne = 0;
    int rc = point_fast(b, bubble != 0, z[i], st, es, ei, se, ne);
    status[i] = (uint8_t)rc;
    evals[i] = ne;
    if (rc == kStatusOk) {
      double re[4];
      point_densities(st, bubble != 0, z[i], *reinterpret_cast<double(*)[4]>(rho + 4 * i));
      point_densities(se, bubble != 0, z[i], re);
      double spec[2], inc[2];
      if (bubble) { spec[0] = re[2]; spec[1] = re[3]; inc[0] = re[0]; inc[1] = re[1]; }
      else { spec[0] = re[0]; spec[1] = re[1]; inc[0] = re[2]; inc[1] = re[3]; }
      value[i] = point_pressure(es, ei.a, spec, inc) * T[i] * kPaPerRed;
    }
  }
}

// grad_seg [n_seg,8] and grad_kab [n_rec] accumulate over states; grad_phi [n,2], grad_t [n]
// the k_ab-only backward path (gc_point_bwd_kab_kernel)
void sim_gc_point_bwd_kab(const GcPair* pairs, const double* seg, const double* kab, int n_rec, const int32_t* idx,
                          const double* phi, const double* T, const double* rho, int64_t n, int bubble, const double* g,
                          double* grad_kab) {
  for (int64_t i = 0; i < n; ++i) {
    GcSource s = gc_source(pairs, seg, kab, idx, phi, T, i);
    GcPairCoef pc;  // (the kernels compute this once per pair and call: gc_pair_coef_kernel)
    gc_pair_coef(pc, seg, kab, *s.pair);
    s.pc = &pc;
    const GcPair& pr = *s.pair;
    const double* r = rho + 4 * i;
    double spec[2], inc[2];
    if (bubble) { spec[0] = r[2]; spec[1] = r[3]; inc[0] = r[0]; inc[1] = r[1]; }
    else { spec[0] = r[0]; spec[1] = r[1]; inc[0] = r[2]; inc[1] = r[3]; }
    double cb1, cb2;
    if (gc_is_slim(s, pr)) {  // as gc_point_bwd_kab_kernel
      double slot[kGcSlimWords];
      GcSlim b;
      gc_slim_init(b, slot, s);
      PointWeights q = point_weights(b.eval(spec[0], spec[1]), spec, inc);
      point_q_adjoints(GcSlimView(slot, s.pair), spec, inc, q, cb1, cb2);
    } else {
      GcBinary b;
      gc_binary_init(b, s);
      PointWeights q = point_weights(b.eval(spec[0], spec[1]), spec, inc);
      point_q_adjoints(b.c, spec, inc, q, cb1, cb2);
    }
    double scale = g[i] * T[i] * kPaPerRed;
    for (int a = 0; a < pr.nseg; ++a)
      for (int c2 = 0; c2 < pr.nseg; ++c2) {
        int rec = pr.krec[a][c2];
        if (rec < 0 || pr.cnt[0][a] * pr.cnt[1][c2] == 0.0) continue;
        grad_kab[rec] += gc_kab_term(s, pr, a, c2, cb1, cb2) * scale;
      }
  }
  (void)n_rec;
}

void sim_gc_point_bwd(const GcPair* pairs, const double* seg, int n_seg, const double* kab, int n_rec,
                      const int32_t* idx, const double* phi, const double* T, const double* rho, int64_t n, int bubble,
                      const double* g, double* grad_seg, double* grad_kab, double* grad_phi, double* grad_t,
                      int generic_all) {
  for (int64_t i = 0; i < n; ++i) {
    GcSource s = gc_source(pairs, seg, kab, idx, phi, T, i);
    const GcPair& pr = *s.pair;
    GcBinary b;
    gc_binary_init(b, s);
    const double* r = rho + 4 * i;
    double spec[2], inc[2];
    if (bubble) { spec[0] = r[2]; spec[1] = r[3]; inc[0] = r[0]; inc[1] = r[1]; }
    else { spec[0] = r[0]; spec[1] = r[1]; inc[0] = r[2]; inc[1] = r[3]; }
    MixEval es = b.eval(spec[0], spec[1]);
    PointWeights q = point_weights(es, spec, inc);
    double p_red = point_pressure(es, gc_binary_value(b.c, inc[0], inc[1]), spec, inc);
    double cbar[kGcMaxCoef];
    int ncoef = coef_count(b.c);
    if (generic_all) {
      for (int f = 0; f < ncoef; f += 3) gc_coef_adjoint_pass<3>(b.c, spec, inc, q, f, ncoef, cbar);
    } else {
      GcDualCoef cd;
      double* seed[kGcMaxCoef];
      gc_coef_adjoints(b.c, spec, inc, q, cd, seed, cbar);
      if (getenv("FEOS_SIM_CHECK_COEF")) {
        // debugging aid: the closed-form / restricted adjoints against one full dual pass per coefficient
        gc_dual_coef(cd, b.c, seed);
        for (int f = 0; f < ncoef; ++f) {
          *seed[f] = 1.0;
          double ref = gc_coef_adjoint(cd, spec, inc, q);
          *seed[f] = 0.0;
          if (fabs(ref - cbar[f]) > 1e-9 * (fabs(ref) + 1e-300)) printf("state %lld coef %d: %.15g vs %.15g\n", (long long)i, f, cbar[f], ref);
        }
      }
    }
    double scale = g[i] * T[i] * kPaPerRed;
    double gseg[8 * kGcMaxSeg];
    for (int k = 0; k < 8 * pr.nseg; ++k) gseg[k] = 0.0;
    double gp2[2] = {0.0, 0.0}, gtr = 0.0;
    if (generic_all) {
      // cross-check path: every coefficient through dual passes over ALL input directions
      int n_all = 8 * pr.nseg + n_rec + 3;
      for (int f = 0; f < n_all; f += 4) {
        int ids[4];
        for (int k = 0; k < 4; ++k) ids[k] = (f + k < n_all) ? f + k : -1;
        double g4[4];
        gc_input_grad_pass<4, true>(s, n_rec, ids, cbar, g4);
        for (int k = 0; k < 4 && ids[k] >= 0; ++k) {
          int d = ids[k];
          if (d < 8 * pr.nseg) gseg[d] += g4[k];
          else if (d < 8 * pr.nseg + n_rec) grad_kab[d - 8 * pr.nseg] += g4[k] * scale;
          else if (d < 8 * pr.nseg + n_rec + 2) gp2[d - 8 * pr.nseg - n_rec] += g4[k];
          else gtr += g4[k];
        }
      }
    } else {
      double gkab[kGcMaxSeg * kGcMaxSeg] = {0.0};
      gc_base_adjoint(s, b.c, cbar, gseg, gp2, &gtr, gkab);
      if (b.c.dipolar) gc_dipole_adjoint(s, b.c, cbar + 16 + b.c.nbond, gseg, &gtr);
      for (int a = 0; a < pr.nseg; ++a)
        for (int c2 = 0; c2 < pr.nseg; ++c2)
          if (pr.krec[a][c2] >= 0) grad_kab[pr.krec[a][c2]] += gkab[kGcMaxSeg * a + c2] * scale;
      int dirs[kGcMaxDirs + 4];
      int n_dir = gc_active_directions(s, b.c, n_rec, dirs);
      for (int k = 0; k < 4; ++k) dirs[n_dir + k] = -1;
      for (int f = 0; f < n_dir; f += 4) {
        double g4[4];
        gc_input_grad_pass<4, false>(s, n_rec, dirs + f, cbar, g4);
        for (int k = 0; k < 4 && dirs[f + k] >= 0; ++k) {
          int d = dirs[f + k];
          if (d < 8 * pr.nseg) gseg[d] += g4[k];
          else gtr += g4[k];
        }
      }
    }
    for (int k = 0; k < 8 * pr.nseg; ++k) grad_seg[8 * pr.seg[k >> 3] + (k & 7)] += gseg[k] * scale;
    grad_phi[2 * i] = gp2[0] * scale;
    grad_phi[2 * i + 1] = gp2[1] * scale;
    grad_t[i] = gtr * scale + g[i] * p_red * kPaPerRed;
  }
}
}

// ---------------------------------------------------------------------------------------------
// gradients of the public state functions (state_grad.cuh)
// ---------------------------------------------------------------------------------------------
#include "../../feos_torch_b200/csrc/state_grad.cuh"

extern "C" {

// grad [n,10]: d(ga a + gp p + gdp dp)/d(8 parameters, T, rho)
void sim_pure_state_bwd(const double* params, const double* T, const double* rho, int64_t n, const double* ga,
                        const double* gp, const double* gdp, double* grad) {
  for (int64_t i = 0; i < n; ++i)
    for (int f = 0; f < 10; f += 2) {
      double o[2];
      pure_state_grad_pass<2>(params + 8 * i, T[i], rho[i], f, ga[i], gp[i], gdp[i], o);
      grad[10 * i + f] = o[0];
      grad[10 * i + f + 1] = o[1];
    }
}

// gw [n,6] = weights of (a, p, mu0, mu1, v0, v1); grad [n,21]
void sim_mix_state_bwd(const double* params, const double* kij, const double* T, const double* rho, int64_t n,
                       const double* gw, double* grad) {
  for (int64_t i = 0; i < n; ++i) {
    double prm[16], kk[2] = {kij[2 * i], kij[2 * i + 1]};
    for (int k = 0; k < 16; ++k) prm[k] = params[16 * i + k];
    for (int f = 0; f < 21; f += 3) {
      int ids[3] = {f, f + 1, f + 2};
      double o[3];
      mix_state_grad_pass<3>(prm, kk, T[i], rho[2 * i], rho[2 * i + 1], ids, gw + 6 * i, o);
      for (int k = 0; k < 3; ++k) grad[21 * i + f + k] = o[k];
    }
  }
}

// grad_seg [n_seg,8], grad_kab [n_rec] accumulate; grad_phi [n,2], grad_t [n], grad_rho [n,2]
void sim_gc_state_bwd(const GcPair* pairs, const double* seg, int n_seg, const double* kab, int n_rec,
                      const int32_t* idx, const double* phi, const double* T, const double* rho, int64_t n,
                      const double* gw, double* grad_seg, double* grad_kab, double* grad_phi, double* grad_t,
                      double* grad_rho) {
  for (int64_t i = 0; i < n; ++i) {
    GcSource s = gc_source(pairs, seg, kab, idx, phi, T, i);
    const GcPair& pr = *s.pair;
    int n_all = 8 * pr.nseg + n_rec + 5;
    for (int f = 0; f < n_all; f += 3) {
      int ids[3];
      for (int k = 0; k < 3; ++k) ids[k] = (f + k < n_all) ? f + k : -1;
      double o[3];
      gc_state_grad_pass<3>(s, n_rec, rho[2 * i], rho[2 * i + 1], ids, gw + 6 * i, o);
      for (int k = 0; k < 3 && ids[k] >= 0; ++k) {
        int d = ids[k];
        if (d < 8 * pr.nseg) grad_seg[8 * pr.seg[d >> 3] + (d & 7)] += o[k];
        else if (d < 8 * pr.nseg + n_rec) grad_kab[d - 8 * pr.nseg] += o[k];
        else if (d < 8 * pr.nseg + n_rec + 2) grad_phi[2 * i + d - 8 * pr.nseg - n_rec] = o[k];
        else if (d == 8 * pr.nseg + n_rec + 2) grad_t[i] = o[k];
        else grad_rho[2 * i + d - 8 * pr.nseg - n_rec - 3] = o[k];
      }
    }
  }
}
}

// ---------------------------------------------------------------------------------------------
// closed-form base evaluation (mix_eval_fast.cuh) against the generic H2 evaluation, PARTS = 1
// ---------------------------------------------------------------------------------------------

extern "C" {

// out [n,12]: (a, g0, g1, h00, h01, h11) fast, then generic
void sim_mix_base_eval(const double* params, const double* kij, const double* T, const double* rho, int64_t n, double* out) {
  for (int64_t i = 0; i < n; ++i) {
    double prm[16], kk[2] = {kij[2 * i], kij[2 * i + 1]};
    for (int k = 0; k < 16; ++k) prm[k] = params[16 * i + k];
    MixCoef<double> c;
    mix_coef<double>(c, prm, kk, T[i], false);
    BaseH2 f = mix_base_eval<2, false>(c, rho[2 * i], rho[2 * i + 1]);
    H2<double> r0(rho[2 * i]), r1(rho[2 * i + 1]);
    r0.g0 = 1.0;
    r1.g1 = 1.0;
    H2<double> g = mix_helmholtz<double, H2<double>, 2, false, 1>(c, r0, r1);
    double* o = out + 12 * i;
    o[0] = f.a; o[1] = f.g0; o[2] = f.g1; o[3] = f.h00; o[4] = f.h01; o[5] = f.h11;
    o[6] = g.re; o[7] = g.g0; o[8] = g.g1; o[9] = g.h00; o[10] = g.h01; o[11] = g.h11;
  }
}

void sim_gc_base_eval(const GcPair* pairs, const double* seg, const double* kab, const int32_t* idx, const double* phi,
                      const double* T, const double* rho, int64_t n, double* out) {
  for (int64_t i = 0; i < n; ++i) {
    GcSource s = gc_source(pairs, seg, kab, idx, phi, T, i);
    GcBinary b;
    gc_binary_init(b, s);
    BaseH2 f = mix_base_eval<kGcMaxBond, true>(b.c, rho[2 * i], rho[2 * i + 1]);
    H2<double> r0(rho[2 * i]), r1(rho[2 * i + 1]);
    r0.g0 = 1.0;
    r1.g1 = 1.0;
    H2<double> g = mix_helmholtz<double, H2<double>, kGcMaxBond, true, 1>(b.c, r0, r1);
    double* o = out + 12 * i;
    o[0] = f.a; o[1] = f.g0; o[2] = f.g1; o[3] = f.h00; o[4] = f.h01; o[5] = f.h11;
    o[6] = g.re; o[7] = g.g0; o[8] = g.g1; o[9] = g.h00; o[10] = g.h01; o[11] = g.h11;
  }
}
}

extern "C" {
// out [n,12]: (a, mu0, mu1, h00, h01, h11) of mix_eval (closed-form base + dipole) and of mix_eval_generic
void sim_mix_eval_pair(const double* params, const double* kij, const double* T, const double* rho, int64_t n, double* out) {
  for (int64_t i = 0; i < n; ++i) {
    double prm[16], kk[2] = {kij[2 * i], kij[2 * i + 1]};
    for (int k = 0; k < 16; ++k) prm[k] = params[16 * i + k];
    MixCoef<double> c;
    mix_coef<double>(c, prm, kk, T[i], false);
    MixEval f = mix_eval<2, false>(c, rho[2 * i], rho[2 * i + 1]);
    MixEval g = mix_eval_generic<2, false>(c, rho[2 * i], rho[2 * i + 1]);
    double* o = out + 12 * i;
    o[0] = f.a; o[1] = f.mu[0]; o[2] = f.mu[1]; o[3] = f.h00; o[4] = f.h01; o[5] = f.h11;
    o[6] = g.a; o[7] = g.mu[0]; o[8] = g.mu[1]; o[9] = g.h00; o[10] = g.h01; o[11] = g.h11;
  }
}
}

// ---- sort keys of the bubble / dew state order (csrc/order.cuh) ------------------------------------------
#include "../../feos_torch_b200/csrc/order.cuh"

extern "C" {
void sim_mix_order_key(const double* params, const double* T, const double* z, int64_t n, uint64_t* keys) {
  for (int64_t i = 0; i < n; ++i) {
    double prm[16];
    for (int k = 0; k < 16; ++k) prm[k] = params[16 * i + k];
    keys[i] = mix_order_key(prm, T[i], z[i]);
  }
}
void sim_gc_order_key(const int32_t* pair, const double* T, const double* z, int64_t n, uint64_t* keys) {
  for (int64_t i = 0; i < n; ++i) keys[i] = gc_order_key(pair[i], T[i], z[i]);
}
int sim_order_bits(int which) { return which == 0 ? kMixOrderBits : kOrderStateBits; }
}

extern "C" void sim_pure_order_key(const double* params, const double* T, int64_t n, uint64_t* keys) {
  for (int64_t i = 0; i < n; ++i) {
    double prm[8];
    for (int k = 0; k < 8; ++k) prm[k] = params[8 * i + k];
    keys[i] = feos::pure_order_key(prm, T[i]);
  }
}
extern "C" int sim_pure_order_bits() { return feos::kPureOrderBits; }
