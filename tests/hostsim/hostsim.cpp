// CPU build of the kernels' per-state device functions (same headers, -DFEOS_HOST_SIM).
// Debugging / CPU-test aid only: lets the solver logic that runs inside the CUDA kernels be
// exercised in the GPU-less build container.  Not part of the product; nothing in feos_torch_b200
// loads it.
#include <stdint.h>
#include "../../feos_torch_b200/csrc/pure_solver.cuh"
#include "../../feos_torch_b200/csrc/pure_grad.cuh"

using namespace feos;

static void load_coef(const double* p, double T, PureCoef<double>& c, double& rho_max) {
  pure_coef<double>(c, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], T, false);
  rho_max = kMaxEta / (kPi / 6.0 * p[0] * p[1] * p[1] * p[1]);
}

extern "C" {

void sim_pure_vle(const double* params, const double* T, int64_t n, int use_robust, double* rho, double* vp,
                  double* rleq, uint8_t* status, int32_t* evals) {
  for (int64_t i = 0; i < n; ++i) {
    PureCoef<double> c;
    double rho_max;
    load_coef(params + 8 * i, T[i], c, rho_max);
    double rl, u, rle, rve;
    PureEval el, ev;
    int ne = 0;
    int st = pure_vle_fast(c, rho_max, rl, u, el, ev, rle, rve, ne);
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
    int st = pure_liquid_density(c, rho_max, p_t, rho, e, ne);
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
