// TEST / BASELINE INFRASTRUCTURE ONLY -- never loaded by the product path (feos_torch_b200/).
//
// Compiled CPU restatement of what the reference's native layer does for the pure-component path:
//   PcSaft.vapor_pressure  -> vapor_pressure_  -> PhaseEquilibrium::pure(eos, T, None, default)
//                             (/root/reference/src/pcsaft.rs:82-103), rayon over states (:86-92)
//   PcSaft.liquid_density  -> liquid_density_  -> State::new_npt(eos, T, p, [1 mol], Liquid)
//                             (/root/reference/src/pcsaft.rs:105-129)
// one state at a time, all host cores (OpenMP `parallel for` = the rayon `par_map_collect`).  The feos crate
// (Cargo.toml:18, version 0.6, no lock file) is not under /root/reference, so the solver follows its
// published algorithm shape exactly as oracle/solver.py states it: ideal-gas
// start -> density iterations -> Maxwell / Newton loop -> spinodal restart; every branch below names
// the solver.py function it mirrors, and tests/test_oracle.py checks masks and densities of the two
// against each other.  The Helmholtz energy follows /root/reference/feos_torch/pcsaft_pure.py:106-178
// term by term, evaluated with second-order dual numbers in the density (what feos does with
// num-dual's Dual2; the reference's `Dual3`, feos_torch/dual.py:5-78, is the same object).
//
// Used as: the CPU baseline / reference arm of bench.py (solve stage), with the reference's
// Newton-step gradient stage (oracle/eos.py, torch autograd on the CPU) on top.
//
// Build: see oracle/build_cpu.py (g++ -O3 -fopenmp -shared -fPIC).
#include <math.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---- second-order dual number in one variable: value, d/drho, d2/drho2 -------------------------
struct D2 {
  double v, d1, d2;
};
inline D2 mk(double v) { return {v, 0.0, 0.0}; }
inline D2 operator+(D2 a, D2 b) { return {a.v + b.v, a.d1 + b.d1, a.d2 + b.d2}; }
inline D2 operator-(D2 a, D2 b) { return {a.v - b.v, a.d1 - b.d1, a.d2 - b.d2}; }
inline D2 operator-(D2 a) { return {-a.v, -a.d1, -a.d2}; }
inline D2 operator*(D2 a, D2 b) { return {a.v * b.v, a.d1 * b.v + a.v * b.d1, a.d2 * b.v + 2.0 * a.d1 * b.d1 + a.v * b.d2}; }
inline D2 operator+(D2 a, double b) { return {a.v + b, a.d1, a.d2}; }
inline D2 operator+(double b, D2 a) { return a + b; }
inline D2 operator-(D2 a, double b) { return {a.v - b, a.d1, a.d2}; }
inline D2 operator-(double b, D2 a) { return {b - a.v, -a.d1, -a.d2}; }
inline D2 operator*(D2 a, double b) { return {a.v * b, a.d1 * b, a.d2 * b}; }
inline D2 operator*(double b, D2 a) { return a * b; }
// f(a) with derivatives f0, f1, f2 at a.v
inline D2 chain(D2 a, double f0, double f1, double f2) { return {f0, f1 * a.d1, f2 * a.d1 * a.d1 + f1 * a.d2}; }
inline D2 recip(D2 a) {
  double r = 1.0 / a.v;
  return chain(a, r, -r * r, 2.0 * r * r * r);
}
inline D2 operator/(D2 a, D2 b) { return a * recip(b); }
inline D2 operator/(double a, D2 b) { return recip(b) * a; }
inline D2 operator/(D2 a, double b) { return a * (1.0 / b); }
inline D2 log(D2 a) {
  double r = 1.0 / a.v;
  return chain(a, ::log(a.v), r, -r * r);
}
inline D2 sqrt(D2 a) {
  double s = ::sqrt(a.v);
  return chain(a, s, 0.5 / s, -0.25 / (s * a.v));
}

// ---- universal model constants (pcsaft_pure.py:10-86) ------------------------------------------
const double A[7][3] = {
    {0.91056314451539, -0.30840169182720, -0.09061483509767}, {0.63612814494991, 0.18605311591713, 0.45278428063920},
    {2.68613478913903, -2.50300472586548, 0.59627007280101},  {-26.5473624914884, 21.4197936296668, -1.72418291311787},
    {97.7592087835073, -65.2558853303492, -4.13021125311661}, {-159.591540865600, 83.3186804808856, 13.7766318697211},
    {91.2977740839123, -33.7469229297323, -8.67284703679646}};
const double B[7][3] = {
    {0.72409469413165, -0.57554980753450, 0.09768831158356},   {2.23827918609380, 0.69950955214436, -0.25575749816100},
    {-4.00258494846342, 3.89256733895307, -9.15585615297321},  {-21.00357681484648, -17.21547164777212, 20.64207597439724},
    {26.8556413626615, 192.6722644652495, -38.80443005206285}, {206.5513384066188, -161.8264616487648, 93.6267740770146},
    {-355.60235612207947, -165.2076934555607, -29.66690558514725}};
const double AD[5][3] = {{0.30435038064, 0.95346405973, -1.16100802773},
                         {-0.13585877707, -1.83963831920, 4.52586067320},
                         {1.44933285154, 2.01311801180, 0.97512223853},
                         {0.35569769252, -7.37249576667, -12.2810377713},
                         {-2.06533084541, 8.23741345333, 5.93975747420}};
const double BD[5][3] = {{0.21879385627, -0.58731641193, 3.48695755800},
                         {-1.18964307357, 1.24891317047, -14.9159739347},
                         {1.16268885692, -0.50852797392, 15.3720218600},
                         {0.0, 0.0, 0.0},
                         {0.0, 0.0, 0.0}};
const double CD[4][3] = {{-0.06467735252, -0.95208758351, -0.62609792333},
                         {0.19758818347, 2.99242575222, 1.29246858189},
                         {-0.80875619458, -2.38026356489, 1.65427830900},
                         {0.69028490492, -0.27012609786, -3.43967436378}};
const double PI = 3.14159265358979323846;
const double KB = 1.380649e-23;
const double PA_PER_RED = KB / 1e-30;  // p[Pa] = p_red T kB / A^3   (pcsaft_pure.py:215)
const double MAX_ETA = 0.5;            // feos: max_density of PC-SAFT at packing fraction 0.5 (sigma based)

// ---- per-state model: everything that depends on (parameters, T) only ---------------------------
struct Pure {
  double m, s3, e, d3, mu2, delta0, na, nb;
  double c1[7], c2[7];     // dispersion series coefficients (pcsaft_pure.py:129-133)
  double ja[5], jb[5], jc[4];  // dipole series coefficients (:150-155)
  double rho_max;
};

void pure_init(Pure& q, const double* p, double T) {
  // pcsaft_pure.py:90-103
  double m = p[0], sigma = p[1], eps = p[2], mu = p[3], kappa = p[4], eps_ab = p[5];
  q.m = m;
  q.s3 = sigma * sigma * sigma;
  q.e = eps / T;
  double d = sigma * (1.0 - 0.12 * exp(-3.0 * eps / T));  // :108
  q.d3 = d * d * d;
  double mu2 = mu * mu / (m * q.s3 * eps) * 1e-19 / KB;   // :94-99
  q.mu2 = mu2 * q.e * q.s3;                               // :145
  q.delta0 = (exp(eps_ab / T) - 1.0) * q.s3 * kappa;      // :163-167
  q.na = p[6];
  q.nb = p[7];
  double m1 = (m - 1.0) / m, m2 = (m - 2.0) / m;
  for (int i = 0; i < 7; ++i) {
    q.c1[i] = m1 * (m2 * A[i][2] + A[i][1]) + A[i][0];
    q.c2[i] = m1 * (m2 * B[i][2] + B[i][1]) + B[i][0];
  }
  double mc = m < 2.0 ? m : 2.0;  // :146
  double w1 = (mc - 1.0) / mc, w2 = w1 * (mc - 2.0) / mc;
  for (int i = 0; i < 5; ++i) {
    q.ja[i] = AD[i][0] + w1 * AD[i][1] + w2 * AD[i][2];
    q.jb[i] = BD[i][0] + w1 * BD[i][1] + w2 * BD[i][2];
  }
  for (int i = 0; i < 4; ++i) q.jc[i] = CD[i][0] + w1 * CD[i][1] + w2 * CD[i][2];
  q.rho_max = MAX_ETA / (PI / 6.0 * m * q.s3);
}

// a(T, rho) with first and second density derivative -- pcsaft_pure.py:106-178
D2 helmholtz(const Pure& q, double density) {
  D2 rho = {density, 1.0, 0.0};
  D2 eta = rho * (PI / 6.0 * q.m * q.d3);  // :110
  D2 eta2 = eta * eta, eta3 = eta2 * eta;
  D2 eta_m1 = recip(1.0 - eta);
  D2 eta_m2 = eta_m1 * eta_m1;
  D2 etas[7] = {mk(1.0), eta, eta2, eta3, eta2 * eta2, eta2 * eta3, eta3 * eta3};
  D2 hs = rho * q.m * (4.0 * eta - 3.0 * eta2) * eta_m2;                      // :118
  D2 hc = -(rho * (q.m - 1.0) * log((1.0 - eta * 0.5) * eta_m1 * eta_m2));    // :121-122
  D2 I1 = mk(0.0), I2 = mk(0.0);
  for (int i = 0; i < 7; ++i) {
    I1 = I1 + etas[i] * q.c1[i];
    I2 = I2 + etas[i] * q.c2[i];
  }
  D2 om = 1.0 - eta, tm = 2.0 - eta;
  D2 C1 = recip(1.0 + q.m * (8.0 * eta - 2.0 * eta2) * eta_m2 * eta_m2 +
                (1.0 - q.m) * (20.0 * eta - 27.0 * eta2 + 12.0 * eta2 * eta - 2.0 * eta2 * eta2) / (om * om * tm * tm));  // :134-140
  D2 I = 2.0 * I1 + C1 * I2 * (q.m * q.e);
  D2 a = hs + hc + rho * rho * (-PI * q.m * q.m * q.e * q.s3) * I;  // :141-142
  if (q.mu2 != 0.0) {  // :145-160 (the term is exactly zero for mu = 0)
    D2 J1 = mk(0.0), J2 = mk(0.0);
    for (int i = 0; i < 5; ++i) J1 = J1 + etas[i] * (q.ja[i] + q.jb[i] * q.e);
    for (int i = 0; i < 4; ++i) J2 = J2 + etas[i] * q.jc[i];
    D2 phi2 = -(rho * rho * J1) * (PI / q.s3);
    D2 phi3 = -(rho * rho * rho * J2) * (4.0 / 3.0 * PI * PI / q.s3);
    a = a + phi2 * phi2 * (q.mu2 * q.mu2) / (phi2 - phi3 * q.mu2);
  }
  if (q.delta0 != 0.0 && (q.na != 0.0 || q.nb != 0.0)) {  // :163-176 (exactly zero otherwise)
    D2 k = eta * eta_m1;
    D2 delta = (1.0 + k * (1.5 + 0.5 * k)) * eta_m1 * q.delta0;
    D2 rhoa = rho * q.na, rhob = rho * q.nb;
    D2 aux = 1.0 + (rhoa - rhob) * delta;
    D2 root = sqrt(aux * aux + 4.0 * rhob * delta);
    D2 xa = 2.0 / (root + 1.0 + (rhob - rhoa) * delta);
    D2 xb = 2.0 / (root + 1.0 - (rhob - rhoa) * delta);
    a = a + rhoa * (log(xa) - 0.5 * xa + 0.5) + rhob * (log(xb) - 0.5 * xb + 0.5);
  }
  return a;
}

struct Props {
  double p, dp, a, mu;  // p/kT, d(p/kT)/drho, a, da/drho   (pcsaft_pure.py:180-182)
};
inline Props props(const Pure& q, double rho) {
  D2 a = helmholtz(q, rho);
  return {rho + rho * a.d1 - a.v, 1.0 + rho * a.d2, a.v, a.d1};
}

// ---- solver.py: _maxwell_loop --------------------------------------------------------------------
bool maxwell_loop(const Pure& q, double& rv, double& rl, int max_iter, double tol) {
  for (int it = 0; it < max_iter; ++it) {
    Props v = props(q, rv), l = props(q, rl);
    double dv = 1.0 / rv - 1.0 / rl;
    double p_new = -(v.a / rv - l.a / rl + ::log(rv / rl)) / dv;
    if (!(p_new > 0.0)) p_new = rl * exp(l.mu - l.p / rl);  // negative Maxwell pressure: ideal vapour
    double nv = rv + (p_new - v.p) / v.dp, nl = rl + (p_new - l.p) / l.dp;
    bool fail = nv != nv || nl != nl || nv <= 0.0 || nl <= 0.0 || v.dp <= 0.0 || l.dp <= 0.0 || fabs(nl / nv - 1.0) < 1e-5;
    if (fail) return false;
    bool conv = fabs(nv - rv) < tol * rv && fabs(nl - rl) < tol * rl;
    rv = nv;
    rl = nl;
    if (conv) return true;
  }
  return false;
}

// ---- solver.py: _spinodal_start ------------------------------------------------------------------
bool spinodal_start(const Pure& q, double& p_start, double& r_sv, double& r_sl) {
  const int n_grid = 96, n_bisect = 45;
  double rho_g[n_grid];
  bool neg[n_grid];
  int first = n_grid, last = -1;
  for (int i = 0; i < n_grid; ++i) {
    double f = (i + 0.5) / n_grid;
    rho_g[i] = q.rho_max * f * f * 1.05;
    double dp = props(q, rho_g[i]).dp;
    neg[i] = dp < 0.0 || dp != dp;
    if (neg[i]) {
      if (first == n_grid) first = i;
      last = i;
    }
  }
  bool ok = last >= 0 && !neg[0];
  if (first < 1) first = 1;
  if (first > n_grid - 1) first = n_grid - 1;
  if (last < 0) last = 0;
  if (last > n_grid - 2) last = n_grid - 2;
  double lo = rho_g[first - 1], hi = rho_g[first];
  for (int k = 0; k < n_bisect; ++k) {  // stable side of the vapour spinodal
    double mid = 0.5 * (lo + hi);
    if (props(q, mid).dp > 0.0) lo = mid; else hi = mid;
  }
  r_sv = lo;
  lo = rho_g[last];
  hi = rho_g[last + 1];
  for (int k = 0; k < n_bisect; ++k) {  // stable side of the liquid spinodal
    double mid = 0.5 * (lo + hi);
    if (!(props(q, mid).dp > 0.0)) lo = mid; else hi = mid;
  }
  r_sl = hi;
  double p_sv = props(q, r_sv).p, p_sl = props(q, r_sl).p;
  p_start = 0.5 * ((p_sl > 0.0 ? p_sl : 0.0) + p_sv);
  return ok && p_sv > 0.0 && p_sv > p_sl;
}

// ---- solver.py: _bracketed_density ---------------------------------------------------------------
double bracketed_density(const Pure& q, double p_target, double lo, double hi) {
  double rho = 0.5 * (lo + hi);
  for (int it = 0; it < 60; ++it) {
    Props e = props(q, rho);
    if (e.p < p_target) lo = rho; else hi = rho;
    double next = rho - (e.p - p_target) / e.dp;
    bool bad = !(next > lo) || !(next < hi) || !(e.dp > 0.0);
    next = bad ? 0.5 * (lo + hi) : next;
    if (next == rho) break;
    rho = next;
  }
  return rho;
}

// ---- solver.py: pure_vle --------------------------------------------------------------------------
bool pure_vle(const Pure& q, double& rho_v, double& rho_l) {
  // 1) ideal-gas start from a liquid-like state at 0.75 rho_max, then density iterations at that pressure
  double r0 = 0.75 * q.rho_max;
  Props e0 = props(q, r0);
  double p_start = r0 * exp(e0.mu - e0.p / r0);
  bool good = isfinite(p_start) && p_start > 0.0;
  if (good) {
    double rv = p_start, rl = q.rho_max;
    for (int it = 0; it < 40 && good; ++it) {
      Props v = props(q, rv), l = props(q, rl);
      good = v.dp > 0.0 && l.dp > 0.0;
      if (!good) break;
      double sv = (p_start - v.p) / v.dp;
      double sl = (p_start - l.p) / l.dp;
      double lim = 0.075 * q.rho_max;
      sl = sl > lim ? lim : (sl < -lim ? -lim : sl);
      double nv = rv + sv;
      if (nv < 0.05 * rv) nv = 0.05 * rv;
      bool done = fabs(nv - rv) <= 1e-15 * rv && fabs(sl) <= 1e-15 * rl;
      rv = nv;
      rl += sl;
      if (done) break;
    }
    good = good && fabs(rl / rv - 1.0) > 1e-5 && rl > rv;
    if (good && maxwell_loop(q, rv, rl, 60, 1e-13)) {
      rho_v = rv;
      rho_l = rl;
      return true;
    }
  }
  // 2) spinodal-based restart
  double p_s, r_sv, r_sl;
  if (!spinodal_start(q, p_s, r_sv, r_sl)) return false;
  double v0 = bracketed_density(q, p_s, 0.0, r_sv);
  double l0 = bracketed_density(q, p_s, r_sl, 1.6 * q.rho_max);
  if (!maxwell_loop(q, v0, l0, 200, 1e-13)) return false;
  rho_v = v0;
  rho_l = l0;
  return true;
}

// ---- solver.py: pure_liquid_density ---------------------------------------------------------------
bool liquid_density(const Pure& q, double p_t, double& rho_out) {
  double rho = q.rho_max;
  int unstable_hits = 0;
  for (int it = 0; it < 80; ++it) {
    Props e = props(q, rho);
    double lim = 0.075 * q.rho_max;
    double step = -(e.p - p_t) / e.dp;
    step = step > lim ? lim : (step < -lim ? -lim : step);
    if (step < -0.95 * rho) step = -0.95 * rho;
    bool unstable = !(e.dp > 0.0);
    if (unstable) ++unstable_hits;
    double jump = e.p > p_t ? 0.5 * rho : (1.1 * rho < q.rho_max ? 1.1 * rho : q.rho_max);
    double next = unstable ? jump : rho + step;
    bool conv = !unstable && fabs(next - rho) <= 1e-13 * rho;
    rho = next;
    if (conv) {
      rho_out = rho;
      return true;
    }
    if (next != next || unstable_hits > 20) return false;
  }
  return false;
}

}  // namespace

extern "C" {

int oracle_cpu_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// params [n,8], temperature [n] -> rho_v [n], rho_l [n] (1/A^3, NaN where failed), failed [n]
void oracle_pure_vle(const double* params, const double* temperature, int64_t n, double* rho_v, double* rho_l,
                     uint8_t* failed) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < n; ++i) {
    Pure q;
    pure_init(q, params + 8 * i, temperature[i]);
    double rv = NAN, rl = NAN;
    bool ok = pure_vle(q, rv, rl);
    rho_v[i] = ok ? rv : NAN;
    rho_l[i] = ok ? rl : NAN;
    failed[i] = ok ? 0 : 1;
  }
}

// pressure in Pa -> rho [n] (1/A^3), failed [n]
void oracle_pure_liquid_density(const double* params, const double* temperature, const double* pressure, int64_t n,
                                double* rho, uint8_t* failed) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < n; ++i) {
    Pure q;
    pure_init(q, params + 8 * i, temperature[i]);
    double r = NAN;
    bool ok = liquid_density(q, pressure[i] / temperature[i] / PA_PER_RED, r);
    rho[i] = r;
    failed[i] = ok ? 0 : 1;
  }
}

// a, da/drho, d2a/drho2 at given (T, rho): lets the tests pin this file's Helmholtz energy to the golden vectors
void oracle_pure_helmholtz(const double* params, const double* temperature, const double* density, int64_t n,
                           double* a, double* a1, double* a2) {
  for (int64_t i = 0; i < n; ++i) {
    Pure q;
    pure_init(q, params + 8 * i, temperature[i]);
    D2 r = helmholtz(q, density[i]);
    a[i] = r.v;
    a1[i] = r.d1;
    a2[i] = r.d2;
  }
}
}
