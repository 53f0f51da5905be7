// In-register forward-mode automatic differentiation types for the PC-SAFT kernels.
//
// Replaces the reference's Python dual-number classes:
//   Dual3      (feos_torch/dual.py:5-78)           -> D2   : value, d/dx, d2/dx2 in one variable
//   DualTensor (feos_torch/dual_torch.py:4-162)    -> H2<T>: value, gradient and Hessian in two
//                                                            variables (partial densities)
//   Dual2      (feos_torch/dual_torch.py:165-208)  -> not needed: site-fraction Newton systems are
//                                                            solved with explicit 2x2 Jacobians
// plus DN<N> (value + N first-order directions) used by the backward kernels for the parameter
// gradients.  All types live entirely in registers; every operator is __forceinline__.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define FD __device__ __forceinline__

namespace feos {

// ---------------------------------------------------------------------------------------------
// scalar helpers so that templated model code works for plain double as well
// ---------------------------------------------------------------------------------------------
FD double ad_exp(double x) { return exp(x); }
FD double ad_log(double x) { return log(x); }
FD double ad_sqrt(double x) { return sqrt(x); }
FD double ad_recip(double x) { return 1.0 / x; }
FD double ad_cbrt(double x) { return cbrt(x); }
FD double ad_value(double x) { return x; }

// ---------------------------------------------------------------------------------------------
// DN<N>: first order, N directions
// ---------------------------------------------------------------------------------------------
template <int N>
struct DN {
  double re;
  double e[N];
  FD DN() {}
  FD DN(double v) : re(v) {
#pragma unroll
    for (int i = 0; i < N; ++i) e[i] = 0.0;
  }
  static FD DN seed(double v, int dir) {
    DN r(v);
#pragma unroll
    for (int i = 0; i < N; ++i) r.e[i] = (i == dir) ? 1.0 : 0.0;
    return r;
  }
  // f(x) with f0 = f(re), f1 = f'(re)
  FD DN chain(double f0, double f1) const {
    DN r;
    r.re = f0;
#pragma unroll
    for (int i = 0; i < N; ++i) r.e[i] = f1 * e[i];
    return r;
  }
  friend FD DN operator+(const DN& a, const DN& b) {
    DN r;
    r.re = a.re + b.re;
#pragma unroll
    for (int i = 0; i < N; ++i) r.e[i] = a.e[i] + b.e[i];
    return r;
  }
  friend FD DN operator-(const DN& a, const DN& b) {
    DN r;
    r.re = a.re - b.re;
#pragma unroll
    for (int i = 0; i < N; ++i) r.e[i] = a.e[i] - b.e[i];
    return r;
  }
  friend FD DN operator*(const DN& a, const DN& b) {
    DN r;
    r.re = a.re * b.re;
#pragma unroll
    for (int i = 0; i < N; ++i) r.e[i] = a.e[i] * b.re + a.re * b.e[i];
    return r;
  }
  friend FD DN operator-(const DN& a) {
    DN r;
    r.re = -a.re;
#pragma unroll
    for (int i = 0; i < N; ++i) r.e[i] = -a.e[i];
    return r;
  }
  friend FD DN operator+(const DN& a, double b) { DN r = a; r.re += b; return r; }
  friend FD DN operator+(double b, const DN& a) { DN r = a; r.re += b; return r; }
  friend FD DN operator-(const DN& a, double b) { DN r = a; r.re -= b; return r; }
  friend FD DN operator-(double b, const DN& a) { DN r = -a; r.re += b; return r; }
  friend FD DN operator*(const DN& a, double b) {
    DN r;
    r.re = a.re * b;
#pragma unroll
    for (int i = 0; i < N; ++i) r.e[i] = a.e[i] * b;
    return r;
  }
  friend FD DN operator*(double b, const DN& a) { return a * b; }
  friend FD DN operator/(const DN& a, double b) { return a * (1.0 / b); }
  friend FD DN operator/(const DN& a, const DN& b) { return a * ad_recip(b); }
  friend FD DN operator/(double a, const DN& b) { return ad_recip(b) * a; }
  friend FD DN ad_recip(const DN& a) { double r = 1.0 / a.re; return a.chain(r, -r * r); }
  friend FD DN ad_exp(const DN& a) { double f = exp(a.re); return a.chain(f, f); }
  friend FD DN ad_log(const DN& a) { return a.chain(log(a.re), 1.0 / a.re); }
  friend FD DN ad_sqrt(const DN& a) { double s = sqrt(a.re); return a.chain(s, 0.5 / s); }
  friend FD DN ad_cbrt(const DN& a) { double c = cbrt(a.re); return a.chain(c, c / (3.0 * a.re)); }
  friend FD double ad_value(const DN& a) { return a.re; }
};

// ---------------------------------------------------------------------------------------------
// D2T<T>: second order in one variable over a scalar field T (double or DN<N>).
// D2 = D2T<double> is the kernel-side equivalent of the reference's Dual3.
// ---------------------------------------------------------------------------------------------
template <class T>
struct D2T {
  T re, d1, d2;
  FD D2T() {}
  FD D2T(double v) : re(v), d1(0.0), d2(0.0) {}
  FD D2T(const T& v, int) : re(v), d1(0.0), d2(0.0) {}
  FD D2T(const T& r, const T& a, const T& b) : re(r), d1(a), d2(b) {}
  static FD D2T variable(const T& v) { return D2T(v, T(1.0), T(0.0)); }
  // f(x): f0 = f, f1 = f', f2 = f''
  FD D2T chain(const T& f0, const T& f1, const T& f2) const {
    return D2T(f0, f1 * d1, f2 * (d1 * d1) + f1 * d2);
  }
  friend FD D2T operator+(const D2T& a, const D2T& b) { return D2T(a.re + b.re, a.d1 + b.d1, a.d2 + b.d2); }
  friend FD D2T operator-(const D2T& a, const D2T& b) { return D2T(a.re - b.re, a.d1 - b.d1, a.d2 - b.d2); }
  friend FD D2T operator-(const D2T& a) { return D2T(-a.re, -a.d1, -a.d2); }
  friend FD D2T operator*(const D2T& a, const D2T& b) {
    return D2T(a.re * b.re, a.d1 * b.re + a.re * b.d1, a.d2 * b.re + 2.0 * (a.d1 * b.d1) + a.re * b.d2);
  }
  friend FD D2T operator+(const D2T& a, double b) { return D2T(a.re + b, a.d1, a.d2); }
  friend FD D2T operator+(double b, const D2T& a) { return D2T(a.re + b, a.d1, a.d2); }
  friend FD D2T operator-(const D2T& a, double b) { return D2T(a.re - b, a.d1, a.d2); }
  friend FD D2T operator-(double b, const D2T& a) { return D2T(b - a.re, -a.d1, -a.d2); }
  friend FD D2T operator*(const D2T& a, double b) { return D2T(a.re * b, a.d1 * b, a.d2 * b); }
  friend FD D2T operator*(double b, const D2T& a) { return D2T(a.re * b, a.d1 * b, a.d2 * b); }
  friend FD D2T operator/(const D2T& a, double b) { return a * (1.0 / b); }
  friend FD D2T operator/(const D2T& a, const D2T& b) { return a * ad_recip(b); }
  friend FD D2T operator/(double a, const D2T& b) { return ad_recip(b) * a; }
  friend FD D2T ad_recip(const D2T& a) {
    T r = ad_recip(a.re);
    T r2 = r * r;
    return a.chain(r, -r2, 2.0 * (r2 * r));
  }
  friend FD D2T ad_exp(const D2T& a) { T f = ad_exp(a.re); return a.chain(f, f, f); }
  friend FD D2T ad_log(const D2T& a) {
    T r = ad_recip(a.re);
    return a.chain(ad_log(a.re), r, -(r * r));
  }
  friend FD D2T ad_sqrt(const D2T& a) {
    T s = ad_sqrt(a.re);
    T h = 0.5 / s;
    return a.chain(s, h, -0.5 * (h / a.re));
  }
};
using D2 = D2T<double>;

// scalar-field (T) mixed operators for D2T<DN<N>> are reached through the implicit
// D2T(const T&, int) path only explicitly: use lift() to promote a T to D2T<T>.
template <class T>
FD D2T<T> lift(const T& v) { return D2T<T>(v, 0); }

// ---------------------------------------------------------------------------------------------
// H2<T>: value, gradient (g0,g1) and Hessian (h00,h01,h11) in two variables over a field T.
// H2<double> carries exactly the information the reference extracts from its hyper-dual
// DualTensor evaluation (pcsaft_mix.py:395-420): a, mu_i = a_i, p and v_i from a_ij.
// ---------------------------------------------------------------------------------------------
template <class T>
struct H2 {
  T re, g0, g1, h00, h01, h11;
  FD H2() {}
  FD H2(double v) : re(v), g0(0.0), g1(0.0), h00(0.0), h01(0.0), h11(0.0) {}
  FD H2(const T& v, int) : re(v), g0(0.0), g1(0.0), h00(0.0), h01(0.0), h11(0.0) {}
  static FD H2 variable(const T& v, int which) {
    H2 r(v, 0);
    if (which == 0) r.g0 = T(1.0); else r.g1 = T(1.0);
    return r;
  }
  FD H2 chain(const T& f0, const T& f1, const T& f2) const {
    H2 r;
    r.re = f0;
    r.g0 = f1 * g0;
    r.g1 = f1 * g1;
    T f2g0 = f2 * g0;
    r.h00 = f2g0 * g0 + f1 * h00;
    r.h01 = f2g0 * g1 + f1 * h01;
    r.h11 = f2 * (g1 * g1) + f1 * h11;
    return r;
  }
  friend FD H2 operator+(const H2& a, const H2& b) {
    H2 r; r.re = a.re + b.re; r.g0 = a.g0 + b.g0; r.g1 = a.g1 + b.g1;
    r.h00 = a.h00 + b.h00; r.h01 = a.h01 + b.h01; r.h11 = a.h11 + b.h11; return r;
  }
  friend FD H2 operator-(const H2& a, const H2& b) {
    H2 r; r.re = a.re - b.re; r.g0 = a.g0 - b.g0; r.g1 = a.g1 - b.g1;
    r.h00 = a.h00 - b.h00; r.h01 = a.h01 - b.h01; r.h11 = a.h11 - b.h11; return r;
  }
  friend FD H2 operator-(const H2& a) {
    H2 r; r.re = -a.re; r.g0 = -a.g0; r.g1 = -a.g1; r.h00 = -a.h00; r.h01 = -a.h01; r.h11 = -a.h11; return r;
  }
  friend FD H2 operator*(const H2& a, const H2& b) {
    H2 r;
    r.re = a.re * b.re;
    r.g0 = a.g0 * b.re + a.re * b.g0;
    r.g1 = a.g1 * b.re + a.re * b.g1;
    r.h00 = a.h00 * b.re + 2.0 * (a.g0 * b.g0) + a.re * b.h00;
    r.h01 = a.h01 * b.re + a.g0 * b.g1 + a.g1 * b.g0 + a.re * b.h01;
    r.h11 = a.h11 * b.re + 2.0 * (a.g1 * b.g1) + a.re * b.h11;
    return r;
  }
  friend FD H2 operator+(const H2& a, double b) { H2 r = a; r.re = a.re + b; return r; }
  friend FD H2 operator+(double b, const H2& a) { H2 r = a; r.re = a.re + b; return r; }
  friend FD H2 operator-(const H2& a, double b) { H2 r = a; r.re = a.re - b; return r; }
  friend FD H2 operator-(double b, const H2& a) { H2 r = -a; r.re = r.re + b; return r; }
  friend FD H2 operator*(const H2& a, double b) {
    H2 r; r.re = a.re * b; r.g0 = a.g0 * b; r.g1 = a.g1 * b; r.h00 = a.h00 * b; r.h01 = a.h01 * b; r.h11 = a.h11 * b; return r;
  }
  friend FD H2 operator*(double b, const H2& a) { return a * b; }
  friend FD H2 operator/(const H2& a, double b) { return a * (1.0 / b); }
  friend FD H2 operator/(const H2& a, const H2& b) { return a * ad_recip(b); }
  friend FD H2 operator/(double a, const H2& b) { return ad_recip(b) * a; }
  friend FD H2 ad_recip(const H2& a) {
    T r = ad_recip(a.re);
    T r2 = r * r;
    return a.chain(r, -r2, 2.0 * (r2 * r));
  }
  friend FD H2 ad_exp(const H2& a) { T f = ad_exp(a.re); return a.chain(f, f, f); }
  friend FD H2 ad_log(const H2& a) {
    T r = ad_recip(a.re);
    return a.chain(ad_log(a.re), r, -(r * r));
  }
  friend FD H2 ad_sqrt(const H2& a) {
    T s = ad_sqrt(a.re);
    T h = 0.5 / s;
    return a.chain(s, h, -0.5 * (h / a.re));
  }
};

template <class T>
FD H2<T> lift2(const T& v) { return H2<T>(v, 0); }

// x^3 and small integer powers without pow()
template <class X>
FD X ad_cube(const X& x) { return x * x * x; }

}  // namespace feos
