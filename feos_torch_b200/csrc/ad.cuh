// In-register forward-mode automatic differentiation types for the PC-SAFT kernels.
//
// Replaces the reference's Python dual-number classes:
//   Dual3      (feos_torch/dual.py:5-78)           -> D2T<T>: value, d/dx, d2/dx2 in one variable
//   DualTensor (feos_torch/dual_torch.py:4-162)    -> H2<T> : value, gradient and Hessian in the two
//                                                             partial densities of a binary mixture
//   Dual2      (feos_torch/dual_torch.py:165-208)  -> not needed: the site-fraction Newton systems
//                                                             are solved with explicit Jacobians and
//                                                             differentiated implicitly
// plus D1T<T> (value + first derivative) and DN<N> (value + N first-order parameter directions) used
// by the backward kernels.  All types live entirely in registers; every operator is forced inline.
//
// The same header compiles for the device (nvcc) and, with -DFEOS_HOST_SIM, as plain C++ for the
// CPU test harness in tests/hostsim (a debugging aid, not a product path).
#pragma once
#include <math.h>

#ifdef FEOS_HOST_SIM
#define FD inline
#define FD_NOINLINE static inline
#define FEOS_CONSTANT static const
#else
#include <cuda_runtime.h>
#define FD __device__ __forceinline__
#define FD_NOINLINE static __device__ __noinline__
#define FEOS_CONSTANT __constant__
#endif

namespace feos {

// ---------------------------------------------------------------------------------------------
// scalar helpers so that templated model code works for plain double as well
// ---------------------------------------------------------------------------------------------
FD double ad_exp(double x) { return exp(x); }
FD double ad_log(double x) { return log(x); }
FD double ad_sqrt(double x) { return sqrt(x); }
FD double ad_recip(double x) { return 1.0 / x; }
FD double ad_primal(double x) { return x; }

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
  // f(x) with f0 = f(re), f1 = f'(re)
  FD DN chain(double f0, double f1) const {
    DN r;
    r.re = f0;
#pragma unroll
    for (int i = 0; i < N; ++i) r.e[i] = f1 * e[i];
    return r;
  }
};
template <int N>
FD DN<N> operator+(const DN<N>& a, const DN<N>& b) {
  DN<N> r;
  r.re = a.re + b.re;
#pragma unroll
  for (int i = 0; i < N; ++i) r.e[i] = a.e[i] + b.e[i];
  return r;
}
template <int N>
FD DN<N> operator-(const DN<N>& a, const DN<N>& b) {
  DN<N> r;
  r.re = a.re - b.re;
#pragma unroll
  for (int i = 0; i < N; ++i) r.e[i] = a.e[i] - b.e[i];
  return r;
}
template <int N>
FD DN<N> operator*(const DN<N>& a, const DN<N>& b) {
  DN<N> r;
  r.re = a.re * b.re;
#pragma unroll
  for (int i = 0; i < N; ++i) r.e[i] = a.e[i] * b.re + a.re * b.e[i];
  return r;
}
template <int N>
FD DN<N> operator-(const DN<N>& a) {
  DN<N> r;
  r.re = -a.re;
#pragma unroll
  for (int i = 0; i < N; ++i) r.e[i] = -a.e[i];
  return r;
}
template <int N>
FD DN<N> operator+(const DN<N>& a, double b) { DN<N> r = a; r.re += b; return r; }
template <int N>
FD DN<N> operator+(double b, const DN<N>& a) { DN<N> r = a; r.re += b; return r; }
template <int N>
FD DN<N> operator-(const DN<N>& a, double b) { DN<N> r = a; r.re -= b; return r; }
template <int N>
FD DN<N> operator-(double b, const DN<N>& a) { DN<N> r = -a; r.re += b; return r; }
template <int N>
FD DN<N> operator*(const DN<N>& a, double b) {
  DN<N> r;
  r.re = a.re * b;
#pragma unroll
  for (int i = 0; i < N; ++i) r.e[i] = a.e[i] * b;
  return r;
}
template <int N>
FD DN<N> operator*(double b, const DN<N>& a) { return a * b; }
template <int N>
FD DN<N> ad_recip(const DN<N>& a) { double r = 1.0 / a.re; return a.chain(r, -r * r); }
template <int N>
FD DN<N> ad_exp(const DN<N>& a) { double f = exp(a.re); return a.chain(f, f); }
template <int N>
FD DN<N> ad_log(const DN<N>& a) { return a.chain(log(a.re), 1.0 / a.re); }
template <int N>
FD DN<N> ad_sqrt(const DN<N>& a) { double s = sqrt(a.re); return a.chain(s, 0.5 / s); }
template <int N>
FD double ad_primal(const DN<N>& a) { return a.re; }
template <int N>
FD DN<N> operator/(const DN<N>& a, double b) { return a * (1.0 / b); }
template <int N>
FD DN<N> operator/(const DN<N>& a, const DN<N>& b) { return a * ad_recip(b); }
template <int N>
FD DN<N> operator/(double a, const DN<N>& b) { return ad_recip(b) * a; }

// ---------------------------------------------------------------------------------------------
// D1T<T>: first order in one variable over a scalar field T (double or DN<N>).
// ---------------------------------------------------------------------------------------------
template <class T>
struct D1T {
  T re, d1;
  FD D1T() {}
  FD D1T(double v) : re(v), d1(0.0) {}
  FD D1T(const T& r, const T& a) : re(r), d1(a) {}
  FD D1T chain(const T& f0, const T& f1) const { return D1T(f0, f1 * d1); }
};
template <class T>
FD D1T<T> operator+(const D1T<T>& a, const D1T<T>& b) { return D1T<T>(a.re + b.re, a.d1 + b.d1); }
template <class T>
FD D1T<T> operator-(const D1T<T>& a, const D1T<T>& b) { return D1T<T>(a.re - b.re, a.d1 - b.d1); }
template <class T>
FD D1T<T> operator-(const D1T<T>& a) { return D1T<T>(-a.re, -a.d1); }
template <class T>
FD D1T<T> operator*(const D1T<T>& a, const D1T<T>& b) { return D1T<T>(a.re * b.re, a.d1 * b.re + a.re * b.d1); }
template <class T>
FD D1T<T> operator+(const D1T<T>& a, double b) { return D1T<T>(a.re + b, a.d1); }
template <class T>
FD D1T<T> operator+(double b, const D1T<T>& a) { return D1T<T>(a.re + b, a.d1); }
template <class T>
FD D1T<T> operator-(const D1T<T>& a, double b) { return D1T<T>(a.re - b, a.d1); }
template <class T>
FD D1T<T> operator-(double b, const D1T<T>& a) { return D1T<T>(b - a.re, -a.d1); }
template <class T>
FD D1T<T> operator*(const D1T<T>& a, double b) { return D1T<T>(a.re * b, a.d1 * b); }
template <class T>
FD D1T<T> operator*(double b, const D1T<T>& a) { return D1T<T>(a.re * b, a.d1 * b); }
template <class T>
FD D1T<T> ad_recip(const D1T<T>& a) {
  T r = ad_recip(a.re);
  return a.chain(r, -(r * r));
}
template <class T>
FD D1T<T> ad_exp(const D1T<T>& a) { T f = ad_exp(a.re); return a.chain(f, f); }
template <class T>
FD D1T<T> ad_log(const D1T<T>& a) { return a.chain(ad_log(a.re), ad_recip(a.re)); }
template <class T>
FD D1T<T> ad_sqrt(const D1T<T>& a) {
  T s = ad_sqrt(a.re);
  return a.chain(s, 0.5 * ad_recip(s));
}
template <class T>
FD double ad_primal(const D1T<T>& a) { return ad_primal(a.re); }
template <class T>
FD D1T<T> operator/(const D1T<T>& a, double b) { return a * (1.0 / b); }
template <class T>
FD D1T<T> operator/(const D1T<T>& a, const D1T<T>& b) { return a * ad_recip(b); }
template <class T>
FD D1T<T> operator/(double a, const D1T<T>& b) { return ad_recip(b) * a; }
// ---------------------------------------------------------------------------------------------
// D2T<T>: second order in one variable over a scalar field T.
// D2 = D2T<double> is the kernel-side equivalent of the reference's Dual3.
// ---------------------------------------------------------------------------------------------
template <class T>
struct D2T {
  T re, d1, d2;
  FD D2T() {}
  FD D2T(double v) : re(v), d1(0.0), d2(0.0) {}
  FD D2T(const T& r, const T& a, const T& b) : re(r), d1(a), d2(b) {}
  // f(x): f0 = f, f1 = f', f2 = f''
  FD D2T chain(const T& f0, const T& f1, const T& f2) const {
    return D2T(f0, f1 * d1, f2 * (d1 * d1) + f1 * d2);
  }
};
template <class T>
FD D2T<T> operator+(const D2T<T>& a, const D2T<T>& b) { return D2T<T>(a.re + b.re, a.d1 + b.d1, a.d2 + b.d2); }
template <class T>
FD D2T<T> operator-(const D2T<T>& a, const D2T<T>& b) { return D2T<T>(a.re - b.re, a.d1 - b.d1, a.d2 - b.d2); }
template <class T>
FD D2T<T> operator-(const D2T<T>& a) { return D2T<T>(-a.re, -a.d1, -a.d2); }
template <class T>
FD D2T<T> operator*(const D2T<T>& a, const D2T<T>& b) {
  return D2T<T>(a.re * b.re, a.d1 * b.re + a.re * b.d1, a.d2 * b.re + 2.0 * (a.d1 * b.d1) + a.re * b.d2);
}
template <class T>
FD D2T<T> operator+(const D2T<T>& a, double b) { return D2T<T>(a.re + b, a.d1, a.d2); }
template <class T>
FD D2T<T> operator+(double b, const D2T<T>& a) { return D2T<T>(a.re + b, a.d1, a.d2); }
template <class T>
FD D2T<T> operator-(const D2T<T>& a, double b) { return D2T<T>(a.re - b, a.d1, a.d2); }
template <class T>
FD D2T<T> operator-(double b, const D2T<T>& a) { return D2T<T>(b - a.re, -a.d1, -a.d2); }
template <class T>
FD D2T<T> operator*(const D2T<T>& a, double b) { return D2T<T>(a.re * b, a.d1 * b, a.d2 * b); }
template <class T>
FD D2T<T> operator*(double b, const D2T<T>& a) { return D2T<T>(a.re * b, a.d1 * b, a.d2 * b); }
template <class T>
FD D2T<T> ad_recip(const D2T<T>& a) {
  T r = ad_recip(a.re);
  T r2 = r * r;
  return a.chain(r, -r2, 2.0 * (r2 * r));
}
template <class T>
FD D2T<T> ad_exp(const D2T<T>& a) { T f = ad_exp(a.re); return a.chain(f, f, f); }
template <class T>
FD D2T<T> ad_log(const D2T<T>& a) {
  T r = ad_recip(a.re);
  return a.chain(ad_log(a.re), r, -(r * r));
}
template <class T>
FD D2T<T> ad_sqrt(const D2T<T>& a) {
  T s = ad_sqrt(a.re);
  T h = 0.5 * ad_recip(s);
  return a.chain(s, h, -0.5 * (h * ad_recip(a.re)));
}
template <class T>
FD double ad_primal(const D2T<T>& a) { return ad_primal(a.re); }
template <class T>
FD D2T<T> operator/(const D2T<T>& a, double b) { return a * (1.0 / b); }
template <class T>
FD D2T<T> operator/(const D2T<T>& a, const D2T<T>& b) { return a * ad_recip(b); }
template <class T>
FD D2T<T> operator/(double a, const D2T<T>& b) { return ad_recip(b) * a; }
using D2 = D2T<double>;
using D1 = D1T<double>;

// ---------------------------------------------------------------------------------------------
// H2<T>: value, gradient (g0,g1) and Hessian (h00,h01,h11) in two variables over a field T.
// H2<double> carries exactly the information the reference extracts from its hyper-dual
// DualTensor evaluation (pcsaft_mix.py:395-420): a, mu_i = a_i, and p, v_i from a_ij.
// ---------------------------------------------------------------------------------------------
template <class T>
struct H2 {
  T re, g0, g1, h00, h01, h11;
  FD H2() {}
  FD H2(double v) : re(v), g0(0.0), g1(0.0), h00(0.0), h01(0.0), h11(0.0) {}
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
};
template <class T>
FD H2<T> operator+(const H2<T>& a, const H2<T>& b) {
  H2<T> r; r.re = a.re + b.re; r.g0 = a.g0 + b.g0; r.g1 = a.g1 + b.g1;
  r.h00 = a.h00 + b.h00; r.h01 = a.h01 + b.h01; r.h11 = a.h11 + b.h11; return r;
}
template <class T>
FD H2<T> operator-(const H2<T>& a, const H2<T>& b) {
  H2<T> r; r.re = a.re - b.re; r.g0 = a.g0 - b.g0; r.g1 = a.g1 - b.g1;
  r.h00 = a.h00 - b.h00; r.h01 = a.h01 - b.h01; r.h11 = a.h11 - b.h11; return r;
}
template <class T>
FD H2<T> operator-(const H2<T>& a) {
  H2<T> r; r.re = -a.re; r.g0 = -a.g0; r.g1 = -a.g1; r.h00 = -a.h00; r.h01 = -a.h01; r.h11 = -a.h11; return r;
}
template <class T>
FD H2<T> operator*(const H2<T>& a, const H2<T>& b) {
  H2<T> r;
  r.re = a.re * b.re;
  r.g0 = a.g0 * b.re + a.re * b.g0;
  r.g1 = a.g1 * b.re + a.re * b.g1;
  r.h00 = a.h00 * b.re + 2.0 * (a.g0 * b.g0) + a.re * b.h00;
  r.h01 = a.h01 * b.re + a.g0 * b.g1 + a.g1 * b.g0 + a.re * b.h01;
  r.h11 = a.h11 * b.re + 2.0 * (a.g1 * b.g1) + a.re * b.h11;
  return r;
}
template <class T>
FD H2<T> operator+(const H2<T>& a, double b) { H2<T> r = a; r.re = a.re + b; return r; }
template <class T>
FD H2<T> operator+(double b, const H2<T>& a) { H2<T> r = a; r.re = a.re + b; return r; }
template <class T>
FD H2<T> operator-(const H2<T>& a, double b) { H2<T> r = a; r.re = a.re - b; return r; }
template <class T>
FD H2<T> operator-(double b, const H2<T>& a) { H2<T> r = -a; r.re = r.re + b; return r; }
template <class T>
FD H2<T> operator*(const H2<T>& a, double b) {
  H2<T> r; r.re = a.re * b; r.g0 = a.g0 * b; r.g1 = a.g1 * b; r.h00 = a.h00 * b; r.h01 = a.h01 * b; r.h11 = a.h11 * b; return r;
}
template <class T>
FD H2<T> operator*(double b, const H2<T>& a) { return a * b; }
template <class T>
FD H2<T> ad_recip(const H2<T>& a) {
  T r = ad_recip(a.re);
  T r2 = r * r;
  return a.chain(r, -r2, 2.0 * (r2 * r));
}
template <class T>
FD H2<T> ad_exp(const H2<T>& a) { T f = ad_exp(a.re); return a.chain(f, f, f); }
template <class T>
FD H2<T> ad_log(const H2<T>& a) {
  T r = ad_recip(a.re);
  return a.chain(ad_log(a.re), r, -(r * r));
}
template <class T>
FD H2<T> ad_sqrt(const H2<T>& a) {
  T s = ad_sqrt(a.re);
  T h = 0.5 * ad_recip(s);
  return a.chain(s, h, -0.5 * (h * ad_recip(a.re)));
}
template <class T>
FD double ad_primal(const H2<T>& a) { return ad_primal(a.re); }
template <class T>
FD H2<T> operator/(const H2<T>& a, double b) { return a * (1.0 / b); }
template <class T>
FD H2<T> operator/(const H2<T>& a, const H2<T>& b) { return a * ad_recip(b); }
template <class T>
FD H2<T> operator/(double a, const H2<T>& b) { return ad_recip(b) * a; }

// ---------------------------------------------------------------------------------------------
// Mixed arithmetic between a density-type R (D1T<P>, D2T<P>, H2<P>) and its coefficient field P.
// Model code multiplies density-dependent quantities with (parameter, T)-only coefficients; for
// P = double these are the double overloads above, for P = DN<N> they are defined here.
// ---------------------------------------------------------------------------------------------
template <int N>
FD D1T<DN<N>> operator*(const D1T<DN<N>>& a, const DN<N>& b) { return D1T<DN<N>>(a.re * b, a.d1 * b); }
template <int N>
FD D1T<DN<N>> operator*(const DN<N>& b, const D1T<DN<N>>& a) { return a * b; }
template <int N>
FD D1T<DN<N>> operator+(const D1T<DN<N>>& a, const DN<N>& b) { return D1T<DN<N>>(a.re + b, a.d1); }
template <int N>
FD D1T<DN<N>> operator+(const DN<N>& b, const D1T<DN<N>>& a) { return a + b; }
template <int N>
FD D2T<DN<N>> operator*(const D2T<DN<N>>& a, const DN<N>& b) { return D2T<DN<N>>(a.re * b, a.d1 * b, a.d2 * b); }
template <int N>
FD D2T<DN<N>> operator*(const DN<N>& b, const D2T<DN<N>>& a) { return a * b; }
template <int N>
FD D2T<DN<N>> operator+(const D2T<DN<N>>& a, const DN<N>& b) { return D2T<DN<N>>(a.re + b, a.d1, a.d2); }
template <int N>
FD D2T<DN<N>> operator+(const DN<N>& b, const D2T<DN<N>>& a) { return a + b; }
template <int N>
FD H2<DN<N>> operator*(const H2<DN<N>>& a, const DN<N>& b) {
  H2<DN<N>> r; r.re = a.re * b; r.g0 = a.g0 * b; r.g1 = a.g1 * b; r.h00 = a.h00 * b; r.h01 = a.h01 * b; r.h11 = a.h11 * b; return r;
}
template <int N>
FD H2<DN<N>> operator*(const DN<N>& b, const H2<DN<N>>& a) { return a * b; }
template <int N>
FD H2<DN<N>> operator+(const H2<DN<N>>& a, const DN<N>& b) { H2<DN<N>> r = a; r.re = a.re + b; return r; }
template <int N>
FD H2<DN<N>> operator+(const DN<N>& b, const H2<DN<N>>& a) { return a + b; }

// torch.clamp(max=c): passes the derivative where x <= c (pcsaft_pure.py:146)
template <class X>
FD X ad_clamp_max(const X& x, double c) { return (ad_primal(x) <= c) ? x : X(c); }

}  // namespace feos
