// Sort keys for the bubble / dew point kernels' state order (include/feos_b200.h: feos_dev_*_point_order).
//
// One thread solves one state, so a warp runs as long as its slowest lane: with states in arbitrary order
// the forward kernels keep 16 of 32 lanes busy per issued instruction (profiles/README.md).  The number of
// evaluations a state needs varies smoothly with its reduced temperature and composition, and the code
// path (dipole term, association term, segment / bond loops of a GC pair) depends on the parameters only.
// The kernels therefore take an optional permutation `order` (thread t works on state order[t]) that puts
// states with the same code path and neighbouring (T, x) into the same warp.  The key is
//     [ code path | temperature bin (64 per octave) | mole fraction (12 bits) ]
// and the permutation is one radix sort of (key, index) pairs.  Results do not depend on the order: every
// state is computed by the same instructions either way.
#pragma once
#include <math.h>
#include <stdint.h>

#include "ad.cuh"

namespace feos {

constexpr int kOrderFracBits = 12;
constexpr int kOrderTempBits = 9;
constexpr int kOrderStateBits = kOrderFracBits + kOrderTempBits;

// floor(64 log2 x) + offset, clamped to the 9-bit range (x <= 0 or NaN -> 0)
FD uint32_t order_log_bin(double x, int offset) {
  if (!(x > 0.0) || !(x < 1e300)) return 0u;
  int b = (int)floor(64.0 * log2(x)) + offset;
  return (uint32_t)(b < 0 ? 0 : (b > 511 ? 511 : b));
}

FD uint32_t order_frac_bin(double z) {
  if (!(z > 0.0)) return 0u;
  if (!(z < 1.0)) return (1u << kOrderFracBits) - 1u;
  return (uint32_t)(z * (double)(1u << kOrderFracBits));
}

// Binary PC-SAFT state: params [2][8] rows (m, sigma, eps, mu, kappa, eps_ab, na, nb), T, x_1 of the specified phase.
// Reduced temperature T / eps_bar with the segment-fraction average of eps_i; T / eps in 1/8 .. 8 covers 6 octaves.
FD uint64_t mix_order_key(const double (&prm)[16], double T, double z) {
  const double w0 = z * prm[0], w1 = (1.0 - z) * prm[8];
  const double eps = (w0 * prm[2] + w1 * prm[10]) / (w0 + w1);
  uint32_t path = 0;
  if (prm[3] != 0.0 || prm[11] != 0.0) path |= 1u;          // dipole term present
  if (prm[6] + prm[7] > 0.0) path |= 2u;                    // association sites on component 1
  if (prm[14] + prm[15] > 0.0) path |= 4u;                  // ... on component 2
  return ((uint64_t)path << kOrderStateBits) | ((uint64_t)order_log_bin(T / eps, 256) << kOrderFracBits) |
         (uint64_t)order_frac_bin(z);
}
constexpr int kMixOrderBits = kOrderStateBits + 3;

// Pure-component state: params row (m, sigma, eps, mu, kappa, eps_ab, na, nb), T.  The evaluation counts of the
// VLE and density iterations follow the reduced temperature tau = T / T_c; for non-associating, non-polar PC-SAFT
// T_c / eps depends on m only (0.638 + 0.707 m - 0.0642 m^2 on 1 <= m <= 4); association
// raises it by up to 2.8x (order_assoc_tc_shift below).  Polar states are grouped by code path and ordered by
// the same smooth function of (m, T / eps) inside their group.
//     [ code path (2) | chain length bin, 8 per octave (5) | tau, 1/512 steps on [0, 8) (12) ]
// ln(T_c / T_c0) of a 2B-associating fluid: cubic in a = 1/m, b = ln kappa_AB, c = eps_AB / (10 eps), fitted
// (least squares, rms 1.6 %, max 8 %) to the critical temperatures of 6000 random parameter sets on 1 <= m <= 6,
// 1e-3 <= kappa <= 0.3, 3 <= eps_AB / eps <= 20, located by bisection on the converged / failed mask of the
// VLE solver itself (T_c / T_c0 = 1 .. 2.8 there).  Only the ORDER of the states depends on it.
FD double order_assoc_tc_shift(double m, double kappa, double eps_ratio) {
  const double w[20] = {1.307605e-01, -1.411176e-02, 1.187721e-01, -3.571486e-02, -1.378836e-02,
                        7.843463e-02, -2.414582e-02, 8.443805e-03, -3.136799e-03, -3.472380e-05,
                        -8.881037e-01, 1.178740e+00, -1.917897e-01, 2.525275e-01, -4.864911e-02,
                        -1.632931e-04, 2.552692e+00, -4.434713e-01, -1.037485e-01, -1.479482e+00};
  const double a = 1.0 / (m < 1.0 ? 1.0 : (m > 6.0 ? 6.0 : m));
  const double b = log(kappa < 1e-3 ? 1e-3 : (kappa > 0.3 ? 0.3 : kappa));
  const double c = 0.1 * (eps_ratio < 3.0 ? 3.0 : (eps_ratio > 20.0 ? 20.0 : eps_ratio));
  const double pa[4] = {1.0, a, a * a, a * a * a}, pb[4] = {1.0, b, b * b, b * b * b}, pc[4] = {1.0, c, c * c, c * c * c};
  double y = 0.0;
  int q = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; i + j < 4; ++j)
      for (int k = 0; i + j + k < 4; ++k) y += w[q++] * pa[i] * pb[j] * pc[k];
  return y < 0.0 ? 0.0 : (y > 1.2 ? 1.2 : y);
}

FD uint64_t pure_order_key(const double (&prm)[8], double T) {
  const double m = prm[0] < 1.0 ? 1.0 : (prm[0] > 5.5 ? 5.5 : prm[0]);
  double tc = prm[2] * (0.6380 + m * (0.7070 - 0.06418 * m));
  if (prm[4] > 0.0 && prm[5] > 0.0 && prm[6] + prm[7] > 0.0 && prm[2] > 0.0)
    tc *= exp(order_assoc_tc_shift(prm[0], prm[4], prm[5] / prm[2]));
  double tau = T / tc;
  uint32_t tq = !(tau > 0.0) ? 0u : (tau < 7.99 ? (uint32_t)(tau * 512.0) : 4095u);
  uint32_t mb = 0;
  if (prm[0] > 1.0) {
    int b = (int)(8.0 * log2(prm[0]));
    mb = (uint32_t)(b > 31 ? 31 : b);
  }
  uint32_t path = 0;
  if (prm[3] != 0.0) path |= 1u;
  if (prm[4] != 0.0 && prm[5] != 0.0 && prm[6] + prm[7] > 0.0) path |= 2u;
  return ((uint64_t)path << 17) | ((uint64_t)mb << 12) | (uint64_t)tq;
}
constexpr int kPureOrderBits = 19;

// GC state: the molecule pair fixes parameters and code path; T in 16 K .. 4096 K covers 8 octaves.
FD uint64_t gc_order_key(int32_t pair, double T, double z) {
  return ((uint64_t)(uint32_t)pair << kOrderStateBits) | ((uint64_t)order_log_bin(T, -256) << kOrderFracBits) |
         (uint64_t)order_frac_bin(z);
}

}  // namespace feos
