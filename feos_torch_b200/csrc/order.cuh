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
// T_c / eps depends on m only (0.638 + 0.707 m - 0.0642 m^2 on 1 <= m <= 4).  Polar /
// associating states have a higher T_c: they are grouped by code path first and ordered by the same smooth
// function of (m, T / eps) inside their group.
//     [ code path (2) | chain length bin, 8 per octave (5) | tau, 1/512 steps on [0, 8) (12) ]
FD uint64_t pure_order_key(const double (&prm)[8], double T) {
  const double m = prm[0] < 1.0 ? 1.0 : (prm[0] > 5.5 ? 5.5 : prm[0]);
  const double tc = prm[2] * (0.6380 + m * (0.7070 - 0.06418 * m));
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
