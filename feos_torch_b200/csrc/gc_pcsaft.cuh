// Heterosegmented group-contribution PC-SAFT for binary mixtures: coefficient stage.
//
// Device-side replacement for the per-state set-up of GcPcSaftMix (/root/reference/feos_torch/
// gc_pcsaft.py:14-86 constructor tensors, :116-253 Helmholtz energy) and of the native per-state
// `GcPcSaftEosParameters::from_segments(..).phi(..)` (/root/reference/src/gc_pcsaft.rs:116-123).
//
// The reference carries dense [N,2,S] segment-count and [N,2,S,S] bond-count tensors over the whole
// segment table (S = 23: 8.5 KB per state) and loops over all S^2 segment pairs.  Here the host
// de-duplicates the molecule pairs of a batch into a small table of `GcPair` records that list only
// the segment types and bonded type pairs that are present (identical value),
// and every state refers to its pair by index.  Everything that depends on (segment parameters, k_ab,
// phi, T) is folded into the same coefficient structure the PC-SAFT binary uses (MixCoef with a bond
// list and the GC association formulas), so the density-dependent evaluation, the bubble/dew solver and
// the Newton-step pressure are shared with the PcSaftMix path.
#pragma once
#include <stdint.h>

#include "mix_grad.cuh"

namespace feos {

constexpr int kGcMaxSeg = 12;   // distinct segment types of one molecule pair
constexpr int kGcMaxBond = 16;  // distinct bonded segment-type pairs of one molecule pair

struct GcPair {
  int32_t nseg, nbond;
  int32_t seg[kGcMaxSeg];                // index into the global segment table
  double cnt[2][kGcMaxSeg];              // number of segments of that type per molecule (gc_pcsaft.py:26-30)
  int32_t ba[kGcMaxBond], bb[kGcMaxBond];  // bonded type pairs, local indices (gc_pcsaft.py:32-52)
  double bcnt[2][kGcMaxBond];            // number of such bonds per molecule
  int8_t krec[kGcMaxSeg][kGcMaxSeg];     // index of the binary segment record of (a, b) or -1
};

using GcCoef = MixCoef<double, kGcMaxBond, true>;

// Plain-double source of the inputs of the coefficient stage.
struct GcPairCoef;

struct GcSource {
  const GcPair* pair;
  const GcPairCoef* pc = nullptr;  // per-pair part of the slim coefficient stage (gc_pair_coef), if the caller has one
  const double* seg_params;  // [S, 8] global table
  const double* kab;         // [R] values of the binary segment records
  double phi_[2], T_;
  FD double seg(int a, int j) const { return seg_params[8 * pair->seg[a] + j]; }
  FD double kab_of(int a, int b) const {
    int r = pair->krec[a][b];
    return r < 0 ? 0.0 : kab[r];
  }
  FD double phi(int c) const { return phi_[c]; }
  FD double T() const { return T_; }
};

// The same inputs carrying K dual directions [first, first + K) of the flat direction list
//   8 a + j (segment a, parameter j) | 8 nseg + r (record r) | 8 nseg + R + c (phi_c) | 8 nseg + R + 2 (T)
template <int K>
struct GcDualSource {
  GcSource s;
  int ids[K];  // flat direction ids carried by this pass (-1 = unused slot)
  int nrec;
  FD DN<K> make(double v, int dir) const {
    DN<K> x(v);
#pragma unroll
    for (int k = 0; k < K; ++k) x.e[k] = (dir == ids[k]) ? 1.0 : 0.0;
    return x;
  }
  FD DN<K> seg(int a, int j) const { return make(s.seg(a, j), 8 * a + j); }
  FD DN<K> kab_of(int a, int b) const {
    int r = s.pair->krec[a][b];
    return r < 0 ? DN<K>(0.0) : make(s.kab[r], 8 * s.pair->nseg + r);
  }
  FD DN<K> phi(int c) const { return make(s.phi_[c], 8 * s.pair->nseg + nrec + c); }
  FD DN<K> T() const { return make(s.T_, 8 * s.pair->nseg + nrec + 2); }
};

// BASE = false computes the association group only (the gradients of the hard-sphere / hard-chain /
// dispersion and dipole groups are taken by the reverse sweeps gc_base_adjoint / gc_dipole_adjoint); the
// structure flags (nbond, dipolar, cat) are always set so that the flat coefficient order is the same.
// bond counts: stored in the coefficient structure, or (slim view below) read from the pair table in place
template <class P>
FD void gc_set_bn(MixCoef<P, kGcMaxBond, true>& c, int b, double n0, double n1) {
  c.bn[b][0] = P(n0);
  c.bn[b][1] = P(n1);
}

template <class P, class Src, bool BASE = true, bool ASSOC = true, class Coef = MixCoef<P, kGcMaxBond, true>>
FD void gc_coef(Coef& c, const Src& src, const GcPair& pr) {
  P T = src.T();
  P ti = ad_recip(T);
  const int ns = pr.nseg;
  P d[kGcMaxSeg];
  // per-segment diameter (gc_pcsaft.py:118-120) and zeta coefficients (:122-131)
#pragma unroll
  for (int k = 0; k < 4; ++k) c.z[k][0] = c.z[k][1] = P(0.0);
  c.m[0] = c.m[1] = P(0.0);
  for (int a = 0; a < ns; ++a) {
    P m = src.seg(a, 0), s = src.seg(a, 1), e = src.seg(a, 2);
    P da = s * (1.0 - 0.12 * ad_exp(-3.0 * (e * ti)));
    d[a] = da;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      double n = pr.cnt[i][a];
      if (n == 0.0) continue;
      P md = (kPi / 6.0 * n) * m;
      c.m[i] = c.m[i] + n * m;
      c.z[0][i] = c.z[0][i] + md;
      md = md * da;
      c.z[1][i] = c.z[1][i] + md;
      md = md * da;
      c.z[2][i] = c.z[2][i] + md;
      c.z[3][i] = c.z[3][i] + md * da;
    }
  }
  // hard chain over bonded type pairs (:156-165)
  c.nbond = pr.nbond;
  for (int b = 0; b < pr.nbond; ++b) {
    P da = d[pr.ba[b]], db = d[pr.bb[b]];
    c.cd[b] = da * db / (da + db);
    gc_set_bn(c, b, pr.bcnt[0][b], pr.bcnt[1][b]);
  }
  // dispersion (:177-194): eps_ab = sqrt(eps_a eps_b phi_i phi_j)/T (1 - k_ab [i != j])
  if (!BASE) {
#pragma unroll
    for (int k = 0; k < 3; ++k) c.q1[k] = c.q2[k] = P(0.0);
  } else {
    // sqrt(eps_a eps_b phi_i phi_j) = sqrt(eps_a) sqrt(eps_b) sqrt(phi_i phi_j): one square root per segment and
    // per (i, j) instead of three per segment pair
    P phi0 = src.phi(0), phi1 = src.phi(1);
    P sp[3] = {phi0 * ti, ad_sqrt(phi0 * phi1) * ti, phi1 * ti};  // sqrt(phi_i phi_j) / T
    P se[kGcMaxSeg];
    for (int a = 0; a < ns; ++a) {
      P e = src.seg(a, 2);
      // eps = 0 segments (">C<" of the Sauer table): the term and its derivatives are taken as zero
      // (the reference's sqrt(0) yields NaN gradients there)
      se[a] = (ad_primal(e) == 0.0) ? P(0.0) : ad_sqrt(e);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) c.q1[k] = c.q2[k] = P(0.0);
    for (int a = 0; a < ns; ++a) {
      P ma = src.seg(a, 0), sa = src.seg(a, 1);
      if (ad_primal(se[a]) == 0.0) continue;
      for (int b = 0; b < ns; ++b) {
        if (ad_primal(se[b]) == 0.0) continue;
        P mb = src.seg(b, 0), sb = src.seg(b, 1);
        P sab = 0.5 * (sa + sb);
        P ms3 = (ma * mb) * (sab * sab * sab);
        P see = se[a] * se[b];
        // ij = 00, 01, 11 (the 10 term equals the 01 term by symmetry of k_ab)
        const int ci[3] = {0, 0, 1}, cj[3] = {0, 1, 1};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double nn = pr.cnt[ci[k]][a] * pr.cnt[cj[k]][b];
          if (nn == 0.0) continue;
          P eab = see * sp[k];
          if (k == 1) eab = eab * (1.0 - src.kab_of(a, b));
          P t = (nn * ms3) * eab;
          c.q1[k] = c.q1[k] + t;
          c.q2[k] = c.q2[k] + t * eab;
        }
      }
    }
  }
  // molecule-level dipole parameters (:66-73, :255-262): mu2_term = mu2 / T
  {
    bool dipolar = false;
    P sg[2], ep[2], mu2t[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      P ms3(0.0), me(0.0), mu2(0.0);
      for (int a = 0; a < ns; ++a) {
        double n = pr.cnt[i][a];
        if (n == 0.0) continue;
        P m = src.seg(a, 0), s = src.seg(a, 1), e = src.seg(a, 2), mu = src.seg(a, 3);
        ms3 = ms3 + (n * m) * (s * s * s);
        me = me + (n * m) * e;
        mu2 = mu2 + n * (mu * mu);
        if (ad_primal(mu) != 0.0) dipolar = true;
      }
      P mi = ad_recip(c.m[i]);
      sg[i] = ad_cbrt_pos(ms3 * mi);
      ep[i] = me * mi;
      mu2t[i] = mu2 * mi * kDipoleScale * ti;
    }
    const P mm[2] = {c.m[0], c.m[1]};
    if (BASE) dipole_coef(c, mm, sg, ep, mu2t, ti, dipolar);
    else c.dipolar = dipolar;
  }
  // association: at most one associating segment per molecule (:76-86), categories (:212-251)
  if constexpr (ASSOC) {
    P sa[2], ea[2], kap[2], eab[2];
    int n_assoc = 0, n_self = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      sa[i] = ea[i] = kap[i] = eab[i] = c.na[i] = c.nb[i] = P(0.0);
      for (int a = 0; a < ns; ++a) {
        double n = pr.cnt[i][a];
        if (n == 0.0) continue;
        P k = src.seg(a, 4), e = src.seg(a, 5);
        double sign = (ad_primal(k) * ad_primal(e) > 0.0) ? 1.0 : ((ad_primal(k) * ad_primal(e) < 0.0) ? -1.0 : 0.0);
        sa[i] = sa[i] + (n * sign) * src.seg(a, 1);
        ea[i] = ea[i] + (n * sign) * src.seg(a, 2);
        kap[i] = kap[i] + n * k;
        eab[i] = eab[i] + n * e;
        c.na[i] = c.na[i] + n * src.seg(a, 6);
        c.nb[i] = c.nb[i] + n * src.seg(a, 7);
      }
      double ke = ad_primal(kap[i]) * ad_primal(eab[i]);
      c.sgn[i] = ke > 0.0 ? 1.0 : (ke < 0.0 ? -1.0 : 0.0);
      n_assoc += (ke != 0.0);
      n_self += (ad_primal(c.na[i]) * ad_primal(c.nb[i]) != 0.0);
    }
    c.cat = kAssocNone;
    if (n_assoc == 1 && n_self == 1) c.cat = kAssocSelf;
    if (n_assoc == 2 && n_self == 2) c.cat = kAssocCross;
    if (n_assoc == 2 && n_self == 1) c.cat = kAssocInduced;
    if (c.cat == kAssocSelf) {
      // :309-330 -- sigma, eps, kappa, eps_ab summed over the two molecules (one of them contributes)
      P s = sa[0] + sa[1], e = ea[0] + ea[1], k = kap[0] + kap[1], eb = eab[0] + eab[1];
      P dg = s * (1.0 - 0.12 * ad_exp(-3.0 * (e * ti)));
      c.dlt0[0] = ad_sqrt(s * s) * (s * s) * ad_sqrt(k * k) * (ad_exp(eb * ti) - 1.0);
      c.dd[0] = dg * dg / (dg + dg);
    } else if (c.cat != kAssocNone) {
      P dg[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) dg[i] = sa[i] * (1.0 - 0.12 * ad_exp(-3.0 * (ea[i] * ti)));
      const int di_[4] = {0, 0, 1, 1}, dj_[4] = {0, 1, 0, 1};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int i = di_[k], j = dj_[k];
        P ss = sa[i] * sa[j];
        c.dlt0[k] = ad_sqrt(ss) * ss * ad_sqrt(kap[i] * kap[j]) * (ad_exp((0.5 * (eab[i] + eab[j])) * ti) - 1.0);
        c.dd[k] = dg[i] * dg[j] / (dg[i] + dg[j]);
      }
    }
  }
}

// The scalar coefficients of a MixCoef in a fixed order (the flat coefficient index used by the
// gradient kernels).  FEOS_VISIT_COEF(c, X) expands X(field) for every coefficient of `c`.
#define FEOS_VISIT_COEF(c, X)                                  \
  {                                                            \
    for (int k_ = 0; k_ < 4; ++k_) {                           \
      X((c).z[k_][0]);                                         \
      X((c).z[k_][1]);                                         \
    }                                                          \
    X((c).m[0]);                                               \
    X((c).m[1]);                                               \
    for (int k_ = 0; k_ < 3; ++k_) X((c).q1[k_]);              \
    for (int k_ = 0; k_ < 3; ++k_) X((c).q2[k_]);              \
    for (int b_ = 0; b_ < (c).nbond; ++b_) X((c).cd[b_]);      \
    if ((c).dipolar) {                                         \
      X((c).mu2t[0]);                                          \
      X((c).mu2t[1]);                                          \
      for (int k_ = 0; k_ < 3; ++k_) {                         \
        X((c).pw1[k_]);                                        \
        X((c).pw2[k_]);                                        \
        X((c).pe[k_]);                                         \
        X((c).ps[k_]);                                         \
      }                                                        \
      for (int k_ = 0; k_ < 4; ++k_) {                         \
        X((c).tw1[k_]);                                        \
        X((c).tw2[k_]);                                        \
        X((c).ts[k_]);                                         \
      }                                                        \
    }                                                          \
    if ((c).cat != kAssocNone) {                               \
      int n_ = ((c).cat == kAssocSelf) ? 1 : 4;                \
      for (int k_ = 0; k_ < n_; ++k_) {                        \
        X((c).dlt0[k_]);                                       \
        X((c).dd[k_]);                                         \
      }                                                        \
      X((c).na[0]);                                            \
      X((c).na[1]);                                            \
      X((c).nb[0]);                                            \
      X((c).nb[1]);                                            \
    }                                                          \
  }
constexpr int kGcMaxCoef = 8 + 2 + 6 + kGcMaxBond + 26 + 12;

FD int coef_count(const GcCoef& c) {
  int n = 16 + c.nbond;
  if (c.dipolar) n += 26;
  if (c.cat != kAssocNone) n += ((c.cat == kAssocSelf) ? 2 : 8) + 4;
  return n;
}

// copy of the primal coefficients carrying K dual directions over the flat coefficient index
template <int K>
FD void seed_coef(MixCoef<DN<K>, kGcMaxBond, true>& cd, const GcCoef& c, int first) {
  cd.nbond = c.nbond;
  cd.dipolar = c.dipolar;
  cd.cat = c.cat;
  cd.sgn[0] = c.sgn[0];
  cd.sgn[1] = c.sgn[1];
  for (int b = 0; b < c.nbond; ++b) {
    cd.bn[b][0] = DN<K>(c.bn[b][0]);
    cd.bn[b][1] = DN<K>(c.bn[b][1]);
  }
  // walk both structures in the same order
  double vals[kGcMaxCoef];
  int idx = 0;
#define FEOS_X(f) vals[idx++] = f
  FEOS_VISIT_COEF(c, FEOS_X)
#undef FEOS_X
  idx = 0;
#define FEOS_X(f)                                                  \
  {                                                                \
    DN<K> t_(vals[idx]);                                           \
    for (int k = 0; k < K; ++k) t_.e[k] = (idx == first + k) ? 1.0 : 0.0; \
    f = t_;                                                        \
    ++idx;                                                         \
  }
  FEOS_VISIT_COEF(cd, FEOS_X)
#undef FEOS_X
}

FD_NOINLINE void gc_binary_eval(const GcCoef& c, double r0, double r1, MixEval& e) { e = mix_eval<kGcMaxBond, true>(c, r0, r1); }

struct GcBinary {
  GcCoef c;
  double msig3[2];
  FD MixEval eval(double r0, double r1) const {
    MixEval e;
    gc_binary_eval(c, r0, r1, e);
    return e;
  }
  FD double max_density(double x0) const { return kMaxEta / (x0 * msig3[0] + (1.0 - x0) * msig3[1]); }
  FD double start_density(double x0) const { return one_fluid_start_density(c, msig3, x0, max_density(x0)); }
};

FD void gc_binary_init(GcBinary& b, const GcSource& src) {
  const GcPair& pr = *src.pair;
  gc_coef<double, GcSource>(b.c, src, pr);
  b.msig3[0] = b.msig3[1] = 0.0;
  for (int a = 0; a < pr.nseg; ++a) {
    double s = src.seg(a, 1);
    double ms3 = kPi / 6.0 * src.seg(a, 0) * s * s * s;
    b.msig3[0] += pr.cnt[0][a] * ms3;
    b.msig3[1] += pr.cnt[1][a] * ms3;
  }
}

// ---- slim model for the forward kernel -----------------------------------------------------------------
// The forward kernels keep the coefficients of a state in shared memory; GcBinary (MixCoef with 16 bond slots and
// the association group: 122 doubles) allows 2 warps x 3 blocks per SM.  Pairs WITHOUT an association term and with
// at most kGcSlimBond bond types -- the usual case -- are solved by an instantiation that keeps 45 doubles per state:
//   z[4][2] m[2] q1[3] q2[3] msig3[2] flag | mu2t[2] c2[3][5] (dipole polynomials, DipolePoly) | cd[8]
// and reads the bond counts and the triplet polynomials c3 from the pair tables in place.  GcSlimView gives the record
// the member names of MixCoef, so the evaluation (mix_base_eval_impl) and the solver are the same code.
constexpr int kGcSlimBond = 8;
constexpr int kGcSlimMu = 19, kGcSlimC2 = 21, kGcSlimCd = 36;
constexpr int kGcSlimWords = (kGcSlimCd + kGcSlimBond) | 1;  // 45: odd, conflict-free as a per-thread stride

// Everything of the slim coefficient stage that does not depend on the state (T, phi): computed once per molecule
// PAIR and call (gc_pair_coef_kernel) instead of once per state.  The O(n_seg^2) dispersion sums factor as
//   q1[ij] = S1[ij] sqrt(phi_i phi_j) / T,   q2[ij] = S2[ij] (sqrt(phi_i phi_j) / T)^2
// (gc_pcsaft.py:177-194: eps_ab = sqrt(eps_a eps_b phi_i phi_j) / T (1 - k_ab)), the dipole polynomials as
//   c2[k][n] = alpha[k][n] + (epp[k] / T) beta[k][n],   c3 = constant,   mu2t_i = mu2_i / T   (DipolePoly).
struct GcPairCoef {
  double S1[3], S2[3];
  double m[2], msig3[2];
  double mu2[2], epp[3];
  double alpha[3][5], beta[3][3];
  double c3[4][4];
  double dipolar, slim;
};

FD GcSource gc_pair_source(const double* seg_params, const double* kab, const GcPair& pr) {
  GcSource src;
  src.pair = &pr;
  src.seg_params = seg_params;
  src.kab = kab;
  src.phi_[0] = src.phi_[1] = 1.0;
  src.T_ = 1.0;
  return src;
}

// dispersion sums (the loop of gc_coef with sqrt(phi_i phi_j) / T = 1) over the segment pairs (a, b) with
// (a n_seg + b) % step == first: the whole sum for (0, 1), one lane's share of it for (lane, 32)
FD void gc_pair_dispersion(const double* seg_params, const double* kab, const GcPair& pr, int first, int step,
                           double (&S1)[3], double (&S2)[3]) {
  const GcSource src = gc_pair_source(seg_params, kab, pr);
  const int ns = pr.nseg;
  for (int k = 0; k < 3; ++k) S1[k] = S2[k] = 0.0;
  for (int ab = first; ab < ns * ns; ab += step) {
    const int a = ab / ns, b = ab - a * ns;
    const double ea = src.seg(a, 2), eb = src.seg(b, 2);
    if (ea == 0.0 || eb == 0.0) continue;
    const double sab = 0.5 * (src.seg(a, 1) + src.seg(b, 1));
    const double ms3 = (src.seg(a, 0) * src.seg(b, 0)) * (sab * sab * sab);
    const double see = sqrt(ea) * sqrt(eb);
    const int ci[3] = {0, 0, 1}, cj[3] = {0, 1, 1};
    for (int k = 0; k < 3; ++k) {
      const double nn = pr.cnt[ci[k]][a] * pr.cnt[cj[k]][b];
      if (nn == 0.0) continue;
      double eab = see;
      if (k == 1) eab = eab * (1.0 - src.kab_of(a, b));
      const double t = (nn * ms3) * eab;
      S1[k] += t;
      S2[k] += t * eab;
    }
  }
}

// everything but S1, S2
FD void gc_pair_molecules(GcPairCoef& o, const double* seg_params, const double* kab, const GcPair& pr) {
  const GcSource src = gc_pair_source(seg_params, kab, pr);
  const int ns = pr.nseg;
  // molecule level: segment numbers, packing-fraction weights, dipole parameters, association test (gc_coef)
  bool dipolar = false, assoc = false;
  double sg[2], ep[2], mu2[2];
  for (int i = 0; i < 2; ++i) {
    double m = 0.0, ms3 = 0.0, me = 0.0, mu2s = 0.0, pf = 0.0, kap = 0.0, eab = 0.0;
    for (int a = 0; a < ns; ++a) {
      double n = pr.cnt[i][a];
      if (n == 0.0) continue;
      double ma = src.seg(a, 0), s = src.seg(a, 1), e = src.seg(a, 2), mu = src.seg(a, 3);
      m += n * ma;
      ms3 += (n * ma) * (s * s * s);
      me += (n * ma) * e;
      mu2s += n * (mu * mu);
      pf += n * (kPi / 6.0 * ma * s * s * s);
      kap += n * src.seg(a, 4);
      eab += n * src.seg(a, 5);
      if (mu != 0.0) dipolar = true;
    }
    if (kap * eab != 0.0) assoc = true;
    o.m[i] = m;
    o.msig3[i] = pf;
    double mi = 1.0 / m;
    sg[i] = cbrt(ms3 * mi);
    ep[i] = me * mi;
    mu2[i] = mu2s * mi * kDipoleScale;
    o.mu2[i] = mu2[i];
  }
  o.dipolar = dipolar ? 1.0 : 0.0;
  o.slim = (!assoc && pr.nbond <= kGcSlimBond) ? 1.0 : 0.0;
  for (int k = 0; k < 3; ++k) {
    o.epp[k] = 0.0;
    for (int n = 0; n < 5; ++n) o.alpha[k][n] = 0.0;
    for (int n = 0; n < 3; ++n) o.beta[k][n] = 0.0;
  }
  for (int t = 0; t < 4; ++t)
    for (int n = 0; n < 4; ++n) o.c3[t][n] = 0.0;
  if (dipolar) {
    struct {
      bool dipolar;
      double mu2t[2], pw1[3], pw2[3], pe[3], ps[3], tw1[4], tw2[4], ts[4];
    } raw;
    dipole_coef(raw, o.m, sg, ep, mu2, 1.0, true);  // with 1 / T = 1: pe = sqrt(ep_i ep_j)
    for (int k = 0; k < 3; ++k) {
      const double w1 = raw.pw1[k], w2 = raw.pw2[k], f = -kPi * raw.ps[k];
      o.epp[k] = raw.pe[k];
      for (int n = 0; n < 5; ++n) o.alpha[k][n] = f * (cAD[n][0] + w1 * cAD[n][1] + w2 * cAD[n][2]);
      for (int n = 0; n < 3; ++n) o.beta[k][n] = f * (cBD[n][0] + w1 * cBD[n][1] + w2 * cBD[n][2]);
    }
    for (int t = 0; t < 4; ++t) {
      const double w1 = raw.tw1[t], w2 = raw.tw2[t], f = (-4.0 / 3.0 * kPi * kPi) * raw.ts[t];
      for (int n = 0; n < 4; ++n) o.c3[t][n] = f * (cCD[n][0] + w1 * cCD[n][1] + w2 * cCD[n][2]);
    }
  }
}

struct GcSlimView {
  struct BnRow {
    double v0, v1;
    FD double operator[](int i) const { return i ? v1 : v0; }
  };
  struct Bn {
    const GcPair* pr;
    FD BnRow operator[](int b) const { return BnRow{pr->bcnt[0][b], pr->bcnt[1][b]}; }
  };
  double (*z)[2];
  double *m, *q1, *q2, *msig3, *cd;
  Bn bn;
  int nbond;
  bool dipolar;
  static constexpr int cat = kAssocNone;
  FD GcSlimView(double* p, const GcPair* pr_)
      : z(reinterpret_cast<double(*)[2]>(p)), m(p + 8), q1(p + 10), q2(p + 13), msig3(p + 16), cd(p + kGcSlimCd),
        bn{pr_}, nbond(pr_->nbond), dipolar(p[18] != 0.0) {}
};

// non-associating and few enough bond types (the association test of gc_coef, on the primal parameters)
FD bool gc_is_slim(const GcSource& src, const GcPair& pr) {
  if (src.pc) return src.pc->slim != 0.0;
  if (pr.nbond > kGcSlimBond) return false;
  for (int i = 0; i < 2; ++i) {
    double kap = 0.0, eab = 0.0;
    for (int a = 0; a < pr.nseg; ++a) {
      double n = pr.cnt[i][a];
      if (n == 0.0) continue;
      kap += n * src.seg(a, 4);
      eab += n * src.seg(a, 5);
    }
    if (kap * eab != 0.0) return false;
  }
  return true;
}

FD MixEval gc_slim_eval_inline(double* slot, const GcPair* pr, const GcPairCoef* pc, double r0, double r1) {
  const GcSlimView c(slot, pr);
  BaseH2 b = mix_base_eval_impl<kGcMaxBond>(c, r0, r1);
  if (c.dipolar)
    dipole_poly_eval(slot + kGcSlimMu, reinterpret_cast<const double(*)[5]>(slot + kGcSlimC2), pc->c3, c.z[3][0], c.z[3][1],
                     r0, r1, b);
  H2<double> a;
  a.re = b.a; a.g0 = b.g0; a.g1 = b.g1; a.h00 = b.h00; a.h01 = b.h01; a.h11 = b.h11;
  return mix_eval_from(a, r0, r1);
}
FD_NOINLINE void gc_slim_eval(double* slot, const GcPair* pr, const GcPairCoef* pc, double r0, double r1, MixEval& e) {
  e = gc_slim_eval_inline(slot, pr, pc, r0, r1);
}

struct GcSlim {
  double* slot;
  const GcPair* pr;
  const GcPairCoef* pc;
  FD MixEval eval(double r0, double r1) const {
    MixEval e;
    gc_slim_eval(slot, pr, pc, r0, r1, e);
    return e;
  }
  FD double max_density(double x0) const { return kMaxEta / (x0 * slot[16] + (1.0 - x0) * slot[17]); }
  FD double start_density(double x0) const {
    const double ms[2] = {slot[16], slot[17]};
    return one_fluid_start_density(GcSlimView(slot, pr), ms, x0, max_density(x0));
  }
};

FD void gc_pair_coef(GcPairCoef& o, const double* seg_params, const double* kab, const GcPair& pr) {
  gc_pair_dispersion(seg_params, kab, pr, 0, 1, o.S1, o.S2);
  gc_pair_molecules(o, seg_params, kab, pr);
}

// per-state part of the slim coefficient stage; src.pc = the pair's GcPairCoef
FD void gc_slim_init(GcSlim& b, double* slot, const GcSource& src) {
  const GcPair& pr = *src.pair;
  const GcPairCoef& pc = *src.pc;
  b.slot = slot;
  b.pr = src.pair;
  b.pc = src.pc;
  const double ti = 1.0 / src.T_;
  // per-segment diameter (gc_pcsaft.py:118-120) and zeta coefficients (:122-131)
  double d[kGcMaxSeg];
  double z[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
  for (int a = 0; a < pr.nseg; ++a) {
    const double m = src.seg(a, 0), s = src.seg(a, 1), e = src.seg(a, 2);
    const double da = s * (1.0 - 0.12 * exp(-3.0 * (e * ti)));
    d[a] = da;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const double n = pr.cnt[i][a];
      if (n == 0.0) continue;
      double md = (kPi / 6.0 * n) * m;
      z[0][i] += md;
      md *= da;
      z[1][i] += md;
      md *= da;
      z[2][i] += md;
      z[3][i] += md * da;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    slot[2 * k] = z[k][0];
    slot[2 * k + 1] = z[k][1];
  }
  slot[8] = pc.m[0];
  slot[9] = pc.m[1];
  // dispersion: the pair's sums times sqrt(phi_i phi_j) / T (and its square)
  const double sp[3] = {src.phi_[0] * ti, sqrt(src.phi_[0] * src.phi_[1]) * ti, src.phi_[1] * ti};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    slot[10 + k] = pc.S1[k] * sp[k];
    slot[13 + k] = pc.S2[k] * (sp[k] * sp[k]);
  }
  slot[16] = pc.msig3[0];
  slot[17] = pc.msig3[1];
  slot[18] = pc.dipolar;
  if (pc.dipolar != 0.0) {
    slot[kGcSlimMu] = pc.mu2[0] * ti;
    slot[kGcSlimMu + 1] = pc.mu2[1] * ti;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double pe = pc.epp[k] * ti;
#pragma unroll
      for (int n = 0; n < 5; ++n) slot[kGcSlimC2 + 5 * k + n] = n < 3 ? pc.alpha[k][n] + pe * pc.beta[k][n] : pc.alpha[k][n];
    }
  }
  // hard chain over bonded type pairs (:156-165)
  for (int bd = 0; bd < pr.nbond; ++bd) {
    const double da = d[pr.ba[bd]], db = d[pr.bb[bd]];
    slot[kGcSlimCd + bd] = da * db / (da + db);
  }
}

// ---- gradient of the Newton-step bubble/dew pressure -------------------------------------------
// Step A: adjoints of the reduced pressure with respect to the coefficients (same formula as
// mix_grad.cuh), ONE coefficient direction per pass so that D1T<DN<1>> (4 doubles per value) stays in
// registers.  The dual copy of the coefficients is built once per state (gc_dual_coef); a pass only flips
// the seed of its coefficient through the pointer table `seed`.  cbar[idx] = d p_red / d coef[idx].
using GcDualCoef = MixCoef<DN<1>, kGcMaxBond, true>;

FD int gc_dual_coef(GcDualCoef& cd, const GcCoef& c, double** seed) {
  cd.nbond = c.nbond;
  cd.dipolar = c.dipolar;
  cd.cat = c.cat;
  cd.sgn[0] = c.sgn[0];
  cd.sgn[1] = c.sgn[1];
  for (int b = 0; b < c.nbond; ++b) {
    cd.bn[b][0] = DN<1>(c.bn[b][0]);
    cd.bn[b][1] = DN<1>(c.bn[b][1]);
  }
  double vals[kGcMaxCoef];
  int idx = 0;
#define FEOS_X(f) vals[idx++] = f
  FEOS_VISIT_COEF(c, FEOS_X)
#undef FEOS_X
  idx = 0;
#define FEOS_X(f)            \
  {                          \
    f = DN<1>(vals[idx]);    \
    seed[idx] = &f.e[0];     \
    ++idx;                   \
  }
  FEOS_VISIT_COEF(cd, FEOS_X)
#undef FEOS_X
  return idx;
}

FD_NOINLINE double gc_coef_adjoint(const GcDualCoef& cd, const double (&spec)[2], const double (&inc)[2],
                                   const PointWeights& q) {
  typedef DN<1> P;
  P a_inc = mix_helmholtz<P, P, kGcMaxBond, true>(cd, P(inc[0]), P(inc[1]));
  D1T<P> r0(P(spec[0]), P(q.dir[0])), r1(P(spec[1]), P(q.dir[1]));
  D1T<P> A = mix_helmholtz<P, D1T<P>, kGcMaxBond, true>(cd, r0, r1);
  P N = a_inc * (1.0 / q.rho_inc) - A.re * q.v + A.d1 + q.cst;
  return -N.e[0] / q.denom;
}

// The same for the association part of the Helmholtz energy only (PARTS = 4): used for the coefficients
// of that group and for the zeta_2 / zeta_3 coefficients it also depends on.
FD_NOINLINE double gc_coef_adjoint_extra(const GcDualCoef& cd, const double (&spec)[2], const double (&inc)[2],
                                         const PointWeights& q) {
  typedef DN<1> P;
  P a_inc = mix_helmholtz<P, P, kGcMaxBond, true, 4>(cd, P(inc[0]), P(inc[1]));
  D1T<P> r0(P(spec[0]), P(q.dir[0])), r1(P(spec[1]), P(q.dir[1]));
  D1T<P> A = mix_helmholtz<P, D1T<P>, kGcMaxBond, true, 4>(cd, r0, r1);
  P N = a_inc * (1.0 / q.rho_inc) - A.re * q.v + A.d1;
  return -N.e[0] / q.denom;
}

// Step A as the kernels run it: the base group (hard sphere, hard chain, dispersion: flat indices
// 0 .. 15 + nbond) and the dipole group (26 coefficients) in closed form (mix_base_coef_adjoint,
// mix_dipole_coef_adjoint, mix_grad.cuh); only associating pairs pay forward-mode passes, restricted to the
// association term (PARTS = 4), for its coefficients and for z[2][i], z[3][i] (flat indices 4 .. 7), to which
// they ADD.  `cd` / `seed` are caller-provided storage for the dual coefficient copy.  Compiler note (nvcc
// 12.9, sm_100a, -O3): ptxas merges the frames of the out-of-line pass functions into the kernel's frame and
// overlapped `cd` with the caller's `cbar` when they were separate locals (the base adjoints read back as
// coefficient values).  The kernels therefore keep every array whose address is passed to a pass function
// in ONE scratch struct, which cannot be split.
FD void gc_coef_adjoints(const GcCoef& c, const double (&spec)[2], const double (&inc)[2], const PointWeights& q,
                         GcDualCoef& cd, double** seed, double* cbar) {
  mix_base_coef_adjoint<kGcMaxBond, true>(c, spec, inc, q, cbar);
  int first = 16 + c.nbond;
  if (c.dipolar) {
    mix_dipole_coef_adjoint<kGcMaxBond, true>(c, spec, inc, q, cbar, first);
    first += 26;
  }
  if (c.cat == kAssocNone) return;
  const int ncoef = gc_dual_coef(cd, c, seed);
#pragma unroll 1
  for (int f = 4; f < ncoef; ++f) {
    if (f == 8) f = first;
    *seed[f] = 1.0;
    double v = gc_coef_adjoint_extra(cd, spec, inc, q);
    *seed[f] = 0.0;
    if (f < 8) cbar[f] += v;
    else cbar[f] = v;
  }
}

FD_NOINLINE double gc_binary_value(const GcCoef& c, double r0, double r1) {
  return mix_helmholtz<double, double, kGcMaxBond, true>(c, r0, r1);
}

// K coefficient directions per pass with a freshly seeded copy (kept as the cross-check of the above)
template <int K>
FD_NOINLINE void gc_coef_adjoint_pass(const GcCoef& c, const double (&spec)[2], const double (&inc)[2], const PointWeights& q,
                             int first, int ncoef, double* cbar) {
  MixCoef<DN<K>, kGcMaxBond, true> cd;
  seed_coef<K>(cd, c, first);
  DN<K> a_inc = mix_helmholtz<DN<K>, DN<K>, kGcMaxBond, true>(cd, DN<K>(inc[0]), DN<K>(inc[1]));
  D1T<DN<K>> r0(DN<K>(spec[0]), DN<K>(q.dir[0])), r1(DN<K>(spec[1]), DN<K>(q.dir[1]));
  D1T<DN<K>> A = mix_helmholtz<DN<K>, D1T<DN<K>>, kGcMaxBond, true>(cd, r0, r1);
  DN<K> N = a_inc * (1.0 / q.rho_inc) - A.re * q.v + A.d1 + q.cst;
#pragma unroll
  for (int k = 0; k < K; ++k)
    if (first + k < ncoef) cbar[first + k] = -N.e[k] / q.denom;
}

// The input directions of the ASSOCIATION group that can have a non-zero gradient for this state (every other
// group is handled by a reverse sweep): sigma, eps of the associating segments; kappa, eps_ab, na, nb of every
// present segment (they are summed over the molecule, gc_pcsaft.py:76-86); T.
constexpr int kGcMaxDirs = 6 * kGcMaxSeg + 1;
FD int gc_active_directions(const GcSource& src, const GcCoef& c, int nrec, int* dirs) {
  const GcPair& pr = *src.pair;
  int n = 0;
  if (c.cat == kAssocNone) return 0;
  for (int a = 0; a < pr.nseg; ++a) {
    if (src.seg(a, 4) * src.seg(a, 5) != 0.0) {
      dirs[n++] = 8 * a + 1;
      dirs[n++] = 8 * a + 2;
    }
    for (int j = 4; j < 8; ++j) dirs[n++] = 8 * a + j;
  }
  dirs[n++] = 8 * pr.nseg + nrec + 2;  // T
  return n;
}

// Step B, base group: reverse-mode sweep through the part of the coefficient stage that produces
// z[k][i], m[i], cd[b], q1[k], q2[k] (flat coefficient indices 0 .. 15 + nbond).  Given the adjoints
// cbar it ADDS  sum_idx cbar[idx] d coef[idx] / d theta  to
//   gseg[8 a + j]  (j = 0, 1, 2: m, sigma, eps of local segment a),  gphi[2],  gT,  and
//   gkab[kGcMaxSeg a + b] for the k_ab of the segment pair (a in molecule 0, b in molecule 1).
// (__noinline__ like the pass functions: see the compiler note in gc_kernels.cu)
FD_NOINLINE void gc_base_adjoint(const GcSource& src, const GcCoef& c, const double* cbar, double* gseg, double* gphi,
                                 double* gT_out, double* gkab) {
  double gT = 0.0;
  const GcPair& pr = *src.pair;
  const int ns = pr.nseg;
  const double T = src.T(), ti = 1.0 / T;
  double d[kGcMaxSeg], dbar[kGcMaxSeg], ex[kGcMaxSeg];
  // z[k][i] = pi/6 sum_a n_ia m_a d_a^k (indices 2k + i), m[i] = sum_a n_ia m_a (8 + i)
  for (int a = 0; a < ns; ++a) {
    double m = src.seg(a, 0), s = src.seg(a, 1), e = src.seg(a, 2);
    ex[a] = exp(-3.0 * e * ti);
    double da = s * (1.0 - 0.12 * ex[a]);
    d[a] = da;
    double mb = 0.0, db = 0.0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      double n = pr.cnt[i][a];
      if (n == 0.0) continue;
      double w = kPi / 6.0 * n;
      double z0 = cbar[i], z1 = cbar[2 + i], z2 = cbar[4 + i], z3 = cbar[6 + i];
      mb += w * (z0 + da * (z1 + da * (z2 + da * z3))) + cbar[8 + i] * n;
      db += w * m * (z1 + da * (2.0 * z2 + 3.0 * da * z3));
    }
    gseg[8 * a] += mb;
    dbar[a] = db;
  }
  // cd[b] = d_a d_b / (d_a + d_b) (indices 16 + b)
  for (int b = 0; b < pr.nbond; ++b) {
    int ia = pr.ba[b], ib = pr.bb[b];
    double sum = d[ia] + d[ib], cb = cbar[16 + b] / (sum * sum);
    if (ia == ib) {
      dbar[ia] += cbar[16 + b] * 0.5;
    } else {
      dbar[ia] += cb * d[ib] * d[ib];
      dbar[ib] += cb * d[ia] * d[ia];
    }
  }
  // d_a = sigma_a (1 - 0.12 exp(-3 eps_a / T))
  for (int a = 0; a < ns; ++a) {
    double s = src.seg(a, 1), e = src.seg(a, 2);
    gseg[8 * a + 1] += dbar[a] * d[a] / s;
    double de = dbar[a] * s * 0.36 * ex[a] * ti;
    gseg[8 * a + 2] += de;
    gT -= de * e * ti;
  }
  // q1[k] = sum_ab nn_k m_a m_b eab s3ab (indices 10 + k), q2[k] = sum ... eab^2 (13 + k),
  // eab = sqrt(eps_a) sqrt(eps_b) sqrt(phi_i phi_j) / T (1 - k_ab [i != j]).  Reciprocals and square roots are
  // taken once per segment; the pair loop holds one division (1 / sigma_ab).
  const double phi0 = src.phi(0), phi1 = src.phi(1);
  const double sp[3] = {phi0 * ti, sqrt(phi0 * phi1) * ti, phi1 * ti};
  double se[kGcMaxSeg], im[kGcMaxSeg], ie[kGcMaxSeg], ga[kGcMaxSeg], gs[kGcMaxSeg], ge[kGcMaxSeg];
  for (int a = 0; a < ns; ++a) {
    double m = src.seg(a, 0), e = src.seg(a, 2);
    se[a] = sqrt(e);
    im[a] = 1.0 / m;
    ie[a] = e == 0.0 ? 0.0 : 0.5 / e;
    ga[a] = gs[a] = ge[a] = 0.0;
  }
  double eacc[3] = {0.0, 0.0, 0.0};  // sum of eab * (adjoint of eab) per (i, j)
  for (int a = 0; a < ns; ++a) {
    if (se[a] == 0.0) continue;
    double ma = src.seg(a, 0), sa = src.seg(a, 1);
    for (int b = 0; b < ns; ++b) {
      if (se[b] == 0.0) continue;
      double mb = src.seg(b, 0), sb = src.seg(b, 1);
      double sab = 0.5 * (sa + sb), s3 = sab * sab * sab;
      double isab = 1.5 / sab;  // d s3 / d sigma_a = 1.5 s3 / sab
      double ms3 = ma * mb * s3, see = se[a] * se[b];
      const int ci[3] = {0, 0, 1}, cj[3] = {0, 1, 1};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double nn = pr.cnt[ci[k]][a] * pr.cnt[cj[k]][b];
        if (nn == 0.0) continue;
        double kab = (k == 1) ? src.kab_of(a, b) : 0.0;
        double eab = see * sp[k] * (1.0 - kab);
        double t = nn * ms3 * eab;
        double tb = cbar[10 + k] + cbar[13 + k] * eab;       // adjoint of t
        double eb_ = (tb + cbar[13 + k] * eab) * t;          // eab * (adjoint of eab)
        double tt = tb * t;
        ga[a] += tt * im[a];
        ga[b] += tt * im[b];
        double sg = tt * isab;
        gs[a] += sg;
        gs[b] += sg;
        ge[a] += eb_ * ie[a];
        ge[b] += eb_ * ie[b];
        eacc[k] += eb_;
        if (k == 1 && pr.krec[a][b] >= 0) gkab[kGcMaxSeg * a + b] += -eb_ / (1.0 - kab);
      }
    }
  }
  for (int a = 0; a < ns; ++a) {
    gseg[8 * a] += ga[a];
    gseg[8 * a + 1] += gs[a];
    gseg[8 * a + 2] += ge[a];
  }
  gT -= (eacc[0] + eacc[1] + eacc[2]) * ti;
  // eab ~ sqrt(phi_i phi_j): d/d phi_0 = (eacc_00 + eacc_01 / 2) / phi_0, d/d phi_1 likewise
  double ppb[3] = {eacc[0], eacc[1], eacc[2]};
  gphi[0] += (ppb[0] + 0.5 * ppb[1]) / phi0;
  gphi[1] += (0.5 * ppb[1] + ppb[2]) / phi1;
  *gT_out += gT;
  (void)c;
}

// d p_red / d k_ab of the record of the segment pair (a in molecule 0, b in molecule 1) from the adjoints cb1, cb2 of
// the cross dispersion coefficients q1[01], q2[01] -- the k == 1 branch of gc_base_adjoint on its own, for the
// k_ab-only backward kernel:  t = n_0a n_1b m_a m_b sigma_ab^3 e_ab,  e_ab = sqrt(eps_a eps_b phi_0 phi_1) / T (1 - k_ab)
FD double gc_kab_term(const GcSource& src, const GcPair& pr, int a, int b, double cb1, double cb2) {
  double nn = pr.cnt[0][a] * pr.cnt[1][b];
  double sab = 0.5 * (src.seg(a, 1) + src.seg(b, 1));
  double kab = src.kab_of(a, b);
  double eab = sqrt(src.seg(a, 2)) * sqrt(src.seg(b, 2)) * (sqrt(src.phi(0) * src.phi(1)) / src.T()) * (1.0 - kab);
  double t = nn * (src.seg(a, 0) * src.seg(b, 0) * (sab * sab * sab)) * eab;
  return -((cb1 + 2.0 * cb2 * eab) * t) / (1.0 - kab);
}

// Step B, dipole group: reverse-mode sweep through the molecule-level dipole parameters (gc_coef, :146-168)
// and dipole_coef (pcsaft_mix.cuh).  cb = adjoints of the 26 dipole coefficients in the flat order
// (mu2t[2]; pw1, pw2, pe, ps per pair; tw1, tw2, ts per triplet); ADDS to gseg (m, sigma, eps, mu) and gT.
FD_NOINLINE void gc_dipole_adjoint(const GcSource& src, const GcCoef& c, const double* cb, double* gseg, double* gT_out) {
  const GcPair& pr = *src.pair;
  const int ns = pr.nseg;
  const double T = src.T(), ti = 1.0 / T;
  double m[2], ms3[2], me[2], mu2[2], sg[2], ep[2], mc[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    m[i] = c.m[i];
    ms3[i] = me[i] = mu2[i] = 0.0;
    for (int a = 0; a < ns; ++a) {
      double n = pr.cnt[i][a];
      if (n == 0.0) continue;
      double ma = src.seg(a, 0), s = src.seg(a, 1), e = src.seg(a, 2), mu = src.seg(a, 3);
      ms3[i] += n * ma * s * s * s;
      me[i] += n * ma * e;
      mu2[i] += n * mu * mu;
    }
    sg[i] = cbrt(ms3[i] / m[i]);
    ep[i] = me[i] / m[i];
    mc[i] = m[i] <= 2.0 ? m[i] : 2.0;
  }
  double mcb[2] = {0.0, 0.0}, sgb[2] = {0.0, 0.0}, epb[2] = {0.0, 0.0}, gT = 0.0;
  dipole_coef_adjoint(c, mc, sg, ep, ti, cb, mcb, sgb, epb, gT);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    double mi = 1.0 / m[i];
    double mu2t = c.mu2t[i];
    double mu2b = cb[i] * mi * kDipoleScale * ti;  // adjoint of sum n mu^2
    gT -= cb[i] * mu2t * ti;
    double mb = -cb[i] * mu2t * mi;
    if (m[i] <= 2.0) mb += mcb[i];
    double u = ms3[i] * mi;
    double ub = sgb[i] * sg[i] / (3.0 * u);  // sg = u^(1/3)
    double ms3b = ub * mi;
    mb -= ub * u * mi;
    double meb = epb[i] * mi;
    mb -= epb[i] * ep[i] * mi;
    for (int a = 0; a < ns; ++a) {
      double n = pr.cnt[i][a];
      if (n == 0.0) continue;
      double ma = src.seg(a, 0), s = src.seg(a, 1), e = src.seg(a, 2), mu = src.seg(a, 3);
      gseg[8 * a] += n * (mb + ms3b * s * s * s + meb * e);
      gseg[8 * a + 1] += ms3b * n * ma * 3.0 * s * s;
      gseg[8 * a + 2] += meb * n * ma;
      gseg[8 * a + 3] += mu2b * n * 2.0 * mu;
    }
  }
  *gT_out += gT;
}

// Step B: K input directions ids[0..K): out[k] = sum_idx cbar[idx] d coef[idx] / d direction ids[k]
// ALL = true: every coefficient (cross-check of the reverse sweeps in the CPU tests); ALL = false: only the
// association group (flat index >= 16 + nbond [+ 26 if dipolar]), the others being handled by the sweeps.
template <int K, bool ALL>
FD_NOINLINE void gc_input_grad_pass(const GcSource& src, int nrec, const int* ids, const double* cbar, double (&out)[K]) {
  GcDualSource<K> ds;
  ds.s = src;
#pragma unroll
  for (int k = 0; k < K; ++k) ds.ids[k] = ids[k];
  ds.nrec = nrec;
  MixCoef<DN<K>, kGcMaxBond, true> cd;
  gc_coef<DN<K>, GcDualSource<K>, ALL>(cd, ds, *src.pair);
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = 0.0;
  int idx = 0;
  const int skip = ALL ? 0 : 16 + cd.nbond + (cd.dipolar ? 26 : 0);
#define FEOS_X(f)                                                 \
  {                                                               \
    if (idx >= skip)                                              \
      for (int k = 0; k < K; ++k) out[k] += cbar[idx] * f.e[k];   \
    ++idx;                                                        \
  }
  FEOS_VISIT_COEF(cd, FEOS_X)
#undef FEOS_X
}

}  // namespace feos
