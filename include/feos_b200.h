/* feos_b200 -- C ABI of the B200-native batched PC-SAFT phase-equilibrium hot path.
 *
 * Drop-in boundary for the native layer of feos-torch (Rust/PyO3, /root/reference/src/):
 * every entry point below replaces one function a maintainer's FFI for this path would bind.
 *
 * Conventions
 *   - plain pointers and sizes only; no exceptions cross the ABI.  Return value 0 = success,
 *     negative = error (argument / CUDA runtime / launch); feos_b200_last_error() gives the text
 *     (thread local).
 *   - `feos_dev_*`  : all array pointers are DEVICE pointers (current device), work is enqueued on
 *                     `stream` (a cudaStream_t passed as void*, NULL = legacy default stream) and
 *                     the call returns without synchronising.  Outputs are FULL LENGTH (n rows):
 *                     rows of failed states hold NaN and `failed[i] = 1`.
 *   - `feos_host_*` : array pointers are HOST pointers with exactly the reference's native
 *                     signature semantics (compacted outputs, bool mask); the call copies to the
 *                     device, runs the kernels, compacts, copies back and synchronises.
 *   - parameter rows are [m, sigma/A, epsilon_k/K, mu/D, kappa_ab, epsilon_k_ab/K, na, nb]
 *     (README.md:12), densities are number densities in 1/Angstrom^3, T in K, pressures in Pa.
 *   - status / failed: 1 = the solver failed for this state (reference: `status = true`,
 *     src/pcsaft.rs:93).
 */
#ifndef FEOS_B200_H
#define FEOS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------ */
/* library                                                                                     */
/* ------------------------------------------------------------------------------------------ */
const char* feos_b200_version(void);    /* replaces __version__ registered at src/lib.rs:13 */
const char* feos_b200_last_error(void); /* text of the last error on this thread ("" if none) */
/* number of kernel launches issued by this process through the library (bench accounting) */
int64_t feos_b200_launch_count(void);
/* FP64 FMA-pipe peak of the current device in TFLOP/s from a register-resident DFMA chain kernel on
 * all SMs (bench.py roofline denominator; blocks until measured). */
int feos_b200_measure_fp64_peak(int iters, double* tflops);

/* Order-preserving removal of the rows of failed states (the compaction of src/pcsaft.rs:94-101,
 * :127-128 and filter_binary :216-231, and of the mask indexing at pcsaft_pure.py:192-194).
 *   failed [n]                                in
 *   pos [n]     row i -> output row, -1 if failed (may be NULL)            out
 *   n_ok        number of converged states: a device scalar, or a word of pinned (mapped) host memory that
 *               the caller polls instead of synchronising the stream                          out
 *   in [n,in_cols] -> out [n_ok,out_cols], columns >= in_cols are zero (in may be NULL) */
int feos_dev_compact(const uint8_t* failed, int64_t n, int32_t* pos, int64_t* n_ok, const double* in,
                     int in_cols, double* out, int out_cols, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* pure component: PcSaft.vapor_pressure / PcSaft.liquid_density                               */
/* ------------------------------------------------------------------------------------------ */

/* State order of the iterative (forward) kernels.  One thread solves one state, so a warp runs as long as its
 * slowest lane; the forward kernels take an optional permutation `order` [n] (thread t works on state order[t],
 * NULL = input order) that puts states with the same code path (dipole / association terms, GC molecule pair)
 * and neighbouring reduced temperature (and composition) into the same warp: a key kernel + one radix sort of
 * (key, index) pairs (csrc/order.cuh).  Results do not depend on the order.  No counterpart in the reference
 * (rayon schedules states one by one, src/pcsaft.rs:86-92); the feos_host_* entry points apply it themselves
 * to batches of >= 16384 states.
 *   params [n,8], temperature [n]     in
 *   order [n]   permutation of 0 .. n-1, keyed on (code path, chain length bin, T / T_c estimate)   out */
int feos_dev_pure_order(const double* params, const double* temperature, int64_t n, int32_t* order, void* stream);

/* Pure-component vapour-liquid equilibrium at T.
 * Replaces PyPcSaft::vapor_pressure -> vapor_pressure_ (src/pcsaft.rs:19-26, :82-103) and fuses the
 * Newton-step property formulas of PcSaftPure.vapor_pressure / equilibrium_liquid_density
 * (feos_torch/pcsaft_pure.py:201-233).
 *   params [n,8], temperature [n]                       in
 *   rho [n,2]   (rho_V, rho_L) in 1/A^3                  out
 *   vapor_pressure [n] in Pa            (may be NULL)    out
 *   liquid_density [n] in kmol/m^3      (may be NULL)    out  (saturated liquid density)
 *   failed [n]                                           out
 *   evals [n]  Helmholtz evaluations used (may be NULL)  out  (roofline accounting)
 *   order [n]  state order from feos_dev_pure_order (may be NULL)  in */
int feos_dev_pure_vle_fwd(const double* params, const double* temperature, int64_t n, double* rho,
                          double* vapor_pressure, double* liquid_density, uint8_t* failed, int32_t* evals,
                          const int32_t* order, void* stream);

/* Gradient of  sum_i g_vp[i] * vapor_pressure[i] + g_rho[i] * liquid_density[i]  with respect to the
 * parameters and the temperature at the converged densities (implicit-function theorem); replaces
 * torch autograd through pcsaft_pure.py:210-215 / :226-233.  g_vp or g_rho may be NULL (= 0).
 * `pos` (from feos_dev_compact, may be NULL) maps state i to its row in the COMPACTED g arrays;
 * with pos == NULL the g arrays are full length.  g_broadcast != 0 (all *_bwd entry points): the g arrays hold ONE
 * value that applies to every state -- the upstream gradient of `loss = values.sum()`, which torch hands over as
 * a stride-0 expansion; reading it in place saves materialising it.
 *   grad_params [n,8], grad_temperature [n] (may be NULL)   out; rows of failed states are 0 */
int feos_dev_pure_vle_bwd(const double* params, const double* temperature, const double* rho,
                          const uint8_t* failed, const int32_t* pos, const double* g_vp, const double* g_rho,
                          int g_broadcast, int64_t n, double* grad_params, double* grad_temperature, void* stream);

/* Liquid density at (T, p).  Replaces PyPcSaft::liquid_density -> liquid_density_
 * (src/pcsaft.rs:29-41, :105-129) fused with pcsaft_pure.py:196-199.
 *   rho [n] converged density in 1/A^3, liquid_density [n] in kmol/m^3 (may be NULL) */
int feos_dev_pure_rho_tp_fwd(const double* params, const double* temperature, const double* pressure,
                             int64_t n, double* rho, double* liquid_density, uint8_t* failed, int32_t* evals,
                             const int32_t* order, void* stream);

/* Gradient of sum_i g[i] * liquid_density[i] w.r.t. parameters, T and p (any output may be NULL);
 * `pos` as above. */
int feos_dev_pure_rho_tp_bwd(const double* params, const double* temperature, const double* pressure,
                             const double* rho, const uint8_t* failed, const int32_t* pos, const double* g,
                             int g_broadcast, int64_t n, double* grad_params, double* grad_temperature,
                             double* grad_pressure, void* stream);

/* a, p/kT, d(p/kT)/drho at given (T, rho): PcSaftPure.derivatives (pcsaft_pure.py:180-182);
 * any output may be NULL. */
int feos_dev_pure_state(const double* params, const double* temperature, const double* density, int64_t n,
                        double* a, double* p, double* dp, void* stream);

/* Gradient of  sum_i g_a[i] a[i] + g_p[i] p[i] + g_dp[i] dp[i]  of feos_dev_pure_state with respect to the
 * parameters [n,8], T [n] and the density [n]; replaces torch autograd through the reference's Dual3
 * evaluation (pcsaft_pure.py:180-182, feos_torch/dual.py).  Any g_* may be NULL (= 0, not all), any
 * output may be NULL. */
int feos_dev_pure_state_bwd(const double* params, const double* temperature, const double* density, int64_t n,
                            const double* g_a, const double* g_p, const double* g_dp, double* grad_params,
                            double* grad_temperature, double* grad_density, void* stream);

/* Host-buffer entry points with the reference's native signatures.
 * PcSaft.vapor_pressure(parameters f64[n,8], temperature f64[n]) -> (f64[n_ok,4], bool[n])
 * (src/pcsaft.rs:19-26): rho_out must hold n*4 doubles, the first *n_ok rows are filled with
 * (rho_V, rho_L, 0, 0) of the converged states in input order (src/pcsaft.rs:94-101). */
int feos_host_pcsaft_vapor_pressure(const double* params, const double* temperature, int64_t n,
                                    double* rho_out, uint8_t* status, int64_t* n_ok);
/* PcSaft.liquid_density(parameters, temperature, pressure) -> (f64[n_ok], bool[n]) (src/pcsaft.rs:29-41) */
int feos_host_pcsaft_liquid_density(const double* params, const double* temperature, const double* pressure,
                                    int64_t n, double* rho_out, uint8_t* status, int64_t* n_ok);

/* ------------------------------------------------------------------------------------------ */
/* binary mixture with k_ij: PcSaft.bubble_point / PcSaft.dew_point                            */
/* ------------------------------------------------------------------------------------------ */

#define FEOS_SHARED_PARAMS 1
#define FEOS_SHARED_KIJ 2
/* the caller asserts that NO state of the batch has a dipole or association term (mu = na = nb = 0 in every row):
 * the instantiation for such pairs is not launched at all (it would find nothing to do; the launch costs ~20 us of
 * a 2.4 ms step).  States that break the promise are left unsolved. */
#define FEOS_PLAIN_ONLY 4

/* State order of the bubble / dew kernels (see feos_dev_pure_order), keyed on (code path, T / eps_bar bin, x).
 *   params [n,2,8] (or one [2,8] row with FEOS_SHARED_PARAMS in `flags`), temperature [n], molefracs [n]      in
 *   order [n]   permutation of 0 .. n-1                  out */
int feos_dev_mix_point_order(const double* params, const double* temperature, const double* molefracs, int64_t n,
                             int flags, int32_t* order, void* stream);
/*   pair_index [n] (state -> molecule pair, 0 <= pair_index < n_pairs), temperature [n], molefracs [n];
 *   pair_only != 0: grouped by molecule pair, input order inside a pair (for the backward kernel, which runs
 *   no iteration: only the code path matters and neighbouring states keep its memory accesses local) */
int feos_dev_gc_point_order(const int32_t* pair_index, int32_t n_pairs, const double* temperature,
                            const double* molefracs, int64_t n, int pair_only, int32_t* order, void* stream);

/* Binary bubble (bubble = 1, `molefracs` = liquid x_1) or dew (bubble = 0, vapour y_1) point at T.
 * Replaces PyPcSaft::bubble_point / dew_point -> bubble_point_ / dew_point_ (src/pcsaft.rs:44-79,
 * :150-214) and fuses the Newton-step pressure of PcSaftMix.bubble_point / dew_point
 * (feos_torch/pcsaft_mix.py:422-468).
 *   params [n,2,8], kij [n,2] = (k_ij, eps_AiBj; 0 = combining rule), temperature [n], molefracs [n]
 *   rho [n,4] = (rho_V1, rho_V2, rho_L1, rho_L2) in 1/A^3 (filter_binary order, src/pcsaft.rs:225-228)
 *   start_pressure [n] in Pa (may be NULL = 1e5 Pa): the `pressure` argument of the reference's bubble_point /
 *     dew_point, handed to feos as Some(p0) (src/pcsaft.rs:170-177, :203-210).  The fast path builds its own start
 *     (zero-pressure liquid + ideal vapour); the retry path starts its successive substitution from p0 where no such
 *     start exists (near-critical states), as the reference's solver starts from it.
 *   flags: FEOS_SHARED_PARAMS = `params` is ONE [2,8] row shared by all states, FEOS_SHARED_KIJ = `kij` is one [2]
 *     row (fit of shared parameters, BASELINE config C4: the reference expands them to [n,2,8] / [n,2])
 *   pressure [n] in Pa (may be NULL), failed [n], evals [n] (may be NULL), order [n] (may be NULL, see above) */
int feos_dev_mix_point_fwd(const double* params, const double* kij, const double* temperature,
                           const double* molefracs, const double* start_pressure, int64_t n, int bubble, int flags,
                           double* rho, double* pressure, uint8_t* failed, int32_t* evals, const int32_t* order,
                           void* stream);

/* Gradient of sum_i g[i] * pressure[i] w.r.t. params [n,2,8], kij [n,2] and T [n] (any may be NULL);
 * replaces torch autograd through pcsaft_mix.py:435-444 / :459-468.  `pos` as for the pure path.
 * With FEOS_SHARED_PARAMS / FEOS_SHARED_KIJ in `flags`, grad_params is [2,8] / grad_kij is [2] and receives the SUM
 * over the states, ACCUMULATED into (zero first): the reduction torch would run over an expanded [n,2,8] gradient
 * is fused into the kernel (warp shuffle -> shared memory -> one atomic per slot and block).
 * With grad_kij alone (grad_params = grad_temperature = NULL: a k_ij fit) the k_ij column comes from the two dispersion
 * partials d a / d Q1, d a / d Q2 (k_ij enters through the cross dispersion coefficients only) instead of the full
 * coefficient adjoint + reverse sweep; the eps_AiBj column keeps its dual pass where a cross-associating pair uses it. */
int feos_dev_mix_point_bwd(const double* params, const double* kij, const double* temperature, const double* rho,
                           const uint8_t* failed, const int32_t* pos, const double* g, int g_broadcast, int64_t n,
                           int bubble, int flags, double* grad_params, double* grad_kij, double* grad_temperature,
                           const int32_t* order, void* stream);

/* a, p/kT, mu_i^res/kT [n,2], v_i [n,2] at (T, rho [n,2]): PcSaftMix.derivatives (pcsaft_mix.py:395-420). */
int feos_dev_mix_state(const double* params, const double* kij, const double* temperature, const double* density,
                       int64_t n, double* a, double* p, double* mu, double* v, void* stream);

/* Gradient of  sum_i g_a a + g_p p + g_mu . mu + g_v . v  of feos_dev_mix_state w.r.t. params [n,2,8], kij [n,2],
 * T [n] and rho [n,2]; replaces torch autograd through the reference's hyper-dual evaluation
 * (pcsaft_mix.py:395-420, feos_torch/dual_torch.py).  g_mu, g_v are [n,2]; NULL conventions as above. */
int feos_dev_mix_state_bwd(const double* params, const double* kij, const double* temperature, const double* density,
                           int64_t n, const double* g_a, const double* g_p, const double* g_mu, const double* g_v,
                           double* grad_params, double* grad_kij, double* grad_temperature, double* grad_density,
                           void* stream);

/* PcSaft.bubble_point / dew_point(parameters f64[n,2,8], kij f64[n,2], temperature, molefracs, pressure)
 * -> (f64[n_ok,4], bool[n]) with host buffers (src/pcsaft.rs:44-79).  `pressure` = the reference's starting
 * value (may be NULL), see `start_pressure` of feos_dev_mix_point_fwd. */
int feos_host_pcsaft_bubble_point(const double* params, const double* kij, const double* temperature,
                                  const double* liquid_molefracs, const double* pressure, int64_t n, double* rho_out,
                                  uint8_t* status, int64_t* n_ok);
int feos_host_pcsaft_dew_point(const double* params, const double* kij, const double* temperature,
                               const double* vapor_molefracs, const double* pressure, int64_t n, double* rho_out,
                               uint8_t* status, int64_t* n_ok);

/* ------------------------------------------------------------------------------------------ */
/* group-contribution PC-SAFT: GcPcSaft(...).bubble_point / dew_point                          */
/* ------------------------------------------------------------------------------------------ */

/* One de-duplicated molecule pair of a batch: the segment types present in either molecule and the
 * bonded type pairs, with per-molecule counts.  Replaces the per-state ChemicalRecord pair and the
 * per-state `GcPcSaftEosParameters::from_segments` of src/gc_pcsaft.rs:25-69, :116-123 (and the dense
 * [N,2,S] / [N,2,S,S] count tensors of feos_torch/gc_pcsaft.py:26-52). */
#define FEOS_GC_MAX_SEG 12
#define FEOS_GC_MAX_BOND 16
typedef struct feos_gc_pair {
  int32_t nseg, nbond;
  int32_t seg[FEOS_GC_MAX_SEG];                    /* rows of the segment table */
  double cnt[2][FEOS_GC_MAX_SEG];                  /* segments of that type in molecule 0 / 1 */
  int32_t ba[FEOS_GC_MAX_BOND], bb[FEOS_GC_MAX_BOND]; /* bonded type pairs (local segment indices) */
  double bcnt[2][FEOS_GC_MAX_BOND];                /* bonds of that type in molecule 0 / 1 */
  int8_t krec[FEOS_GC_MAX_SEG][FEOS_GC_MAX_SEG];   /* binary segment record of (a, b), -1 = none */
} feos_gc_pair;

/* Common inputs of the GC entry points:
 *   pairs [n_pairs], seg_params [n_seg,8] (rows as the PC-SAFT parameter rows), kab [n_rec] values of
 *   the binary segment records, pair_index [n] (state -> pair), phi [n,2] (may be NULL = 1),
 *   temperature [n].  Device pointers for feos_dev_*, host pointers for feos_host_*.
 *   order [n] (may be NULL): state order from feos_dev_gc_point_order, see feos_dev_mix_point_order.
 *   start_pressure [n] (may be NULL): as for feos_dev_mix_point_fwd (src/gc_pcsaft.rs:125-132, :160-167).
 *   flags: 0, or FEOS_PLAIN_ONLY = the caller asserts that no molecule pair has an association term and none has more
 *     than 8 bond types (every pair is of the "slim" class): the general instantiation is not launched.
 * Replaces PyGcPcSaft::bubble_point / dew_point -> bubble_point_ / dew_point_ (src/gc_pcsaft.rs:71-99,
 * :103-171) fused with GcPcSaftMix.bubble_point / dew_point (feos_torch/gc_pcsaft.py:470-512). */
int feos_dev_gc_point_fwd(const feos_gc_pair* pairs, int32_t n_pairs, const double* seg_params, int32_t n_seg,
                          const double* kab, int32_t n_rec, const int32_t* pair_index, const double* phi,
                          const double* temperature, const double* molefracs, const double* start_pressure, int64_t n,
                          int bubble, int flags, double* rho, double* pressure, uint8_t* failed, int32_t* evals,
                          const int32_t* order, void* stream);

/* Gradient of sum_i g[i] * pressure[i].  grad_seg_params [n_seg,8] and grad_kab [n_rec] are shared by
 * all states and are ACCUMULATED into (zero them first); grad_phi [n,2] and grad_temperature [n] are
 * written.  Any output may be NULL.  With grad_kab alone the gradient comes from the two dispersion partials
 * (as for feos_dev_mix_point_bwd with grad_kij alone).
 * (Both GC entry points first compute the state-independent part of the coefficient stage once per molecule pair --
 * one small kernel over n_pairs -- so n_pairs should be much smaller than n: de-duplicate the pairs of a batch.) */
int feos_dev_gc_point_bwd(const feos_gc_pair* pairs, int32_t n_pairs, const double* seg_params, int32_t n_seg,
                          const double* kab, int32_t n_rec, const int32_t* pair_index, const double* phi,
                          const double* temperature, const double* rho, const uint8_t* failed, const int32_t* pos,
                          const double* g, int g_broadcast, int64_t n, int bubble, double* grad_seg_params,
                          double* grad_kab, double* grad_phi, double* grad_temperature, const int32_t* order,
                          void* stream);

/* a, p/kT, mu_i, v_i at (T, rho [n,2]): GcPcSaftMix.derivatives (gc_pcsaft.py:443-468). */
int feos_dev_gc_state(const feos_gc_pair* pairs, int32_t n_pairs, const double* seg_params, int32_t n_seg,
                      const double* kab, int32_t n_rec, const int32_t* pair_index, const double* phi,
                      const double* temperature, const double* density, int64_t n, double* a, double* p, double* mu,
                      double* v, void* stream);

/* Gradient of the outputs of feos_dev_gc_state (weights as feos_dev_mix_state_bwd); replaces torch autograd
 * through gc_pcsaft.py:443-468.  grad_seg_params [n_seg,8] and grad_kab [n_rec] are ACCUMULATED into (zero them
 * first); grad_phi [n,2], grad_temperature [n], grad_density [n,2] are written.  Any output may be NULL. */
int feos_dev_gc_state_bwd(const feos_gc_pair* pairs, int32_t n_pairs, const double* seg_params, int32_t n_seg,
                          const double* kab, int32_t n_rec, const int32_t* pair_index, const double* phi,
                          const double* temperature, const double* density, int64_t n, const double* g_a,
                          const double* g_p, const double* g_mu, const double* g_v, double* grad_seg_params,
                          double* grad_kab, double* grad_phi, double* grad_temperature, double* grad_density,
                          void* stream);

/* GcPcSaft.bubble_point / dew_point with host buffers -> (f64[n_ok,4], bool[n]) (src/gc_pcsaft.rs:71-99). */
int feos_host_gc_point(const feos_gc_pair* pairs, int32_t n_pairs, const double* seg_params, int32_t n_seg,
                       const double* kab, int32_t n_rec, const int32_t* pair_index, const double* phi,
                       const double* temperature, const double* molefracs, const double* pressure, int64_t n,
                       int bubble, double* rho_out, uint8_t* status, int64_t* n_ok);

#ifdef __cplusplus
}
#endif
#endif /* FEOS_B200_H */
