/*
 * lsspa.h -- C ABI of the B200-native LS-SPA hot path (libls_spa_b200.so).
 *
 * The reference (cvxgrp/ls-spa) is pure Python and has no FFI of its own; the
 * entry points below are the device-side replacements for the NumPy/SciPy/LAPACK
 * calls on its hot path.  Each one cites the reference lines (paths relative to the
 * reference checkout) it replaces.  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add to call them.
 *
 * Conventions
 *   - every function returns an int status: 0 = ok, <0 = LSSPA_E_* below, and
 *     -(1000 + cudaError_t) when a CUDA call failed (see lsspa_status_string);
 *   - all data pointers are DEVICE pointers unless the name ends in _host;
 *   - matrices are row-major float64 unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *     nothing synchronises the host -- the caller owns synchronisation;
 *   - no hidden global state: scratch memory is passed in by the caller and its
 *     size is obtained from the matching *_workspace_bytes function.
 */
#ifndef LSSPA_H_
#define LSSPA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LSSPA_API __attribute__((visibility("default")))
#else
#define LSSPA_API
#endif

#define LSSPA_ABI_VERSION 1

#define LSSPA_OK 0
#define LSSPA_E_BADARG (-1)     /* null pointer / non-positive size / p out of range */
#define LSSPA_E_WORKSPACE (-2)  /* workspace too small                               */
#define LSSPA_E_NODEVICE (-3)   /* no CUDA device / wrong architecture               */
#define LSSPA_E_UNSUPPORTED (-4)

#define LSSPA_ERR_DRAWS 1024 /* ls_spa/ls_spa.py:334  size=2**10 */

LSSPA_API int lsspa_abi_version(void);
LSSPA_API const char *lsspa_status_string(int status);
/* number of SMs / bytes of opt-in shared memory of the current device (0 on failure) */
LSSPA_API int lsspa_device_sm_count(void);
LSSPA_API int lsspa_device_smem_optin(void);

/* ------------------------------------------------------------------------
 * 1. Tall-skinny reduction            replaces reduce_data, ls_spa/ls_spa.py:290-318
 *    (np.linalg.qr of the (N+p) x p train block and the M x p test block, :314-315,
 *    and the Q^T y products, :316-317).
 *
 * lsspa_tsqr_rows:  rows [0,nrows) of [X / divisor | y / divisor] (X row-major with
 *    leading dimension ldx, p columns) are reduced to `nparts` upper-triangular
 *    (p+1) x (p+1) factors, one per CTA, written to parts[nparts][slot] where
 *    slot = lsspa_tsqr_slot_doubles(p); the last 8 doubles of a slot hold the
 *    CTA's partial sum of (y/divisor)^2 in [0].  nparts = lsspa_tsqr_num_parts().
 * lsspa_tsqr_merge: stacks `count` factors (same slot layout) in groups of `group`
 *    and re-triangularises each group: out[ceil(count/group)][slot].  Repeated until
 *    one factor is left; its leading p x p block is R, its last column holds c
 *    (and the (p,p) entry the residual norm), out[..][(p+1)^2] the sum of squares.
 *    The sqrt(reg)*I ridge rows (:310) are one more "factor" in the stack.
 * ------------------------------------------------------------------------ */
LSSPA_API int64_t lsspa_tsqr_slot_doubles(int p);
LSSPA_API int lsspa_tsqr_num_parts(int p, int64_t nrows);
LSSPA_API int lsspa_tsqr_rows(const double *X, int64_t ldx, const double *y, int64_t nrows, int p,
                    double divisor, double *parts, int nparts, void *stream);
LSSPA_API int lsspa_tsqr_merge(const double *parts, int count, int group, int p, double *out,
                     void *stream);

/* CholeskyQR2 variant of the same reduction for p <= 111 (gram.cu): Gram matrices with fp64 tensor
 * tiles instead of one Householder reflector per column per row block.
 *   pass 1: lsspa_gram_rows(Rinv = NULL) -> partial Grams; lsspa_gram_finish sums them (fixed
 *           order) and scales by 1/divisor^2 -> G1 ((8*ceil((p+1)/8))^2 doubles, row-major);
 *           lsspa_chol_factor -> R1 (q x q row-major), R1^-1 (padded, the layout pass 2 wants),
 *           info[2] = {0 ok / 1 bad pivot, |R'|_F |R'^-1|_F >= cond_2 of the column-equilibrated
 *           factor R' = R diag(|R[:,j]|)^-1 -- the conditioning that governs Cholesky's accuracy};
 *           when that bound is small (<= 1e3) the caller may keep R1 as the factor (one pass);
 *   pass 2: lsspa_gram_rows(Rinv = R1^-1) accumulates (Z R1^-1)^T (Z R1^-1) -> G2 -> R2;
 *   lsspa_tri_product: factor = R2 R1 in the slot layout of lsspa_tsqr_merge.
 * The caller falls back to lsspa_tsqr_rows when info reports a bad pivot or cond > ~1e6. */
LSSPA_API int lsspa_gram_supported(int p);
LSSPA_API int64_t lsspa_gram_slot_doubles(int p);
LSSPA_API int64_t lsspa_gram_rinv_doubles(int p);
LSSPA_API int lsspa_gram_num_parts(int p, int64_t nrows, int pass2);
LSSPA_API int lsspa_gram_rows(const double *X, int64_t ldx, const double *y, int64_t nrows, int p,
                              const double *Rinv_or_null, double *parts, int nparts, void *stream);
LSSPA_API int lsspa_gram_finish(const double *parts, int count, int p, double scale, double *G_out,
                                void *stream);
LSSPA_API int lsspa_chol_factor(const double *G, int p, double *R_out, double *Rinv_out, double *info,
                                void *stream);
/* lsspa_chol_factor that also leaves, in gram_out (layout of lsspa_lifts_gram, may be NULL), what the
 * Cholesky lift route needs from the TRAIN factor -- the Gram matrix with unit feature diagonal, the column
 * scales and the condition bound of the equilibrated leading p x p block -- straight from G (= R^T R),
 * so that a job does not run lsspa_lifts_gram on the factor it has just computed */
LSSPA_API int lsspa_chol_factor_gram(const double *G, int p, double *R_out, double *Rinv_out, double *info,
                                     double *gram_out_or_null, void *stream);

/* ridge rows of the train block (:310) as reg added to the first p diagonal entries of G (in place) */
LSSPA_API int lsspa_gram_add_ridge(double *G, int p, double reg, void *stream);
LSSPA_API int lsspa_tri_product(const double *R2, const double *R1, int p, const double *G1, double *out_slot,
                                void *stream);

/* Wide problems (p + 1 > 112, up to p = 2047; gram_big.cu + lifts_big.cu): the same one-pass Gram
 * reduction with 128-column blocks.
 *   lsspa_gram_big_rows        partial Gram blocks of rows [0, nrows) of [X | y]:
 *                              parts[nsplit][lsspa_gram_big_part_doubles(p)], nsplit = lsspa_gram_big_num_splits;
 *   lsspa_gram_big_accumulate  G_acc ((p+1) x (p+1) row-major, upper blocks) += sum of the partial blocks;
 *                              G_acc is what a multi-GPU job all-reduces;
 *   lsspa_gram_big_factor      slot (layout of lsspa_tsqr_merge) = Cholesky factor of
 *                              scale * G_acc + reg * diag(I_p, 0)  (blocked, fp64 tensor tiles);
 *                              status_flag (device int, zero-initialised) raised on a non-positive feature pivot. */
LSSPA_API int lsspa_gram_big_supported(int p);
LSSPA_API int lsspa_gram_big_num_splits(int p, int64_t nrows);
LSSPA_API int64_t lsspa_gram_big_part_doubles(int p);
LSSPA_API int lsspa_gram_big_rows(const double *X, int64_t ldx, const double *y, int64_t nrows, int p,
                                  double *parts, int nsplit, void *stream);
LSSPA_API int lsspa_gram_big_accumulate(const double *parts, int nsplit, int p, double *G_acc, void *stream);
LSSPA_API size_t lsspa_gram_big_factor_workspace_bytes(int p);
LSSPA_API int lsspa_gram_big_factor(const double *G_acc, int p, double scale, double reg, double *slot_out,
                                    void *workspace, size_t workspace_bytes, int *status_flag, void *stream);

/* ------------------------------------------------------------------------
 * 2. Permutation sources (int32 indices, perms_out[count][p])
 *    exact          itertools.permutations(range(p)), ls_spa/ls_spa.py:171
 *                   (lexicographic rank first_rank .. first_rank+count-1)
 *    pcg64          default_rng(seed).permutation(p) repeated, ls_spa/ls_spa.py:168,175
 *                   (numpy PCG64 XSL-RR + buffered next_uint32 + masked-rejection
 *                   Fisher-Yates).  `gen_state` is 6 x uint64 on the device:
 *                   {state_hi, state_lo, inc_hi, inc_lo, has_uint32, uinteger}; it is
 *                   advanced in place so that calls chain.  status_flag (device int)
 *                   is set non-zero if the raw-draw budget was exhausted.
 *    sobol_argsort  np.argsort(Sobol(p, seed).random(n), axis=1),
 *                   experiments/ground_truth_medium.py:70-71; sv[p][30], shift[p] are
 *                   scipy's scrambled direction numbers / digital shift (uint32)
 *    permutohedron  experiments/ground_truth_medium.py:56-67 fed by
 *                   MultivariateNormalQMC(zeros(p-1), inv_transform=False); sv/shift
 *                   belong to its Sobol engine of dimension 2*ceil((p-1)/2)
 * ------------------------------------------------------------------------ */
LSSPA_API int lsspa_perms_exact(int p, uint64_t first_rank, int64_t count, int32_t *perms_out, void *stream);
LSSPA_API size_t lsspa_perms_pcg64_workspace_bytes(int p, int64_t count);
LSSPA_API int lsspa_perms_pcg64(int p, uint64_t *gen_state, int64_t count, int32_t *perms_out,
                      void *workspace, size_t workspace_bytes, int *status_flag, void *stream);
LSSPA_API int lsspa_perms_sobol_argsort(int p, const uint32_t *sv, const uint32_t *shift, int bits,
                              uint64_t first_index, int64_t count, int32_t *perms_out, void *stream);
LSSPA_API int lsspa_perms_permutohedron(int p, const uint32_t *sv, const uint32_t *shift, int bits,
                              uint64_t first_index, int64_t count, int32_t *perms_out, void *stream);

/* Caller-supplied permutations (perms=, ls_spa/ls_spa.py:165-167,176-177: the reference does not
 * validate them and numpy raises IndexError on a bad index): every row must be a bijection of
 * {0..p-1}.  bad_flag (device int, zero-initialised by the caller) receives 1 + index of an
 * offending row, 0 if all rows are valid. */
LSSPA_API int lsspa_perms_validate(int p, const int32_t *perms, int64_t count, int *bad_flag, void *stream);

/* ------------------------------------------------------------------------
 * 3. Per-permutation core             replaces square_shapley, ls_spa/ls_spa.py:256-287
 *    and the antithetic pair average, :205-208.
 *
 * R_tr_cm / R_te_cm are the p x p reduced factors stored COLUMN-major (column j
 * contiguous, leading dimension p), c_tr / c_te the reduced targets (p).  For each of
 * the `count` permutations the lift vector (p) is written to lifts_out[count][p]; with
 * antithetical != 0 the row is the mean of the lifts of perm and perm[::-1].
 * ------------------------------------------------------------------------ */
LSSPA_API size_t lsspa_lifts_workspace_bytes(int p, int64_t count);
LSSPA_API int lsspa_lifts(int p, const double *R_tr_cm, const double *c_tr, const double *R_te_cm,
                const double *c_te, double y_norm_sq, const int32_t *perms, int64_t count,
                int antithetical, double *lifts_out, void *workspace, size_t workspace_bytes,
                void *stream);

/* Fast route of the same function for well-conditioned reduced problems (17 <= p <= 152):
 * the triangular factor of R_tr[:, perm] (np.linalg.qr, ls_spa/ls_spa.py:268) is computed as
 * the Cholesky factor of the permuted Gram matrix of [R_tr | c_tr], so the forward error grows
 * like eps * cond(R_tr)^2 instead of eps * cond(R_tr).
 *   lsspa_lifts_gram   once per reduced problem.  With D = diag(column norms of R_tr) and
 *                      R' = R_tr D^-1 (the lifts do not change when train and test features are
 *                      scaled alike, and unit columns remove the conditioning that is only units):
 *                      gram_out[0 .. (p+1)^2)       = [R'|c_tr]^T [R'|c_tr],
 *                      gram_out[(p+1)^2 + 0]        = a bound >= cond_2(R') (inf if singular): the smaller
 *                                                     of [+2] = |R'|_F |R'^-1|_F and, for p <= 128,
 *                                                     [+3] = sqrt(max row sum of |G|) *
 *                                                     sqrt(|R'^-1|_1 |R'^-1|_inf),
 *                      gram_out[(p+1)^2 + 1]        = min|R'_kk| / max|R'_kk|,
 *                      gram_out[(p+1)^2 + 8 .. +8+p) = D; the rest is scratch.
 *   lsspa_lifts_chol   same outputs as lsspa_lifts, reading gram_out instead of R_tr / c_tr;
 *                      R_te_cm must be the test factor with its columns divided by D.
 * The caller decides from the condition estimate which route to take (ls_spa_b200/ops.py uses
 * the Cholesky route when the estimate is <= 1e3, i.e. an expected error <= 1e-10). */
LSSPA_API int lsspa_lifts_chol_supported(int p);
LSSPA_API int64_t lsspa_lifts_gram_doubles(int p);
LSSPA_API int lsspa_lifts_gram(int p, const double *R_tr_cm, const double *c_tr, double *gram_out,
                     void *stream);
LSSPA_API int lsspa_lifts_chol(int p, const double *gram, const double *R_te_cm, const double *c_te,
                     double y_norm_sq, const int32_t *perms, int64_t count, int antithetical,
                     double *lifts_out, void *stream);

/* Wide problems (152 < p <= 2047; BASELINE config 5 is p = 1000): the same Cholesky route as a batched
 * blocked factorisation over a tile workspace in device memory (lifts_big.cu).  lsspa_lifts_gram accepts
 * these widths too (same gram_out layout).  workspace: lsspa_lifts_big_workspace_bytes(p, count,
 * antithetical, budget) bytes -- as many samples as fit into `budget` are processed per pass (at least
 * one); status_flag (device int, zero-initialised) is raised if a feature pivot was not positive. */
LSSPA_API int lsspa_lifts_big_supported(int p);
LSSPA_API size_t lsspa_lifts_big_workspace_bytes(int p, int64_t count, int antithetical, size_t budget_bytes);
LSSPA_API int lsspa_lifts_big(int p, const double *gram, const double *R_te_cm, const double *c_te,
                              double y_norm_sq, const int32_t *perms, int64_t count, int antithetical,
                              double *lifts_out, void *workspace, size_t workspace_bytes, int *status_flag,
                              void *stream);

/* The same route in two launches (49 <= p <= 128), for jobs whose inputs arrive over PCIe: the
 * factorisation of R_tr[:, perm] (np.linalg.qr, ls_spa/ls_spa.py:268-270) needs the train side
 * only, so it can run while the test rows are still being copied; the elimination against the
 * test factor (:279-283) follows.  factors: lsspa_lifts_chol_factor_doubles(p) doubles per
 * permutation evaluation (count, or 2 * count with antithetical != 0: perm and its reverse). */
LSSPA_API int64_t lsspa_lifts_chol_factor_doubles(int p);
LSSPA_API int lsspa_lifts_chol_factor(int p, const double *gram, const int32_t *perms, int64_t count,
                            int antithetical, double *factors_out, void *stream);
LSSPA_API int lsspa_lifts_chol_eliminate(int p, const double *factors, const double *R_te_cm,
                               const double *c_te, double y_norm_sq, const int32_t *perms,
                               int64_t count, int antithetical, double *lifts_out, void *stream);

/* ------------------------------------------------------------------------
 * 4. Estimator                        replaces ls_spa/ls_spa.py:186-236
 *    (merge_sample_mean/cov :103-119, error_estimates :321-341, stop test :229).
 *
 * Work is organised per super-batch (a run of consecutive batches of lift rows):
 *   lsspa_estimator_partials   per-batch partial moments of the rows this rank computed:
 *                              {n, mean, sum (l-mean)(l-mean)^T, G = sum_k g_k, S = sum_k g_k (l_k - mean)}
 *                              with g the counter-based N(0,1) stream keyed by the GLOBAL sample index
 *                              (so the result does not depend on how batches are sharded over GPUs);
 *   lsspa_estimator_absorb     folds nb consecutive batches (blocks partials[slot_map[b]]) into the
 *                              state (mean, biased cov, draw sums) with the Chan merge of
 *                              merge_sample_mean/cov, in order, and for the batches [own0, own1) writes
 *                              the squared error draws z_sf^2, z_s = sum_k g_ks (l_k - mean) / sqrt(n (n-1)),
 *                              which have exactly the covariance unbiased_cov / n that error_estimates
 *                              samples from (factor-free: that covariance is singular);
 *   lsspa_estimator_quantiles  0.95 quantiles of |z_sf| per feature and of |z_s|_2 for every owned batch.
 * The caller scans the per-batch errors for the first one below the tolerance (the reference's `break`);
 * if it falls inside a super-batch it restores its snapshot of the state and replays absorb up to it.
 * State = lsspa_estimator_state_bytes(p) bytes, zero-initialised by the caller:
 * mean[2][p] and G[2][1024] ping-pong (`cur` selects the live copy; absorb writes the other one),
 * cov[p][p], S[p][1024].
 * ------------------------------------------------------------------------ */
LSSPA_API size_t lsspa_estimator_state_bytes(int p);
LSSPA_API int64_t lsspa_estimator_partial_doubles(int p);
/* largest nb one absorb call accepts (running means of all its batches sit in shared memory) */
LSSPA_API int lsspa_estimator_max_batches(int p);
/* batch_desc (device, int64[nbatch][3]) = {first row in `lifts`, row count (may be 0), global index of the
 * first sample}; partials[nbatch][partial_doubles] */
LSSPA_API int lsspa_estimator_partials(int p, const double *lifts, const int64_t *batch_desc, int nbatch,
                             uint64_t seed, int estimate_errors, double *partials, void *stream);
/* zsq: device [own1-own0][p+1][1024] doubles (row p of every batch is scratch of lsspa_estimator_quantiles), or NULL when no error draws are wanted; with_draws = 0
 * skips the draw sums altogether (partials made with estimate_errors = 0 carry none) */
LSSPA_API int lsspa_estimator_absorb(void *state, int p, int cur, double n_before, const double *partials,
                           const int32_t *slot_map, int nb, int own0, int own1, double *zsq, int with_draws,
                           void *stream);
LSSPA_API int lsspa_estimator_quantiles(int p, double *zsq, int nown, double *overall_out,
                              double *feat_out, void *stream);
/* absorb + quantiles in the form the sample loop consumes (ls_spa/ls_spa.py:222-230 keeps the overall error of
 * every batch for the stop test / error_history, but the per-feature errors only of the batch it ends on):
 * the same fold as lsspa_estimator_absorb, overall_out[own1-own0] = the overall error after every owned
 * batch, feat_out[p] = the per-feature errors after batch feat_batch (own0 <= feat_batch < own1; -1: none).
 * The squared draws never leave the kernel: per owned batch only [groups of 4 features][1024] partial norms
 * are written.  workspace: lsspa_estimator_errors_workspace_doubles(p, own1 - own0) doubles. */
LSSPA_API int64_t lsspa_estimator_errors_workspace_doubles(int p, int nown);
LSSPA_API int lsspa_estimator_absorb_errors(void *state, int p, int cur, double n_before, const double *partials,
                                            const int32_t *slot_map, int nb, int own0, int own1, int feat_batch,
                                            double *overall_out, double *feat_out, double *workspace,
                                            void *stream);
/* multi-GPU: one partial block equal to the merge of nb consecutive partial blocks (the run total a rank
 * ships to the others), as parallel sums -- merge_sample_mean/cov are associative (test/test_ls_spa.py:20-44) */
LSSPA_API int lsspa_estimator_block_total(int p, const double *partials, int nb, int with_draws, double *out_block,
                                          void *stream);
/* running means after each sample (attribution_history, :217-219): hist[k] =
 * (carry_sum + sum_{r<=k} lifts[r]) / (carry_count + k + 1); carry is updated */
LSSPA_API int lsspa_prefix_means(int p, const double *lifts, int64_t rows, double *carry_sum,
                       double carry_count, double *hist_out, void *stream);
/* error_estimates(rng, cov) as a standalone call (ls_spa/ls_spa.py:321-341): 1024 draws of N(0, cov) for a
 * caller-supplied positive semi-definite cov (p x p row-major) through a Cholesky factorisation that
 * skips vanishing pivots (cov = L L^T also when cov is singular), Gaussians from the counter-based
 * device stream keyed by `seed`.  zsq [p+1][1024] receives the squared draws (input of
 * lsspa_estimator_quantiles with nown = 1); workspace: p*p doubles. */
LSSPA_API int lsspa_error_draws(int p, const double *cov, uint64_t seed, double *zsq, double *workspace,
                                void *stream);
/* standalone merge, the device twin of merge_sample_mean / merge_sample_cov */
LSSPA_API int lsspa_merge_moments(int p, double *mean, double *cov, double old_n, const double *new_mean,
                        const double *new_cov_or_null, double new_n, void *stream);

/* ------------------------------------------------------------------------
 * 5. Epilogue                          replaces ls_spa/ls_spa.py:240-243
 *    theta = np.linalg.lstsq(R_tr, c_tr, rcond=None)[0]: minimum-norm solution through a
 *    one-sided Jacobi SVD (singular values <= eps * p * sigma_max dropped, as numpy does);
 *    r_squared = (|c_te|^2 - |c_te - R_te theta|^2) / y_norm_sq
 *    out[p+1] = {theta[0..p), r_squared}
 * ------------------------------------------------------------------------ */
LSSPA_API size_t lsspa_theta_r2_workspace_bytes(int p);
LSSPA_API int lsspa_theta_r2(int p, const double *R_tr_cm, const double *c_tr, const double *R_te_cm,
                             const double *c_te, double y_norm_sq, double *out, void *workspace,
                             size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LSSPA_H_ */
