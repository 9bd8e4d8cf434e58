/*
 * pxo_oracle.h -- CPU restatement of the Progressive-X hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This library is the parity oracle for the sm_100a kernels. Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it. The product (libpxb200.so and
 * the pyprogressivex host mirror) never links, imports or calls anything under oracle/.
 *
 * PARITY PINNING STATUS
 *   - a10/a11 (greedy UFL label sweep, alpha-expansion, BK max-flow, LO st-cut): PINNED against the reference's
 *     own gco-v3 + maxflow sources compiled unchanged into oracle/_ref/libgco_ref.so (oracle/Makefile).
 *   - a1, a2, a3 (residuals), a4 (getScore incl. early exit / int exponent), a5 (setPreferenceVector), a6 (four-point
 *     solver, gaussElimination, isValidSample, isValidModel), the oriented-epipolar test of a7, a9 (PEARL data
 *     costs): PINNED BIT-EXACT against the reference's own function bodies, compiled verbatim from
 *     /root/reference by oracle/extract_ref.py into oracle/_ref/libpx_refbodies.so (tests/test_oracle_pinned.py).
 *   - a7 null space + cubic, a8 (P3P), the Tanimoto reduction order of a5, a12, a13: restated, "parity unpinned"
 *     beyond tolerance: their arithmetic lives in Eigen (FullPivLU::kernel, PolynomialSolver, redux) / libm, which are
 *     not on disk; the reference ships no golden vectors for them. Known-answer tests on planted structures
 *     (tests/test_oracle_cpu.py) cover them.
 *   Everything is IEEE double, left-to-right operation order, built with -O3 -ffp-contract=off (no FMA contraction,
 *   like the reference's own -O3 x86-64 build).
 *
 * All matrices are row-major doubles. Points: H/F rows are [x1 y1 x2 y2]; PnP rows are
 * [u v X Y Z] with (u,v) already K^-1-normalised. Models: H/F 9 doubles, PnP 12 doubles (3x4).
 * "/root/reference" paths are abbreviated:  gcr/ = graph-cut-ransac/src/pygcransac/include/,
 * px/ = src/pyprogressivex/.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { PXO_MODEL_H = 0, PXO_MODEL_F = 1, PXO_MODEL_PNP = 2, PXO_MODEL_VP = 3, PXO_MODEL_LINE = 4 };

int pxo_point_dim(int model_type);   /* 4, 4, 5, 4, 2 */
int pxo_model_size(int model_type);  /* 9, 9, 12, 3, 3 */
int pxo_sample_size(int model_type); /* 4, 7, 3, 2, 2 */

/* a1/a2/a3: squared residual of one point w.r.t. one model. */
double pxo_squared_residual(int model_type, const double *point, const double *model);

/* N x K residual-and-inlier matrix, hypothesis-major: r2[k*N + i]; mask bit i of hypothesis k is
 * bit (i & 31) of mask[k*words + (i >> 5)], words = (N + 31) / 32; set iff r2 < T2. */
void pxo_residual_matrix(int model_type, const double *pts, int64_t N, const double *models, int64_t K,
                         double T2, double *r2 /*K*N or NULL*/, uint32_t *mask /*K*words or NULL*/);

/* a4: MSACScoringFunctionWithCompoundModel::getScore, px/include/scoring_function_with_compound_model.h:61-125.
 * Sequential sums in point order. compound_pref may be NULL (= empty compound model).
 * Outputs: count, value_sum = sum max(0, 1 - r2/T2) over inliers, shared = sum min(compound_pref, pref).
 * Returns the final Score::value = value_sum - pow(shared, exponent) (value_sum when compound_pref == NULL),
 * or 0 with *count = 0 when the reference's early exit fires (count + 1 < best_inlier_number). */
double pxo_get_score(int model_type, const double *pts, int64_t N, const double *model, double T2,
                     const double *compound_pref, int exponent, int64_t best_inlier_number,
                     int64_t *count, double *value_sum, double *shared, int64_t *inliers /*N or NULL*/);

/* Same quantities for K hypotheses without the early exit (what the GPU batch returns). */
void pxo_score_batch(int model_type, const double *pts, int64_t N, const double *models, int64_t K, double T2,
                     const double *compound_pref, int64_t *count, double *value_sum, double *shared,
                     int threads);

/* a5: progx::Model::setPreferenceVector, px/include/progx_model.h:70-87 */
void pxo_preference_vector(int model_type, const double *pts, int64_t N, const double *model, double T,
                           double *pref);
/* a5: tanimoto similarity, px/include/progressive_x.h:583-588 (Eigen dot/squaredNorm reduction order
 * restated from Eigen 3.3/3.4 redux.h, SSE2 packets of two doubles). */
double pxo_tanimoto(const double *a, const double *b, int64_t N);
/* a5: updateCompoundModel, px/include/progressive_x.h:597-624: out[i] = max_k prefs[k*N+i], starting from 0 */
void pxo_compound_max(const double *prefs, int64_t L, int64_t N, double *out);

/* a6: H four-point minimal solver + validity tests. Returns 1 and writes 9 doubles if a model was produced
 * (no NaN), else 0.  gcr/estimators/solver_homography_four_point.h:109-190, gcr/math_utils.h:45-87 */
int pxo_h4_solve(const double *pts, const int64_t *sample, double *H);
/* gcr/estimators/homography_estimator.h:346-381 */
int pxo_h4_is_valid_sample(const double *pts, const int64_t *sample);
/* gcr/estimators/homography_estimator.h:326-342 (|det| >= 1e-2; determinant via partial-pivot LU as Eigen
 * does for a dynamic-size MatrixXd) */
int pxo_h_is_valid_model(const double *H);

/* a7: seven-point solver + oriented epipolar test. Returns the number of models kept (0..3), each 9 doubles.
 * gcr/estimators/solver_fundamental_matrix_seven_point.h:91-291, gcr/estimators/fundamental_estimator.h:161-184,737-800.
 * The null space follows Eigen FullPivLU::kernel(); the cubic is solved in closed form + Newton polish instead
 * of Eigen::PolynomialSolver (companion-matrix eigenvalues): roots agree to ~1e-12, order is ascending. */
int pxo_f7_solve(const double *pts, const int64_t *sample, double *F_out /*27*/, int apply_orientation_test);

/* a8: P3P (Lambda-twist). Returns number of poses (0..4), each 12 doubles row-major 3x4.
 * gcr/estimators/solver_p3p.h:108-385 */
int pxo_p3p_solve(const double *pts /*rows of 5*/, const int64_t *sample, double *P_out /*48*/);

/* a9: PEARL data cost matrix, px/include/PEARL.h:41-128. D is N x (L+1) row-major (site-major, as gco's
 * setDataCost(array) expects). thr is the inlier-outlier threshold (T = 9/4*thr*thr). */
void pxo_pearl_datacost(int model_type, const double *pts, int64_t N, const double *models, int64_t L,
                        double thr, double lambda, double *D);

/* a12: per-instance sums of residual (= sqrt r2) over the points carrying that label, point order.
 * px/include/PEARL.h:369-371,388-390 */
void pxo_segment_residual_sums(int model_type, const double *pts, int64_t N, const double *models, int64_t L,
                               const int32_t *labels, double *sums /*L*/, int64_t *counts /*L*/);

/* a13: GC-RANSAC LO unary terms (gcr/GCRANSAC.h:937-962): d[i] = clamp(r2/T',0,1), T' = thr*thr*9/4;
 * e0[i], e1[i] are the two arguments of add_term1(i, e0, e1). */
void pxo_lo_unary_terms(int model_type, const double *pts, int64_t N, const double *model, double thr,
                        double lambda, double *d, double *e0, double *e1);
/* a13: Tukey bisquare weights (gcr/GCRANSAC.h:658-669): w = max(0, 1 - r2/T2)^2 on the listed inliers */
void pxo_tukey_weights(int model_type, const double *pts, const int64_t *inliers, int64_t n, const double *model,
                       double T2, double *weights /*indexed by point*/);

/* f-1 (next row): non-minimal homography fit, gcr/estimators/homography_estimator.h:140-173 +
 * solver_homography_four_point.h:192-264 (Hartley normalisation, 2n x 8 least squares by column-pivoted Householder
 * QR, denormalisation). weights_by_row may be NULL. Returns 1 on success. */
int pxo_fit_h_nonminimal(const double *pts, const int64_t *idx, int64_t n, const double *weights_by_row, double *H);

/* f-4 (next row): vanishing points (rows [xs ys xe ye] = line segments, model = homogeneous point) and 2D lines
 * (rows [x y], model = (nx, ny, c)). Residuals: px/include/vanishing_point_estimator.h:127-189,
 * gcr/estimators/linear_model_estimator.h:121-131 (reachable through pxo_squared_residual and every operator above).
 * Minimal solvers: px/include/solver_vanishing_point_two_lines.h:146-186, gcr/estimators/solver_linear_model.h:143-171
 * (with the reference's `nx = y1 - x2`). Non-minimal: solver_vanishing_point_two_lines.h:187-233 (weights indexed by
 * point), linear_model_estimator.h:152-250 + solver_linear_model.h:198-239. */
int pxo_vp2_solve(const double *pts, const int64_t *sample, double *v /*3*/);
int pxo_line2_solve(const double *pts, const int64_t *sample, double *l /*3*/);
int pxo_fit_vp_nonminimal(const double *pts, const int64_t *idx, int64_t n, const double *weights_by_point, double *v);
int pxo_fit_line_nonminimal(const double *pts, const int64_t *idx, int64_t n, double *l);

/* a10 restated (used when oracle/_ref is not available and to cross-check it):
 * GCoptimization::solveGreedy, gcr/GCoptimization.cpp:608-751, for dense data costs and one uniform
 * per-label cost. init_labels gives the start labelling whose energy the greedy result must beat. */
double pxo_greedy_ufl(const double *D, int64_t N, int32_t L1, double label_cost, const int32_t *init_labels,
                      int32_t *labels_out);

#ifdef __cplusplus
}
#endif
