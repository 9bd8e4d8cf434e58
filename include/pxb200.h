/*
 * pxb200.h -- C ABI of libpxb200.so, the B200 (sm_100a) implementation of the Progressive-X hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types. Every entry point names the
 * reference interface it replaces (paths relative to the danini/progressive-x tree; gcr/ =
 * graph-cut-ransac/src/pygcransac/include/, px/ = src/pyprogressivex/). INTEGRATION.md shows the binding a
 * reference maintainer would add (a gcransac::ScoringFunction subclass and the pybind/ctypes stubs).
 *
 * Conventions
 *   - All functions return 0 on success, a negative pxb_status on failure; pxb_last_error() gives the message
 *     (thread-local). Nothing throws across the boundary and nothing calls exit().
 *   - There is NO CPU fallback: if no CUDA device / wrong architecture is present, pxb_ctx_create fails.
 *   - Host entry points take HOST pointers and perform the H2D/D2H copies themselves on the context's stream,
 *     synchronising before they return. "_dev" entry points take DEVICE pointers, are asynchronous on the
 *     context's stream and need pxb_sync() before results are read.
 *   - Matrices are row-major float64 exactly as the reference holds them: correspondences [N,4] = x1 y1 x2 y2
 *     (cv::Mat(N,4,CV_64F) view, px/src/progressivex_python.cpp:203); 2D-3D matches [N,5] = u v X Y Z with (u,v)
 *     K^-1-normalised (progressivex_python.cpp:64-98); models 3x3 (9) or 3x4 (12) row-major, the order in which
 *     the reference exports descriptors (progressivex_python.cpp:284-301).
 *   - Arithmetic is IEEE float64 without FMA contraction in the reference's operation order, so residuals,
 *     inlier masks and labels are bit-identical to the CPU reference; sums use a fixed reduction topology
 *     (DESIGN.md "Summation order").
 */
#ifndef PXB200_H
#define PXB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pxb_ctx pxb_ctx;

typedef enum pxb_status {
	PXB_OK = 0,
	PXB_ERR_CUDA = -1,        /* CUDA runtime error (message has the CUDA string) */
	PXB_ERR_NO_DEVICE = -2,   /* no usable sm_100 device */
	PXB_ERR_ARGUMENT = -3,    /* bad shape / null pointer / unknown enum */
	PXB_ERR_STATE = -4,       /* e.g. points not uploaded */
	PXB_ERR_UNSUPPORTED = -5
} pxb_status;

/* Estimator families of the reference (gcr/types.h:78-80,93-95,136-138). */
typedef enum pxb_model_type {
	PXB_MODEL_HOMOGRAPHY = 0,  /* RobustHomographyEstimator, 4-point minimal solver */
	PXB_MODEL_FUNDAMENTAL = 1, /* FundamentalMatrixEstimator, 7-point minimal solver */
	PXB_MODEL_PNP = 2,         /* PerspectiveNPointEstimator, P3P minimal solver */
	/* px/include/vanishing_point_estimator.h: rows [xs ys xe ye] are line segments, the model is a homogeneous point */
	PXB_MODEL_VANISHING_POINT = 3,
	/* gcr/estimators/linear_model_estimator.h (Default2DLineEstimator, gcr/types.h:146-149): rows [x y], model (nx, ny, c) */
	PXB_MODEL_LINE2D = 4
} pxb_model_type;

const char *pxb_last_error(void);
const char *pxb_version(void);

/* ---- context ------------------------------------------------------------------------------------------- */
int pxb_ctx_create(int device, pxb_ctx **out);
void pxb_ctx_destroy(pxb_ctx *ctx);
/* The CUDA stream (cudaStream_t as void*) every kernel of this context is launched on; lets a caller record
 * events around "_dev" calls. */
void *pxb_ctx_stream(pxb_ctx *ctx);
int pxb_sync(pxb_ctx *ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t pxb_launch_count(pxb_ctx *ctx);

/* Device memory helpers so that non-CUDA hosts (ctypes, cgo) can keep buffers resident. */
int pxb_dev_alloc(pxb_ctx *ctx, size_t bytes, void **dev_ptr);
int pxb_dev_free(pxb_ctx *ctx, void *dev_ptr);
int pxb_memcpy_h2d(pxb_ctx *ctx, void *dev_dst, const void *host_src, size_t bytes);
int pxb_memcpy_d2h(pxb_ctx *ctx, void *host_dst, const void *dev_src, size_t bytes);
int pxb_host_alloc_pinned(size_t bytes, void **host_ptr);
int pxb_host_free_pinned(void *host_ptr);

/* ---- data ---------------------------------------------------------------------------------------------- */
/* Replaces the cv::Mat view the reference builds over the caller's buffer (progressivex_python.cpp:203).
 * Copies [N, dim] host doubles to the device and re-tiles them to the kernels' SoA layout. dim is 4 (H, F, vanishing
 * points), 5 (PnP) or 2 (2D lines). */
int pxb_upload_points(pxb_ctx *ctx, int model_type, const double *pts_host, int64_t N);
int64_t pxb_point_count(pxb_ctx *ctx);

/* ---- a1/a2/a3: residual-and-inlier matrix ---------------------------------------------------------------- */
/* Replaces K calls of Estimator::squaredResidual over all points (gcr/estimators/homography_estimator.h:181-199,
 * fundamental_estimator.h:195-222, perspective_n_point_estimator.h:148-184).
 * r2 is hypothesis-major: r2[k*N + i]. mask is a bit matrix: bit (i & 31) of mask[k*words + (i >> 5)],
 * words = (N+31)/32, set iff r2 < T2 (the reference's inlier test, scoring_function_with_compound_model.h:85).
 * Either output may be NULL. */
int pxb_residual_matrix(pxb_ctx *ctx, const double *models_host, int64_t K, double T2, double *r2_host,
                        uint32_t *mask_host);
int pxb_residual_matrix_dev(pxb_ctx *ctx, const double *models_dev, int64_t K, double T2, double *r2_dev,
                            uint32_t *mask_dev);
/* With r2 == NULL only the bit matrix is produced, by the float32-SCREENED kernel: pairs that are provably outliers
 * never take the float64 path (the mask is bit-identical to the one written beside r2). A float32-r2 output variant
 * existed in round 1 (pxb_residual_matrix_f32_dev); it had to evaluate every pair in float64 before rounding, ran at the
 * speed of the float64 matrix (30 % of its own 4.125 B/eval roofline) and was removed. */

/* ---- a4: compound-aware MSAC score ----------------------------------------------------------------------- */
/* Replaces MSACScoringFunctionWithCompoundModel::getScore (px/include/scoring_function_with_compound_model.h:61-125)
 * for K hypotheses at once. Outputs per hypothesis: count (Score::inlier_number), value_sum = sum max(0,1-r2/T2),
 * shared = sum min(compound_pref_i, pref_i) (0 when compound_pref is NULL). The caller finishes the score as the
 * reference does: value = value_sum - pow(shared, exponent), and Score() when count + 1 < best.inlier_number
 * (:105-106). compound_pref has N entries or is NULL (empty compound set, :110). */
int pxb_score_compound(pxb_ctx *ctx, const double *models_host, int64_t K, double T2,
                       const double *compound_pref_host, int64_t *count_host, double *value_sum_host,
                       double *shared_host);
int pxb_score_compound_dev(pxb_ctx *ctx, const double *models_dev, int64_t K, double T2,
                           const double *compound_pref_dev, int64_t *count_dev, double *value_sum_dev,
                           double *shared_dev);
/* Inlier index list of one model (the std::vector<size_t>& inliers_ of getScore), ascending point order.
 * inliers_host must hold N entries; *n_inliers receives the count. */
int pxb_inliers(pxb_ctx *ctx, const double *model_host, double T2, int64_t *inliers_host, int64_t *n_inliers);

/* ---- a5: preference vectors ------------------------------------------------------------------------------ */
/* progx::Model::setPreferenceVector (px/include/progx_model.h:70-87): pref[i] = max(0, 1 - r2_i / T). */
int pxb_preference_vector(pxb_ctx *ctx, const double *model_host, double T, double *pref_host);
/* Tanimoto similarity of isPutativeModelValid (px/include/progressive_x.h:583-588). */
int pxb_tanimoto(pxb_ctx *ctx, const double *a_host, const double *b_host, int64_t N, double *similarity);
/* updateCompoundModel (px/include/progressive_x.h:597-624): out[i] = max(0, max_k prefs[k*N + i]). */
int pxb_compound_max(pxb_ctx *ctx, const double *prefs_host, int64_t L, int64_t N, double *out_host);

/* ---- a6/a7/a8: batched minimal solvers ------------------------------------------------------------------- */
/* samples: K rows of m point indices (m = 4, 7, 3). models_out: K * max_solutions * model_size doubles
 * (max_solutions = 1, 3, 4); n_models[k] receives how many leading slots of sample k are filled.
 * H: HomographyFourPointSolver::estimateMinimalModel + gaussElimination<8> (gcr/estimators/
 *    solver_homography_four_point.h:109-190, gcr/math_utils.h:45-87); sample_valid[k] =
 *    RobustHomographyEstimator::isValidSample (homography_estimator.h:346-381); model_valid[k] = isValidModel's
 *    determinant test (:326-342).
 * F: FundamentalMatrixSevenPointSolver::estimateModel + the oriented-epipolar filter of
 *    FundamentalMatrixEstimator::estimateModel (solver_fundamental_matrix_seven_point.h:91-291,
 *    fundamental_estimator.h:161-184,737-800). Roots in ascending order.
 * PnP: P3PSolver::estimateModel (solver_p3p.h:177-385).
 * VP: VanishingPointTwoLineSolver::estimateModel, minimal branch (px/include/solver_vanishing_point_two_lines.h:146-186).
 * 2D line: LinearModelSolver<2>::estimate2DLine (gcr/estimators/solver_linear_model.h:152-188), including the
 *    reference's `nx = y1 - x2`.
 * m = 4, 7, 3, 2, 2 indices per sample; max_solutions = 1, 3, 4, 1, 1; model size 9, 9, 12, 3, 3.
 * sample_valid / model_valid may be NULL. */
int pxb_solve_minimal(pxb_ctx *ctx, const int64_t *samples_host, int64_t K, double *models_out_host,
                      int32_t *n_models_host, uint8_t *sample_valid_host, uint8_t *model_valid_host);

/* DEGENSAC's minimal solver (FundamentalMatrixPlaneParallaxSolver::estimateModel,
 * gcr/estimators/solver_fundamental_matrix_plane_and_parallax.h:107-162): a fixed homography H (row-major 3x3) and
 * two off-plane correspondences give F = [e]_x H with e = ((H x1_a) x x2_a) x ((H x1_b) x x2_b); no model when
 * |e_z| < DBL_EPSILON. samples: [K, 2] indices; models_out: [K, 9]; n_models[k] in {0, 1}. Points must have been
 * uploaded as PXB_MODEL_FUNDAMENTAL. */
int pxb_solve_plane_parallax(pxb_ctx *ctx, const int64_t *samples_host, int64_t K, const double *H_host,
                             double *models_out_host, int32_t *n_models_host);

/* The H-degeneracy test of a seven-point sample (FundamentalMatrixEstimator::applyDegensac,
 * gcr/estimators/fundamental_estimator.h:341-476): for the five point triplets {0,1,2},{3,4,5},{0,1,6},{3,4,6},{2,5,6}
 * the homography compatible with F through the triplet is formed; the sample is degenerate when, for one of them, at
 * least five of the seven correspondences have a transfer error below 2 px. Host arithmetic (seven points): no
 * context needed. rows: [N, 4] correspondences, sample7: 7 indices, F: row-major 3x3. *degenerate receives 0/1 and,
 * when 1, H_out (9 doubles) the homography of the first such triplet. */
int pxb_h_degenerate_sample(const double *rows, const int64_t *sample7, const double *F, double *H_out, int32_t *degenerate);

/* ---- a9/a10/a11/a12: PEARL ------------------------------------------------------------------------------- */
/* dataEnergyFunctor + EnergyDataStructure (px/include/PEARL.h:17-56,82-128) evaluated densely:
 * D[i*(L+1) + l], l < L: 2(1-lambda) if r2 > T else (1-lambda) r2 / T, T = 9/4 thr^2; D[i*(L+1)+L] = 1-lambda. */
int pxb_pearl_datacost(pxb_ctx *ctx, const double *models_host, int64_t L, double thr, double lambda,
                       double *D_host);
/* One PEARL::labeling call (px/include/PEARL.h:476-555) = GCoptimizationGeneralGraph::expansion(it, 1000)
 * (gcr/GCoptimization.cpp:1003-1086) on N sites and L1 = L+1 labels with data costs D [N, L1], Potts smooth
 * cost lambda on the neighbour lists, and one uniform per-label cost. csr_off [N+1] / csr_idx are the DIRECTED
 * neighbour lists as getNeighbors(i) returns them (duplicates preserved; each entry adds an undirected edge as
 * setNeighbors does). lambda == 0 or no edges dispatches to the greedy facility-location solver like
 * solveSpecialCases/solveGreedy (GCoptimization.cpp:483-555,608-751). init_labels may be NULL (all zero, the
 * state of a fresh GCoptimization object). */
int pxb_pearl_label(pxb_ctx *ctx, const double *D_host, int64_t N, int32_t L1, double lambda, double label_cost,
                    const int32_t *csr_off_host, const int32_t *csr_idx_host, const int32_t *init_labels_host,
                    int32_t *labels_out_host, double *energy_out);
/* Sums of residual() over the points of each instance and their counts (px/include/PEARL.h:342-352,369-371,
 * 388-390). labels: int32 [N]; entries outside [0, L) are ignored (outlier label). */
int pxb_segment_residual_sums(pxb_ctx *ctx, const double *models_host, int64_t L, const int32_t *labels_host,
                              double *sums_host, int64_t *counts_host);

/* ---- a13: GC-RANSAC local optimisation terms ------------------------------------------------------------- */
/* Unary terms of GCRANSAC::labeling (gcr/GCRANSAC.h:937-962): d[i] = clamp(r2/T',0,1), T' = thr*thr*9/4 and the
 * (E0,E1) pair passed to add_term1. */
int pxb_lo_unary_terms(pxb_ctx *ctx, const double *model_host, double thr, double lambda, double *d_host,
                       double *e0_host, double *e1_host);
/* Tukey bisquare weights of iteratedLeastSquaresFitting (gcr/GCRANSAC.h:658-669) for all points:
 * w[i] = max(0, 1 - r2_i/T2)^2. */
int pxb_tukey_weights(pxb_ctx *ctx, const double *model_host, double T2, double *weights_host);

/* The st-cut of GCRANSAC::labeling (gcr/GCRANSAC.h:964-1018) given the unary terms above: pairwise terms
 * e00 = 0.5 (d_i + d_j) lambda, e01 = e10 = lambda, e11 = 0 on every undirected pair of the directed neighbour lists
 * (first occurrence only, like the reference's used_edges matrix), Boykov-Kolmogorov labelling rule
 * (inlier = SINK = the node can still reach the sink in the residual graph). inlier_out: N bytes (0/1). */
int pxb_lo_graph_cut(pxb_ctx *ctx, const double *e0_host, const double *e1_host, const double *d_host, int64_t N,
                     double lambda, const int32_t *csr_off_host, const int32_t *csr_idx_host, uint8_t *inlier_out);

/* GCRANSAC::labeling as a whole (gcr/GCRANSAC.h:914-1022): unary terms, the pairwise graph and the st-cut without
 * leaving the device (the arc skeleton of a neighbourhood graph is cached on the device after the first call). Same
 * result as pxb_lo_unary_terms followed by pxb_lo_graph_cut. */
int pxb_lo_labeling(pxb_ctx *ctx, const double *model_host, double thr, double lambda, const int32_t *csr_off_host,
                    const int32_t *csr_idx_host, uint8_t *inlier_out);

/* ---- "next" rows the task-level driver needs -------------------------------------------------- */
/* Neighbourhood graph (replaces FlannNeighborhoodGraph, gcr/neighborhood/flann_neighborhood_graph.h:100-139): the
 * k nearest points within `radius` of every point (self excluded, ties by index), all coordinates of the uploaded
 * rows. nbr_out: N*k int32 (-1 padded), deg_out: N int32. */
int pxb_knn_graph(pxb_ctx *ctx, double radius, int k, int32_t *nbr_out_host, int32_t *deg_out_host);
/* Batched non-minimal homography fits (RobustHomographyEstimator::estimateModelNonminimal,
 * gcr/estimators/homography_estimator.h:140-173 + solver_homography_four_point.h:192-264): P problems given as CSR
 * index lists; weights_by_row may be NULL. H_out: P*9, ok_out: P. */
int pxb_fit_homographies(pxb_ctx *ctx, int32_t P, const int32_t *off_host, const int32_t *idx_host,
                         const double *weights_by_row_host, double *H_out_host, int32_t *ok_out_host);
/* Same for whatever estimator family the uploaded points belong to (Estimator::estimateModelNonminimal). H: as above.
 * F: normalised 8-point + rank-2 projection (gcr/estimators/fundamental_estimator.h:574-618) followed by a Levenberg-
 * Marquardt polish of the Sampson error (the objective of solver_fundamental_matrix_bundle_adjustment.h:114-178), n >= 8. PnP: normalised DLT + Levenberg-Marquardt on the reprojection
 * error (stands in for solver_pnp_bundle_adjustment.h:108-225), n >= 6, weights ignored like the reference does.
 * Vanishing point: smallest eigenvector of the weighted 3x3 normal matrix (px/include/solver_vanishing_point_two_lines.h:
 * 187-233); here `weights_by_row_host` holds N entries and is read BY POINT, the reference's indexing for this solver (:203).
 * 2D line: mass point + sqrt2 scaling + 2x2 full-pivot Householder QR (gcr/estimators/linear_model_estimator.h:152-250,
 * solver_linear_model.h:198-239), weights unused like in the reference.
 * models_out: P * 9 (H, F), P * 12 (PnP) or P * 3 (vanishing point, line) doubles. */
int pxb_fit_nonminimal(pxb_ctx *ctx, int32_t P, const int32_t *off_host, const int32_t *idx_host,
                       const double *weights_by_row_host, double *models_out_host, int32_t *ok_out_host);

/* ---- self-test --------------------------------------------------------------------------------------------- */
/* Compares the hot loop's shared-reciprocal double division (two quotients, one Newton reciprocal) with the
 * IEEE div.rn.f64 on n_triples pseudo-random (a1, a2, b) operands; mode 0 = arbitrary bit patterns, 1 = magnitudes
 * typical for the residual kernels. *mismatches must come back 0. */
int pxb_selftest_division(pxb_ctx *ctx, uint64_t seed, int64_t n_triples, int mode, int64_t *mismatches);

/* ---- task level (px/include/progressivex_python.h:4-95) --------------------------------------------------- */
/* Same scalar argument lists as findHomographies_ / findTwoViewMotions_ / find6DPoses_. Return value = number
 * of models (>= 0) or a negative pxb_status. labeling_out: N int64 (the reference's std::vector<size_t>),
 * models_out: capacity max_models_out * 9 (or 12) doubles, row-major.
 * Extra arguments the reference does not have: `seed` (the reference seeds from std::random_device,
 * gcr/uniform_random_generator.h:50-54; 0 asks for a time-based seed) . */
int pxb_find_homographies(pxb_ctx *ctx, const double *correspondences, int64_t N, int64_t *labeling_out,
                          double *models_out, int64_t max_models_out, size_t source_image_width,
                          size_t source_image_height, size_t destination_image_width,
                          size_t destination_image_height, double spatial_coherence_weight, double threshold,
                          double confidence, double neighborhood_ball_radius, double maximum_tanimoto_similarity,
                          size_t max_iters, size_t minimum_point_number, int maximum_model_number,
                          size_t sampler_id, double scoring_exponent, int do_logging, uint64_t seed);
int pxb_find_two_view_motions(pxb_ctx *ctx, const double *correspondences, int64_t N, int64_t *labeling_out,
                              double *models_out, int64_t max_models_out, size_t source_image_width,
                              size_t source_image_height, size_t destination_image_width,
                              size_t destination_image_height, double spatial_coherence_weight, double threshold,
                              double confidence, double neighborhood_ball_radius,
                              double maximum_tanimoto_similarity, size_t max_iters, size_t minimum_point_number,
                              int maximum_model_number, size_t sampler_id, double scoring_exponent,
                              int do_logging, uint64_t seed);
int pxb_find_6d_poses(pxb_ctx *ctx, const double *image_points, const double *world_points,
                      const double *intrinsics, int64_t N, int64_t *labeling_out, double *poses_out,
                      int64_t max_models_out, double spatial_coherence_weight, double threshold, double confidence,
                      double neighborhood_ball_radius, double maximum_tanimoto_similarity, size_t max_iters,
                      size_t minimum_point_number, int maximum_model_number, uint64_t seed);

/* findVanishingPoints_ (px/src/progressivex_python.cpp:306-423): lines [N,4] = segments xs ys xe ye; weights (N doubles or
 * NULL) are the per-segment weights of the least-squares fits inside PEARL (settings.point_weights, PEARL.h:373-380).
 * Only samplers 0 and 1 exist for this entry (:353-367; any other id prints the reference's message and returns 0 --
 * including the Python default sampler_id = 3). vanishing_points_out: max_models_out * 3 doubles. */
int pxb_find_vanishing_points(pxb_ctx *ctx, const double *lines, const double *weights, int64_t N, int64_t *labeling_out,
                              double *vanishing_points_out, int64_t max_models_out, size_t image_width,
                              size_t image_height, double spatial_coherence_weight, double threshold, double confidence,
                              double neighborhood_ball_radius, double maximum_tanimoto_similarity, size_t max_iters,
                              size_t minimum_point_number, int maximum_model_number, size_t sampler_id,
                              double scoring_exponent, int do_logging, uint64_t seed);
/* findLines_ (px/src/progressivex_python.cpp:425-535): points [N,2]; the weights argument of the reference is accepted and
 * unused there as well (:425-535 never reads it). Samplers 0, 1, 2. lines_out: max_models_out * 3 doubles (nx, ny, c). */
int pxb_find_lines(pxb_ctx *ctx, const double *points, const double *weights, int64_t N, int64_t *labeling_out,
                   double *lines_out, int64_t max_models_out, size_t image_width, size_t image_height,
                   double spatial_coherence_weight, double threshold, double confidence, double neighborhood_ball_radius,
                   double maximum_tanimoto_similarity, size_t max_iters, size_t minimum_point_number,
                   int maximum_model_number, size_t sampler_id, double scoring_exponent, int do_logging, uint64_t seed);

/* ---- settings and statistics (px/include/progressive_x.h:32-104,210-217) ------------------------------------------------ */
/* progx::MultiModelSettings with the gcransac::utils::Settings of the proposal engine it carries (gcr/settings.h:66-86 as
 * overridden by progressive_x.h:64-71). pxb_settings_default fills in the constructor's values. The task-level entry
 * points overwrite the fields their argument lists carry (threshold, confidence, spatial_coherence_weight,
 * maximum_tanimoto_similarity, minimum_number_of_inliers, maximum_model_number, max_iteration_number, scoring_exponent) --
 * exactly as findHomographies_ does (progressivex_python.cpp:262-276); the remaining fields are what
 * ProgressiveX::getMutableSettings() (progressive_x.h:214-217) lets a C++ caller change: pxb_ctx_set_settings installs them
 * for all later find* calls on the context (NULL restores the defaults). */
typedef struct pxb_multi_model_settings {
	size_t minimum_number_of_inliers;          /* 20 */
	size_t max_proposal_number_without_change; /* 10 */
	size_t cell_number_in_neighborhood_graph;  /* 8 (unused by the Python entry points: they build a radius graph) */
	size_t maximum_model_number;               /* SIZE_MAX */
	double maximum_tanimoto_similarity;        /* 0.5 */
	double confidence;                         /* 0.95 */
	double inlier_outlier_threshold;           /* 2.0 */
	double spatial_coherence_weight;           /* 0.14 */
	/* proposal_engine_settings */
	size_t max_iteration_number;               /* 5000 */
	size_t min_iteration_number;               /* 20 */
	size_t min_iteration_number_before_lo;     /* 20 */
	size_t max_local_optimization_number;      /* 50 */
	size_t max_graph_cut_number;               /* 10 */
	size_t max_least_squares_iterations;       /* 10 */
	size_t max_unsuccessful_model_generations; /* 100 */
	int scoring_exponent;                      /* 2 (scoring_function_with_compound_model.h:20) */
} pxb_multi_model_settings;
int pxb_settings_default(pxb_multi_model_settings *out);
int pxb_ctx_set_settings(pxb_ctx *ctx, const pxb_multi_model_settings *settings_or_null);

/* progx::IterationStatistics / MultiModelStatistics (progressive_x.h:78-104) of the LAST find* call on the context
 * (ProgressiveX::getStatistics, progressive_x.h:210-213). The four per-round times are measured on the context's CUDA
 * stream with events recorded where the reference reads its chrono clock (progressive_x.h:299-437); every round ends in a
 * synchronisation, so they are wall times of the round's phases as the device saw them, in seconds. Rounds whose proposal
 * was rejected add no entry, like the reference (:334-346 `continue` before addIterationStatistics). The labeling is
 * what the find* call returned; the reference's inliers_of_each_model only ever receives the first instance's inliers
 * (:375-381) -- its size is reported. */
#define PXB_MAX_ROUNDS 10 /* the outer loop is capped at 10 proposals (progressive_x.h:272) */
typedef struct pxb_iteration_statistics {
	double time_of_proposal_engine, time_of_model_validation, time_of_optimization, time_of_compound_model_update;
	size_t number_of_instances;
	/* RANSACStatistics of the round's proposal (gcr/statistics.h): */
	size_t ransac_iteration_number, local_optimization_number, graph_cut_number, proposal_inlier_number;
} pxb_iteration_statistics;
typedef struct pxb_multi_model_statistics {
	double processing_time, total_time_of_proposal_engine, total_time_of_model_validation, total_time_of_optimization,
	    total_time_of_compound_model_calculation;
	size_t iteration_statistics_size;
	pxb_iteration_statistics iteration_statistics[PXB_MAX_ROUNDS];
	size_t model_number, inliers_of_each_model_size;
	size_t kernel_launches; /* kernels this call launched (no counterpart in the reference) */
} pxb_multi_model_statistics;
int pxb_ctx_get_statistics(pxb_ctx *ctx, pxb_multi_model_statistics *out);

/* Many independent problems on ONE GPU (BASELINE config C4; with pxb_allgather_instances below: over several GPUs).
 * Equivalent to calling pxb_find_homographies once per pair -- correspondences[p] is [n_points[p], 4], labeling_out[p]
 * holds n_points[p] int64, models_out[p] max_models_out * 9 doubles, n_models_out[p] receives the instance count -- with
 * seed (+ p when per_pair_seed != 0). The library runs the problems on `host_threads` threads with `in_flight` problems
 * each: every problem is a fiber with its own stream and scratch that yields wherever the sequential driver would block
 * on its stream, so a couple of host threads keep the GPU busy (pxb_batch.cu). Results are identical to the one-by-one
 * calls. The reference has no counterpart (it is one problem, one thread: progressivex_python.cpp:173-304). */
int pxb_find_homographies_batch(int device, int64_t n_pairs, const double *const *correspondences, const int64_t *n_points,
                                int64_t *const *labeling_out, double *const *models_out, int64_t max_models_out,
                                int32_t *n_models_out, size_t source_image_width, size_t source_image_height,
                                size_t destination_image_width, size_t destination_image_height,
                                double spatial_coherence_weight, double threshold, double confidence,
                                double neighborhood_ball_radius, double maximum_tanimoto_similarity, size_t max_iters,
                                size_t minimum_point_number, int maximum_model_number, size_t sampler_id,
                                double scoring_exponent, uint64_t seed, int per_pair_seed, int host_threads, int in_flight);
/* Destroys the contexts the batch driver keeps per device between calls. */
void pxb_batch_release(void);

/* ---- multi-GPU exchange steps ----------------------------------------------------------------- */
/* One process per GPU; NCCL over NVLink / NVSwitch. The reference is single-process and has no counterpart: these are
 * the two places where the sharded path has a real exchange step. NCCL is resolved at run time (dlopen of the libnccl.so.2
 * the process already holds, else the system copy); without it every entry below returns PXB_ERR_UNSUPPORTED.
 * `nccl_comm` is an ncclComm_t passed as void*: either one the caller already owns (all ranks must hold the same
 * communicator with one rank per GPU) or one made by pxb_nccl_comm_init from an id that rank 0 obtained with
 * pxb_nccl_unique_id and distributed by any means (torch.distributed object broadcast in pyprogressivex.sharding). */
int pxb_nccl_version(int *version);
int pxb_nccl_unique_id(void *id128 /* 128 bytes out */);
int pxb_nccl_comm_init(pxb_ctx *ctx, const void *id128, int world, int rank, void **nccl_comm_out);
int pxb_nccl_comm_destroy(void *nccl_comm);

/* Hypothesis-block sharding of ONE problem (BASELINE configs C3 / C5). After pxb_ctx_set_shard(ctx, comm) the task-level
 * entry points pxb_find_* on this context are COLLECTIVE: every rank of the communicator calls the same function with the
 * same arguments (the points are replicated). Rank 0 runs GCRANSAC::run's control flow (gcr/GCRANSAC.h:283-518); for every
 * block of minimal samples it broadcasts the sample indices, rank r solves and scores samples [r*S, (r+1)*S) of the block
 * against all N points, and one ncclAllGather returns every slice's models, validity flags and (count, value, shared) in
 * sample order. Scores do not depend on the batch or device that evaluated them, so models and labels are bit-identical
 * to the single-GPU run with the same seed. Local optimisation, IRLS and PEARL (N x <= 50 work) stay on rank 0; the final
 * (models, labeling) is broadcast, so every rank returns the same result. NULL detaches the communicator. */
int pxb_ctx_set_shard(pxb_ctx *ctx, void *nccl_comm);
int pxb_shard_info(pxb_ctx *ctx, int *world, int *rank);

/* Independent problems sharded over ranks (BASELINE config C4): merges the surviving instances of every rank's pairs.
 * Every rank passes `pairs_per_rank` records -- counts [P] int32 (-1 = empty slot), models [P, max_models, model_size]
 * float64, labels [P, n_points] int32 -- and receives all ranks' records rank-major: counts_out [world * P], models_out
 * [world * P, max_models, model_size], labels_out [world * P, n_points]. One ncclAllGather. */
int pxb_allgather_instances(pxb_ctx *ctx, void *nccl_comm, int64_t pairs_per_rank, int64_t n_points, int32_t model_size,
                            int32_t max_models, const int32_t *counts_host, const double *models_host,
                            const int32_t *labels_host, int32_t *counts_out_host, double *models_out_host,
                            int32_t *labels_out_host);

#ifdef __cplusplus
}
#endif
#endif /* PXB200_H */
