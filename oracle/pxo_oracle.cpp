// pxo_oracle.cpp -- CPU restatement of the Progressive-X hot path. TEST INFRASTRUCTURE ONLY; see pxo_oracle.h
// for the rules about who may load this library and for the parity-pinning status.
//
// Build: g++ -O3 -ffp-contract=off -fopenmp (oracle/Makefile). -ffp-contract=off keeps every product and sum
// separately rounded, which is what the reference's x86-64 -O3 build (no -march, no -ffast-math,
// /root/reference/CMakeLists.txt:19-24) executes.
//
// Reference path abbreviations: gcr/ = graph-cut-ransac/src/pygcransac/include/, px/ = src/pyprogressivex/.
#include "pxo_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace {

// OpenCV's MIN/MAX macros (opencv2/core/cvdef.h), which is what the reference's MAX()/MIN() expand to.
// The NaN behaviour of these exact spellings matters: MAX(0, NaN) == 0.
inline double cvMAX(double a, double b) { return (a < b) ? b : a; }
inline double cvMIN(double a, double b) { return (a > b) ? b : a; }

// gcr/estimators/homography_estimator.h:181-199 (one-way transfer error, not symmetric)
inline double residual_h(const double *s, const double *h) {
	const double x1 = s[0], y1 = s[1], x2 = s[2], y2 = s[3];
	const double t1 = h[0] * x1 + h[1] * y1 + h[2];
	const double t2 = h[3] * x1 + h[4] * y1 + h[5];
	const double t3 = h[6] * x1 + h[7] * y1 + h[8];
	const double d1 = x2 - (t1 / t3);
	const double d2 = y2 - (t2 / t3);
	return d1 * d1 + d2 * d2;
}

// gcr/estimators/fundamental_estimator.h:195-222 (squared Sampson distance)
inline double residual_f(const double *s, const double *f) {
	const double x1 = s[0], y1 = s[1], x2 = s[2], y2 = s[3];
	const double e11 = f[0], e12 = f[1], e13 = f[2], e21 = f[3], e22 = f[4], e23 = f[5];
	double rxc = e11 * x2 + e21 * y2 + f[6];
	double ryc = e12 * x2 + e22 * y2 + f[7];
	double rwc = e13 * x2 + e23 * y2 + f[8];
	double r = (x1 * rxc + y1 * ryc + rwc);
	double rx = e11 * x1 + e12 * y1 + e13;
	double ry = e21 * x1 + e22 * y1 + e23;
	return r * r / (rxc * rxc + ryc * ryc + rx * rx + ry * ry);
}

// gcr/estimators/perspective_n_point_estimator.h:148-184 (squared reprojection error, normalised coords)
inline double residual_pnp(const double *s, const double *p) {
	const double u = s[0], v = s[1], x = s[2], y = s[3], z = s[4];
	const double r11 = p[0], r12 = p[1], r13 = p[2], tx = p[3];
	const double r21 = p[4], r22 = p[5], r23 = p[6], ty = p[7];
	const double r31 = p[8], r32 = p[9], r33 = p[10], tz = p[11];
	const double px = r11 * x + r12 * y + r13 * z + tx, py = r21 * x + r22 * y + r23 * z + ty,
	             pz = r31 * x + r32 * y + r33 * z + tz;
	const double pu = px / pz, pv = py / pz;
	const double du = pu - u, dv = pv - v;
	return du * du + dv * dv;
}

// px/include/vanishing_point_estimator.h:127-189: distance of the segment's start point from the line through the
// segment's midpoint and the vanishing point; squaredResidual = residual * residual (:135-137)
inline double residual_vp(const double *s, const double *d) {
	const double xs = s[0], ys = s[1], xe = s[2], ye = s[3];
	double lx, ly, lz, mx = (xs + xe) / 2.0, my = (ys + ye) / 2.0;
	lx = my * d[2] - d[1];
	ly = -(mx * d[2] - d[0]);
	lz = mx * d[1] - my * d[0];
	const double dist = std::fabs(lx * xs + ly * ys + lz) / std::sqrt(lx * lx + ly * ly);
	return dist * dist;
}

// gcr/estimators/linear_model_estimator.h:121-131 with _DimensionNumber = 2: (x nx + y ny + c)^2, accumulated from 0
inline double residual_line(const double *s, const double *d) {
	double residual = 0;
	residual += s[0] * d[0];
	residual += s[1] * d[1];
	residual += d[2];
	return residual * residual;
}

inline double residual(int type, const double *s, const double *m) {
	switch (type) {
	case PXO_MODEL_H: return residual_h(s, m);
	case PXO_MODEL_F: return residual_f(s, m);
	case PXO_MODEL_VP: return residual_vp(s, m);
	case PXO_MODEL_LINE: return residual_line(s, m);
	default: return residual_pnp(s, m);
	}
}

// Eigen 3.3/3.4 Core/Redux.h, redux_impl<Func, Evaluator, LinearVectorizedTraversal, NoUnrolling> with SSE2
// packets of two doubles, for sum(f(i)) over a dynamic-size vector whose storage is 16-byte aligned.
template <class F> double eigen_redux_sum(int64_t n, F f) {
	if (n == 0) return 0.0;
	const int64_t ps = 2;
	const int64_t aligned2 = (n / (2 * ps)) * (2 * ps);
	const int64_t aligned = (n / ps) * ps;
	double res;
	if (aligned) {
		double p0[2] = {f(0), f(1)};
		if (aligned > ps) {
			double p1[2] = {f(2), f(3)};
			for (int64_t i = 2 * ps; i < aligned2; i += 2 * ps) {
				p0[0] = p0[0] + f(i);
				p0[1] = p0[1] + f(i + 1);
				p1[0] = p1[0] + f(i + 2);
				p1[1] = p1[1] + f(i + 3);
			}
			p0[0] = p0[0] + p1[0];
			p0[1] = p0[1] + p1[1];
			if (aligned > aligned2) {
				p0[0] = p0[0] + f(aligned2);
				p0[1] = p0[1] + f(aligned2 + 1);
			}
		}
		res = p0[0] + p0[1];
		for (int64_t i = aligned; i < n; ++i) res = res + f(i);
	} else {
		res = f(0);
		for (int64_t i = 1; i < n; ++i) res = res + f(i);
	}
	return res;
}

} // namespace

extern "C" {

int pxo_point_dim(int t) { return t == PXO_MODEL_PNP ? 5 : (t == PXO_MODEL_LINE ? 2 : 4); }
int pxo_model_size(int t) { return t == PXO_MODEL_PNP ? 12 : (t >= PXO_MODEL_VP ? 3 : 9); }
int pxo_sample_size(int t) {
	switch (t) {
	case PXO_MODEL_H: return 4;
	case PXO_MODEL_F: return 7;
	case PXO_MODEL_PNP: return 3;
	default: return 2;
	}
}

double pxo_squared_residual(int type, const double *point, const double *model) {
	return residual(type, point, model);
}

void pxo_residual_matrix(int type, const double *pts, int64_t N, const double *models, int64_t K, double T2,
                         double *r2, uint32_t *mask) {
	const int d = pxo_point_dim(type), ms = pxo_model_size(type);
	const int64_t words = (N + 31) / 32;
	if (mask) std::memset(mask, 0, sizeof(uint32_t) * (size_t)(K * words));
	for (int64_t k = 0; k < K; ++k) {
		const double *m = models + k * ms;
		for (int64_t i = 0; i < N; ++i) {
			const double r = residual(type, pts + i * d, m);
			if (r2) r2[k * N + i] = r;
			if (mask && r < T2) mask[k * words + (i >> 5)] |= (1u << (i & 31));
		}
	}
}

// px/include/scoring_function_with_compound_model.h:61-125
double pxo_get_score(int type, const double *pts, int64_t N, const double *model, double T2,
                     const double *compound_pref, int exponent, int64_t best_inlier_number, int64_t *count_out,
                     double *value_sum_out, double *shared_out, int64_t *inliers) {
	const int d = pxo_point_dim(type);
	int64_t count = 0;
	double value = 0.0;
	std::vector<double> pref((size_t)N, 0.0); // :75
	for (int64_t i = 0; i < N; ++i) {
		const double r2 = residual(type, pts + i * d, model); // :81
		if (r2 < T2) {                                         // :85
			if (inliers) inliers[count] = i;
			++count;
			const double sv = cvMAX(0.0, 1.0 - r2 / T2); // :94
			value += sv;                                 // :97
			pref[(size_t)i] = sv;                        // :100
		}
		// :105-106, size_t arithmetic in the reference; N - i + count never underflows
		if ((uint64_t)(N - i + count) < (uint64_t)best_inlier_number) {
			*count_out = 0;
			if (value_sum_out) *value_sum_out = 0.0;
			if (shared_out) *shared_out = 0.0;
			return 0.0;
		}
	}
	double shared = 0.0;
	double final_value = value;
	if (compound_pref) { // :110-121
		for (int64_t i = 0; i < N; ++i) shared += cvMIN(compound_pref[i], pref[(size_t)i]);
		final_value = value - std::pow(shared, exponent);
	}
	*count_out = count;
	if (value_sum_out) *value_sum_out = value;
	if (shared_out) *shared_out = shared;
	return final_value;
}

void pxo_score_batch(int type, const double *pts, int64_t N, const double *models, int64_t K, double T2,
                     const double *compound_pref, int64_t *count, double *value_sum, double *shared,
                     int threads) {
	const int d = pxo_point_dim(type), ms = pxo_model_size(type);
	(void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
	for (int64_t k = 0; k < K; ++k) {
		const double *m = models + k * ms;
		int64_t c = 0;
		double v = 0.0, sh = 0.0;
		// one pass: shared support only receives contributions from inliers (pref is 0 elsewhere and
		// MIN(cp, 0) adds +0.0 for cp >= 0), so folding the two reference loops keeps the same sums
		// whenever compound_pref has no negative / NaN entries -- which holds by construction (:94).
		for (int64_t i = 0; i < N; ++i) {
			const double r2 = residual(type, pts + i * d, m);
			if (r2 < T2) {
				++c;
				const double sv = cvMAX(0.0, 1.0 - r2 / T2);
				v += sv;
				if (compound_pref) sh += cvMIN(compound_pref[i], sv);
			}
		}
		count[k] = c;
		value_sum[k] = v;
		shared[k] = sh;
	}
}

// px/include/progx_model.h:70-87
void pxo_preference_vector(int type, const double *pts, int64_t N, const double *model, double T, double *pref) {
	const int d = pxo_point_dim(type);
	for (int64_t i = 0; i < N; ++i) {
		const double r2 = residual(type, pts + i * d, model);
		pref[i] = cvMAX(0, 1.0 - r2 / T);
	}
}

// px/include/progressive_x.h:583-588
double pxo_tanimoto(const double *a, const double *b, int64_t N) {
	const double dot = eigen_redux_sum(N, [&](int64_t i) { return a[i] * b[i]; });
	const double na = eigen_redux_sum(N, [&](int64_t i) { return a[i] * a[i]; });
	const double nb = eigen_redux_sum(N, [&](int64_t i) { return b[i] * b[i]; });
	return dot / (na + nb - dot);
}

// px/include/progressive_x.h:597-624
void pxo_compound_max(const double *prefs, int64_t L, int64_t N, double *out) {
	for (int64_t i = 0; i < N; ++i) out[i] = 0.0;
	for (int64_t k = 0; k < L; ++k)
		for (int64_t i = 0; i < N; ++i) out[i] = cvMAX(out[i], prefs[k * N + i]);
}

// ---------------------------------------------------------------------------------------------
// a6: homography four-point solver
// ---------------------------------------------------------------------------------------------

// gcr/math_utils.h:45-87, gaussElimination<8> on an 8x9 augmented matrix
static void gauss_elimination8(double m[8][9], double result[8]) {
	const int S = 8;
	int i, j, k;
	double temp;
	for (i = 0; i < S; i++) // "Pivotisation": pre-pass of row swaps only (:54-62)
		for (k = i + 1; k < S; k++)
			if (std::fabs(m[i][i]) < std::fabs(m[k][i]))
				for (j = 0; j <= S; j++) {
					temp = m[i][j];
					m[i][j] = m[k][j];
					m[k][j] = temp;
				}
	for (i = 0; i < S - 1; i++) // elimination without further pivoting (:65-72)
		for (k = i + 1; k < S; k++) {
			double t = m[k][i] / m[i][i];
			for (j = 0; j <= S; j++) m[k][j] = m[k][j] - t * m[i][j];
		}
	for (i = S - 1; i >= 0; i--) { // back-substitution (:75-86)
		result[i] = m[i][S];
		for (j = i + 1; j < S; j++)
			if (j != i) result[i] = result[i] - m[i][j] * result[j];
		result[i] = result[i] / m[i][i];
	}
}

// gcr/estimators/solver_homography_four_point.h:109-190 (weights_ == nullptr => weight = 1.0)
int pxo_h4_solve(const double *pts, const int64_t *sample, double *H) {
	double c[8][9];
	int row = 0;
	const double weight = 1.0;
	for (int i = 0; i < 4; ++i) {
		const double *p = pts + 4 * sample[i];
		const double x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
		const double mwx1 = -weight * x1, mwy1 = -weight * y1, wx2 = weight * x2, wy2 = weight * y2;
		c[row][0] = mwx1; c[row][1] = mwy1; c[row][2] = -weight;
		c[row][3] = 0; c[row][4] = 0; c[row][5] = 0;
		c[row][6] = wx2 * x1; c[row][7] = wx2 * y1; c[row][8] = -wx2;
		++row;
		c[row][0] = 0; c[row][1] = 0; c[row][2] = 0;
		c[row][3] = mwx1; c[row][4] = mwy1; c[row][5] = -weight;
		c[row][6] = wy2 * x1; c[row][7] = wy2 * y1; c[row][8] = -wy2;
		++row;
	}
	double h[8];
	gauss_elimination8(c, h);
	for (int i = 0; i < 8; ++i)
		if (std::isnan(h[i])) return 0; // :181 hasNaN
	for (int i = 0; i < 8; ++i) H[i] = h[i];
	H[8] = 1.0;
	return 1;
}

// gcr/estimators/homography_estimator.h:312-322 (cross_product with st_ = 1) and :346-381
static inline void h_cross(double r[3], const double *v1, const double *v2) {
	r[0] = v1[1] - v2[1];
	r[1] = v2[0] - v1[0];
	r[2] = v1[0] * v2[1] - v1[1] * v2[0];
}
int pxo_h4_is_valid_sample(const double *pts, const int64_t *sample) {
	const double *a = pts + 4 * sample[0], *b = pts + 4 * sample[1], *c = pts + 4 * sample[2],
	             *d = pts + 4 * sample[3];
	double p[3], q[3];
	h_cross(p, a, b);
	h_cross(q, a + 2, b + 2);
	if ((p[0] * c[0] + p[1] * c[1] + p[2]) * (q[0] * c[2] + q[1] * c[3] + q[2]) < 0) return 0;
	if ((p[0] * d[0] + p[1] * d[1] + p[2]) * (q[0] * d[2] + q[1] * d[3] + q[2]) < 0) return 0;
	h_cross(p, c, d);
	h_cross(q, c + 2, d + 2);
	if ((p[0] * a[0] + p[1] * a[1] + p[2]) * (q[0] * a[2] + q[1] * a[3] + q[2]) < 0) return 0;
	if ((p[0] * b[0] + p[1] * b[1] + p[2]) * (q[0] * b[2] + q[1] * b[3] + q[2]) < 0) return 0;
	return 1;
}

// gcr/estimators/homography_estimator.h:326-342. model.descriptor is a dynamic-size Eigen::MatrixXd, for which
// Eigen's determinant() goes through PartialPivLU (Eigen/src/LU/Determinant.h, determinant_impl<Derived, Dynamic>);
// restated here: unblocked partial-pivot LU (first maximal |entry| in the column wins), det = sign * (u00*u11)*u22.
static double det3_partial_piv_lu(const double *M) {
	double a[3][3] = {{M[0], M[1], M[2]}, {M[3], M[4], M[5]}, {M[6], M[7], M[8]}};
	int sign = 1;
	for (int k = 0; k < 3; ++k) {
		int piv = k;
		double best = std::fabs(a[k][k]);
		for (int i = k + 1; i < 3; ++i)
			if (std::fabs(a[i][k]) > best) {
				best = std::fabs(a[i][k]);
				piv = i;
			}
		if (best == 0.0) continue; // Eigen skips the elimination step for a zero pivot column
		if (piv != k) {
			for (int j = 0; j < 3; ++j) std::swap(a[k][j], a[piv][j]);
			sign = -sign;
		}
		for (int i = k + 1; i < 3; ++i) a[i][k] = a[i][k] / a[k][k];
		for (int i = k + 1; i < 3; ++i)
			for (int j = k + 1; j < 3; ++j) a[i][j] = a[i][j] - a[i][k] * a[k][j];
	}
	return (double)sign * ((a[0][0] * a[1][1]) * a[2][2]);
}
int pxo_h_is_valid_model(const double *H) {
	const double det = det3_partial_piv_lu(H);
	return !(std::fabs(det) < 1e-2);
}

// ---------------------------------------------------------------------------------------------
// a7: fundamental matrix seven-point solver
// ---------------------------------------------------------------------------------------------

// Real roots of c0 + c1 x + c2 x^2 + c3 x^3 (c3 != 0), ascending. Stands in for Eigen::PolynomialSolver<double,3>
// ::realRoots (companion-matrix eigenvalues, |imag| < 1e-12): closed form + Newton polish on the original cubic.
static int cubic_real_roots(const double c[4], double roots[3]) {
	const double a2 = c[2] / c[3], a1 = c[1] / c[3], a0 = c[0] / c[3];
	const double Q = (3.0 * a1 - a2 * a2) / 9.0;
	const double R = (9.0 * a2 * a1 - 27.0 * a0 - 2.0 * a2 * a2 * a2) / 54.0;
	const double D = Q * Q * Q + R * R;
	int n = 0;
	if (D > 0) {
		const double sD = std::sqrt(D);
		const double S = std::cbrt(R + sD), T = std::cbrt(R - sD);
		roots[n++] = S + T - a2 / 3.0;
	} else {
		const double sq = std::sqrt(-Q);
		double ct = (sq > 0) ? R / (sq * sq * sq) : 0.0;
		ct = std::min(1.0, std::max(-1.0, ct));
		const double theta = std::acos(ct);
		const double kPi = 3.14159265358979323846;
		for (int k = 0; k < 3; ++k) roots[n++] = 2.0 * sq * std::cos((theta + 2.0 * kPi * k) / 3.0) - a2 / 3.0;
	}
	for (int i = 0; i < n; ++i) { // Newton polish on the monic cubic
		double x = roots[i];
		for (int it = 0; it < 8; ++it) {
			const double f = ((x + a2) * x + a1) * x + a0;
			const double df = (3.0 * x + 2.0 * a2) * x + a1;
			if (df == 0.0) break;
			const double step = f / df;
			x -= step;
			if (std::fabs(step) <= 1e-16 * std::fabs(x)) break;
		}
		roots[i] = x;
	}
	std::sort(roots, roots + n);
	return n;
}

// gcr/estimators/fundamental_estimator.h:737-800
static void f_epipole(double e[3], const double *F) {
	const double eps = 1.9984e-15;
	const double *r0 = F, *r1 = F + 3, *r2 = F + 6;
	e[0] = r0[1] * r2[2] - r0[2] * r2[1];
	e[1] = r0[2] * r2[0] - r0[0] * r2[2];
	e[2] = r0[0] * r2[1] - r0[1] * r2[0];
	for (int i = 0; i < 3; ++i)
		if (e[i] > eps || e[i] < -eps) return;
	e[0] = r1[1] * r2[2] - r1[2] * r2[1];
	e[1] = r1[2] * r2[0] - r1[0] * r2[2];
	e[2] = r1[0] * r2[1] - r1[1] * r2[0];
}
static inline double f_signum(const double *F, const double *e, const double *p) {
	const double s1 = F[0] * p[2] + F[3] * p[3] + F[6], s2 = e[1] - e[2] * p[1];
	return s1 * s2;
}
static int f_orientation_valid(const double *F, const double *pts, const int64_t *sample, int n) {
	double e[3];
	f_epipole(e, F);
	const double s2 = f_signum(F, e, pts + 4 * sample[0]);
	for (int i = 1; i < n; ++i) {
		const double s1 = f_signum(F, e, pts + 4 * sample[i]);
		if (s2 * s1 < 0) return 0;
	}
	return 1;
}

int pxo_f7_solve(const double *pts, const int64_t *sample, double *F_out, int apply_orientation_test) {
	// :98-155 coefficient matrix (weights_ == nullptr)
	double A[7][9];
	for (int i = 0; i < 7; ++i) {
		const double *p = pts + 4 * sample[i];
		const double x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
		A[i][0] = x1 * x0; A[i][1] = x1 * y0; A[i][2] = x1;
		A[i][3] = y1 * x0; A[i][4] = y1 * y0; A[i][5] = y1;
		A[i][6] = x0; A[i][7] = y0; A[i][8] = 1;
	}
	// :161-169 Eigen::FullPivLU<MatrixXd>(7x9): dimensionOfKernel() must be 2, kernel() basis.
	// Restated from Eigen/src/LU/FullPivLU.h (computeInPlace, rank, kernel_retval::evalTo).
	int colidx[9];
	for (int j = 0; j < 9; ++j) colidx[j] = j;
	double maxpivot = 0.0;
	int nonzero_pivots = 7;
	for (int k = 0; k < 7; ++k) {
		int pr = k, pc = k;
		double biggest = -1.0;
		for (int j = k; j < 9; ++j) // column-major visitor, first strictly-greater wins
			for (int i = k; i < 7; ++i)
				if (std::fabs(A[i][j]) > biggest) {
					biggest = std::fabs(A[i][j]);
					pr = i;
					pc = j;
				}
		if (biggest == 0.0) {
			nonzero_pivots = k;
			break;
		}
		if (biggest > maxpivot) maxpivot = biggest;
		if (pr != k)
			for (int j = 0; j < 9; ++j) std::swap(A[k][j], A[pr][j]);
		if (pc != k) {
			for (int i = 0; i < 7; ++i) std::swap(A[i][k], A[i][pc]);
			std::swap(colidx[k], colidx[pc]);
		}
		for (int i = k + 1; i < 7; ++i) A[i][k] = A[i][k] / A[k][k];
		for (int i = k + 1; i < 7; ++i)
			for (int j = k + 1; j < 9; ++j) A[i][j] = A[i][j] - A[i][k] * A[k][j];
	}
	const double thresh = std::fabs(maxpivot) * (std::numeric_limits<double>::epsilon() * 7.0);
	int rank = 0;
	for (int i = 0; i < nonzero_pivots; ++i)
		if (std::fabs(A[i][i]) > thresh) ++rank;
	if (9 - rank != 2) return 0;
	// kernel: solve U(7x7) X = B(7x2) (upper triangular, column-oriented back substitution), kernel = [-X; I]
	double X[7][2];
	for (int k = 0; k < 2; ++k) {
		double x[7];
		for (int i = 0; i < 7; ++i) x[i] = A[i][7 + k];
		for (int i = 6; i >= 0; --i) {
			x[i] = x[i] / A[i][i];
			for (int j = 0; j < i; ++j) x[j] = x[j] - x[i] * A[j][i];
		}
		for (int i = 0; i < 7; ++i) X[i][k] = x[i];
	}
	double f1[9], f2[9];
	for (int i = 0; i < 7; ++i) {
		f1[colidx[i]] = -X[i][0];
		f2[colidx[i]] = -X[i][1];
	}
	f1[colidx[7]] = 1.0; f2[colidx[7]] = 0.0;
	f1[colidx[8]] = 0.0; f2[colidx[8]] = 1.0;

	// :194-242 cubic det(lambda*f1 + (1-lambda)*f2) = 0
	for (int i = 0; i < 9; ++i) f1[i] -= f2[i];
	double c[4], t0, t1, t2;
	t0 = f2[4] * f2[8] - f2[5] * f2[7];
	t1 = f2[3] * f2[8] - f2[5] * f2[6];
	t2 = f2[3] * f2[7] - f2[4] * f2[6];
	c[0] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
	c[1] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) +
	       f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) - f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) +
	       f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
	       f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
	t0 = f1[4] * f1[8] - f1[5] * f1[7];
	t1 = f1[3] * f1[8] - f1[5] * f1[6];
	t2 = f1[3] * f1[7] - f1[4] * f1[6];
	c[2] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) +
	       f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) - f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) +
	       f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
	       f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
	c[3] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
	// :246-251
	const double eps = std::numeric_limits<double>::epsilon();
	if (std::fabs(c[0] + c[1] + c[2] + c[3]) < 1e-9 || std::fabs(c[0]) < eps || std::fabs(c[1]) < eps ||
	    std::fabs(c[2]) < eps || std::fabs(c[3]) < eps)
		return 0;
	double roots[3];
	const int n = cubic_real_roots(c, roots);
	if (n < 1 || n > 3) return 0;
	int kept = 0;
	for (int r = 0; r < n; ++r) { // :266-287
		double lambda = roots[r], mu = 1.;
		const double s = f1[8] * roots[r] + f2[8];
		if (std::fabs(s) > eps) {
			mu = 1.0f / s;
			lambda *= mu;
			double F[9];
			for (int i = 0; i < 9; ++i) F[i] = f1[i] * lambda + f2[i] * mu;
			F[8] = 1.0;
			// fundamental_estimator.h:175-181 erases models failing the oriented epipolar constraint
			if (apply_orientation_test && !f_orientation_valid(F, pts, sample, 7)) continue;
			for (int i = 0; i < 9; ++i) F_out[kept * 9 + i] = F[i];
			++kept;
		}
	}
	return kept;
}

// ---------------------------------------------------------------------------------------------
// a8: P3P, gcr/estimators/solver_p3p.h
// ---------------------------------------------------------------------------------------------
namespace {
struct V3 {
	double x, y, z;
};
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// :108-142 (M symmetric 3x3 given as M[r][c])
void eig3x3known0(const double M[3][3], double E[3][2], double &sig1, double &sig2) {
	double p1 = -M[0][0] - M[1][1] - M[2][2];
	double p0 = -M[0][1] * M[0][1] - M[0][2] * M[0][2] - M[1][2] * M[1][2] + M[0][0] * (M[1][1] + M[2][2]) +
	            M[1][1] * M[2][2];
	double disc = std::sqrt(p1 * p1 / 4.0 - p0);
	double tmp = -p1 / 2.0;
	sig1 = tmp + disc;
	sig2 = tmp - disc;
	if (std::fabs(sig1) < std::fabs(sig2)) std::swap(sig1, sig2);
	double c = sig1 * sig1 + M[0][0] * M[1][1] - sig1 * (M[0][0] + M[1][1]) - M[0][1] * M[0][1];
	double a1 = (sig1 * M[0][2] + M[0][1] * M[1][2] - M[0][2] * M[1][1]) / c;
	double a2 = (sig1 * M[1][2] + M[0][1] * M[0][2] - M[0][0] * M[1][2]) / c;
	double n = 1.0 / std::sqrt(1 + a1 * a1 + a2 * a2);
	E[0][0] = a1 * n; E[1][0] = a2 * n; E[2][0] = n;
	c = sig2 * sig2 + M[0][0] * M[1][1] - sig2 * (M[0][0] + M[1][1]) - M[0][1] * M[0][1];
	a1 = (sig2 * M[0][2] + M[0][1] * M[1][2] - M[0][2] * M[1][1]) / c;
	a2 = (sig2 * M[1][2] + M[0][1] * M[0][2] - M[0][0] * M[1][2]) / c;
	n = 1.0 / std::sqrt(1 + a1 * a1 + a2 * a2);
	E[0][1] = a1 * n; E[1][1] = a2 * n; E[2][1] = n;
}

// :145-175
void refine_lambda(double &l1, double &l2, double &l3, double a12, double a13, double a23, double b12, double b13,
                   double b23) {
	for (int iter = 0; iter < 5; ++iter) {
		double r1 = (l1 * l1 - 2.0 * l1 * l2 * b12 + l2 * l2 - a12);
		double r2 = (l1 * l1 - 2.0 * l1 * l3 * b13 + l3 * l3 - a13);
		double r3 = (l2 * l2 - 2.0 * l2 * l3 * b23 + l3 * l3 - a23);
		if (std::fabs(r1) + std::fabs(r2) + std::fabs(r3) < 1e-10) return;
		double x11 = l1 - l2 * b12, x12 = l2 - l1 * b12, x21 = l1 - l3 * b13, x23 = l3 - l1 * b13,
		       x32 = l2 - l3 * b23, x33 = l3 - l2 * b23;
		double detJ = 0.5 / (x11 * x23 * x32 + x12 * x21 * x33);
		l1 += (-x23 * x32 * r1 - x12 * x33 * r2 + x12 * x23 * r3) * detJ;
		l2 += (-x21 * x33 * r1 + x11 * x33 * r2 - x11 * x23 * r3) * detJ;
		l3 += (x21 * x32 * r1 - x11 * x32 * r2 - x12 * x21 * r3) * detJ;
	}
}

// 3x3 matrices as columns
struct M3 {
	V3 c0, c1, c2;
};
inline double at(const M3 &m, int r, int c) {
	const V3 &v = c == 0 ? m.c0 : (c == 1 ? m.c1 : m.c2);
	return r == 0 ? v.x : (r == 1 ? v.y : v.z);
}
// Eigen Matrix3d::inverse(): cofactor formula (Eigen/src/LU/InverseImpl.h, compute_inverse<.,.,3>)
M3 inverse3(const M3 &m) {
	auto cof = [&](int i, int j) {
		const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
		return at(m, i1, j1) * at(m, i2, j2) - at(m, i1, j2) * at(m, i2, j1);
	};
	// cofactors_col0 = (cof(0,0), cof(1,0), cof(2,0)); det = sum(cofactors_col0 .* matrix.col(0))
	const double c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
	const double det = c00 * at(m, 0, 0) + c10 * at(m, 1, 0) + c20 * at(m, 2, 0);
	const double invdet = 1.0 / det;
	M3 r;
	// result(r,c) = cofactor(c,r) * invdet
	r.c0 = {c00 * invdet, cof(0, 1) * invdet, cof(0, 2) * invdet};
	r.c1 = {c10 * invdet, cof(1, 1) * invdet, cof(1, 2) * invdet};
	r.c2 = {c20 * invdet, cof(2, 1) * invdet, cof(2, 2) * invdet};
	return r;
}
inline V3 mulv(const M3 &m, V3 v) { // M*v, sum over k in order
	return {at(m, 0, 0) * v.x + at(m, 0, 1) * v.y + at(m, 0, 2) * v.z,
	        at(m, 1, 0) * v.x + at(m, 1, 1) * v.y + at(m, 1, 2) * v.z,
	        at(m, 2, 0) * v.x + at(m, 2, 1) * v.y + at(m, 2, 2) * v.z};
}
inline M3 mulm(const M3 &a, const M3 &b) { return {mulv(a, b.c0), mulv(a, b.c1), mulv(a, b.c2)}; }
} // namespace

int pxo_p3p_solve(const double *pts, const int64_t *sample, double *P_out) {
	V3 y[3], X[3];
	for (int i = 0; i < 3; ++i) { // :192-203
		const double *p = pts + 5 * sample[i];
		V3 v = {p[0], p[1], 1.0};
		const double z = dot(v, v); // Eigen normalize(): v /= sqrt(squaredNorm) when > 0
		if (z > 0) {
			const double nrm = std::sqrt(z);
			v = {v.x / nrm, v.y / nrm, v.z / nrm};
		}
		y[i] = v;
		X[i] = {p[2], p[3], p[4]};
	}
	V3 dX12 = X[0] - X[1], dX13 = X[0] - X[2], dX23 = X[1] - X[2];
	double a12 = dot(dX12, dX12), b12 = dot(y[0], y[1]);
	double a13 = dot(dX13, dX13), b13 = dot(y[0], y[2]);
	double a23 = dot(dX23, dX23), b23 = dot(y[1], y[2]);
	double a23b12 = a23 * b12, a12b23 = a12 * b23, a23b13 = a23 * b13, a13b23 = a13 * b23;
	// :224-225 (row-wise comma initialisers; symmetric)
	M3 D1 = {{a23, -a23b12, 0.0}, {-a23b12, a23 - a12, a12b23}, {0.0, a12b23, -a12}};
	M3 D2 = {{a23, 0.0, -a23b13}, {0.0, -a13, a13b23}, {-a23b13, a13b23, a23 - a13}};
	M3 DX1 = {cross(D1.c1, D1.c2), cross(D1.c2, D1.c0), cross(D1.c0, D1.c1)};
	M3 DX2 = {cross(D2.c1, D2.c2), cross(D2.c2, D2.c0), cross(D2.c0, D2.c1)};
	auto sumprod = [](const M3 &a, const M3 &b) { // (A.array()*B.array()).sum(), column-major order
		return dot(a.c0, b.c0) + dot(a.c1, b.c1) + dot(a.c2, b.c2);
	};
	double c3 = dot(D2.c0, DX2.c0);
	double c2 = sumprod(D1, DX2);
	double c1 = sumprod(D2, DX1);
	double c0 = dot(D1.c0, DX1.c0);
	const double c3inv = 1.0 / c3;
	c2 *= c3inv; c1 *= c3inv; c0 *= c3inv;
	double a = c1 - c2 * c2 / 3.0;
	double b = (2.0 * c2 * c2 * c2 - 9.0 * c2 * c1) / 27.0 + c0;
	double c = b * b / 4.0 + a * a * a / 27.0;
	double gamma;
	if (c > 0) {
		c = std::sqrt(c);
		b *= -0.5;
		gamma = std::cbrt(b + c) + std::cbrt(b - c) - c2 / 3.0;
	} else {
		c = 3.0 * b / (2.0 * a) * std::sqrt(-3.0 / a);
		gamma = 2.0 * std::sqrt(-a / 3.0) * std::cos(std::acos(c) / 3.0) - c2 / 3.0;
	}
	double f = gamma * gamma * gamma + c2 * gamma * gamma + c1 * gamma + c0;
	double df = 3.0 * gamma * gamma + 2.0 * c2 * gamma + c1;
	gamma = gamma - f / df;

	double D0[3][3];
	for (int r = 0; r < 3; ++r)
		for (int cc = 0; cc < 3; ++cc) D0[r][cc] = at(D1, r, cc) + gamma * at(D2, r, cc);
	double E[3][2], sig1, sig2;
	eig3x3known0(D0, E, sig1, sig2);
	double s = std::sqrt(-sig2 / sig1);
	double lambda1, lambda2, lambda3;
	M3 XX = {dX12, dX13, cross(dX12, dX13)};
	XX = inverse3(XX);
	const double TOL_DOUBLE_ROOT = 1e-12;
	int nsol = 0;
	auto emit = [&](double l1, double l2, double l3) {
		refine_lambda(l1, l2, l3, a12, a13, a23, b12, b13, b23);
		V3 v1 = l1 * y[0] - l2 * y[1];
		V3 v2 = l1 * y[0] - l3 * y[2];
		M3 YY = {v1, v2, cross(v1, v2)};
		M3 R = mulm(YY, XX);
		V3 t = l1 * y[0] - mulv(R, X[0]);
		double *P = P_out + 12 * nsol;
		for (int r = 0; r < 3; ++r) {
			for (int cc = 0; cc < 3; ++cc) P[4 * r + cc] = at(R, r, cc);
		}
		P[3] = t.x; P[7] = t.y; P[11] = t.z;
		++nsol;
	};
	for (int s_flip = 0; s_flip < 2; ++s_flip, s = -s) {
		double u1 = E[0][0] - s * E[0][1];
		double u2 = E[1][0] - s * E[1][1];
		double u3 = E[2][0] - s * E[2][1];
		bool switch_12 = std::fabs(u1) < std::fabs(u2);
		double qa, qb, qc, w0, w1;
		if (switch_12) {
			w0 = -u1 / u2;
			w1 = -u3 / u2;
			qa = -a13 * w1 * w1 + 2 * a13b23 * w1 - a13 + a23;
			qb = 2 * a13b23 * w0 - 2 * a23b13 - 2 * a13 * w0 * w1;
			qc = -a13 * w0 * w0 + a23;
			double b2m4ac = qb * qb - 4.0 * qa * qc;
			if (b2m4ac < -TOL_DOUBLE_ROOT) continue;
			double sq = std::sqrt(std::max(0.0, b2m4ac));
			double tau = (qb > 0) ? (2.0 * qc) / (-qb - sq) : (2.0 * qc) / (-qb + sq);
			for (int tau_flip = 0; tau_flip < 2; ++tau_flip, tau = qc / (qa * tau)) {
				if (tau > 0) {
					lambda1 = std::sqrt(a13 / (tau * (tau - 2.0 * b13) + 1.0));
					lambda3 = tau * lambda1;
					lambda2 = w0 * lambda1 + w1 * lambda3;
					if (lambda2 < 0) continue;
					emit(lambda1, lambda2, lambda3);
				}
				if (b2m4ac < TOL_DOUBLE_ROOT) break;
			}
		} else {
			w0 = -u2 / u1;
			w1 = -u3 / u1;
			qa = (a13 - a12) * w1 * w1 + 2.0 * a12 * b13 * w1 - a12;
			qb = -2.0 * a13 * b12 * w1 + 2.0 * a12 * b13 * w0 - 2.0 * w0 * w1 * (a12 - a13);
			qc = (a13 - a12) * w0 * w0 - 2.0 * a13 * b12 * w0 + a13;
			double b2m4ac = qb * qb - 4.0 * qa * qc;
			if (b2m4ac < -TOL_DOUBLE_ROOT) continue;
			double sq = std::sqrt(std::max(0.0, b2m4ac));
			double tau = (qb > 0) ? (2.0 * qc) / (-qb - sq) : (2.0 * qc) / (-qb + sq);
			for (int tau_flip = 0; tau_flip < 2; ++tau_flip, tau = qc / (qa * tau)) {
				if (tau > 0) {
					lambda2 = std::sqrt(a23 / (tau * (tau - 2.0 * b23) + 1.0));
					lambda3 = tau * lambda2;
					lambda1 = w0 * lambda2 + w1 * lambda3;
					if (lambda1 < 0) continue;
					emit(lambda1, lambda2, lambda3);
				}
				if (b2m4ac < TOL_DOUBLE_ROOT) break;
			}
		}
	}
	return nsol;
}

// ---------------------------------------------------------------------------------------------
// a9 / a12 / a13
// ---------------------------------------------------------------------------------------------

// px/include/PEARL.h:41-56 (thresholds) and :82-128 (dataEnergyFunctor)
void pxo_pearl_datacost(int type, const double *pts, int64_t N, const double *models, int64_t L, double thr,
                        double lambda, double *D) {
	const int d = pxo_point_dim(type), ms = pxo_model_size(type);
	const double one_minus = 1.0 - lambda;
	const double T = 9.0 / 4.0 * thr * thr;
	for (int64_t i = 0; i < N; ++i) {
		for (int64_t l = 0; l < L; ++l) {
			const double r2 = residual(type, pts + i * d, models + l * ms);
			D[i * (L + 1) + l] = (r2 > T) ? 2.0 * one_minus : one_minus * r2 / T;
		}
		D[i * (L + 1) + L] = one_minus;
	}
}

// px/include/PEARL.h:369-371 / :388-390
void pxo_segment_residual_sums(int type, const double *pts, int64_t N, const double *models, int64_t L,
                               const int32_t *labels, double *sums, int64_t *counts) {
	const int d = pxo_point_dim(type), ms = pxo_model_size(type);
	for (int64_t l = 0; l < L; ++l) {
		sums[l] = 0.0;
		counts[l] = 0;
	}
	for (int64_t i = 0; i < N; ++i) {
		const int32_t l = labels[i];
		if (l < 0 || l >= L) continue;
		sums[l] += std::sqrt(residual(type, pts + i * d, models + l * ms));
		counts[l]++;
	}
}

// gcr/GCRANSAC.h:937-962
void pxo_lo_unary_terms(int type, const double *pts, int64_t N, const double *model, double thr, double lambda,
                        double *dd, double *e0, double *e1) {
	const int d = pxo_point_dim(type);
	const double T = thr * thr * 9 / 4;
	const double one_minus_lambda = 1.0 - lambda;
	for (int64_t i = 0; i < N; ++i) {
		const double r2 = residual(type, pts + i * d, model);
		const double dist = std::clamp(r2 / T, 0.0, 1.0);
		dd[i] = dist;
		const double tmp_energy = 1 - dist;
		if (r2 <= T) {
			e0[i] = one_minus_lambda * tmp_energy;
			e1[i] = 0;
		} else {
			e0[i] = 0;
			e1[i] = one_minus_lambda * (1 - tmp_energy);
		}
	}
}

// gcr/GCRANSAC.h:658-669
void pxo_tukey_weights(int type, const double *pts, const int64_t *inliers, int64_t n, const double *model,
                       double T2, double *weights) {
	const int d = pxo_point_dim(type);
	for (int64_t j = 0; j < n; ++j) {
		const int64_t i = inliers[j];
		const double r2 = residual(type, pts + i * d, model);
		const double w = cvMAX(0.0, 1.0 - r2 / T2);
		weights[i] = w * w;
	}
}

// Non-minimal homography fit: RobustHomographyEstimator::estimateModelNonminimal (gcr/estimators/homography_estimator.h:
// 140-173) = normalizePoints (:201-309) + HomographyFourPointSolver::estimateNonMinimalModel (solver_homography_four_point.h:
// 192-264, Eigen colPivHouseholderQr().solve on the 2n x 8 system, restated as an unblocked Householder QR with column
// pivoting) + denormalisation H = T2^-1 Hn T1. weights_by_row follows the reference's indexing (weights_[i], i = row of
// the gathered sample). Returns 1 on success.
int pxo_fit_h_nonminimal(const double *pts, const int64_t *idx, int64_t n, const double *weights_by_row, double *H) {
	if (n < 4) return 0;
	double m1x = 0, m1y = 0, m2x = 0, m2y = 0;
	for (int64_t i = 0; i < n; ++i) {
		const double *p = pts + 4 * idx[i];
		m1x += p[0]; m1y += p[1]; m2x += p[2]; m2y += p[3];
	}
	m1x /= n; m1y /= n; m2x /= n; m2y /= n;
	double a1 = 0, a2 = 0;
	for (int64_t i = 0; i < n; ++i) {
		const double *p = pts + 4 * idx[i];
		const double dx1 = m1x - p[0], dy1 = m1y - p[1], dx2 = m2x - p[2], dy2 = m2y - p[3];
		a1 += std::sqrt(dx1 * dx1 + dy1 * dy1);
		a2 += std::sqrt(dx2 * dx2 + dy2 * dy2);
	}
	a1 /= n; a2 /= n;
	const double r1 = M_SQRT2 / a1, r2 = M_SQRT2 / a2;
	const int64_t rows = 2 * n;
	std::vector<double> A((size_t)rows * 8), b((size_t)rows);
	for (int64_t i = 0; i < n; ++i) {
		const double *p = pts + 4 * idx[i];
		const double x1 = (p[0] - m1x) * r1, y1 = (p[1] - m1y) * r1, x2 = (p[2] - m2x) * r2, y2 = (p[3] - m2y) * r2;
		const double w = weights_by_row ? weights_by_row[i] : 1.0;
		const double mwx1 = -w * x1, mwy1 = -w * y1, wx2 = w * x2, wy2 = w * y2;
		double *ra = &A[(size_t)(2 * i) * 8], *rb = &A[(size_t)(2 * i + 1) * 8];
		ra[0] = mwx1; ra[1] = mwy1; ra[2] = -w; ra[3] = 0; ra[4] = 0; ra[5] = 0; ra[6] = wx2 * x1; ra[7] = wx2 * y1;
		b[2 * i] = -wx2;
		rb[0] = 0; rb[1] = 0; rb[2] = 0; rb[3] = mwx1; rb[4] = mwy1; rb[5] = -w; rb[6] = wy2 * x1; rb[7] = wy2 * y1;
		b[2 * i + 1] = -wy2;
	}
	// Householder QR with column pivoting, then back substitution on the leading rank x rank block
	int perm[8];
	for (int j = 0; j < 8; ++j) perm[j] = j;
	double cn[8];
	for (int j = 0; j < 8; ++j) {
		cn[j] = 0;
		for (int64_t i = 0; i < rows; ++i) cn[j] += A[i * 8 + j] * A[i * 8 + j];
	}
	int rank = 0;
	double maxnorm = 0;
	for (int k = 0; k < 8 && k < rows; ++k) {
		int piv = k;
		for (int j = k; j < 8; ++j) { // recompute trailing norms exactly (small problem)
			cn[j] = 0;
			for (int64_t i = k; i < rows; ++i) cn[j] += A[i * 8 + j] * A[i * 8 + j];
			if (cn[j] > cn[piv]) piv = j;
		}
		if (k == 0) maxnorm = cn[piv];
		if (!(cn[piv] > maxnorm * 1e-28)) break;
		if (piv != k) {
			for (int64_t i = 0; i < rows; ++i) std::swap(A[i * 8 + k], A[i * 8 + piv]);
			std::swap(perm[k], perm[piv]);
			std::swap(cn[k], cn[piv]);
		}
		double alpha = std::sqrt(cn[k]);
		if (A[(size_t)k * 8 + k] > 0) alpha = -alpha;
		std::vector<double> v((size_t)(rows - k));
		for (int64_t i = k; i < rows; ++i) v[i - k] = A[i * 8 + k];
		v[0] -= alpha;
		double vnorm2 = 0;
		for (double t : v) vnorm2 += t * t;
		if (vnorm2 > 0) {
			for (int j = k; j < 8; ++j) {
				double s = 0;
				for (int64_t i = k; i < rows; ++i) s += v[i - k] * A[i * 8 + j];
				s = 2.0 * s / vnorm2;
				for (int64_t i = k; i < rows; ++i) A[i * 8 + j] -= s * v[i - k];
			}
			double s = 0;
			for (int64_t i = k; i < rows; ++i) s += v[i - k] * b[i];
			s = 2.0 * s / vnorm2;
			for (int64_t i = k; i < rows; ++i) b[i] -= s * v[i - k];
		}
		++rank;
	}
	double y[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	for (int k = rank - 1; k >= 0; --k) {
		double s = b[k];
		for (int j = k + 1; j < rank; ++j) s -= A[(size_t)k * 8 + j] * y[j];
		y[k] = s / A[(size_t)k * 8 + k];
	}
	double h[9];
	for (int j = 0; j < 8; ++j) h[perm[j]] = y[j];
	h[8] = 1.0;
	for (int j = 0; j < 8; ++j)
		if (!std::isfinite(h[j])) return 0;
	// H = T2^-1 * Hn * T1
	const double T1[9] = {r1, 0, -r1 * m1x, 0, r1, -r1 * m1y, 0, 0, 1};
	const double T2i[9] = {1.0 / r2, 0, m2x, 0, 1.0 / r2, m2y, 0, 0, 1};
	double tmp[9];
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c) tmp[3 * r + c] = T2i[3 * r] * h[c] + T2i[3 * r + 1] * h[3 + c] + T2i[3 * r + 2] * h[6 + c];
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c) H[3 * r + c] = tmp[3 * r] * T1[c] + tmp[3 * r + 1] * T1[3 + c] + tmp[3 * r + 2] * T1[6 + c];
	return 1;
}

// gcr/GCoptimization.cpp:608-751 for dense costs + one cost per label (restated; the true reference is
// oracle/_ref/libgco_ref.so and tests cross-check the two)
double pxo_greedy_ufl(const double *D, int64_t N, int32_t L1, double label_cost, const int32_t *init_labels,
                      int32_t *labels_out) {
	// estart = compute_energy() of the initial labelling: data + label costs of used labels (:614, :960-984)
	double estart;
	{
		double de = 0.0;
		std::vector<char> used((size_t)L1, 0);
		for (int64_t i = 0; i < N; ++i) {
			const int32_t l = init_labels ? init_labels[i] : 0;
			de += D[i * L1 + l];
			used[(size_t)l] = 1;
		}
		double le = 0.0;
		// m_labelcostsAll is a linked list built by prepending: iteration order is label L1-1 .. 0; all costs equal
		for (int32_t l = L1 - 1; l >= 0; --l)
			if (used[(size_t)l]) le += label_cost;
		estart = de + 0.0 + le;
	}
	std::vector<double> e((size_t)L1), cur((size_t)N);
	std::vector<int32_t> order((size_t)L1), lab((size_t)N);
	std::vector<char> active((size_t)L1, 0);
	int32_t alpha = 0;
	for (int32_t l = 0; l < L1; ++l) { // :634-650
		e[l] = 0;
		e[l] += label_cost;
		e[l] += (double)(N - N) * 10000000.0;
		for (int64_t i = 0; i < N; ++i) {
			e[l] += D[i * L1 + l];
			if (e[l] > e[alpha]) break;
		}
		if (e[l] < e[alpha]) alpha = l;
	}
	for (int64_t i = 0; i < N; ++i) {
		lab[i] = alpha;
		cur[i] = D[i * L1 + alpha];
	}
	active[alpha] = 1;
	for (int32_t l = 0; l < L1; ++l) order[l] = l;
	order[alpha] = 0;
	order[0] = alpha;
	for (int32_t alpha_count = 1; alpha_count <= L1; ++alpha_count) { // :667-722
		const int32_t alpha_prev = alpha;
		for (int32_t li = alpha_count; li < L1; ++li) {
			const int32_t l = order[li];
			e[l] = e[alpha_prev];
			if (!active[l]) e[l] += label_cost;
		}
		if (L1 - alpha_count > 0)
			for (int64_t i = 0; i < N; ++i)
				for (int32_t li = alpha_count; li < L1; ++li) {
					const int32_t l = order[li];
					const double delta = D[i * L1 + l] - cur[i];
					if (delta < 0) e[l] += delta;
				}
		int32_t alpha_index = alpha_count - 1;
		for (int32_t li = alpha_count; li < L1; ++li) {
			const int32_t l = order[li];
			if (e[l] < e[alpha]) {
				alpha = l;
				alpha_index = li;
			}
		}
		if (alpha == alpha_prev) break;
		std::swap(order[alpha_count], order[alpha_index]);
		for (int64_t i = 0; i < N; ++i) {
			const double dc_l = D[i * L1 + alpha];
			if (dc_l - cur[i] < 0) {
				lab[i] = alpha;
				cur[i] = dc_l;
			}
		}
		active[alpha] = 1;
	}
	const double efinal = e[alpha];
	if (efinal < estart) {
		for (int64_t i = 0; i < N; ++i) labels_out[i] = lab[i];
		return efinal;
	}
	for (int64_t i = 0; i < N; ++i) labels_out[i] = init_labels ? init_labels[i] : 0;
	return estart;
}

// ---- vanishing points and 2D lines -------------------------------------------------------------------

// VanishingPointTwoLineSolver::estimateModel, minimal branch (px/include/solver_vanishing_point_two_lines.h:146-186):
// intersection of the two segments' lines, normalised to unit length.
int pxo_vp2_solve(const double *pts, const int64_t *sample, double *v) {
	const double *a = pts + 4 * sample[0], *b = pts + 4 * sample[1];
	const double xs0 = a[0], ys0 = a[1], xe0 = a[2], ye0 = a[3], xs1 = b[0], ys1 = b[1], xe1 = b[2], ye1 = b[3];
	double l0[3], l1[3];
	// vec_cross(a1,b1,c1, a2,b2,c2): a3 = b1*c2 - c1*b2; b3 = -(a1*c2 - c1*a2); c3 = a1*b2 - b1*a2   (:100-114)
	l0[0] = ys0 * 1 - 1 * ye0; l0[1] = -(xs0 * 1 - 1 * xe0); l0[2] = xs0 * ye0 - ys0 * xe0;
	l1[0] = ys1 * 1 - 1 * ye1; l1[1] = -(xs1 * 1 - 1 * xe1); l1[2] = xs1 * ye1 - ys1 * xe1;
	v[0] = l0[1] * l1[2] - l0[2] * l1[1];
	v[1] = -(l0[0] * l1[2] - l0[2] * l1[0]);
	v[2] = l0[0] * l1[1] - l0[1] * l1[0];
	const double len = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); // vec_norm (:116-125)
	v[0] /= len; v[1] /= len; v[2] /= len;
	return 1;
}

// LinearModelSolver<2>::estimate2DLine (gcr/estimators/solver_linear_model.h:143-171), INCLUDING its typo
// `nx = y1 - x2` (a correct normal would be y1 - y2): the reference's minimal line hypotheses are what they are.
int pxo_line2_solve(const double *pts, const int64_t *sample, double *l) {
	const double *a = pts + 2 * sample[0], *b = pts + 2 * sample[1];
	const double x1 = a[0], y1 = a[1], x2 = b[0];
	double nx = y1 - x2, ny = x2 - x1;
	const double magnitude = std::sqrt(nx * nx + ny * ny);
	nx /= magnitude;
	ny /= magnitude;
	l[0] = nx; l[1] = ny; l[2] = -nx * x1 - ny * y1;
	return 1;
}

// Cyclic Jacobi on a symmetric 3x3 (stands in for Eigen::SelfAdjointEigenSolver<Matrix3d>, which is not on disk):
// eigenvalues in w, eigenvectors in the columns of V.
static void jacobi_eig3(double A[3][3], double V[3][3], double w[3]) {
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) V[i][j] = i == j;
	for (int sweep = 0; sweep < 60; ++sweep) {
		const double off = std::fabs(A[0][1]) + std::fabs(A[0][2]) + std::fabs(A[1][2]);
		if (off == 0.0) break;
		for (int p = 0; p < 2; ++p)
			for (int q = p + 1; q < 3; ++q) {
				if (A[p][q] == 0.0) continue;
				const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
				const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
				const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
				for (int k = 0; k < 3; ++k) { // A <- A J
					const double akp = A[k][p], akq = A[k][q];
					A[k][p] = c * akp - sn * akq;
					A[k][q] = sn * akp + c * akq;
				}
				for (int k = 0; k < 3; ++k) { // A <- J^T A
					const double apk = A[p][k], aqk = A[q][k];
					A[p][k] = c * apk - sn * aqk;
					A[q][k] = sn * apk + c * aqk;
				}
				for (int k = 0; k < 3; ++k) {
					const double vkp = V[k][p], vkq = V[k][q];
					V[k][p] = c * vkp - sn * vkq;
					V[k][q] = sn * vkp + c * vkq;
				}
			}
	}
	for (int i = 0; i < 3; ++i) w[i] = A[i][i];
}

// VanishingPointTwoLineSolver::estimateModel, non-minimal branch (solver_vanishing_point_two_lines.h:187-233): rows
// [y0 - my, mx - x0, x0 my - y0 mx] * weight, eigenvector of A^T A with the smallest eigenvalue, normalised. The
// weights are indexed BY POINT when a sample is given (weights_[sample_[i]], :203). The sign of an eigenvector is
// implementation defined in Eigen; here the largest-magnitude component is made positive ("parity unpinned" beyond
// tolerance and sign -- the residual is sign invariant).
int pxo_fit_vp_nonminimal(const double *pts, const int64_t *idx, int64_t n, const double *weights_by_point, double *v) {
	if (n < 2) return 0;
	double M[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
	for (int64_t i = 0; i < n; ++i) {
		const double *p = pts + 4 * idx[i];
		const double w = weights_by_point ? weights_by_point[idx[i]] : 1.0;
		const double x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
		const double mx = (x0 + x1) / 2.0, my = (y0 + y1) / 2.0, mz = 1.0;
		const double r[3] = {(y0 * mz - my) * w, (mx - x0 * mz) * w, (x0 * my - y0 * mx) * w};
		for (int a = 0; a < 3; ++a)
			for (int b = 0; b < 3; ++b) M[a][b] += r[a] * r[b];
	}
	double V[3][3], w3[3];
	jacobi_eig3(M, V, w3);
	int k = 0;
	for (int i = 1; i < 3; ++i)
		if (w3[i] < w3[k]) k = i;
	double e[3] = {V[0][k], V[1][k], V[2][k]};
	const double len = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
	int big = 0;
	for (int i = 1; i < 3; ++i)
		if (std::fabs(e[i]) > std::fabs(e[big])) big = i;
	const double sgn = e[big] < 0 ? -1.0 : 1.0;
	for (int i = 0; i < 3; ++i) v[i] = sgn * e[i] / len;
	return 1;
}

// LinearModelEstimator<..., 2>::estimateModelNonminimal (gcr/estimators/linear_model_estimator.h:152-186) =
// normalizePoints (:189-250: subtract the mass point, scale so that the mean distance is sqrt 2) +
// LinearModelSolver<2>::estimateModel non-minimal branch (solver_linear_model.h:198-239: C^T C, 2x2,
// FullPivHouseholderQR, last column of Q, normalised) + w = -mass . n. Weights are not used by this solver.
// Eigen's FullPivHouseholderQR of a 2x2 restated: pivot = entry of largest magnitude (first maximum in column-major
// scan order), one Householder reflection, Q = P_rows H.
int pxo_fit_line_nonminimal(const double *pts, const int64_t *idx, int64_t n, double *l) {
	if (n < 2) return 0;
	double mx = 0, my = 0;
	for (int64_t i = 0; i < n; ++i) { mx += pts[2 * idx[i]]; my += pts[2 * idx[i] + 1]; }
	mx /= (double)n; my /= (double)n;
	double avg = 0;
	for (int64_t i = 0; i < n; ++i) {
		const double dx = pts[2 * idx[i]] - mx, dy = pts[2 * idx[i] + 1] - my;
		avg += std::sqrt(dx * dx + dy * dy);
	}
	avg /= (double)n;
	const double ratio = std::sqrt(2.0) / avg;
	double a = 0, b = 0, c = 0; // C^T C = [[a, b], [b, c]]
	for (int64_t i = 0; i < n; ++i) {
		const double dx = (pts[2 * idx[i]] - mx) * ratio, dy = (pts[2 * idx[i] + 1] - my) * ratio;
		a += dx * dx; b += dx * dy; c += dy * dy;
	}
	double Mx[2][2] = {{a, b}, {b, c}};
	int pr = 0, pc = 0;
	double best = std::fabs(Mx[0][0]);
	for (int col = 0; col < 2; ++col)
		for (int row = 0; row < 2; ++row)
			if (std::fabs(Mx[row][col]) > best) { best = std::fabs(Mx[row][col]); pr = row; pc = col; }
	if (best == 0.0 || !(best <= 1e300)) return 0;
	// after the row swap (0 <-> pr) and column swap (0 <-> pc) the first column is x = (x0, x1)
	const double x0 = Mx[pr][pc], x1 = Mx[1 - pr][pc];
	double q[2]; // last column of H = I - tau v v^T, v = (1, ess)
	if (x1 == 0.0) {
		q[0] = 0.0; q[1] = 1.0;
	} else {
		double beta = std::sqrt(x0 * x0 + x1 * x1);
		if (x0 >= 0) beta = -beta;
		const double ess = x1 / (x0 - beta), tau = (beta - x0) / beta;
		q[0] = -tau * ess;
		q[1] = 1.0 - tau * ess * ess;
	}
	if (pr == 1) std::swap(q[0], q[1]); // Q = P_rows H: undo the row transposition
	const double len = std::sqrt(q[0] * q[0] + q[1] * q[1]);
	l[0] = q[0] / len; l[1] = q[1] / len;
	l[2] = -mx * l[0] - my * l[1];
	return 1;
}

} // extern "C"
