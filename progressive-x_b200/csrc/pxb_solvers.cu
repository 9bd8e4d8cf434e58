// pxb_solvers.cu -- batched minimal solvers, one thread per minimal sample (rows a6/a7/a8 of the scope table).
//
// Thousands of independent tiny dense problems: no shared state, straight-line float64 without FMA contraction so
// that the four-point solver reproduces the reference bit for bit (same pivot decisions, same rounding). The
// seven-point and P3P solvers call cbrt/acos/cos, whose CUDA implementations differ from glibc's in the last ulp;
// their outputs agree with the reference to ~1e-12 relative (the contract is 1e-5).
#include <cfloat>

#include "pxb_internal.h"
#include "pxb_residuals.cuh"

namespace pxb {

// ------------------------------------------------------------------------------------------------
// a6: HomographyFourPointSolver::estimateMinimalModel (gcr/estimators/solver_homography_four_point.h:109-190)
//     + gaussElimination<8> (gcr/math_utils.h:45-87)
//     + RobustHomographyEstimator::isValidSample / isValidModel (gcr/estimators/homography_estimator.h:326-381)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void h_cross(double r[3], const double *v1, const double *v2) {
	// homography_estimator.h:312-322 with st_ = 1
	r[0] = sub(v1[1], v2[1]);
	r[1] = sub(v2[0], v1[0]);
	r[2] = sub(mul(v1[0], v2[1]), mul(v1[1], v2[0]));
}
__device__ __forceinline__ double h_side(const double p[3], const double *c) {
	return add(add(mul(p[0], c[0]), mul(p[1], c[1])), p[2]);
}

__device__ double det3_partial_piv_lu(const double *M) {
	// Eigen's determinant() of a dynamic-size MatrixXd goes through PartialPivLU; restated (see oracle).
	double a[3][3] = {{M[0], M[1], M[2]}, {M[3], M[4], M[5]}, {M[6], M[7], M[8]}};
	double sign = 1.0;
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		int piv = k;
		double best = fabs(a[k][k]);
#pragma unroll
		for (int i = k + 1; i < 3; ++i)
			if (fabs(a[i][k]) > best) {
				best = fabs(a[i][k]);
				piv = i;
			}
		if (best == 0.0) continue;
		if (piv != k) {
#pragma unroll
			for (int i = k + 1; i < 3; ++i)
				if (piv == i) {
#pragma unroll
					for (int j = 0; j < 3; ++j) {
						const double t = a[k][j];
						a[k][j] = a[i][j];
						a[i][j] = t;
					}
				}
			sign = -sign;
		}
#pragma unroll
		for (int i = k + 1; i < 3; ++i) a[i][k] = divd(a[i][k], a[k][k]);
#pragma unroll
		for (int i = k + 1; i < 3; ++i)
#pragma unroll
			for (int j = k + 1; j < 3; ++j) a[i][j] = sub(a[i][j], mul(a[i][k], a[k][j]));
	}
	return mul(sign, mul(mul(a[0][0], a[1][1]), a[2][2]));
}

__global__ void __launch_bounds__(128)
    k_solve_h4(const double *__restrict__ aos, const int64_t *__restrict__ samples, int64_t K,
               double *__restrict__ models, int32_t *__restrict__ n_models, uint8_t *__restrict__ sample_valid,
               uint8_t *__restrict__ model_valid) {
	const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= K) return;
	double pt[4][4];
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const double *p = aos + 4 * samples[k * 4 + i];
		const double2 a = *reinterpret_cast<const double2 *>(p);
		const double2 b = *reinterpret_cast<const double2 *>(p + 2);
		pt[i][0] = a.x;
		pt[i][1] = a.y;
		pt[i][2] = b.x;
		pt[i][3] = b.y;
	}
	if (sample_valid) { // homography_estimator.h:346-381
		double p[3], q[3];
		bool ok = true;
		h_cross(p, pt[0], pt[1]);
		h_cross(q, pt[0] + 2, pt[1] + 2);
		if (mul(h_side(p, pt[2]), h_side(q, pt[2] + 2)) < 0) ok = false;
		if (mul(h_side(p, pt[3]), h_side(q, pt[3] + 2)) < 0) ok = false;
		h_cross(p, pt[2], pt[3]);
		h_cross(q, pt[2] + 2, pt[3] + 2);
		if (mul(h_side(p, pt[0]), h_side(q, pt[0] + 2)) < 0) ok = false;
		if (mul(h_side(p, pt[1]), h_side(q, pt[1] + 2)) < 0) ok = false;
		sample_valid[k] = ok ? 1 : 0;
	}
	// 8x9 DLT rows, weight = 1.0 (:147-167)
	double m[8][9];
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const double x1 = pt[i][0], y1 = pt[i][1], x2 = pt[i][2], y2 = pt[i][3];
		const double mwx1 = mul(-1.0, x1), mwy1 = mul(-1.0, y1), wx2 = mul(1.0, x2), wy2 = mul(1.0, y2);
		double *r0 = m[2 * i], *r1 = m[2 * i + 1];
		r0[0] = mwx1; r0[1] = mwy1; r0[2] = -1.0; r0[3] = 0; r0[4] = 0; r0[5] = 0;
		r0[6] = mul(wx2, x1); r0[7] = mul(wx2, y1); r0[8] = -wx2;
		r1[0] = 0; r1[1] = 0; r1[2] = 0; r1[3] = mwx1; r1[4] = mwy1; r1[5] = -1.0;
		r1[6] = mul(wy2, x1); r1[7] = mul(wy2, y1); r1[8] = -wy2;
	}
	// math_utils.h:54-62 "pivotisation": a pre-pass of conditional row swaps only
#pragma unroll
	for (int i = 0; i < 8; ++i)
#pragma unroll
		for (int kk = i + 1; kk < 8; ++kk) {
			const bool sw = fabs(m[i][i]) < fabs(m[kk][i]);
#pragma unroll
			for (int j = 0; j < 9; ++j) {
				const double a = m[i][j], b = m[kk][j];
				m[i][j] = sw ? b : a;
				m[kk][j] = sw ? a : b;
			}
		}
	// :65-72 elimination without further pivoting. Columns j <= i of row kk are never read again, so only the
	// columns that feed later steps are updated (same values as the reference for every entry that is used).
#pragma unroll
	for (int i = 0; i < 7; ++i)
#pragma unroll
		for (int kk = i + 1; kk < 8; ++kk) {
			const double t = divd(m[kk][i], m[i][i]);
#pragma unroll
			for (int j = i + 1; j < 9; ++j) m[kk][j] = sub(m[kk][j], mul(t, m[i][j]));
		}
	// :75-86 back-substitution
	double h[8];
#pragma unroll
	for (int i = 7; i >= 0; --i) {
		double r = m[i][8];
#pragma unroll
		for (int j = i + 1; j < 8; ++j) r = sub(r, mul(m[i][j], h[j]));
		h[i] = divd(r, m[i][i]);
	}
	bool has_nan = false;
#pragma unroll
	for (int i = 0; i < 8; ++i) has_nan |= (h[i] != h[i]);
	double *out = models + k * 9;
#pragma unroll
	for (int i = 0; i < 8; ++i) out[i] = h[i];
	out[8] = 1.0;
	n_models[k] = has_nan ? 0 : 1; // solver_homography_four_point.h:181
	if (model_valid) {
		const double det = det3_partial_piv_lu(out);
		model_valid[k] = (!has_nan && !(fabs(det) < 1e-2)) ? 1 : 0; // homography_estimator.h:338-341
	}
}

// ------------------------------------------------------------------------------------------------
// a7: FundamentalMatrixSevenPointSolver::estimateModel (gcr/estimators/solver_fundamental_matrix_seven_point.h:91-291)
//     + oriented epipolar filter (gcr/estimators/fundamental_estimator.h:161-184,737-800)
// ------------------------------------------------------------------------------------------------
__device__ int cubic_real_roots(const double c[4], double roots[3]) {
	// real roots of c0 + c1 x + c2 x^2 + c3 x^3, ascending; closed form + Newton polish (the reference uses
	// Eigen::PolynomialSolver<double,3>::realRoots, companion-matrix eigenvalues with |imag| < 1e-12)
	const double a2 = c[2] / c[3], a1 = c[1] / c[3], a0 = c[0] / c[3];
	const double Q = (3.0 * a1 - a2 * a2) / 9.0;
	const double R = (9.0 * a2 * a1 - 27.0 * a0 - 2.0 * a2 * a2 * a2) / 54.0;
	const double D = Q * Q * Q + R * R;
	int n = 0;
	if (D > 0) {
		const double sD = sqrt(D);
		roots[n++] = cbrt(R + sD) + cbrt(R - sD) - a2 / 3.0;
	} else {
		const double sq = sqrt(-Q);
		double ct = (sq > 0) ? R / (sq * sq * sq) : 0.0;
		ct = fmin(1.0, fmax(-1.0, ct));
		const double theta = acos(ct);
		const double kPi = 3.14159265358979323846;
		for (int k = 0; k < 3; ++k) roots[n++] = 2.0 * sq * cos((theta + 2.0 * kPi * k) / 3.0) - a2 / 3.0;
	}
	for (int i = 0; i < n; ++i) {
		double x = roots[i];
		for (int it = 0; it < 8; ++it) {
			const double f = ((x + a2) * x + a1) * x + a0;
			const double df = (3.0 * x + 2.0 * a2) * x + a1;
			if (df == 0.0) break;
			const double step = f / df;
			x -= step;
			if (fabs(step) <= 1e-16 * fabs(x)) break;
		}
		roots[i] = x;
	}
	// sort ascending (n <= 3)
	if (n == 3) {
		double t;
		if (roots[0] > roots[1]) { t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
		if (roots[1] > roots[2]) { t = roots[1]; roots[1] = roots[2]; roots[2] = t; }
		if (roots[0] > roots[1]) { t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
	}
	return n;
}

__device__ __forceinline__ double f_signum(const double *F, const double *e, const double *p) {
	const double s1 = F[0] * p[2] + F[3] * p[3] + F[6], s2 = e[1] - e[2] * p[1];
	return s1 * s2;
}

__global__ void __launch_bounds__(64)
    k_solve_f7(const double *__restrict__ aos, const int64_t *__restrict__ samples, int64_t K,
               double *__restrict__ models, int32_t *__restrict__ n_models, int apply_orientation) {
	const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= K) return;
	double P[7][4];
	double A[7][9];
	for (int i = 0; i < 7; ++i) {
		const double *p = aos + 4 * samples[k * 7 + i];
		const double x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
		P[i][0] = x0; P[i][1] = y0; P[i][2] = x1; P[i][3] = y1;
		A[i][0] = x1 * x0; A[i][1] = x1 * y0; A[i][2] = x1;
		A[i][3] = y1 * x0; A[i][4] = y1 * y0; A[i][5] = y1;
		A[i][6] = x0; A[i][7] = y0; A[i][8] = 1;
	}
	n_models[k] = 0;
	// Eigen::FullPivLU<MatrixXd>(7x9): complete pivoting, column-major search, first strictly greater wins
	int colidx[9];
	for (int j = 0; j < 9; ++j) colidx[j] = j;
	double maxpivot = 0.0;
	int nonzero_pivots = 7;
	for (int kk = 0; kk < 7; ++kk) {
		int pr = kk, pc = kk;
		double biggest = -1.0;
		for (int j = kk; j < 9; ++j)
			for (int i = kk; i < 7; ++i)
				if (fabs(A[i][j]) > biggest) {
					biggest = fabs(A[i][j]);
					pr = i;
					pc = j;
				}
		if (biggest == 0.0) {
			nonzero_pivots = kk;
			break;
		}
		if (biggest > maxpivot) maxpivot = biggest;
		if (pr != kk)
			for (int j = 0; j < 9; ++j) {
				const double t = A[kk][j];
				A[kk][j] = A[pr][j];
				A[pr][j] = t;
			}
		if (pc != kk) {
			for (int i = 0; i < 7; ++i) {
				const double t = A[i][kk];
				A[i][kk] = A[i][pc];
				A[i][pc] = t;
			}
			const int t = colidx[kk];
			colidx[kk] = colidx[pc];
			colidx[pc] = t;
		}
		for (int i = kk + 1; i < 7; ++i) A[i][kk] = A[i][kk] / A[kk][kk];
		for (int i = kk + 1; i < 7; ++i)
			for (int j = kk + 1; j < 9; ++j) A[i][j] = A[i][j] - A[i][kk] * A[kk][j];
	}
	const double thresh = fabs(maxpivot) * (DBL_EPSILON * 7.0);
	int rank = 0;
	for (int i = 0; i < nonzero_pivots; ++i)
		if (fabs(A[i][i]) > thresh) ++rank;
	if (9 - rank != 2) return; // :163-164 dimensionOfKernel() != 2
	double f1[9], f2[9];
	for (int kc = 0; kc < 2; ++kc) {
		double x[7];
		for (int i = 0; i < 7; ++i) x[i] = A[i][7 + kc];
		for (int i = 6; i >= 0; --i) {
			x[i] = x[i] / A[i][i];
			for (int j = 0; j < i; ++j) x[j] = x[j] - x[i] * A[j][i];
		}
		double *f = kc == 0 ? f1 : f2;
		for (int i = 0; i < 7; ++i) f[colidx[i]] = -x[i];
		f[colidx[7]] = kc == 0 ? 1.0 : 0.0;
		f[colidx[8]] = kc == 0 ? 0.0 : 1.0;
	}
	for (int i = 0; i < 9; ++i) f1[i] -= f2[i]; // :194
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
	if (fabs(c[0] + c[1] + c[2] + c[3]) < 1e-9 || fabs(c[0]) < DBL_EPSILON || fabs(c[1]) < DBL_EPSILON ||
	    fabs(c[2]) < DBL_EPSILON || fabs(c[3]) < DBL_EPSILON)
		return; // :246-251
	double roots[3];
	const int n = cubic_real_roots(c, roots);
	int kept = 0;
	for (int r = 0; r < n; ++r) { // :266-287
		double lambda = roots[r], mu = 1.0;
		const double s = f1[8] * roots[r] + f2[8];
		if (fabs(s) > DBL_EPSILON) {
			mu = 1.0 / s;
			lambda *= mu;
			double F[9];
			for (int i = 0; i < 9; ++i) F[i] = f1[i] * lambda + f2[i] * mu;
			F[8] = 1.0;
			if (apply_orientation) { // fundamental_estimator.h:737-800
				const double eps = 1.9984e-15;
				double e[3];
				e[0] = F[1] * F[8] - F[2] * F[7];
				e[1] = F[2] * F[6] - F[0] * F[8];
				e[2] = F[0] * F[7] - F[1] * F[6];
				bool big = false;
				for (int i = 0; i < 3; ++i) big |= (e[i] > eps) || (e[i] < -eps);
				if (!big) {
					e[0] = F[4] * F[8] - F[5] * F[7];
					e[1] = F[5] * F[6] - F[3] * F[8];
					e[2] = F[3] * F[7] - F[4] * F[6];
				}
				const double s2 = f_signum(F, e, P[0]);
				bool ok = true;
				for (int i = 1; i < 7; ++i)
					if (s2 * f_signum(F, e, P[i]) < 0) {
						ok = false;
						break;
					}
				if (!ok) continue;
			}
			double *out = models + (k * 3 + kept) * 9;
			for (int i = 0; i < 9; ++i) out[i] = F[i];
			++kept;
		}
	}
	n_models[k] = kept;
}

// ------------------------------------------------------------------------------------------------
// a8: P3PSolver::estimateModel (gcr/estimators/solver_p3p.h:108-385)
// ------------------------------------------------------------------------------------------------
struct V3 {
	double x, y, z;
};
__device__ __forceinline__ V3 v3sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 v3scale(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ double v3dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 v3cross(V3 a, V3 b) {
	return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
struct M3 {
	V3 c0, c1, c2; // columns
};
__device__ __forceinline__ double m3at(const M3 &m, int r, int c) {
	const V3 &v = c == 0 ? m.c0 : (c == 1 ? m.c1 : m.c2);
	return r == 0 ? v.x : (r == 1 ? v.y : v.z);
}
__device__ __forceinline__ V3 m3mulv(const M3 &m, V3 v) {
	return {m.c0.x * v.x + m.c1.x * v.y + m.c2.x * v.z, m.c0.y * v.x + m.c1.y * v.y + m.c2.y * v.z,
	        m.c0.z * v.x + m.c1.z * v.y + m.c2.z * v.z};
}
__device__ M3 m3inverse(const M3 &m) {
	// Eigen Matrix3d::inverse(): cofactors, det from the first column, scale by 1/det
	auto cof = [&](int i, int j) {
		const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
		return m3at(m, i1, j1) * m3at(m, i2, j2) - m3at(m, i1, j2) * m3at(m, i2, j1);
	};
	const double c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
	const double det = c00 * m3at(m, 0, 0) + c10 * m3at(m, 1, 0) + c20 * m3at(m, 2, 0);
	const double invdet = 1.0 / det;
	M3 r;
	r.c0 = {c00 * invdet, cof(0, 1) * invdet, cof(0, 2) * invdet};
	r.c1 = {c10 * invdet, cof(1, 1) * invdet, cof(1, 2) * invdet};
	r.c2 = {c20 * invdet, cof(2, 1) * invdet, cof(2, 2) * invdet};
	return r;
}

__device__ void p3p_refine_lambda(double &l1, double &l2, double &l3, double a12, double a13, double a23, double b12,
                                  double b13, double b23) {
	for (int iter = 0; iter < 5; ++iter) { // solver_p3p.h:145-175
		const double r1 = (l1 * l1 - 2.0 * l1 * l2 * b12 + l2 * l2 - a12);
		const double r2 = (l1 * l1 - 2.0 * l1 * l3 * b13 + l3 * l3 - a13);
		const double r3 = (l2 * l2 - 2.0 * l2 * l3 * b23 + l3 * l3 - a23);
		if (fabs(r1) + fabs(r2) + fabs(r3) < 1e-10) return;
		const double x11 = l1 - l2 * b12, x12 = l2 - l1 * b12, x21 = l1 - l3 * b13, x23 = l3 - l1 * b13,
		             x32 = l2 - l3 * b23, x33 = l3 - l2 * b23;
		const double detJ = 0.5 / (x11 * x23 * x32 + x12 * x21 * x33);
		l1 += (-x23 * x32 * r1 - x12 * x33 * r2 + x12 * x23 * r3) * detJ;
		l2 += (-x21 * x33 * r1 + x11 * x33 * r2 - x11 * x23 * r3) * detJ;
		l3 += (x21 * x32 * r1 - x11 * x32 * r2 - x12 * x21 * r3) * detJ;
	}
}

__global__ void __launch_bounds__(64)
    k_solve_p3p(const double *__restrict__ aos, const int64_t *__restrict__ samples, int64_t K,
                double *__restrict__ models, int32_t *__restrict__ n_models) {
	const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= K) return;
	V3 y[3], X[3];
	for (int i = 0; i < 3; ++i) { // :192-203
		const double *p = aos + 5 * samples[k * 3 + i];
		V3 v = {p[0], p[1], 1.0};
		const double z = v3dot(v, v);
		if (z > 0) {
			const double nrm = sqrt(z);
			v = {v.x / nrm, v.y / nrm, v.z / nrm};
		}
		y[i] = v;
		X[i] = {p[2], p[3], p[4]};
	}
	const V3 dX12 = v3sub(X[0], X[1]), dX13 = v3sub(X[0], X[2]), dX23 = v3sub(X[1], X[2]);
	const double a12 = v3dot(dX12, dX12), b12 = v3dot(y[0], y[1]);
	const double a13 = v3dot(dX13, dX13), b13 = v3dot(y[0], y[2]);
	const double a23 = v3dot(dX23, dX23), b23 = v3dot(y[1], y[2]);
	const double a23b12 = a23 * b12, a12b23 = a12 * b23, a23b13 = a23 * b13, a13b23 = a13 * b23;
	const M3 D1 = {{a23, -a23b12, 0.0}, {-a23b12, a23 - a12, a12b23}, {0.0, a12b23, -a12}};
	const M3 D2 = {{a23, 0.0, -a23b13}, {0.0, -a13, a13b23}, {-a23b13, a13b23, a23 - a13}};
	const M3 DX1 = {v3cross(D1.c1, D1.c2), v3cross(D1.c2, D1.c0), v3cross(D1.c0, D1.c1)};
	const M3 DX2 = {v3cross(D2.c1, D2.c2), v3cross(D2.c2, D2.c0), v3cross(D2.c0, D2.c1)};
	double c3 = v3dot(D2.c0, DX2.c0);
	double c2 = v3dot(D1.c0, DX2.c0) + v3dot(D1.c1, DX2.c1) + v3dot(D1.c2, DX2.c2);
	double c1 = v3dot(D2.c0, DX1.c0) + v3dot(D2.c1, DX1.c1) + v3dot(D2.c2, DX1.c2);
	double c0 = v3dot(D1.c0, DX1.c0);
	const double c3inv = 1.0 / c3;
	c2 *= c3inv;
	c1 *= c3inv;
	c0 *= c3inv;
	double a = c1 - c2 * c2 / 3.0;
	double b = (2.0 * c2 * c2 * c2 - 9.0 * c2 * c1) / 27.0 + c0;
	double c = b * b / 4.0 + a * a * a / 27.0;
	double gamma;
	if (c > 0) {
		c = sqrt(c);
		b *= -0.5;
		gamma = cbrt(b + c) + cbrt(b - c) - c2 / 3.0;
	} else {
		c = 3.0 * b / (2.0 * a) * sqrt(-3.0 / a);
		gamma = 2.0 * sqrt(-a / 3.0) * cos(acos(c) / 3.0) - c2 / 3.0;
	}
	const double f = gamma * gamma * gamma + c2 * gamma * gamma + c1 * gamma + c0;
	const double df = 3.0 * gamma * gamma + 2.0 * c2 * gamma + c1;
	gamma = gamma - f / df;

	// D0 = D1 + gamma * D2 (symmetric); computeEig3x3known0 (:108-142)
	const double M00 = D1.c0.x + gamma * D2.c0.x, M01 = D1.c1.x + gamma * D2.c1.x, M02 = D1.c2.x + gamma * D2.c2.x;
	const double M11 = D1.c1.y + gamma * D2.c1.y, M12 = D1.c2.y + gamma * D2.c2.y, M22 = D1.c2.z + gamma * D2.c2.z;
	double E[3][2], sig1, sig2;
	{
		const double p1 = -M00 - M11 - M22;
		const double p0 = -M01 * M01 - M02 * M02 - M12 * M12 + M00 * (M11 + M22) + M11 * M22;
		const double disc = sqrt(p1 * p1 / 4.0 - p0);
		const double tmp = -p1 / 2.0;
		sig1 = tmp + disc;
		sig2 = tmp - disc;
		if (fabs(sig1) < fabs(sig2)) {
			const double t = sig1;
			sig1 = sig2;
			sig2 = t;
		}
		double cc = sig1 * sig1 + M00 * M11 - sig1 * (M00 + M11) - M01 * M01;
		double a1 = (sig1 * M02 + M01 * M12 - M02 * M11) / cc;
		double a2 = (sig1 * M12 + M01 * M02 - M00 * M12) / cc;
		double n = 1.0 / sqrt(1 + a1 * a1 + a2 * a2);
		E[0][0] = a1 * n; E[1][0] = a2 * n; E[2][0] = n;
		cc = sig2 * sig2 + M00 * M11 - sig2 * (M00 + M11) - M01 * M01;
		a1 = (sig2 * M02 + M01 * M12 - M02 * M11) / cc;
		a2 = (sig2 * M12 + M01 * M02 - M00 * M12) / cc;
		n = 1.0 / sqrt(1 + a1 * a1 + a2 * a2);
		E[0][1] = a1 * n; E[1][1] = a2 * n; E[2][1] = n;
	}
	double s = sqrt(-sig2 / sig1);
	M3 XX = {dX12, dX13, v3cross(dX12, dX13)};
	XX = m3inverse(XX);
	const double TOL_DOUBLE_ROOT = 1e-12;
	int nsol = 0;
	double *outbase = models + k * 4 * 12;
	for (int s_flip = 0; s_flip < 2; ++s_flip, s = -s) {
		const double u1 = E[0][0] - s * E[0][1], u2 = E[1][0] - s * E[1][1], u3 = E[2][0] - s * E[2][1];
		const bool switch_12 = fabs(u1) < fabs(u2);
		double qa, qb, qc, w0, w1;
		if (switch_12) {
			w0 = -u1 / u2;
			w1 = -u3 / u2;
			qa = -a13 * w1 * w1 + 2 * a13b23 * w1 - a13 + a23;
			qb = 2 * a13b23 * w0 - 2 * a23b13 - 2 * a13 * w0 * w1;
			qc = -a13 * w0 * w0 + a23;
		} else {
			w0 = -u2 / u1;
			w1 = -u3 / u1;
			qa = (a13 - a12) * w1 * w1 + 2.0 * a12 * b13 * w1 - a12;
			qb = -2.0 * a13 * b12 * w1 + 2.0 * a12 * b13 * w0 - 2.0 * w0 * w1 * (a12 - a13);
			qc = (a13 - a12) * w0 * w0 - 2.0 * a13 * b12 * w0 + a13;
		}
		const double b2m4ac = qb * qb - 4.0 * qa * qc;
		if (b2m4ac < -TOL_DOUBLE_ROOT) continue;
		const double sq = sqrt(fmax(0.0, b2m4ac));
		double tau = (qb > 0) ? (2.0 * qc) / (-qb - sq) : (2.0 * qc) / (-qb + sq);
		for (int tau_flip = 0; tau_flip < 2; ++tau_flip, tau = qc / (qa * tau)) {
			if (tau > 0) {
				double l1, l2, l3;
				bool neg;
				if (switch_12) {
					l1 = sqrt(a13 / (tau * (tau - 2.0 * b13) + 1.0));
					l3 = tau * l1;
					l2 = w0 * l1 + w1 * l3;
					neg = l2 < 0;
				} else {
					l2 = sqrt(a23 / (tau * (tau - 2.0 * b23) + 1.0));
					l3 = tau * l2;
					l1 = w0 * l2 + w1 * l3;
					neg = l1 < 0;
				}
				if (neg) continue; // the reference's `continue` also skips the double-root break below
				p3p_refine_lambda(l1, l2, l3, a12, a13, a23, b12, b13, b23);
				const V3 v1 = v3sub(v3scale(l1, y[0]), v3scale(l2, y[1]));
				const V3 v2 = v3sub(v3scale(l1, y[0]), v3scale(l3, y[2]));
				const M3 YY = {v1, v2, v3cross(v1, v2)};
				const M3 R = {m3mulv(YY, XX.c0), m3mulv(YY, XX.c1), m3mulv(YY, XX.c2)};
				const V3 t = v3sub(v3scale(l1, y[0]), m3mulv(R, X[0]));
				if (nsol < 4) {
					double *P = outbase + 12 * nsol;
					P[0] = R.c0.x; P[1] = R.c1.x; P[2] = R.c2.x; P[3] = t.x;
					P[4] = R.c0.y; P[5] = R.c1.y; P[6] = R.c2.y; P[7] = t.y;
					P[8] = R.c0.z; P[9] = R.c1.z; P[10] = R.c2.z; P[11] = t.z;
					++nsol;
				}
			}
			if (b2m4ac < TOL_DOUBLE_ROOT) break;
		}
	}
	n_models[k] = nsol;
}

// ------------------------------------------------------------------------------------------------
// f-4: vanishing point from two segments, 2D line from two points (closed forms, thread per sample)
// ------------------------------------------------------------------------------------------------
// VanishingPointTwoLineSolver::estimateModel, minimal branch (px/include/solver_vanishing_point_two_lines.h:146-186):
// l_k = (start_k, 1) x (end_k, 1); v = l_0 x l_1; v /= |v|. vec_cross(a,b) = (b1 c2 - c1 b2, -(a1 c2 - c1 a2),
// a1 b2 - b1 a2) (:100-114). The model is pushed whatever it contains (parallel segments give NaN).
__global__ void __launch_bounds__(128)
    k_solve_vp2(const double *__restrict__ aos, const int64_t *__restrict__ samples, int64_t K, double *__restrict__ models,
                int32_t *__restrict__ n_models) {
	const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= K) return;
	const double *a = aos + 4 * samples[2 * k], *b = aos + 4 * samples[2 * k + 1];
	const double xs0 = a[0], ys0 = a[1], xe0 = a[2], ye0 = a[3], xs1 = b[0], ys1 = b[1], xe1 = b[2], ye1 = b[3];
	double l0[3], l1[3], v[3];
	l0[0] = sub(mul(ys0, 1.0), mul(1.0, ye0));
	l0[1] = -sub(mul(xs0, 1.0), mul(1.0, xe0));
	l0[2] = sub(mul(xs0, ye0), mul(ys0, xe0));
	l1[0] = sub(mul(ys1, 1.0), mul(1.0, ye1));
	l1[1] = -sub(mul(xs1, 1.0), mul(1.0, xe1));
	l1[2] = sub(mul(xs1, ye1), mul(ys1, xe1));
	v[0] = sub(mul(l0[1], l1[2]), mul(l0[2], l1[1]));
	v[1] = -sub(mul(l0[0], l1[2]), mul(l0[2], l1[0]));
	v[2] = sub(mul(l0[0], l1[1]), mul(l0[1], l1[0]));
	const double len = __dsqrt_rn(add(add(mul(v[0], v[0]), mul(v[1], v[1])), mul(v[2], v[2]))); // vec_norm (:116-125)
	models[3 * k] = divd(v[0], len);
	models[3 * k + 1] = divd(v[1], len);
	models[3 * k + 2] = divd(v[2], len);
	n_models[k] = 1;
}

// LinearModelSolver<2>::estimate2DLine (gcr/estimators/solver_linear_model.h:152-188) -- with the reference's
// `nx = y1 - x2` (a true normal would be y1 - y2): drop-in means the same hypotheses.
__global__ void __launch_bounds__(128)
    k_solve_line2(const double *__restrict__ aos, const int64_t *__restrict__ samples, int64_t K, double *__restrict__ models,
                  int32_t *__restrict__ n_models) {
	const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= K) return;
	const double *a = aos + 2 * samples[2 * k], *b = aos + 2 * samples[2 * k + 1];
	const double x1 = a[0], y1 = a[1], x2 = b[0];
	double nx = sub(y1, x2), ny = sub(x2, x1);
	const double magnitude = __dsqrt_rn(add(mul(nx, nx), mul(ny, ny)));
	nx = divd(nx, magnitude);
	ny = divd(ny, magnitude);
	models[3 * k] = nx;
	models[3 * k + 1] = ny;
	models[3 * k + 2] = sub(mul(-nx, x1), mul(ny, y1)); // -nx * x1 - ny * y1
	n_models[k] = 1;
}

// ------------------------------------------------------------------------------------------------
// DEGENSAC's plane-and-parallax solver: F from a fixed homography and two off-plane correspondences
// ------------------------------------------------------------------------------------------------
// FundamentalMatrixPlaneParallaxSolver::estimateModel (gcr/estimators/solver_fundamental_matrix_plane_and_parallax.h):
// line_i = (H x1_i) x x2_i, epipole = line_1 x line_2, F = [epipole]_x H; no model when |epipole_z| < epsilon.
__global__ void __launch_bounds__(128)
    k_solve_fpp(const double *__restrict__ aos, const int64_t *__restrict__ samples, int64_t K, const double *__restrict__ Hm,
                double *__restrict__ models, int32_t *__restrict__ n_models) {
	const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= K) return;
	double H[9];
	for (int i = 0; i < 9; ++i) H[i] = Hm[i];
	double line[2][3];
	for (int j = 0; j < 2; ++j) {
		const double *q = aos + 4 * samples[2 * k + j];
		const double s[3] = {q[0], q[1], 1.0}, d[3] = {q[2], q[3], 1.0};
		double pr[3];
		for (int r = 0; r < 3; ++r) pr[r] = add(add(mul(H[3 * r], s[0]), mul(H[3 * r + 1], s[1])), mul(H[3 * r + 2], s[2]));
		line[j][0] = sub(mul(pr[1], d[2]), mul(pr[2], d[1]));
		line[j][1] = sub(mul(pr[2], d[0]), mul(pr[0], d[2]));
		line[j][2] = sub(mul(pr[0], d[1]), mul(pr[1], d[0]));
	}
	double e[3];
	e[0] = sub(mul(line[0][1], line[1][2]), mul(line[0][2], line[1][1]));
	e[1] = sub(mul(line[0][2], line[1][0]), mul(line[0][0], line[1][2]));
	e[2] = sub(mul(line[0][0], line[1][1]), mul(line[0][1], line[1][0]));
	if (!(fabs(e[2]) >= 2.220446049250313e-16)) { // also rejects NaN
		n_models[k] = 0;
		return;
	}
	const double ex[9] = {0.0, -e[2], e[1], e[2], 0.0, -e[0], -e[1], e[0], 0.0};
	double *F = models + 9 * k;
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c)
			F[3 * r + c] = add(add(mul(ex[3 * r], H[c]), mul(ex[3 * r + 1], H[3 + c])), mul(ex[3 * r + 2], H[6 + c]));
	n_models[k] = 1;
}

int launch_solve_plane_parallax(pxb_ctx *ctx, const int64_t *samples, int64_t K, const double *H_dev, double *models_out,
                                int32_t *n_models, uint8_t *sample_valid, uint8_t *model_valid) {
	if (K <= 0) return PXB_OK;
	k_solve_fpp<<<(unsigned)((K + 127) / 128), 128, 0, ctx->stream>>>(ctx->pts.aos, samples, K, H_dev, models_out, n_models);
	ctx->launches++;
	if (sample_valid) PXB_CUDA(cudaMemsetAsync(sample_valid, 1, (size_t)K, ctx->stream));
	if (model_valid) PXB_CUDA(cudaMemsetAsync(model_valid, 1, (size_t)K, ctx->stream));
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

int launch_solve_minimal(pxb_ctx *ctx, const int64_t *samples, int64_t K, double *models_out, int32_t *n_models,
                         uint8_t *sample_valid, uint8_t *model_valid) {
	if (K <= 0) return PXB_OK;
	const Points &p = ctx->pts;
	switch (p.type) {
	case PXB_MODEL_HOMOGRAPHY:
		k_solve_h4<<<(unsigned)((K + 127) / 128), 128, 0, ctx->stream>>>(p.aos, samples, K, models_out, n_models,
		                                                                 sample_valid, model_valid);
		break;
	case PXB_MODEL_FUNDAMENTAL:
		k_solve_f7<<<(unsigned)((K + 63) / 64), 64, 0, ctx->stream>>>(p.aos, samples, K, models_out, n_models, 1);
		if (sample_valid) PXB_CUDA(cudaMemsetAsync(sample_valid, 1, (size_t)K, ctx->stream));
		if (model_valid) PXB_CUDA(cudaMemsetAsync(model_valid, 1, (size_t)K, ctx->stream));
		break;
	case PXB_MODEL_PNP:
		k_solve_p3p<<<(unsigned)((K + 63) / 64), 64, 0, ctx->stream>>>(p.aos, samples, K, models_out, n_models);
		if (sample_valid) PXB_CUDA(cudaMemsetAsync(sample_valid, 1, (size_t)K, ctx->stream));
		if (model_valid) PXB_CUDA(cudaMemsetAsync(model_valid, 1, (size_t)K, ctx->stream));
		break;
	default: // Estimator::isValidSample / isValidModel are the base-class "true" for both families
		if (p.type == PXB_MODEL_VANISHING_POINT)
			k_solve_vp2<<<(unsigned)((K + 127) / 128), 128, 0, ctx->stream>>>(p.aos, samples, K, models_out, n_models);
		else
			k_solve_line2<<<(unsigned)((K + 127) / 128), 128, 0, ctx->stream>>>(p.aos, samples, K, models_out, n_models);
		if (sample_valid) PXB_CUDA(cudaMemsetAsync(sample_valid, 1, (size_t)K, ctx->stream));
		if (model_valid) PXB_CUDA(cudaMemsetAsync(model_valid, 1, (size_t)K, ctx->stream));
		break;
	}
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

} // namespace pxb
