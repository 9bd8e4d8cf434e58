// pxb_screen.cuh -- rigorous float32 screening of "is this point certainly NOT an inlier of this model?"
//
// The compound score (a4), inlier counts and masks only need the exact float64 residual of the points that can be
// inliers; for every other point the only fact needed is r2 >= T2. The FP64 pipe of a B200 SM retires one warp
// instruction every 2-3 cycles (16 lanes per sub-partition, three 64-bit register reads per DFMA -- measured with
// tools/microbench/fp64_issue.cu), the FP32 pipe one per cycle, so proving "outlier" in float32 is ~4x cheaper than
// evaluating the reference arithmetic. A point that cannot be PROVEN an outlier is re-evaluated with the reference's
// exact float64 arithmetic (pxb_residuals.cuh), so counts, masks and sums stay bit-identical to the CPU reference.
//
// Division-free form of the three inlier tests (exact arithmetic, denominators cancel):
//   H    r2 = (x2 - t1/t3)^2 + (y2 - t2/t3)^2 >= T   <=>   (x2 t3 - t1)^2 + (y2 t3 - t2)^2 >= T t3^2
//   PnP  same with (px, py, pz) and (u, v)
//   F    r2 = r^2 / den >= T                          <=>   r^2 >= T den,  den = rxc^2 + ryc^2 + rx^2 + ry^2
// (t3 = 0 / den = 0 give inf or NaN in the reference, never an inlier; both sides of the rewritten test are then 0 or
// the left side is larger, and the screening below only ever claims "outlier".)
//
// Conditioning. Screening runs in NORMALISED coordinates: image points are mapped to [-1, 1] by the bounding box of
// the uploaded set (common scale s), 3-D points likewise; the model is conjugated accordingly in float64
// (H_n = N2 H N1^-1, F_n = N2^-T F N1^-1, P_n = P N3^-1) and only then rounded to float32. The tests are invariant:
// a_n = a / s, den_n = den s^2, T_n = T / s^2. All model entries are then O(1) for sane hypotheses and the crude
// forward error bound below is ~100x smaller than the threshold.
//
// Error bound. u = 2^-24. With M = max over model entries of the sum of |terms| that formed it, P1 = |x1|+|y1|+1
// (or |X|+|Y|+|Z|+1), P2 = |x2|+|y2|+1 (or |u|+|v|+1) in normalised coordinates, a standard forward analysis of the
// FFMA chains (inputs rounded once, <= 4 operations deep, FFMA keeps products unrounded) gives
//   |t_k,f - t_k| <= 5.1 u M P1,   |a_f - a|, |b_f - b|, |r_f - r| <= 7.1 u M P1 P2.
// The code uses E = 2^-19 M P1 P2 (32 u: a 4x margin; M^2 <= 4 after the rescaling below) for every one of them.
// Then, with the triangle inequality and (p + q)^2 <= (1 + eta) p^2 + (1 + 1/eta) q^2, eta = 1/16:
//   H/PnP  |(a,b)| >= |(a_f,b_f)| - sqrt2 E,  |t3| <= |t3_f| + E:
//          a_f^2 + b_f^2 >= (1+eta) T t3_f^2 + (1+1/eta) (sqrt2 + sqrtT)^2 E^2      ==>  a^2 + b^2 >= T t3^2
//   F      |r| >= |r_f| - E,  sqrt(den) <= sqrt(den_f) (1 + 4u) + 2 E:
//          r_f^2 >= (1+eta) T den_f + (1+1/eta) (2 sqrtT + 1)^2 E^2                 ==>  r^2 >= T den
// i.e. everything farther than ~1.03 thresholds from the model is dismissed in float32, everything closer goes to
// the exact path. Constants are rounded UP when converted to float32 and carry a further (1 + 2^-6) factor
// that covers the roundings of the test itself. Overflow/underflow are excluded by construction: normalised
// coordinates are <= 2^10 in magnitude and the model is rescaled to 1 <= M < 2 (the tests are homogeneous in the
// model), else the point / hypothesis is flagged "wild" (bound = +inf or NaN model) and takes the exact path. NaN anywhere makes the comparison false -> exact path.
#pragma once
#include <cuda_runtime.h>

#include "pxb_residuals.cuh"

namespace pxb {

// Normalisation of the uploaded point set (device memory; written by k_point_stats at upload time).
struct NormDev {
	double c[5]; // centre per coordinate (0 for coordinates that are left alone)
	double s;    // common scale of the normalised block (H/F: all four image coordinates; PnP: X Y Z)
	double inv_s; // RN(1 / s): used instead of divisions (a 2^-53 relative perturbation, far inside the 4x margin of E)
};

constexpr float kScreenU2 = 1.0f / 274877906944.0f; // (2^-19)^2 = 2^-38
constexpr double kScreenSlack = 1.0 + 1.0 / 64.0;

template <int TYPE> struct ScreenTraits;
template <> struct ScreenTraits<PXB_MODEL_HOMOGRAPHY> { static constexpr int kFloats = 12; };
template <> struct ScreenTraits<PXB_MODEL_FUNDAMENTAL> { static constexpr int kFloats = 12; };
template <> struct ScreenTraits<PXB_MODEL_PNP> { static constexpr int kFloats = 12; };
template <> struct ScreenTraits<PXB_MODEL_VANISHING_POINT> { static constexpr int kFloats = 8; };
template <> struct ScreenTraits<PXB_MODEL_LINE2D> { static constexpr int kFloats = 4; };

// Normalised float32 coordinates of one point + its error-scale constant q = ((P1 P2)^2)(1 + 2^-6); +inf marks a
// point that must always take the exact path (non-finite or out-of-range coordinates).
template <int TYPE>
__device__ __forceinline__ void screen_point(const double *p, const NormDev &nd, float *pf, float &q) {
	constexpr int DIM = ModelTraits<TYPE>::kDim;
	bool wild = false;
	double v[5];
	float a[5];
#pragma unroll
	for (int c = 0; c < DIM; ++c) {
		const bool raw = (TYPE == PXB_MODEL_PNP && c < 2); // K^-1-normalised image coordinates are used as they are
		v[c] = raw ? p[c] : (p[c] - nd.c[c]) * nd.inv_s;
		wild |= !(fabs(v[c]) <= 1024.0);
	}
	if (TYPE == PXB_MODEL_VANISHING_POINT) { // stored as (xs, ys, xs + xe, ys + ye): the test uses twice the midpoint
		v[2] = v[0] + v[2];
		v[3] = v[1] + v[3];
	}
#pragma unroll
	for (int c = 0; c < DIM; ++c) a[c] = (float)v[c];
	float P1, P2;
	if (TYPE == PXB_MODEL_PNP) {
		P2 = fabsf(a[0]) + fabsf(a[1]) + 1.0f;
		P1 = fabsf(a[2]) + fabsf(a[3]) + fabsf(a[4]) + 1.0f;
	} else if (TYPE == PXB_MODEL_VANISHING_POINT) {
		P2 = fabsf(a[0]) + fabsf(a[1]) + 1.0f;
		P1 = 2.0f * (fabsf(a[2]) + fabsf(a[3]) + 1.0f); // |l'| <= 2 M (|sx| + |sy| + 1), see screen_sure_outlier<VP>
	} else if (TYPE == PXB_MODEL_LINE2D) {
		P1 = fabsf(a[0]) + fabsf(a[1]) + 1.0f;
		P2 = 1.0f;
	} else {
		P1 = fabsf(a[0]) + fabsf(a[1]) + 1.0f;
		P2 = fabsf(a[2]) + fabsf(a[3]) + 1.0f;
	}
	const float pp = P1 * P2;
	q = wild ? __int_as_float(0x7f800000) : pp * pp * (float)kScreenSlack;
#pragma unroll
	for (int c = 0; c < DIM; ++c) pf[c] = wild ? 0.0f : a[c];
}

// Conjugates one model into normalised coordinates (float64), rescales it to M in [1, 2) and rounds it to float32.
template <int TYPE> __device__ __forceinline__ void screen_model(const double *m, const NormDev &nd, float *mf);

// All three tests are homogeneous in the model (both sides scale with its square), so the normalised model is
// rescaled by an exact power of two to M in [1, 2): the error allowance then uses the constant M^2 <= 4 and no
// hypothesis is out of range unless it is zero or non-finite (those get NaN entries: every comparison is false and
// every point of that hypothesis takes the exact path).
__device__ __forceinline__ int screen_model_finish(const double *out, const double *mag, int n, float *mf, bool &wild) {
	double M = 0.0;
	wild = false;
	for (int i = 0; i < n; ++i) {
		wild |= !(fabs(out[i]) <= 1e300) || !(mag[i] <= 1e300);
		M = fmax(M, mag[i]);
	}
	// The screening argument is scale invariant, the reference's float64 arithmetic is not: outside this window its squares
	// and fourth powers overflow or underflow (a vanishing point scaled by 1e200 makes every residual 0/inf = 0, an inlier
	// everywhere). Such hypotheses take the exact path for every point.
	wild |= !(M >= 0x1p-100) || !(M <= 0x1p100);
	int e = 0;
	frexp(wild ? 1.0 : M, &e); // M = f 2^e, f in [0.5, 1)
	for (int i = 0; i < n; ++i) mf[i] = wild ? __int_as_float(0x7fc00000) : (float)ldexp(out[i], 1 - e);
	return 1 - e; // the model was multiplied by 2^(1-e)
}

template <> __device__ __forceinline__ void screen_model<PXB_MODEL_HOMOGRAPHY>(const double *m, const NormDev &nd, float *mf) {
	const double s = nd.s, c0 = nd.c[0], c1 = nd.c[1], c2 = nd.c[2], c3 = nd.c[3];
	double G[9], A[9]; // G = H N1^-1 and the sums of |terms| behind every entry
	for (int r = 0; r < 3; ++r) {
		G[3 * r] = s * m[3 * r];
		G[3 * r + 1] = s * m[3 * r + 1];
		G[3 * r + 2] = m[3 * r] * c0 + m[3 * r + 1] * c1 + m[3 * r + 2];
		A[3 * r] = fabs(G[3 * r]);
		A[3 * r + 1] = fabs(G[3 * r + 1]);
		A[3 * r + 2] = fabs(m[3 * r] * c0) + fabs(m[3 * r + 1] * c1) + fabs(m[3 * r + 2]);
	}
	double out[9], mag[9];
	for (int j = 0; j < 3; ++j) {
		out[j] = (G[j] - c2 * G[6 + j]) * nd.inv_s;
		out[3 + j] = (G[3 + j] - c3 * G[6 + j]) * nd.inv_s;
		out[6 + j] = G[6 + j];
		mag[j] = (A[j] + fabs(c2) * A[6 + j]) * nd.inv_s;
		mag[3 + j] = (A[3 + j] + fabs(c3) * A[6 + j]) * nd.inv_s;
		mag[6 + j] = A[6 + j];
	}
	bool wild;
	screen_model_finish(out, mag, 9, mf, wild);
}

template <> __device__ __forceinline__ void screen_model<PXB_MODEL_FUNDAMENTAL>(const double *m, const NormDev &nd, float *mf) {
	const double s = nd.s, c0 = nd.c[0], c1 = nd.c[1], c2 = nd.c[2], c3 = nd.c[3];
	double G[9], A[9];
	for (int r = 0; r < 3; ++r) {
		G[3 * r] = s * m[3 * r];
		G[3 * r + 1] = s * m[3 * r + 1];
		G[3 * r + 2] = m[3 * r] * c0 + m[3 * r + 1] * c1 + m[3 * r + 2];
		A[3 * r] = fabs(G[3 * r]);
		A[3 * r + 1] = fabs(G[3 * r + 1]);
		A[3 * r + 2] = fabs(m[3 * r] * c0) + fabs(m[3 * r + 1] * c1) + fabs(m[3 * r + 2]);
	}
	double out[9], mag[9]; // F_n = N2^-T G, N2^-T = [[s,0,0],[0,s,0],[c2,c3,1]]
	for (int j = 0; j < 3; ++j) {
		out[j] = s * G[j];
		out[3 + j] = s * G[3 + j];
		out[6 + j] = c2 * G[j] + c3 * G[3 + j] + G[6 + j];
		mag[j] = s * A[j];
		mag[3 + j] = s * A[3 + j];
		mag[6 + j] = fabs(c2) * A[j] + fabs(c3) * A[3 + j] + A[6 + j];
	}
	bool wild;
	screen_model_finish(out, mag, 9, mf, wild);
}

template <> __device__ __forceinline__ void screen_model<PXB_MODEL_PNP>(const double *m, const NormDev &nd, float *mf) {
	const double s = nd.s;
	double out[12], mag[12]; // P_n = P N3^-1, N3^-1 = [[s I, c],[0, 1]]
	for (int r = 0; r < 3; ++r) {
		for (int j = 0; j < 3; ++j) {
			out[4 * r + j] = s * m[4 * r + j];
			mag[4 * r + j] = fabs(out[4 * r + j]);
		}
		out[4 * r + 3] = m[4 * r] * nd.c[2] + m[4 * r + 1] * nd.c[3] + m[4 * r + 2] * nd.c[4] + m[4 * r + 3];
		mag[4 * r + 3] = fabs(m[4 * r] * nd.c[2]) + fabs(m[4 * r + 1] * nd.c[3]) + fabs(m[4 * r + 2] * nd.c[4]) + fabs(m[4 * r + 3]);
	}
	bool wild;
	screen_model_finish(out, mag, 12, mf, wild);
}

// Vanishing point v (homogeneous): v_n = N v with N = [[1/s, 0, -cx/s], [0, 1/s, -cy/s], [0, 0, 1]] (both segment
// end points share the centre: NormDev.c[0] == c[2], c[1] == c[3] for this family). Layout: d0 d1 d2 2d0 2d1.
template <> __device__ __forceinline__ void screen_model<PXB_MODEL_VANISHING_POINT>(const double *m, const NormDev &nd, float *mf) {
	double out[3], mag[3];
	out[0] = (m[0] - nd.c[0] * m[2]) * nd.inv_s;
	out[1] = (m[1] - nd.c[1] * m[2]) * nd.inv_s;
	out[2] = m[2];
	mag[0] = (fabs(m[0]) + fabs(nd.c[0] * m[2])) * nd.inv_s;
	mag[1] = (fabs(m[1]) + fabs(nd.c[1] * m[2])) * nd.inv_s;
	mag[2] = fabs(m[2]);
	bool wild;
	screen_model_finish(out, mag, 3, mf, wild);
	mf[3] = 2.0f * mf[0];
	mf[4] = 2.0f * mf[1];
	mf[5] = mf[6] = mf[7] = 0.0f;
}

// 2D line l = (nx, ny, c): r = l . (x, y, 1) = l_n . (x_n, y_n, 1) with l_n = (s nx, s ny, nx cx + ny cy + c); r is NOT
// homogeneous against the fixed threshold, so the power-of-two rescale 2^k of l_n is carried into the threshold:
// slot 3 holds 4^k (the squared residual of the rescaled line is compared with T 4^k).
template <> __device__ __forceinline__ void screen_model<PXB_MODEL_LINE2D>(const double *m, const NormDev &nd, float *mf) {
	double out[3], mag[3];
	out[0] = nd.s * m[0];
	out[1] = nd.s * m[1];
	out[2] = m[0] * nd.c[0] + m[1] * nd.c[1] + m[2];
	mag[0] = fabs(out[0]);
	mag[1] = fabs(out[1]);
	mag[2] = fabs(m[0] * nd.c[0]) + fabs(m[1] * nd.c[1]) + fabs(m[2]);
	bool wild;
	const int k = screen_model_finish(out, mag, 3, mf, wild);
	wild |= (k > 60 || k < -60); // 4^k must be a normal float32
	mf[3] = wild ? __int_as_float(0x7fc00000) : (float)ldexp(1.0, 2 * k);
	if (wild) mf[0] = mf[1] = mf[2] = __int_as_float(0x7fc00000);
}

// Launch constants of the test in normalised units (float32, rounded up): cT multiplies the threshold side,
// cE multiplies q and already contains M^2 <= 4.
struct ScreenConsts {
	float cT, cE;
};
template <int TYPE> __device__ __forceinline__ ScreenConsts screen_consts(double T2, const NormDev &nd) {
	const double Tn = (TYPE == PXB_MODEL_PNP || TYPE == PXB_MODEL_LINE2D) ? T2 : T2 * nd.inv_s * nd.inv_s * (1.0 + 1e-15);
	ScreenConsts k;
	if (!(Tn > 0.0) || !(Tn <= 1e30)) { // NaN / non-positive / huge thresholds: everything takes the exact path
		k.cT = k.cE = __int_as_float(0x7f800000);
		return k;
	}
	// (p + q)^2 <= (1 + eta) p^2 + (1 + 1/eta) q^2 with eta = 1/16: the threshold side is inflated by 6 %, the (tiny)
	// error allowance by 17x
	constexpr double eta = 1.0 / 16.0;
	const double rT = sqrt(Tn);
	const double w = (TYPE == PXB_MODEL_FUNDAMENTAL || TYPE == PXB_MODEL_VANISHING_POINT)
	                     ? (2.0 * rT + 1.0)
	                     : (TYPE == PXB_MODEL_LINE2D ? 1.0 : (1.4142135623730951 + rT));
	k.cT = __double2float_ru((1.0 + eta) * Tn * kScreenSlack);
	k.cE = __double2float_ru(4.0 * (1.0 + 1.0 / eta) * w * w * (double)kScreenU2 * kScreenSlack * kScreenSlack);
	return k;
}

// true => the point is certainly not an inlier. p: normalised float32 coordinates, m: normalised float32 model,
// Z = cE * q (per point error allowance), cT as above.
template <int TYPE> __device__ __forceinline__ bool screen_sure_outlier(const float *p, const float *m, float cT, float Z);

template <> __device__ __forceinline__ bool screen_sure_outlier<PXB_MODEL_HOMOGRAPHY>(const float *p, const float *m, float cT, float Z) {
	const float t1 = fmaf(m[0], p[0], fmaf(m[1], p[1], m[2]));
	const float t2 = fmaf(m[3], p[0], fmaf(m[4], p[1], m[5]));
	const float t3 = fmaf(m[6], p[0], fmaf(m[7], p[1], m[8]));
	const float a = fmaf(p[2], t3, -t1), b = fmaf(p[3], t3, -t2);
	const float n = fmaf(a, a, b * b);
	return n >= fmaf(t3 * t3, cT, Z);
}
template <> __device__ __forceinline__ bool screen_sure_outlier<PXB_MODEL_PNP>(const float *p, const float *m, float cT, float Z) {
	const float px = fmaf(m[0], p[2], fmaf(m[1], p[3], fmaf(m[2], p[4], m[3])));
	const float py = fmaf(m[4], p[2], fmaf(m[5], p[3], fmaf(m[6], p[4], m[7])));
	const float pz = fmaf(m[8], p[2], fmaf(m[9], p[3], fmaf(m[10], p[4], m[11])));
	const float a = fmaf(-p[0], pz, px), b = fmaf(-p[1], pz, py);
	const float n = fmaf(a, a, b * b);
	return n >= fmaf(pz * pz, cT, Z);
}
template <> __device__ __forceinline__ bool screen_sure_outlier<PXB_MODEL_FUNDAMENTAL>(const float *p, const float *m, float cT, float Z) {
	const float rxc = fmaf(m[0], p[2], fmaf(m[3], p[3], m[6]));
	const float ryc = fmaf(m[1], p[2], fmaf(m[4], p[3], m[7]));
	const float rwc = fmaf(m[2], p[2], fmaf(m[5], p[3], m[8]));
	const float r = fmaf(p[0], rxc, fmaf(p[1], ryc, rwc));
	const float rx = fmaf(m[0], p[0], fmaf(m[1], p[1], m[2]));
	const float ry = fmaf(m[3], p[0], fmaf(m[4], p[1], m[5]));
	const float den = fmaf(rxc, rxc, fmaf(ryc, ryc, fmaf(rx, rx, ry * ry)));
	return r * r >= fmaf(den, cT, Z);
}

// VP: twice the line through the midpoint and the vanishing point, l' = (sy d2 - 2 d1, 2 d0 - sx d2, sx d1 - sy d0) with
// (sx, sy) = start + end; the segment's start point is tested against it: r^2 >= T (lx^2 + ly^2), the F-type test.
// |l'_k| <= 2 M (|sx| + |sy| + 1) =: M P1; forward errors: l' 6u M P1/2.., r <= 17 u M P1 P2 / 2 (P1 carries the factor 2).
template <> __device__ __forceinline__ bool screen_sure_outlier<PXB_MODEL_VANISHING_POINT>(const float *p, const float *m, float cT, float Z) {
	const float lx = fmaf(p[3], m[2], -m[4]);
	const float ly = fmaf(-p[2], m[2], m[3]);
	const float lz = fmaf(p[2], m[1], -(p[3] * m[0]));
	const float r = fmaf(lx, p[0], fmaf(ly, p[1], lz));
	const float den = fmaf(lx, lx, ly * ly);
	return r * r >= fmaf(den, cT, Z);
}
// 2D line: r_f^2 >= (1 + eta) T 4^k + (1 + 1/eta) E^2  ==>  r^2 >= T   (|r_f - r| <= 4 u M P1 <= E)
template <> __device__ __forceinline__ bool screen_sure_outlier<PXB_MODEL_LINE2D>(const float *p, const float *m, float cT, float Z) {
	const float r = fmaf(m[0], p[0], fmaf(m[1], p[1], m[2]));
	return r * r >= fmaf(m[3], cT, Z);
}

} // namespace pxb
