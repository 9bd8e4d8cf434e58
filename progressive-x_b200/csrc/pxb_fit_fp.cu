// pxb_fit_fp.cu -- batched non-minimal fits for fundamental matrices and poses ("next" row f-1 of SURVEY.md 8f), in
// the reduced form the end-to-end driver needs. One block per problem; the normal equations are block reductions with
// the library's fixed topology; the small dense algebra (9x9 / 12x12 symmetric eigenproblem, 3x3 SVD, 6x6 solve) runs
// in one thread.
//
//   k_fit_f    FundamentalMatrixEstimator::estimateModelNonminimal (gcr/estimators/fundamental_estimator.h:574-618):
//              Hartley normalisation (:636-735), the n x 9 system of FundamentalMatrixEightPointSolver
//              (solver_fundamental_matrix_eight_point.h:93-182) through A^T A, denormalisation F = T2^T Fn T1, unit
//              Frobenius norm, f33 >= 0. DIFFERENCES: the reference takes the last column of a FullPivHouseholderQR of
//              A^T A and then polishes with PoseLib's Levenberg-Marquardt (solver_fundamental_matrix_bundle_adjustment.h:
//              114-178, relative_pose/bundle.cpp); here the null vector is the smallest eigenvector (cyclic Jacobi) and
//              the rank-2 constraint is imposed by a 3x3 SVD instead of the LM parametrisation. n >= 8 required.
//   k_fit_pnp  PerspectiveNPointEstimator::estimateModelNonminimal -> PnPBundleAdjustment (solver_pnp_bundle_adjustment.h:
//              108-225): the reference initialises with cv::solvePnP(EPNP) and refines with PoseLib LM; here the
//              initialisation is a normalised DLT (n >= 6) projected onto SO(3), followed by 10 Levenberg-Marquardt
//              steps on the squared reprojection error (6 parameters, left-multiplicative rotation update).
//   k_f_sym_count  the symmetric-epipolar recount of FundamentalMatrixEstimator::isValidModel (:268-325, :224-252).
//
// These two solvers have no bit-level parity claim (the reference's own versions depend on Eigen/OpenCV internals);
// they are validated by what they must achieve: residuals of the fit on its own inliers (tests/test_gpu_fits.py).
#include <cfloat>

#include "pxb_internal.h"
#include "pxb_residuals.cuh"

namespace pxb {

constexpr int kFitT = 256;

__device__ __forceinline__ double fp_block_sum(double x, double *s_tmp) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) x = add(x, __shfl_xor_sync(0xffffffffu, x, o));
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	__syncthreads();
	if (lane == 0) s_tmp[warp] = x;
	__syncthreads();
	double t = 0.0;
#pragma unroll
	for (int w = 0; w < kFitT / 32; ++w) t = add(t, s_tmp[w]);
	return t;
}

// cyclic Jacobi on a symmetric n x n matrix (row-major A, destroyed); V receives the eigenvectors as columns
template <int NN> __device__ void jacobi_eig(double *A, double *V, double *w) {
	for (int i = 0; i < NN; ++i)
		for (int j = 0; j < NN; ++j) V[i * NN + j] = i == j ? 1.0 : 0.0;
	for (int sweep = 0; sweep < 60; ++sweep) {
		double off = 0.0, diag = 0.0;
		for (int i = 0; i < NN; ++i) {
			diag += A[i * NN + i] * A[i * NN + i];
			for (int j = i + 1; j < NN; ++j) off += A[i * NN + j] * A[i * NN + j];
		}
		if (!(off > 1e-30 * diag) || !(off == off)) break;
		for (int p = 0; p < NN - 1; ++p)
			for (int q = p + 1; q < NN; ++q) {
				const double apq = A[p * NN + q];
				if (apq == 0.0) continue;
				const double theta = (A[q * NN + q] - A[p * NN + p]) / (2.0 * apq);
				const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
				const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
				for (int k = 0; k < NN; ++k) {
					const double akp = A[k * NN + p], akq = A[k * NN + q];
					A[k * NN + p] = c * akp - s * akq;
					A[k * NN + q] = s * akp + c * akq;
				}
				for (int k = 0; k < NN; ++k) {
					const double apk = A[p * NN + k], aqk = A[q * NN + k];
					A[p * NN + k] = c * apk - s * aqk;
					A[q * NN + k] = s * apk + c * aqk;
				}
				for (int k = 0; k < NN; ++k) {
					const double vkp = V[k * NN + p], vkq = V[k * NN + q];
					V[k * NN + p] = c * vkp - s * vkq;
					V[k * NN + q] = s * vkp + c * vkq;
				}
			}
	}
	for (int i = 0; i < NN; ++i) w[i] = A[i * NN + i];
}

// M = U diag(s) V^T for a 3x3 (via the eigen-decomposition of M^T M); singular values descending
__device__ void svd3(const double *M, double *U, double *s, double *V) {
	double MtM[9], w[3];
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) MtM[i * 3 + j] = M[0 + i] * M[0 + j] + M[3 + i] * M[3 + j] + M[6 + i] * M[6 + j];
	jacobi_eig<3>(MtM, V, w);
	int ord[3] = {0, 1, 2};
	for (int a = 0; a < 2; ++a)
		for (int b = a + 1; b < 3; ++b)
			if (w[ord[b]] > w[ord[a]]) {
				const int t = ord[a];
				ord[a] = ord[b];
				ord[b] = t;
			}
	double Vs[9];
	for (int c = 0; c < 3; ++c)
		for (int r = 0; r < 3; ++r) Vs[r * 3 + c] = V[r * 3 + ord[c]];
	for (int i = 0; i < 9; ++i) V[i] = Vs[i];
	for (int c = 0; c < 3; ++c) s[c] = sqrt(fmax(w[ord[c]], 0.0));
	for (int c = 0; c < 2; ++c) { // u_c = M v_c / s_c
		double u[3];
		for (int r = 0; r < 3; ++r) u[r] = M[r * 3 + 0] * V[0 + c] + M[r * 3 + 1] * V[3 + c] + M[r * 3 + 2] * V[6 + c];
		const double nrm = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
		for (int r = 0; r < 3; ++r) U[r * 3 + c] = nrm > 0 ? u[r] / nrm : (r == c ? 1.0 : 0.0);
	}
	U[0 + 2] = U[3 + 0] * U[6 + 1] - U[6 + 0] * U[3 + 1]; // u2 = u0 x u1
	U[3 + 2] = U[6 + 0] * U[0 + 1] - U[0 + 0] * U[6 + 1];
	U[6 + 2] = U[0 + 0] * U[3 + 1] - U[3 + 0] * U[0 + 1];
}

// ------------------------------------------------------------------------------------------------
// fundamental matrix
// ------------------------------------------------------------------------------------------------
// closest rank-2 matrix (smallest singular value zeroed), scaled to unit Frobenius norm
__device__ void fp_rank2_unit(double *F) {
	double U[9], s[3], Vt[9];
	svd3(F, U, s, Vt);
	double nrm = 0.0;
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c) {
			F[r * 3 + c] = U[r * 3 + 0] * s[0] * Vt[c * 3 + 0] + U[r * 3 + 1] * s[1] * Vt[c * 3 + 1];
			nrm += F[r * 3 + c] * F[r * 3 + c];
		}
	nrm = sqrt(nrm);
	if (nrm > 0.0)
		for (int i = 0; i < 9; ++i) F[i] /= nrm;
}

__global__ void __launch_bounds__(kFitT)
    k_fit_f(const double *__restrict__ aos, const int32_t *__restrict__ off, const int32_t *__restrict__ idx,
            const double *__restrict__ weights, double *__restrict__ F_out, int32_t *__restrict__ ok_out) {
	__shared__ double s_tmp[kFitT / 32];
	__shared__ double s_acc[45];
	const int pb = blockIdx.x, tid = threadIdx.x;
	const int beg = off[pb], n = off[pb + 1] - beg;
	if (n < 8) {
		if (tid == 0) ok_out[pb] = 0;
		return;
	}
	double sx1 = 0, sy1 = 0, sx2 = 0, sy2 = 0;
	for (int t = tid; t < n; t += kFitT) {
		const double *q = aos + 4 * (int64_t)idx[beg + t];
		sx1 = add(sx1, q[0]);
		sy1 = add(sy1, q[1]);
		sx2 = add(sx2, q[2]);
		sy2 = add(sy2, q[3]);
	}
	const double mx1 = divd(fp_block_sum(sx1, s_tmp), (double)n), my1 = divd(fp_block_sum(sy1, s_tmp), (double)n);
	const double mx2 = divd(fp_block_sum(sx2, s_tmp), (double)n), my2 = divd(fp_block_sum(sy2, s_tmp), (double)n);
	double d1 = 0, d2 = 0;
	for (int t = tid; t < n; t += kFitT) {
		const double *q = aos + 4 * (int64_t)idx[beg + t];
		const double dx1 = sub(mx1, q[0]), dy1 = sub(my1, q[1]), dx2 = sub(mx2, q[2]), dy2 = sub(my2, q[3]);
		d1 = add(d1, __dsqrt_rn(add(mul(dx1, dx1), mul(dy1, dy1))));
		d2 = add(d2, __dsqrt_rn(add(mul(dx2, dx2), mul(dy2, dy2))));
	}
	const double r1 = divd(1.4142135623730951, divd(fp_block_sum(d1, s_tmp), (double)n));
	const double r2 = divd(1.4142135623730951, divd(fp_block_sum(d2, s_tmp), (double)n));
	double acc[45];
#pragma unroll
	for (int a = 0; a < 45; ++a) acc[a] = 0.0;
	for (int t = tid; t < n; t += kFitT) {
		const double *q = aos + 4 * (int64_t)idx[beg + t];
		const double x0 = mul(sub(q[0], mx1), r1), y0 = mul(sub(q[1], my1), r1);
		const double x1 = mul(sub(q[2], mx2), r2), y1 = mul(sub(q[3], my2), r2);
		const double w = weights ? weights[t] : 1.0; // weights_[i], i = row of the gathered sample (reference quirk)
		const double row[9] = {w * x1 * x0, w * x1 * y0, w * x1, w * y1 * x0, w * y1 * y0, w * y1, w * x0, w * y0, w};
		int a = 0;
#pragma unroll
		for (int r = 0; r < 9; ++r)
#pragma unroll
			for (int c = r; c < 9; ++c, ++a) acc[a] = add(acc[a], mul(row[r], row[c]));
	}
	for (int a = 0; a < 45; ++a) {
		const double v = fp_block_sum(acc[a], s_tmp);
		if (tid == 0) s_acc[a] = v;
	}
	__syncthreads();
	__shared__ double s_F[9], s_cand[9], s_vec[kFitT / 32][56], s_sys[56];
	__shared__ int s_go;
	if (tid == 0) {
		double A[81], V[81], w[9];
		int a = 0;
		for (int r = 0; r < 9; ++r)
			for (int c = r; c < 9; ++c, ++a) A[r * 9 + c] = A[c * 9 + r] = s_acc[a];
		jacobi_eig<9>(A, V, w);
		int best = 0;
		for (int i = 1; i < 9; ++i)
			if (w[i] < w[best]) best = i;
		double Fn[9];
		for (int i = 0; i < 9; ++i) Fn[i] = V[i * 9 + best];
		fp_rank2_unit(Fn);
		// the polish below works with ONE scale for both images (rs = sqrt(r1 r2)), so that its Sampson error is the pixel
		// Sampson error times a constant: y_i = rs (x - m_i) = (rs / r_i) xn_i  =>  F_lm = K2 Fn K1, K_i = diag(r_i/rs, r_i/rs, 1)
		const double rs = sqrt(r1 * r2), k1 = r1 / rs, k2 = r2 / rs;
		for (int r = 0; r < 3; ++r)
			for (int c = 0; c < 3; ++c) s_F[r * 3 + c] = Fn[r * 3 + c] * (r < 2 ? k2 : 1.0) * (c < 2 ? k1 : 1.0);
		fp_rank2_unit(s_F);
	}
	__syncthreads();
	// ---- Levenberg-Marquardt polish of the (weighted) Sampson error, the objective of the reference's bundle adjustment
	// (solver_fundamental_matrix_bundle_adjustment.h:114-178 -> PoseLib refine_fundamental). 9 parameters with Marquardt
	// scaling; the scale gauge is absorbed by the damping, the rank-2 constraint is re-imposed after every step, and a
	// step is kept only if the error decreases.
	const double rs = sqrt(r1 * r2);
	auto accumulate = [&](const double *F, double *sys /*45 JtJ + 9 Jtr + cost*/) {
		double acc2[55];
#pragma unroll
		for (int a = 0; a < 55; ++a) acc2[a] = 0.0;
		for (int t = tid; t < n; t += kFitT) {
			const double *q = aos + 4 * (int64_t)idx[beg + t];
			const double x1[3] = {(q[0] - mx1) * rs, (q[1] - my1) * rs, 1.0}, x2[3] = {(q[2] - mx2) * rs, (q[3] - my2) * rs, 1.0};
			const double wgt = weights ? weights[t] : 1.0;
			double Fx[3], Ftx[3];
			for (int r = 0; r < 3; ++r) Fx[r] = F[r * 3] * x1[0] + F[r * 3 + 1] * x1[1] + F[r * 3 + 2];
			for (int c = 0; c < 3; ++c) Ftx[c] = F[c] * x2[0] + F[3 + c] * x2[1] + F[6 + c];
			const double C = x2[0] * Fx[0] + x2[1] * Fx[1] + Fx[2];
			const double S = Fx[0] * Fx[0] + Fx[1] * Fx[1] + Ftx[0] * Ftx[0] + Ftx[1] * Ftx[1];
			if (!(S > 1e-300)) continue;
			const double inv = 1.0 / sqrt(S), res = wgt * C * inv;
			double J[9];
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c) {
					const double dC = x2[r] * x1[c];
					const double dS = 2.0 * ((r < 2 ? Fx[r] * x1[c] : 0.0) + (c < 2 ? Ftx[c] * x2[r] : 0.0));
					J[r * 3 + c] = wgt * (dC * inv - 0.5 * C * inv * inv * inv * dS);
				}
			int a = 0;
#pragma unroll
			for (int r = 0; r < 9; ++r)
#pragma unroll
				for (int c = r; c < 9; ++c, ++a) acc2[a] += J[r] * J[c];
#pragma unroll
			for (int r = 0; r < 9; ++r) acc2[45 + r] += J[r] * res;
			acc2[54] += res * res;
		}
		// vector block reduction: butterfly inside each warp, then 55 threads add the 8 warp rows in order
#pragma unroll
		for (int a = 0; a < 55; ++a)
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) acc2[a] += __shfl_xor_sync(0xffffffffu, acc2[a], o);
		__syncthreads();
		if ((tid & 31) == 0)
			for (int a = 0; a < 55; ++a) s_vec[tid >> 5][a] = acc2[a];
		__syncthreads();
		if (tid < 55) {
			double v = 0.0;
			for (int wv = 0; wv < kFitT / 32; ++wv) v += s_vec[wv][tid];
			sys[tid] = v;
		}
		__syncthreads();
	};
	accumulate(s_F, s_sys);
	__shared__ double s_sys2[56];
	double mu = 1e-3;
	for (int it = 0; it < 8; ++it) {
		if (tid == 0) { // solve (JtJ + mu diag(JtJ)) delta = -Jtr by Gaussian elimination with partial pivoting
			double M[9][10];
			int a = 0;
			for (int r = 0; r < 9; ++r)
				for (int c = r; c < 9; ++c, ++a) M[r][c] = M[c][r] = s_sys[a];
			for (int r = 0; r < 9; ++r) {
				M[r][r] += mu * fmax(M[r][r], 1e-12);
				M[r][9] = -s_sys[45 + r];
			}
			bool okk = true;
			for (int k = 0; k < 9 && okk; ++k) {
				int piv = k;
				for (int r = k + 1; r < 9; ++r)
					if (fabs(M[r][k]) > fabs(M[piv][k])) piv = r;
				if (!(fabs(M[piv][k]) > 1e-300)) {
					okk = false;
					break;
				}
				if (piv != k)
					for (int c = 0; c < 10; ++c) {
						const double t = M[k][c];
						M[k][c] = M[piv][c];
						M[piv][c] = t;
					}
				for (int r = k + 1; r < 9; ++r) {
					const double f = M[r][k] / M[k][k];
					for (int c = k; c < 10; ++c) M[r][c] -= f * M[k][c];
				}
			}
			double delta[9];
			if (okk)
				for (int r = 8; r >= 0; --r) {
					double v = M[r][9];
					for (int c = r + 1; c < 9; ++c) v -= M[r][c] * delta[c];
					delta[r] = v / M[r][r];
					okk &= fabs(delta[r]) <= 1e300;
				}
			if (okk) {
				for (int i = 0; i < 9; ++i) s_cand[i] = s_F[i] + delta[i];
				fp_rank2_unit(s_cand);
			}
			s_go = okk ? 1 : 0;
		}
		__syncthreads();
		if (!s_go) break; // block-uniform
		accumulate(s_cand, s_sys2);
		if (tid == 0) {
			if (s_sys2[54] < s_sys[54]) {
				const bool small = (s_sys[54] - s_sys2[54]) <= 1e-12 * s_sys[54];
				for (int i = 0; i < 9; ++i) s_F[i] = s_cand[i];
				for (int i = 0; i < 55; ++i) s_sys[i] = s_sys2[i];
				s_go = small ? 0 : 1;
			} else {
				s_go = 2; // rejected: more damping
			}
		}
		__syncthreads();
		if (s_go == 0) break;
		mu = (s_go == 2) ? mu * 10.0 : mu * 0.3;
		if (mu > 1e6) break;
	}
	if (tid != 0) return;
	// F = T2^T F_lm T1 with T_i = [rs 0 -rs m_x; 0 rs -rs m_y; 0 0 1]
	const double T1[9] = {rs, 0, -rs * mx1, 0, rs, -rs * my1, 0, 0, 1};
	const double T2[9] = {rs, 0, -rs * mx2, 0, rs, -rs * my2, 0, 0, 1};
	double tmp[9], F[9];
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c) tmp[r * 3 + c] = T2[0 + r] * s_F[0 + c] + T2[3 + r] * s_F[3 + c] + T2[6 + r] * s_F[6 + c];
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c) F[r * 3 + c] = tmp[r * 3 + 0] * T1[0 + c] + tmp[r * 3 + 1] * T1[3 + c] + tmp[r * 3 + 2] * T1[6 + c];
	double nrm = 0;
	for (int i = 0; i < 9; ++i) nrm += F[i] * F[i];
	nrm = sqrt(nrm);
	bool bad = !(nrm > 0.0) || !(nrm <= DBL_MAX);
	const double sgn = (F[8] < 0) ? -1.0 : 1.0; // fundamental_estimator.h:611-613
	for (int i = 0; i < 9; ++i) F_out[9 * (int64_t)pb + i] = bad ? 0.0 : sgn * F[i] / nrm;
	ok_out[pb] = bad ? 0 : 1;
}

// (count of Sampson inliers, count of those that are also symmetric-epipolar inliers)
__global__ void __launch_bounds__(1024)
    k_f_sym_count(const double *__restrict__ soa, int64_t stride, int64_t N, const double *__restrict__ model, double T2,
                  double Tsym2, long long *__restrict__ out2) {
	__shared__ double m[9];
	__shared__ int s_a[32], s_b[32];
	if (threadIdx.x < 9) m[threadIdx.x] = model[threadIdx.x];
	__syncthreads();
	int ca = 0, cb = 0;
	for (int64_t i = threadIdx.x; i < N; i += 1024) {
		double p[5];
		load_point<4>(soa, stride, i, p);
		if (!(squared_residual<PXB_MODEL_FUNDAMENTAL>(p, m) < T2)) continue;
		++ca;
		const double x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
		const double rxc = add(add(mul(m[0], x2), mul(m[3], y2)), m[6]);
		const double ryc = add(add(mul(m[1], x2), mul(m[4], y2)), m[7]);
		const double rwc = add(add(mul(m[2], x2), mul(m[5], y2)), m[8]);
		const double r = add(add(mul(x1, rxc), mul(y1, ryc)), rwc);
		const double rx = add(add(mul(m[0], x1), mul(m[1], y1)), m[2]);
		const double ry = add(add(mul(m[3], x1), mul(m[4], y1)), m[5]);
		const double a = add(mul(rxc, rxc), mul(ryc, ryc)), b = add(mul(rx, rx), mul(ry, ry));
		const double sym = divd(mul(mul(r, r), add(a, b)), mul(a, b)); // fundamental_estimator.h:224-252
		if (sym < Tsym2) ++cb;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		ca += __shfl_xor_sync(0xffffffffu, ca, o);
		cb += __shfl_xor_sync(0xffffffffu, cb, o);
	}
	if ((threadIdx.x & 31) == 0) {
		s_a[threadIdx.x >> 5] = ca;
		s_b[threadIdx.x >> 5] = cb;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		long long A = 0, B = 0;
		for (int w = 0; w < 32; ++w) {
			A += s_a[w];
			B += s_b[w];
		}
		out2[0] = A;
		out2[1] = B;
	}
}

// ------------------------------------------------------------------------------------------------
// pose (normalised DLT + LM on the reprojection error)
// ------------------------------------------------------------------------------------------------
__device__ void rodrigues_left(const double w[3], double R[9]) { // R <- exp([w]x) R
	const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
	const double th = sqrt(th2);
	double a, b; // exp = I + a K + b K^2
	if (th < 1e-8) {
		a = 1.0;
		b = 0.5;
	} else {
		a = sin(th) / th;
		b = (1.0 - cos(th)) / th2;
	}
	const double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
	double K2[9], E[9], out[9];
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) K2[i * 3 + j] = K[i * 3 + 0] * K[0 + j] + K[i * 3 + 1] * K[3 + j] + K[i * 3 + 2] * K[6 + j];
	for (int i = 0; i < 9; ++i) E[i] = (i % 4 == 0 ? 1.0 : 0.0) + a * K[i] + b * K2[i];
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) out[i * 3 + j] = E[i * 3 + 0] * R[0 + j] + E[i * 3 + 1] * R[3 + j] + E[i * 3 + 2] * R[6 + j];
	for (int i = 0; i < 9; ++i) R[i] = out[i];
}

__global__ void __launch_bounds__(kFitT)
    k_fit_pnp(const double *__restrict__ aos, const int32_t *__restrict__ off, const int32_t *__restrict__ idx,
              double *__restrict__ P_out, int32_t *__restrict__ ok_out) {
	__shared__ double s_tmp[kFitT / 32];
	__shared__ double s_acc[78];
	__shared__ double s_pose[12], s_best[12];
	__shared__ double s_best_cost;
	__shared__ int s_ok, s_skip;
	const int pb = blockIdx.x, tid = threadIdx.x;
	const int beg = off[pb], n = off[pb + 1] - beg;
	if (n < 6) {
		if (tid == 0) ok_out[pb] = 0;
		return;
	}
	// ---- normalise the 3D points (centroid, mean distance sqrt(3)) ----
	double cx = 0, cy = 0, cz = 0;
	for (int t = tid; t < n; t += kFitT) {
		const double *q = aos + 5 * (int64_t)idx[beg + t];
		cx = add(cx, q[2]);
		cy = add(cy, q[3]);
		cz = add(cz, q[4]);
	}
	cx = fp_block_sum(cx, s_tmp) / n;
	cy = fp_block_sum(cy, s_tmp) / n;
	cz = fp_block_sum(cz, s_tmp) / n;
	double md = 0;
	for (int t = tid; t < n; t += kFitT) {
		const double *q = aos + 5 * (int64_t)idx[beg + t];
		const double dx = q[2] - cx, dy = q[3] - cy, dz = q[4] - cz;
		md = add(md, sqrt(dx * dx + dy * dy + dz * dz));
	}
	md = fp_block_sum(md, s_tmp) / n;
	const double sc = md > 0 ? 1.7320508075688772 / md : 1.0;
	// ---- DLT normal equations (12 x 12, 78 unique) ----
	{
		double acc[78];
#pragma unroll
		for (int a = 0; a < 78; ++a) acc[a] = 0.0;
		for (int t = tid; t < n; t += kFitT) {
			const double *q = aos + 5 * (int64_t)idx[beg + t];
			const double u = q[0], v = q[1], X = (q[2] - cx) * sc, Y = (q[3] - cy) * sc, Z = (q[4] - cz) * sc;
			const double ra[12] = {X, Y, Z, 1, 0, 0, 0, 0, -u * X, -u * Y, -u * Z, -u};
			const double rb[12] = {0, 0, 0, 0, X, Y, Z, 1, -v * X, -v * Y, -v * Z, -v};
			int a = 0;
#pragma unroll
			for (int r = 0; r < 12; ++r)
#pragma unroll
				for (int c = r; c < 12; ++c, ++a) acc[a] = add(acc[a], add(mul(ra[r], ra[c]), mul(rb[r], rb[c])));
		}
		for (int a = 0; a < 78; ++a) {
			const double v = fp_block_sum(acc[a], s_tmp);
			if (tid == 0) s_acc[a] = v;
		}
	}
	__syncthreads();
	if (tid == 0) {
		double A[144], V[144], w[12];
		int a = 0;
		for (int r = 0; r < 12; ++r)
			for (int c = r; c < 12; ++c, ++a) A[r * 12 + c] = A[c * 12 + r] = s_acc[a];
		jacobi_eig<12>(A, V, w);
		int best = 0;
		for (int i = 1; i < 12; ++i)
			if (w[i] < w[best]) best = i;
		double Pn[12];
		for (int i = 0; i < 12; ++i) Pn[i] = V[i * 12 + best];
		// project the left 3x3 onto SO(3); scale = mean singular value; sign from det
		double M[9] = {Pn[0], Pn[1], Pn[2], Pn[4], Pn[5], Pn[6], Pn[8], Pn[9], Pn[10]};
		double det = M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
		const double sg = det < 0 ? -1.0 : 1.0;
		for (int i = 0; i < 9; ++i) M[i] *= sg;
		double U[9], s[3], Vv[9];
		svd3(M, U, s, Vv);
		const double scale = (s[0] + s[1] + s[2]) / 3.0;
		double R[9];
		// nearest rotation: U diag(1, 1, det(U V^T)) V^T  (U is right-handed by construction, V may not be)
		const double detV = Vv[0] * (Vv[4] * Vv[8] - Vv[5] * Vv[7]) - Vv[1] * (Vv[3] * Vv[8] - Vv[5] * Vv[6]) +
		                    Vv[2] * (Vv[3] * Vv[7] - Vv[4] * Vv[6]);
		const double d3 = detV < 0 ? -1.0 : 1.0;
		for (int r = 0; r < 3; ++r)
			for (int c = 0; c < 3; ++c)
				R[r * 3 + c] = U[r * 3 + 0] * Vv[c * 3 + 0] + U[r * 3 + 1] * Vv[c * 3 + 1] + d3 * U[r * 3 + 2] * Vv[c * 3 + 2];
		double tn[3] = {sg * Pn[3] / scale, sg * Pn[7] / scale, sg * Pn[11] / scale};
		// undo the 3D normalisation: p = R (sc (X - c)) + tn  ->  R' = R, t' = tn/sc... keep R, rescale depth:
		// the DLT solution is up to scale `scale`; in normalised units p_n = R X_n + tn with X_n = sc (X - c)
		// => p = p_n / sc = R (X - c) + tn / sc
		double t3[3];
		for (int r = 0; r < 3; ++r) t3[r] = tn[r] / sc - (R[r * 3 + 0] * cx + R[r * 3 + 1] * cy + R[r * 3 + 2] * cz);
		for (int r = 0; r < 3; ++r) {
			s_pose[4 * r + 0] = R[r * 3 + 0];
			s_pose[4 * r + 1] = R[r * 3 + 1];
			s_pose[4 * r + 2] = R[r * 3 + 2];
			s_pose[4 * r + 3] = t3[r];
		}
		bool ok = scale > 0;
		for (int i = 0; i < 12; ++i) ok = ok && (fabs(s_pose[i]) <= DBL_MAX);
		s_ok = ok ? 1 : 0;
	}
	__syncthreads();
	if (!s_ok) {
		if (tid == 0) ok_out[pb] = 0;
		return;
	}
	// ---- Levenberg-Marquardt on sum ||proj(R X + t) - (u, v)||^2 (steps that raise the cost are rolled back) ----
	double mu = 1e-4;
	if (tid == 0) {
		s_best_cost = DBL_MAX;
		for (int i = 0; i < 12; ++i) s_best[i] = s_pose[i];
	}
	__syncthreads();
	for (int it = 0; it < 14; ++it) {
		double R[9], tt[3];
		for (int r = 0; r < 3; ++r) {
			R[r * 3 + 0] = s_pose[4 * r + 0];
			R[r * 3 + 1] = s_pose[4 * r + 1];
			R[r * 3 + 2] = s_pose[4 * r + 2];
			tt[r] = s_pose[4 * r + 3];
		}
		double acc[28]; // 21 JtJ + 6 Jtr + cost
#pragma unroll
		for (int a = 0; a < 28; ++a) acc[a] = 0.0;
		for (int t = tid; t < n; t += kFitT) {
			const double *q = aos + 5 * (int64_t)idx[beg + t];
			const double Xr[3] = {R[0] * q[2] + R[1] * q[3] + R[2] * q[4], R[3] * q[2] + R[4] * q[3] + R[5] * q[4],
			                      R[6] * q[2] + R[7] * q[3] + R[8] * q[4]};
			const double x = Xr[0] + tt[0], y = Xr[1] + tt[1], z = Xr[2] + tt[2];
			const double iz = 1.0 / z;
			const double ru = x * iz - q[0], rv = y * iz - q[1];
			// d(res)/dp
			const double a0[3] = {iz, 0, -x * iz * iz}, a1[3] = {0, iz, -y * iz * iz};
			// dp/dw = -[Xr]x, dp/dt = I
			const double Jw[9] = {0, Xr[2], -Xr[1], -Xr[2], 0, Xr[0], Xr[1], -Xr[0], 0};
			double J0[6], J1[6];
			for (int c = 0; c < 3; ++c) {
				J0[c] = a0[0] * Jw[0 + c] + a0[1] * Jw[3 + c] + a0[2] * Jw[6 + c];
				J1[c] = a1[0] * Jw[0 + c] + a1[1] * Jw[3 + c] + a1[2] * Jw[6 + c];
				J0[3 + c] = a0[c];
				J1[3 + c] = a1[c];
			}
			int a = 0;
#pragma unroll
			for (int r = 0; r < 6; ++r)
#pragma unroll
				for (int c = r; c < 6; ++c, ++a) acc[a] += J0[r] * J0[c] + J1[r] * J1[c];
#pragma unroll
			for (int r = 0; r < 6; ++r) acc[21 + r] += J0[r] * ru + J1[r] * rv;
			acc[27] += ru * ru + rv * rv;
		}
		__syncthreads();
		for (int a = 0; a < 28; ++a) {
			const double v = fp_block_sum(acc[a], s_tmp);
			if (tid == 0) s_acc[a] = v;
		}
		__syncthreads();
		if (tid == 0) {
			const double cost = s_acc[27];
			s_skip = 0;
			if (cost <= s_best_cost) { // accept the pose the step led to
				s_best_cost = cost;
				for (int i = 0; i < 12; ++i) s_best[i] = s_pose[i];
				mu *= 0.3;
			} else { // roll back and damp harder; the normal equations are rebuilt at the kept pose next round
				for (int i = 0; i < 12; ++i) s_pose[i] = s_best[i];
				mu *= 10.0;
				s_skip = 1;
			}
		}
		__syncthreads();
		if (s_skip) continue;
		if (tid == 0) {
			double M[6][7];
			int a = 0;
			for (int r = 0; r < 6; ++r)
				for (int c = r; c < 6; ++c, ++a) M[r][c] = M[c][r] = s_acc[a];
			for (int r = 0; r < 6; ++r) {
				M[r][r] *= (1.0 + mu);
				M[r][6] = -s_acc[21 + r];
			}
			bool sing = false;
			for (int c = 0; c < 6 && !sing; ++c) {
				int piv = c;
				for (int r = c + 1; r < 6; ++r)
					if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
				if (!(fabs(M[piv][c]) > 0)) {
					sing = true;
					break;
				}
				if (piv != c)
					for (int j = 0; j < 7; ++j) {
						const double tsw = M[c][j];
						M[c][j] = M[piv][j];
						M[piv][j] = tsw;
					}
				for (int r = c + 1; r < 6; ++r) {
					const double f = M[r][c] / M[c][c];
					for (int j = c; j < 7; ++j) M[r][j] -= f * M[c][j];
				}
			}
			if (!sing) {
				double dlt[6];
				for (int r = 5; r >= 0; --r) {
					double v = M[r][6];
					for (int j = r + 1; j < 6; ++j) v -= M[r][j] * dlt[j];
					dlt[r] = v / M[r][r];
				}
				bool fin = true;
				for (int r = 0; r < 6; ++r) fin = fin && (fabs(dlt[r]) <= 1e6);
				if (fin) {
					double Rn[9];
					for (int r = 0; r < 3; ++r) {
						Rn[r * 3 + 0] = s_pose[4 * r + 0];
						Rn[r * 3 + 1] = s_pose[4 * r + 1];
						Rn[r * 3 + 2] = s_pose[4 * r + 2];
					}
					// left update rotates R X; the translation stays additive
					rodrigues_left(dlt, Rn);
					for (int r = 0; r < 3; ++r) {
						s_pose[4 * r + 0] = Rn[r * 3 + 0];
						s_pose[4 * r + 1] = Rn[r * 3 + 1];
						s_pose[4 * r + 2] = Rn[r * 3 + 2];
						s_pose[4 * r + 3] += dlt[3 + r];
					}
				}
			}
		}
		__syncthreads();
	}
	if (tid == 0) {
		bool ok = true;
		for (int i = 0; i < 12; ++i) {
			ok = ok && (fabs(s_best[i]) <= DBL_MAX);
			P_out[12 * (int64_t)pb + i] = s_best[i];
		}
		ok_out[pb] = ok ? 1 : 0;
	}
}

int launch_fit_f(pxb_ctx *ctx, int P, const int32_t *off, const int32_t *idx, const double *weights, double *F_out,
                 int32_t *ok_out) {
	if (P <= 0) return PXB_OK;
	k_fit_f<<<(unsigned)P, kFitT, 0, ctx->stream>>>(ctx->pts.aos, off, idx, weights, F_out, ok_out);
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

int launch_fit_pnp(pxb_ctx *ctx, int P, const int32_t *off, const int32_t *idx, double *P_out, int32_t *ok_out) {
	if (P <= 0) return PXB_OK;
	k_fit_pnp<<<(unsigned)P, kFitT, 0, ctx->stream>>>(ctx->pts.aos, off, idx, P_out, ok_out);
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

int launch_f_sym_count(pxb_ctx *ctx, const double *model, double T2, double Tsym2, long long *out2) {
	const Points &p = ctx->pts;
	k_f_sym_count<<<1, 1024, 0, ctx->stream>>>(p.soa, p.stride, p.N, model, T2, Tsym2, out2);
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

} // namespace pxb
