// pxb_fit_fp.cu -- batched non-minimal fits for fundamental matrices and poses ("next" row f-1 of SURVEY.md 8f), in
// the reduced form the end-to-end driver needs. One block per problem; the normal equations are block reductions with
// the library's fixed topology; the small dense algebra (9x9 / 12x12 symmetric eigenproblem, 3x3 SVD, 6x6 solve) runs
// in one thread.
//
//   k_fit_f    FundamentalMatrixEstimator::estimateModelNonminimal (gcr/estimators/fundamental_estimator.h:574-618):
//              Hartley normalisation (:636-735), the n x 9 system of FundamentalMatrixEightPointSolver
//              (solver_fundamental_matrix_eight_point.h:93-182) through A^T A, denormalisation F = T2^T Fn T1, unit
//              Frobenius norm, f33 >= 0, and in between PoseLib's Levenberg-Marquardt refinement restated (lm_F_impl,
//              relative_pose/bundle.cpp:253-340: 7-parameter factorised F, truncated loss, LLT steps -- see the kernel).
//              DIFFERENCES: the reference's eight-point step takes the last column of a FullPivHouseholderQR of A^T A,
//              here the null vector is the smallest eigenvector (cyclic Jacobi) -- both only seed the LM; n >= 8 required
//              (for n == 7 the reference refines the up-to-three seven-point solutions).
//   k_fit_pnp  PerspectiveNPointEstimator::estimateModelNonminimal -> PnPBundleAdjustment (solver_pnp_bundle_adjustment.h:
//              108-225): the reference initialises with cv::solvePnP(EPNP) (OpenCV, not on disk) and refines with PoseLib's
//              refine_pnp; here the initialisation is a normalised DLT (n >= 6) projected onto SO(3) and the refinement is
//              lm_pnp_impl restated (relative_pose/bundle.cpp:24-100: truncated loss, LLT steps, right-multiplicative
//              rotation update -- see the kernel).
//   k_f_sym_count  the symmetric-epipolar recount of FundamentalMatrixEstimator::isValidModel (:268-325, :224-252).
//
// These two solvers have no bit-level parity claim (the reference's own versions depend on Eigen/OpenCV internals); they
// are compared with numpy restatements of the same algorithms (oracle/px_sequential.py, 1e-9 / 1e-7) and validated by what
// they must achieve: residuals of the fit on its own inliers (tests/test_gpu_fits.py).
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

// Jacobi executed by ONE WARP on matrices in shared memory, in ROUND-ROBIN order: a sweep is NP - 1 rounds (NP = NN rounded
// up to even) of NP / 2 rotations on disjoint index pairs (the tournament schedule: player NP - 1 stays, the others move
// round a circle). Disjoint plane rotations commute and none reads what another one of its round writes, so a round is
// the sequential application of its rotations -- but their angles (two divisions and two square roots each: ~600 cycles
// of dependent float64 latency) are computed side by side by NP / 2 lanes, and the column / row updates of all of them
// are two phases in which lane k owns row / column k. The serial order cost one such latency chain per rotation (66 per
// sweep of the 12 x 12 pose problem, 36 of the 9 x 9 one): ~200 of k_fit_pnp's 284 us. Rotation arithmetic per element is
// unchanged.
template <int NN> __device__ int jacobi_eig_warp(double *A, double *V, double *w) {
	constexpr int NP = (NN + 1) / 2 * 2, HALF = NP / 2;
	__shared__ double s_c[HALF], s_s[HALF];
	__shared__ int s_p[HALF], s_q[HALF];
	const int lane = threadIdx.x & 31;
	for (int e = lane; e < NN * NN; e += 32) V[e] = (e / NN == e % NN) ? 1.0 : 0.0;
	__syncwarp();
	int sweep = 0;
	for (; sweep < 60; ++sweep) {
		double off = 0.0, diag = 0.0;
		for (int e = lane; e < NN * NN; e += 32) {
			const int i = e / NN, j = e % NN;
			const double a = A[e];
			if (i == j) diag += a * a;
			else if (j > i) off += a * a;
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			off += __shfl_xor_sync(0xffffffffu, off, o);
			diag += __shfl_xor_sync(0xffffffffu, diag, o);
		}
		if (!(off > 1e-30 * diag) || !(off == off)) break;
		for (int round = 0; round < NP - 1; ++round) {
			if (lane < HALF) { // this lane's pair of the round and its rotation
				int a = lane == 0 ? NP - 1 : (round + lane) % (NP - 1);
				int b = lane == 0 ? round : (round - lane + (NP - 1)) % (NP - 1);
				const int p = min(a, b), q = max(a, b);
				double c = 1.0, sn = 0.0;
				if (q < NN) { // (odd NN: the pair with the bye does nothing)
					const double apq = A[p * NN + q];
					if (apq != 0.0) {
						// t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)), theta = (aqq - app) / (2 apq), with numerator and
						// denominator multiplied by |2 apq|: one square root, one division and one reciprocal square root on
						// the critical path instead of two divisions, two square roots and a reciprocal
						const double d = A[q * NN + q] - A[p * NN + p];
						const double r = sqrt(d * d + 4.0 * apq * apq);
						const double t = (d >= 0 ? 2.0 * apq : -2.0 * apq) / (fabs(d) + r);
						c = rsqrt(t * t + 1.0);
						sn = t * c;
					}
				}
				s_p[lane] = p;
				s_q[lane] = q < NN ? q : p; // (identity rotation on (p, p): never applied, see below)
				s_c[lane] = c;
				s_s[lane] = sn;
			}
			__syncwarp();
			// (the pairs of a round touch disjoint columns / rows: all operands are read before anything is written, so the
			// shared-memory round trips of the HALF rotations overlap instead of running one after the other)
			if (lane < NN) { // columns p, q of row `lane` of A and of V
				int pp[HALF], qq[HALF];
				double cc[HALF], ss[HALF], akp[HALF], akq[HALF], vkp[HALF], vkq[HALF];
#pragma unroll
				for (int j = 0; j < HALF; ++j) {
					pp[j] = s_p[j], qq[j] = s_q[j];
					cc[j] = s_c[j], ss[j] = s_s[j];
					akp[j] = A[lane * NN + pp[j]], akq[j] = A[lane * NN + qq[j]];
					vkp[j] = V[lane * NN + pp[j]], vkq[j] = V[lane * NN + qq[j]];
				}
#pragma unroll
				for (int j = 0; j < HALF; ++j) {
					if (pp[j] == qq[j] || (cc[j] == 1.0 && ss[j] == 0.0)) continue;
					A[lane * NN + pp[j]] = cc[j] * akp[j] - ss[j] * akq[j];
					A[lane * NN + qq[j]] = ss[j] * akp[j] + cc[j] * akq[j];
					V[lane * NN + pp[j]] = cc[j] * vkp[j] - ss[j] * vkq[j];
					V[lane * NN + qq[j]] = ss[j] * vkp[j] + cc[j] * vkq[j];
				}
			}
			__syncwarp();
			if (lane < NN) { // rows p, q of column `lane` of A
				int pp[HALF], qq[HALF];
				double cc[HALF], ss[HALF], apk[HALF], aqk[HALF];
#pragma unroll
				for (int j = 0; j < HALF; ++j) {
					pp[j] = s_p[j], qq[j] = s_q[j];
					cc[j] = s_c[j], ss[j] = s_s[j];
					apk[j] = A[pp[j] * NN + lane], aqk[j] = A[qq[j] * NN + lane];
				}
#pragma unroll
				for (int j = 0; j < HALF; ++j) {
					if (pp[j] == qq[j] || (cc[j] == 1.0 && ss[j] == 0.0)) continue;
					A[pp[j] * NN + lane] = cc[j] * apk[j] - ss[j] * aqk[j];
					A[qq[j] * NN + lane] = ss[j] * apk[j] + cc[j] * aqk[j];
				}
			}
			__syncwarp();
		}
	}
	__syncwarp();
	if (lane < NN) w[lane] = A[lane * NN + lane];
	__syncwarp();
	return sweep;
}

// sums of NV per-thread values over the block (butterfly inside the warp, then the warps in order): two barriers for
// the whole vector. s_vec: [kFitT / 32][NV] shared, out: [NV] shared.
template <int NV> __device__ __forceinline__ void fp_block_sums(double (&v)[NV], double *s_vec, double *out) {
#pragma unroll
	for (int a = 0; a < NV; ++a)
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) v[a] = add(v[a], __shfl_xor_sync(0xffffffffu, v[a], o));
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	__syncthreads();
	if (lane == 0)
#pragma unroll
		for (int a = 0; a < NV; ++a) s_vec[warp * NV + a] = v[a];
	__syncthreads();
	if (threadIdx.x < NV) {
		double t = 0.0;
#pragma unroll
		for (int wv = 0; wv < kFitT / 32; ++wv) t = add(t, s_vec[wv * NV + threadIdx.x]);
		out[threadIdx.x] = t;
	}
	__syncthreads();
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
	__shared__ double s_big[(kFitT / 32) * 45];
	fp_block_sums<45>(acc, s_big, s_acc);
	// ---- FundamentalMatrixBundleAdjustmentSolver (solver_fundamental_matrix_bundle_adjustment.h:114-178): the eight-point
	// estimate is refined by PoseLib's refine_fundamental = lm_F_impl (relative_pose/bundle.cpp:253-340,454-500) ON THE
	// NORMALISED points (fundamental_estimator.h:586-603 hands the solver `normalized_points`), restated here:
	//   * F = U diag(1, sigma, 0) V^T (FactorizedFundamentalMatrix, jacobian_impl.h:472-488), 7 parameters: U <- exp([w1]x) U,
	//     V <- exp([w2]x) V, sigma <- sigma + d;
	//   * residual r = C / |J_C| (Sampson), cost = sum w_k min(r^2, loss_scale^2) with the TRUNCATED loss, loss_scale = 1
	//     (normalised units), IRLS weight (r^2 < 1) / n (jacobian_impl.h:504-606);
	//   * Levenberg-Marquardt: lambda0 = 1e-3, +lambda on the diagonal, LLT solve, step accepted iff the cost decreases
	//     (lambda /= 10, Jacobian recomputed) else lambda *= 10; stops at |J^T r| < 1e-8, |step| < 1e-8 or 25 iterations.
	// Column sums use the library's fixed block topology. Not restated: the 7-point initialisation for n == 7 (n >= 8 here).
	__shared__ double s_U[9], s_V[9], s_Fc[9], s_Un[9], s_Vn[9], s_Fn[9], s_vec[kFitT / 32][36], s_sys[36];
	__shared__ double s_sigma, s_sigma_n;
	__shared__ int s_go;
	auto compose = [](const double *U, const double *V, double sigma, double *F) { // U.col(0) V.col(0)^T + sigma U.col(1) V.col(1)^T
		for (int r = 0; r < 3; ++r)
			for (int c = 0; c < 3; ++c) F[r * 3 + c] = U[r * 3 + 0] * V[c * 3 + 0] + sigma * U[r * 3 + 1] * V[c * 3 + 1];
	};
	__shared__ double s_A[81], s_Vv[81], s_w[9];
	if (tid < 32) {
		int a = 0;
		if (tid == 0)
			for (int r = 0; r < 9; ++r)
				for (int c = r; c < 9; ++c, ++a) s_A[r * 9 + c] = s_A[c * 9 + r] = s_acc[a];
		__syncwarp();
		jacobi_eig_warp<9>(s_A, s_Vv, s_w);
	}
	__syncthreads();
	if (tid == 0) {
		int best = 0;
		for (int i = 1; i < 9; ++i)
			if (s_w[i] < s_w[best]) best = i;
		double Fn[9], sv[3];
		for (int i = 0; i < 9; ++i) Fn[i] = s_Vv[i * 9 + best];
		svd3(Fn, s_U, sv, s_V); // FactorizedFundamentalMatrix(F): U, V of the SVD, sigma = s1 / s0
		s_sigma = sv[0] > 0.0 ? sv[1] / sv[0] : 0.0;
		compose(s_U, s_V, s_sigma, s_Fc);
	}
	__syncthreads();
	const double sq_thr = 1.0; // TruncatedLoss(loss_scale = 1.0)
	// cost of F: sum_k w_k min(r_k^2, thr^2)  (FundamentalJacobianAccumulator::residual)
	auto lm_cost = [&](const double *F) -> double {
		double c = 0.0;
		for (int t = tid; t < n; t += kFitT) {
			const double *q = aos + 4 * (int64_t)idx[beg + t];
			const double x1 = (q[0] - mx1) * r1, y1 = (q[1] - my1) * r1, x2 = (q[2] - mx2) * r2, y2 = (q[3] - my2) * r2;
			const double Fx0 = F[0] * x1 + F[1] * y1 + F[2], Fx1 = F[3] * x1 + F[4] * y1 + F[5], Fx2 = F[6] * x1 + F[7] * y1 + F[8];
			const double Ft0 = F[0] * x2 + F[3] * y2 + F[6], Ft1 = F[1] * x2 + F[4] * y2 + F[7];
			const double C = x2 * Fx0 + y2 * Fx1 + Fx2;
			const double nJ = Fx0 * Fx0 + Fx1 * Fx1 + Ft0 * Ft0 + Ft1 * Ft1;
			const double rr = (C * C) / nJ;
			const double l = fmin(rr, sq_thr); // std::min(r2, squared_thr): NaN r2 gives squared_thr
			c += weights ? weights[t] * l : l;
		}
		return fp_block_sum(c, s_tmp);
	};
	// J^T J (lower, 28) and J^T r (7) of F = compose(U, V, sigma)  (FundamentalJacobianAccumulator::accumulate)
	auto lm_accumulate = [&](const double *F, const double *U, const double *V, double *sys /*35*/) {
		double acc2[35];
#pragma unroll
		for (int a = 0; a < 35; ++a) acc2[a] = 0.0;
		// dF/dparams: d/dw1_k = [e_k]x F, d/dw2_k = -F [e_k]x, d/dsigma = U.col(1) V.col(1)^T
		double dP[7][9];
		for (int c = 0; c < 3; ++c) {
			dP[0][0 + c] = 0.0, dP[0][3 + c] = -F[6 + c], dP[0][6 + c] = F[3 + c];
			dP[1][0 + c] = F[6 + c], dP[1][3 + c] = 0.0, dP[1][6 + c] = -F[0 + c];
			dP[2][0 + c] = -F[3 + c], dP[2][3 + c] = F[0 + c], dP[2][6 + c] = 0.0;
		}
		for (int r = 0; r < 3; ++r) {
			dP[3][r * 3 + 0] = 0.0, dP[3][r * 3 + 1] = -F[r * 3 + 2], dP[3][r * 3 + 2] = F[r * 3 + 1];
			dP[4][r * 3 + 0] = F[r * 3 + 2], dP[4][r * 3 + 1] = 0.0, dP[4][r * 3 + 2] = -F[r * 3 + 0];
			dP[5][r * 3 + 0] = -F[r * 3 + 1], dP[5][r * 3 + 1] = F[r * 3 + 0], dP[5][r * 3 + 2] = 0.0;
			for (int c = 0; c < 3; ++c) dP[6][r * 3 + c] = U[r * 3 + 1] * V[c * 3 + 1];
		}
		for (int t = tid; t < n; t += kFitT) {
			const double *q = aos + 4 * (int64_t)idx[beg + t];
			const double p1[3] = {(q[0] - mx1) * r1, (q[1] - my1) * r1, 1.0}, p2[3] = {(q[2] - mx2) * r2, (q[3] - my2) * r2, 1.0};
			double Fx[3], Ft[3];
			for (int r = 0; r < 3; ++r) Fx[r] = F[r * 3] * p1[0] + F[r * 3 + 1] * p1[1] + F[r * 3 + 2];
			for (int c = 0; c < 3; ++c) Ft[c] = F[c] * p2[0] + F[3 + c] * p2[1] + F[6 + c];
			const double C = p2[0] * Fx[0] + p2[1] * Fx[1] + Fx[2];
			const double nJ = sqrt(Ft[0] * Ft[0] + Ft[1] * Ft[1] + Fx[0] * Fx[0] + Fx[1] * Fx[1]);
			const double inv = 1.0 / nJ, res = C * inv;
			double wgt = ((res * res < sq_thr) ? 1.0 : 0.0) / (double)n; // loss_fn.weight(r^2) / sample_size
			if (weights) wgt = weights[t] * wgt;
			if (wgt == 0.0) continue;
			const double sC = C * inv * inv;
			double G[9]; // d r / d F(r, c)
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
					G[r * 3 + c] = (p2[r] * p1[c] - sC * ((c < 2 ? Ft[c] * p2[r] : 0.0) + (r < 2 ? Fx[r] * p1[c] : 0.0))) * inv;
			double J[7];
#pragma unroll
			for (int k = 0; k < 7; ++k) {
				double v = 0.0;
#pragma unroll
				for (int e = 0; e < 9; ++e) v += G[e] * dP[k][e];
				J[k] = v;
			}
			int a = 0;
#pragma unroll
			for (int i = 0; i < 7; ++i)
#pragma unroll
				for (int j = 0; j <= i; ++j, ++a) acc2[a] += wgt * (J[i] * J[j]);
#pragma unroll
			for (int i = 0; i < 7; ++i) acc2[28 + i] += wgt * res * J[i];
		}
#pragma unroll
		for (int a = 0; a < 35; ++a)
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) acc2[a] += __shfl_xor_sync(0xffffffffu, acc2[a], o);
		__syncthreads();
		if ((tid & 31) == 0)
			for (int a = 0; a < 35; ++a) s_vec[tid >> 5][a] = acc2[a];
		__syncthreads();
		if (tid < 35) {
			double v = 0.0;
			for (int wv = 0; wv < kFitT / 32; ++wv) v += s_vec[wv][tid];
			sys[tid] = v;
		}
		__syncthreads();
	};
	double cost = lm_cost(s_Fc);
	double lambda = 1e-3;
	bool recompute = true;
	for (int iter = 0; iter < 25; ++iter) {
		if (recompute) lm_accumulate(s_Fc, s_U, s_V, s_sys);
		if (tid == 0) {
			int go = 1;
			double g2 = 0.0;
			for (int i = 0; i < 7; ++i) g2 += s_sys[28 + i] * s_sys[28 + i];
			if (recompute && sqrt(g2) < 1e-8) go = 0; // gradient_tol
			double sol[7];
			if (go) { // (J^T J + lambda I) sol = -J^T r by LLT (lower triangle)
				double Lm[7][7];
				int a = 0;
				for (int i = 0; i < 7; ++i)
					for (int j = 0; j <= i; ++j, ++a) Lm[i][j] = s_sys[a] + (i == j ? lambda : 0.0);
				for (int j = 0; j < 7 && go; ++j) {
					double d = Lm[j][j];
					for (int k = 0; k < j; ++k) d -= Lm[j][k] * Lm[j][k];
					if (!(d > 0.0)) {
						go = 0; // not positive definite: Eigen's LLT would return garbage; stop with the current estimate
						break;
					}
					Lm[j][j] = sqrt(d);
					for (int i = j + 1; i < 7; ++i) {
						double v = Lm[i][j];
						for (int k = 0; k < j; ++k) v -= Lm[i][k] * Lm[j][k];
						Lm[i][j] = v / Lm[j][j];
					}
				}
				if (go) {
					double y[7];
					for (int i = 0; i < 7; ++i) {
						double v = s_sys[28 + i];
						for (int k = 0; k < i; ++k) v -= Lm[i][k] * y[k];
						y[i] = v / Lm[i][i];
					}
					for (int i = 6; i >= 0; --i) {
						double v = y[i];
						for (int k = i + 1; k < 7; ++k) v -= Lm[k][i] * sol[k];
						sol[i] = v / Lm[i][i];
					}
					for (int i = 0; i < 7; ++i) sol[i] = -sol[i]; // sol = -(J^T J + lambda I)^-1 J^T r
					double s2 = 0.0;
					for (int i = 0; i < 7; ++i) s2 += sol[i] * sol[i];
					if (sqrt(s2) < 1e-8) go = 0; // step_tol
				}
			}
			if (go) { // U <- U + (a sw + (1 - b) sw^2) U, same for V; sigma += sol(6)
				for (int side = 0; side < 2; ++side) {
					double wv[3] = {sol[3 * side], sol[3 * side + 1], sol[3 * side + 2]};
					const double theta = sqrt(wv[0] * wv[0] + wv[1] * wv[1] + wv[2] * wv[2]);
					for (int i = 0; i < 3; ++i) wv[i] /= theta;
					const double a1 = sin(theta), b1 = cos(theta);
					const double sw[9] = {0, -wv[2], wv[1], wv[2], 0, -wv[0], -wv[1], wv[0], 0};
					double sw2[9], R[9];
					for (int i = 0; i < 3; ++i)
						for (int j = 0; j < 3; ++j) sw2[i * 3 + j] = sw[i * 3] * sw[j] + sw[i * 3 + 1] * sw[3 + j] + sw[i * 3 + 2] * sw[6 + j];
					for (int i = 0; i < 9; ++i) R[i] = a1 * sw[i] + (1 - b1) * sw2[i];
					const double *M = side == 0 ? s_U : s_V;
					double *Mn = side == 0 ? s_Un : s_Vn;
					for (int i = 0; i < 3; ++i)
						for (int j = 0; j < 3; ++j)
							Mn[i * 3 + j] = M[i * 3 + j] + (R[i * 3] * M[j] + R[i * 3 + 1] * M[3 + j] + R[i * 3 + 2] * M[6 + j]);
				}
				s_sigma_n = s_sigma + sol[6];
				compose(s_Un, s_Vn, s_sigma_n, s_Fn);
			}
			s_go = go;
		}
		__syncthreads();
		if (!s_go) break; // block-uniform
		const double cost_new = lm_cost(s_Fn);
		__syncthreads();
		if (cost_new < cost) {
			if (tid == 0) {
				for (int i = 0; i < 9; ++i) s_U[i] = s_Un[i], s_V[i] = s_Vn[i], s_Fc[i] = s_Fn[i];
				s_sigma = s_sigma_n;
			}
			lambda /= 10;
			cost = cost_new;
			recompute = true;
		} else {
			lambda *= 10;
			recompute = false;
		}
		__syncthreads();
	}
	if (tid != 0) return;
	// F = T2^T F_lm T1 with T_i = [r_i 0 -r_i m_x; 0 r_i -r_i m_y; 0 0 1] (fundamental_estimator.h:604-613)
	const double T1[9] = {r1, 0, -r1 * mx1, 0, r1, -r1 * my1, 0, 0, 1};
	const double T2[9] = {r2, 0, -r2 * mx2, 0, r2, -r2 * my2, 0, 0, 1};
	double tmp[9], F[9];
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c) tmp[r * 3 + c] = T2[0 + r] * s_Fc[0 + c] + T2[3 + r] * s_Fc[3 + c] + T2[6 + r] * s_Fc[6 + c];
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
// Thread `tid` of the block visits the sample's points tid, tid + kFitT, ... in that order (the summation topology of
// every block sum below); the rows of U consecutive visits are gathered before the first one is consumed, so that the
// dependent pair of L2 round trips (index -> row) is paid once per U points instead of once per point: with one block of
// 256 threads per problem the per-point loops of a 6000-point refit were latency chains (~1500 cycles per point).
template <int U, class Body>
__device__ __forceinline__ void for_sample_rows5(const double *__restrict__ aos, const int32_t *__restrict__ idx, int beg, int n, int tid,
                                                 Body body) {
	for (int t0 = tid; t0 < n; t0 += U * kFitT) {
		double q[U][5];
#pragma unroll
		for (int u = 0; u < U; ++u) {
			const int t = t0 + u * kFitT;
			if (t < n) {
				const double *src = aos + 5 * (int64_t)idx[beg + t];
#pragma unroll
				for (int c = 0; c < 5; ++c) q[u][c] = src[c];
			}
		}
#pragma unroll
		for (int u = 0; u < U; ++u)
			if (t0 + u * kFitT < n) body(q[u]);
	}
}

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
              double *__restrict__ P_out, int32_t *__restrict__ ok_out, int debug) {
	const long long clk0 = clock64();
	long long clk1 = 0, clk2 = 0, clk3 = 0, clk4 = 0;
	int lm_iters = 0, lm_accepted = 0;
	__shared__ double s_tmp[kFitT / 32];
	__shared__ double s_acc[78];
	__shared__ double s_big[(kFitT / 32) * 78], s_A[144], s_Vv[144], s_w[12];
	__shared__ double s_pose[12], s_new[12], s_sys[28];
	__shared__ int s_ok, s_go;
	const int pb = blockIdx.x, tid = threadIdx.x;
	const int beg = off[pb], n = off[pb + 1] - beg;
	if (n < 6) {
		if (tid == 0) ok_out[pb] = 0;
		return;
	}
	// ---- normalise the 3D points (centroid, mean distance sqrt(3)) ----
	double cx = 0, cy = 0, cz = 0;
	for_sample_rows5<4>(aos, idx, beg, n, tid, [&](const double *q) {
		cx = add(cx, q[2]);
		cy = add(cy, q[3]);
		cz = add(cz, q[4]);
	});
	cx = fp_block_sum(cx, s_tmp) / n;
	cy = fp_block_sum(cy, s_tmp) / n;
	cz = fp_block_sum(cz, s_tmp) / n;
	double md = 0;
	for_sample_rows5<4>(aos, idx, beg, n, tid, [&](const double *q) {
		const double dx = q[2] - cx, dy = q[3] - cy, dz = q[4] - cz;
		md = add(md, sqrt(dx * dx + dy * dy + dz * dz));
	});
	md = fp_block_sum(md, s_tmp) / n;
	const double sc = md > 0 ? 1.7320508075688772 / md : 1.0;
	clk1 = clock64();
	// ---- DLT normal equations (12 x 12, 78 unique) ----
	{
		double acc[78];
#pragma unroll
		for (int a = 0; a < 78; ++a) acc[a] = 0.0;
		for_sample_rows5<2>(aos, idx, beg, n, tid, [&](const double *q) {
			const double u = q[0], v = q[1], X = (q[2] - cx) * sc, Y = (q[3] - cy) * sc, Z = (q[4] - cz) * sc;
			const double ra[12] = {X, Y, Z, 1, 0, 0, 0, 0, -u * X, -u * Y, -u * Z, -u};
			const double rb[12] = {0, 0, 0, 0, X, Y, Z, 1, -v * X, -v * Y, -v * Z, -v};
			int a = 0;
#pragma unroll
			for (int r = 0; r < 12; ++r)
#pragma unroll
				for (int c = r; c < 12; ++c, ++a) acc[a] = add(acc[a], add(mul(ra[r], ra[c]), mul(rb[r], rb[c])));
		});
		fp_block_sums<78>(acc, s_big, s_acc);
	}
	clk2 = clock64();
	if (tid < 32) {
		if (tid == 0) {
			int a = 0;
			for (int r = 0; r < 12; ++r)
				for (int c = r; c < 12; ++c, ++a) s_A[r * 12 + c] = s_A[c * 12 + r] = s_acc[a];
		}
		__syncwarp();
		lm_accepted = -jacobi_eig_warp<12>(s_A, s_Vv, s_w); // (debug line: sweeps, until the LM loop counts)
	}
	__syncthreads();
	clk3 = clock64();
	if (tid == 0) {
		const double *V = s_Vv, *w = s_w;
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
	// ---- PoseLib's refine_pnp = lm_pnp_impl (relative_pose/bundle.cpp:24-100,558-597) with the calibrated camera model,
	// restated: residual r = (R X + t).hnormalized() - x over the points in front of the camera (Z_z >= 0), TRUNCATED loss
	// with loss_scale = 1 (normalised image units) -- cost = sum min(|r|^2, 1), IRLS weight (|r|^2 < 1) --, J = [dZ (-[X]x) | dZ]
	// with dZ = [I | -z] R / Z_z (rotation by right multiplication R <- R exp([w]x), t <- t + R dt; jacobian_impl.h:61-155),
	// LM exactly as for F above: lambda0 = 1e-3 added to the diagonal, LLT, accept iff the cost decreases, tolerances 1e-8,
	// at most 25 iterations. The reference's accumulator iterates over correspondences->rows with sample[i] and reads
	// weights[i] unconditionally (jacobian_impl.h:32-56,76-104) -- out-of-bounds / null reads for a sample shorter than the
	// data; restated as intended: over the sample, unit weights.
	clk4 = clock64();
	const double sq_thr = 1.0;
	auto pnp_cost = [&](const double *P) -> double {
		double c = 0.0;
		double Pr[12]; // (the pose is read once: the gathers below must not wait for shared-memory loads in their shadow)
#pragma unroll
		for (int i = 0; i < 12; ++i) Pr[i] = P[i];
		for_sample_rows5<4>(aos, idx, beg, n, tid, [&](const double *q) {
			const double Zx = Pr[0] * q[2] + Pr[1] * q[3] + Pr[2] * q[4] + Pr[3], Zy = Pr[4] * q[2] + Pr[5] * q[3] + Pr[6] * q[4] + Pr[7],
			             Zz = Pr[8] * q[2] + Pr[9] * q[3] + Pr[10] * q[4] + Pr[11];
			if (Zz < 0) return;
			const double iz = 1.0 / Zz, r0 = Zx * iz - q[0], r1 = Zy * iz - q[1];
			c += fmin(r0 * r0 + r1 * r1, sq_thr);
		});
		return fp_block_sum(c, s_tmp);
	};
	auto pnp_accumulate = [&](const double *P, double *sys /*21 lower + 6*/) {
		double acc[27];
#pragma unroll
		for (int a = 0; a < 27; ++a) acc[a] = 0.0;
		double Pr[12];
#pragma unroll
		for (int i = 0; i < 12; ++i) Pr[i] = P[i];
		for_sample_rows5<4>(aos, idx, beg, n, tid, [&](const double *q) {
			const double *P = Pr;
			const double X[3] = {q[2], q[3], q[4]};
			const double Zx = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[3], Zy = P[4] * X[0] + P[5] * X[1] + P[6] * X[2] + P[7],
			             Zz = P[8] * X[0] + P[9] * X[1] + P[10] * X[2] + P[11];
			if (Zz < 0) return;
			const double iz = 1.0 / Zz, zx = Zx * iz, zy = Zy * iz;
			const double r0 = zx - q[0], r1 = zy - q[1];
			if (!(r0 * r0 + r1 * r1 < sq_thr)) return; // loss_fn.weight == 0
			// dZ = [1 0 -zx; 0 1 -zy] / Zz * R
			double d0[3], d1[3];
			for (int c = 0; c < 3; ++c) {
				d0[c] = (P[c] - zx * P[8 + c]) * iz;
				d1[c] = (P[4 + c] - zy * P[8 + c]) * iz;
			}
			// columns of -dZ [X]x: X1 dZ(:,2) - X2 dZ(:,1), X2 dZ(:,0) - X0 dZ(:,2), X0 dZ(:,1) - X1 dZ(:,0)
			const double J0[6] = {X[1] * d0[2] - X[2] * d0[1], X[2] * d0[0] - X[0] * d0[2], X[0] * d0[1] - X[1] * d0[0], d0[0], d0[1], d0[2]};
			const double J1[6] = {X[1] * d1[2] - X[2] * d1[1], X[2] * d1[0] - X[0] * d1[2], X[0] * d1[1] - X[1] * d1[0], d1[0], d1[1], d1[2]};
			int a = 0;
#pragma unroll
			for (int i = 0; i < 6; ++i)
#pragma unroll
				for (int j = 0; j <= i; ++j, ++a) acc[a] += J0[i] * J0[j] + J1[i] * J1[j];
#pragma unroll
			for (int i = 0; i < 6; ++i) acc[21 + i] += J0[i] * r0 + J1[i] * r1;
		});
		fp_block_sums<27>(acc, s_big, sys);
	};
	const int sweeps = -lm_accepted;
	lm_accepted = 0;
	double cost = pnp_cost(s_pose);
	double lambda = 1e-3;
	bool recompute = true;
	for (int iter = 0; iter < 25; ++iter) {
		if (recompute) pnp_accumulate(s_pose, s_sys);
		if (tid == 0) {
			int go = 1;
			double g2 = 0.0;
			for (int i = 0; i < 6; ++i) g2 += s_sys[21 + i] * s_sys[21 + i];
			if (recompute && sqrt(g2) < 1e-8) go = 0; // gradient_tol
			double sol[6];
			if (go) {
				double Lm[6][6];
				int a = 0;
				for (int i = 0; i < 6; ++i)
					for (int j = 0; j <= i; ++j, ++a) Lm[i][j] = s_sys[a] + (i == j ? lambda : 0.0);
				for (int j = 0; j < 6 && go; ++j) {
					double d = Lm[j][j];
					for (int k = 0; k < j; ++k) d -= Lm[j][k] * Lm[j][k];
					if (!(d > 0.0)) {
						go = 0;
						break;
					}
					Lm[j][j] = sqrt(d);
					for (int i = j + 1; i < 6; ++i) {
						double v = Lm[i][j];
						for (int k = 0; k < j; ++k) v -= Lm[i][k] * Lm[j][k];
						Lm[i][j] = v / Lm[j][j];
					}
				}
				if (go) {
					double y[6];
					for (int i = 0; i < 6; ++i) {
						double v = s_sys[21 + i];
						for (int k = 0; k < i; ++k) v -= Lm[i][k] * y[k];
						y[i] = v / Lm[i][i];
					}
					for (int i = 5; i >= 0; --i) {
						double v = y[i];
						for (int k = i + 1; k < 6; ++k) v -= Lm[k][i] * sol[k];
						sol[i] = v / Lm[i][i];
					}
					for (int i = 0; i < 6; ++i) sol[i] = -sol[i]; // sol = -(J^T J + lambda I)^-1 J^T r
					double s2 = 0.0;
					for (int i = 0; i < 6; ++i) s2 += sol[i] * sol[i];
					if (sqrt(s2) < 1e-8) go = 0; // step_tol
				}
			}
			if (go) { // R <- R + R (a sw + (1 - b) sw^2), t <- t + R sol(3:6)
				double wv[3] = {sol[0], sol[1], sol[2]};
				const double theta = sqrt(wv[0] * wv[0] + wv[1] * wv[1] + wv[2] * wv[2]);
				for (int i = 0; i < 3; ++i) wv[i] /= theta;
				const double a1 = sin(theta), b1 = cos(theta);
				const double sw[9] = {0, -wv[2], wv[1], wv[2], 0, -wv[0], -wv[1], wv[0], 0};
				double E[9];
				for (int i = 0; i < 3; ++i)
					for (int j = 0; j < 3; ++j)
						E[i * 3 + j] = a1 * sw[i * 3 + j] + (1 - b1) * (sw[i * 3] * sw[j] + sw[i * 3 + 1] * sw[3 + j] + sw[i * 3 + 2] * sw[6 + j]);
				for (int r = 0; r < 3; ++r) {
					const double R0 = s_pose[4 * r], R1 = s_pose[4 * r + 1], R2 = s_pose[4 * r + 2];
					for (int c = 0; c < 3; ++c) s_new[4 * r + c] = s_pose[4 * r + c] + (R0 * E[c] + R1 * E[3 + c] + R2 * E[6 + c]);
					s_new[4 * r + 3] = s_pose[4 * r + 3] + (R0 * sol[3] + R1 * sol[4] + R2 * sol[5]);
				}
			}
			s_go = go;
		}
		__syncthreads();
		if (!s_go) break; // block-uniform
		++lm_iters;
		const double cost_new = pnp_cost(s_new);
		__syncthreads();
		if (cost_new < cost) {
			++lm_accepted;
			if (tid < 12) s_pose[tid] = s_new[tid];
			lambda /= 10;
			cost = cost_new;
			recompute = true;
		} else {
			lambda *= 10;
			recompute = false;
		}
		__syncthreads();
	}
	if (tid == 0) {
		bool ok = true;
		for (int i = 0; i < 12; ++i) {
			ok = ok && (fabs(s_pose[i]) <= DBL_MAX);
			P_out[12 * (int64_t)pb + i] = s_pose[i];
		}
		ok_out[pb] = ok ? 1 : 0;
		if (debug && pb == 0)
			printf("[k_fit_pnp] n=%d cycles: normalise %lld, DLT sums %lld, Jacobi %lld (%d sweeps), seed %lld, LM %lld (%d steps, %d accepted)\n", n,
			       clk1 - clk0, clk2 - clk1, clk3 - clk2, sweeps, clk4 - clk3, clock64() - clk4, lm_iters, lm_accepted);
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
	static const int debug = getenv("PXB_FIT_STATS") ? atoi(getenv("PXB_FIT_STATS")) : 0;
	k_fit_pnp<<<(unsigned)P, kFitT, 0, ctx->stream>>>(ctx->pts.aos, off, idx, P_out, ok_out, debug);
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
