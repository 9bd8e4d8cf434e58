// pxb_fit_vl.cu -- batched non-minimal fits of the vanishing-point and 2D-line families.
//
//   k_fit_vp    VanishingPointTwoLineSolver::estimateModel, non-minimal branch
//               (px/include/solver_vanishing_point_two_lines.h:187-233): rows [y0 - my, mx - x0, x0 my - y0 mx] * w,
//               eigenvector of A^T A (3x3) with the smallest eigenvalue, normalised. The reference indexes the weights
//               BY POINT when a sample is given (weights_[sample_[i]], :203).
//   k_fit_line  LinearModelEstimator<.., 2>::estimateModelNonminimal (gcr/estimators/linear_model_estimator.h:152-250:
//               mass point, mean distance, sqrt(2)/mean) + LinearModelSolver<2>::estimateModel
//               (solver_linear_model.h:198-239: C^T C, FullPivHouseholderQR, last column of Q) + w = -mass . n.
//
// One block per problem; sums use the block topology of the other fit kernels (thread partial in index order -> xor
// butterfly -> warps in order). The 3x3 symmetric eigenproblem is solved by cyclic Jacobi (Eigen's tridiagonal QL is
// not reproduced: eigenvector to ~1e-13, sign fixed so that the largest-magnitude component is positive); the 2x2
// full-pivot Householder step is restated exactly (oracle/pxo_oracle.cpp has the same restatement).
#include "pxb_internal.h"
#include "pxb_residuals.cuh"

namespace pxb {

constexpr int kVlThreads = 256;

__device__ __forceinline__ double vl_block_sum(double x, double *s_tmp /*kVlThreads/32*/) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) x = add(x, __shfl_xor_sync(0xffffffffu, x, o));
	__syncthreads();
	if ((threadIdx.x & 31) == 0) s_tmp[threadIdx.x >> 5] = x;
	__syncthreads();
	double t = 0.0;
#pragma unroll
	for (int w = 0; w < kVlThreads / 32; ++w) t = add(t, s_tmp[w]);
	return t;
}

__device__ void jacobi_eig3(double A[3][3], double V[3][3], double w[3]) {
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) V[i][j] = i == j;
	for (int sweep = 0; sweep < 60; ++sweep) {
		const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
		if (off == 0.0) break;
		for (int p = 0; p < 2; ++p)
			for (int q = p + 1; q < 3; ++q) {
				if (A[p][q] == 0.0) continue;
				const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
				const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
				const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
				for (int k = 0; k < 3; ++k) {
					const double akp = A[k][p], akq = A[k][q];
					A[k][p] = c * akp - sn * akq;
					A[k][q] = sn * akp + c * akq;
				}
				for (int k = 0; k < 3; ++k) {
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

__global__ void __launch_bounds__(kVlThreads)
    k_fit_vp(const double *__restrict__ aos, const int32_t *__restrict__ off, const int32_t *__restrict__ idx,
             const double *__restrict__ weights_by_point, double *__restrict__ out, int32_t *__restrict__ ok_out) {
	__shared__ double s_tmp[kVlThreads / 32];
	const int pb = blockIdx.x, beg = off[pb], n = off[pb + 1] - beg, tid = threadIdx.x;
	if (n < 2) { // sample_number_ < nonMinimalSampleSize() (vanishing_point_estimator.h:211-212)
		if (tid == 0) ok_out[pb] = 0;
		return;
	}
	double m[6] = {0, 0, 0, 0, 0, 0}; // xx xy xz yy yz zz of A^T A
	for (int t = tid; t < n; t += kVlThreads) {
		const int64_t i = idx[beg + t];
		const double *p = aos + 4 * i;
		const double w = weights_by_point ? weights_by_point[i] : 1.0;
		const double x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
		const double mx = mul(add(x0, x1), 0.5), my = mul(add(y0, y1), 0.5);
		const double r0 = mul(sub(mul(y0, 1.0), my), w), r1 = mul(sub(mx, mul(x0, 1.0)), w),
		             r2 = mul(sub(mul(x0, my), mul(y0, mx)), w);
		m[0] = add(m[0], mul(r0, r0));
		m[1] = add(m[1], mul(r0, r1));
		m[2] = add(m[2], mul(r0, r2));
		m[3] = add(m[3], mul(r1, r1));
		m[4] = add(m[4], mul(r1, r2));
		m[5] = add(m[5], mul(r2, r2));
	}
	double s[6];
	for (int k = 0; k < 6; ++k) s[k] = vl_block_sum(m[k], s_tmp);
	if (tid != 0) return;
	double A[3][3] = {{s[0], s[1], s[2]}, {s[1], s[3], s[4]}, {s[2], s[4], s[5]}}, V[3][3], w3[3];
	jacobi_eig3(A, V, w3);
	int k = 0;
	for (int i = 1; i < 3; ++i)
		if (w3[i] < w3[k]) k = i;
	double e[3] = {V[0][k], V[1][k], V[2][k]};
	const double len = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
	int big = 0;
	for (int i = 1; i < 3; ++i)
		if (fabs(e[i]) > fabs(e[big])) big = i;
	const double sgn = e[big] < 0 ? -1.0 : 1.0;
	bool bad = !(len > 0.0);
	for (int i = 0; i < 3; ++i) {
		const double v = sgn * e[i] / len;
		out[3 * (int64_t)pb + i] = v;
		bad |= !(fabs(v) <= 1e300);
	}
	ok_out[pb] = bad ? 0 : 1; // the reference pushes the model unconditionally; a NaN model never scores
}

__global__ void __launch_bounds__(kVlThreads)
    k_fit_line(const double *__restrict__ aos, const int32_t *__restrict__ off, const int32_t *__restrict__ idx,
               double *__restrict__ out, int32_t *__restrict__ ok_out) {
	__shared__ double s_tmp[kVlThreads / 32];
	const int pb = blockIdx.x, beg = off[pb], n = off[pb + 1] - beg, tid = threadIdx.x;
	if (n < 2) {
		if (tid == 0) ok_out[pb] = 0;
		return;
	}
	double sx = 0, sy = 0;
	for (int t = tid; t < n; t += kVlThreads) {
		const double *p = aos + 2 * (int64_t)idx[beg + t];
		sx = add(sx, p[0]);
		sy = add(sy, p[1]);
	}
	const double mx = divd(vl_block_sum(sx, s_tmp), (double)n), my = divd(vl_block_sum(sy, s_tmp), (double)n);
	double sd = 0;
	for (int t = tid; t < n; t += kVlThreads) {
		const double *p = aos + 2 * (int64_t)idx[beg + t];
		const double dx = sub(p[0], mx), dy = sub(p[1], my);
		sd = add(sd, __dsqrt_rn(add(mul(dx, dx), mul(dy, dy))));
	}
	const double avg = divd(vl_block_sum(sd, s_tmp), (double)n);
	const double ratio = divd(__dsqrt_rn(2.0), avg);
	double a = 0, b = 0, c = 0;
	for (int t = tid; t < n; t += kVlThreads) {
		const double *p = aos + 2 * (int64_t)idx[beg + t];
		const double dx = mul(sub(p[0], mx), ratio), dy = mul(sub(p[1], my), ratio);
		a = add(a, mul(dx, dx));
		b = add(b, mul(dx, dy));
		c = add(c, mul(dy, dy));
	}
	a = vl_block_sum(a, s_tmp);
	b = vl_block_sum(b, s_tmp);
	c = vl_block_sum(c, s_tmp);
	if (tid != 0) return;
	// Eigen FullPivHouseholderQR of [[a, b], [b, c]]: pivot = first maximum of |entry| in column-major order
	const double Mx[2][2] = {{a, b}, {b, c}};
	int pr = 0, pc = 0;
	double best = fabs(Mx[0][0]);
	for (int col = 0; col < 2; ++col)
		for (int row = 0; row < 2; ++row)
			if (fabs(Mx[row][col]) > best) best = fabs(Mx[row][col]), pr = row, pc = col;
	if (best == 0.0 || !(best <= 1e300)) {
		ok_out[pb] = 0;
		return;
	}
	const double x0 = Mx[pr][pc], x1 = Mx[1 - pr][pc];
	double q0, q1; // last column of H = I - tau v v^T, v = (1, ess)
	if (x1 == 0.0) {
		q0 = 0.0, q1 = 1.0;
	} else {
		double beta = sqrt(x0 * x0 + x1 * x1);
		if (x0 >= 0) beta = -beta;
		const double ess = x1 / (x0 - beta), tau = (beta - x0) / beta;
		q0 = -tau * ess;
		q1 = 1.0 - tau * ess * ess;
	}
	if (pr == 1) { // Q = P_rows H
		const double t = q0;
		q0 = q1;
		q1 = t;
	}
	const double len = sqrt(q0 * q0 + q1 * q1);
	const double nx = q0 / len, ny = q1 / len;
	out[3 * (int64_t)pb] = nx;
	out[3 * (int64_t)pb + 1] = ny;
	out[3 * (int64_t)pb + 2] = -mx * nx - my * ny; // linear_model_estimator.h:176-183
	ok_out[pb] = (fabs(nx) <= 1e300 && fabs(ny) <= 1e300) ? 1 : 0;
}

int launch_fit_vp(pxb_ctx *ctx, int P, const int32_t *off, const int32_t *idx, const double *weights_by_point, double *out,
                  int32_t *ok_out) {
	if (P <= 0) return PXB_OK;
	k_fit_vp<<<(unsigned)P, kVlThreads, 0, ctx->stream>>>(ctx->pts.aos, off, idx, weights_by_point, out, ok_out);
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

int launch_fit_line(pxb_ctx *ctx, int P, const int32_t *off, const int32_t *idx, double *out, int32_t *ok_out) {
	if (P <= 0) return PXB_OK;
	k_fit_line<<<(unsigned)P, kVlThreads, 0, ctx->stream>>>(ctx->pts.aos, off, idx, out, ok_out);
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

} // namespace pxb
