// pxb_fit.cu -- the two "next" rows the end-to-end driver cannot live without:
//
//   k_knn_graph   f-3  neighbourhood graph: exact radius search truncated to the k nearest, brute force over all
//                      pairs (N^2 float64 distance evaluations: 2.5e9 at N = 50k, a few ms). Replaces the reference's
//                      randomised FLANN kd-trees (gcr/neighborhood/flann_neighborhood_graph.h:100-139), which return
//                      ~5 approximate neighbours per point and differ between two identical calls; this one is
//                      deterministic (ties broken by index) and yields directed CSR lists.
//   k_fit_h       f-1  batched non-minimal homography fits: Hartley normalisation (gcr/estimators/
//                      homography_estimator.h:201-309) + the 2n x 8 inhomogeneous DLT of
//                      HomographyFourPointSolver::estimateNonMinimalModel (solver_homography_four_point.h:192-264),
//                      solved through the 8x8 normal equations (the reference calls Eigen's colPivHouseholderQr on the
//                      2n x 8 system: same least-squares solution, agreement ~1e-9 relative after normalisation).
//                      One block per problem; A^T A / A^T b are block reductions with the library's fixed topology.
#include <cfloat>
#include <cstdlib>

#include "pxb_internal.h"
#include "pxb_residuals.cuh"

namespace pxb {

// ------------------------------------------------------------------------------------------------
// neighbourhood graph
// ------------------------------------------------------------------------------------------------
constexpr int kKnnMax = 16;
constexpr int kKnnQuery = 64; // query points per block: N = 10^4 gives 157 blocks, one per SM

// Brute force over all pairs, float64 like the reference's distances. A block serves 64 query points; the candidates of
// a query are split into SLICES contiguous index ranges scanned by SLICES threads (64 x SLICES threads per block: a thread
// per query left every SM with two warps of dependent FP64 chains -- 0.92 ms at N = 10^4). A warp is 32 consecutive
// queries of ONE slice, so every candidate is a uniform (broadcast) load. The k best candidates of a thread live in
// registers: arrays of a compile-time size with statically indexed, fully unrolled compare-and-swap insertion
// (dynamically indexed arrays go to local memory; with ~16 % of all points inside a 200 px ball the insertion path is
// hot). Candidates are visited in ascending index order and inserted with strict comparisons, and the slices' lists are
// merged in slice order the same way: the result is the k smallest (distance, index) pairs -- ties go to the lower
// index -- exactly what one sequential scan over all candidates returns.
template <int KMAX>
__device__ __forceinline__ void knn_insert(double (&bd)[KMAX], int (&bi)[KMAX], double &worst, double cd, int ci, int k) {
	// the candidate goes in front of the first strictly larger entry; everything behind it moves down one slot (an entry
	// that has been displaced is never compared again: it would leapfrog an equal neighbour and break the tie rule)
	bool placed = false;
#pragma unroll
	for (int q = 0; q < KMAX; ++q) {
		if (q < k && (placed || cd < bd[q])) {
			const double td = bd[q];
			const int ti = bi[q];
			bd[q] = cd;
			bi[q] = ci;
			cd = td;
			ci = ti;
			placed = true;
		}
		if (q == k - 1) worst = bd[q]; // (k == KMAX in the instantiation the driver's default degree uses: a static index --
		                               //  a runtime k makes ptxas shadow the list in local memory for this one read)
	}
}

template <int DIM> __device__ __forceinline__ double knn_dist2(const double (&me)[DIM], const double *__restrict__ pj) {
	double c[DIM];
	if (DIM == 4) {
		const double2 a = __ldg(reinterpret_cast<const double2 *>(pj)), b = __ldg(reinterpret_cast<const double2 *>(pj) + 1);
		c[0] = a.x, c[1] = a.y, c[2 % DIM] = b.x, c[3 % DIM] = b.y;
	} else if (DIM == 2) {
		const double2 a = __ldg(reinterpret_cast<const double2 *>(pj));
		c[0] = a.x, c[1] = a.y;
	} else {
#pragma unroll
		for (int q = 0; q < DIM; ++q) c[q] = __ldg(pj + q);
	}
	double d2 = 0.0;
#pragma unroll
	for (int q = 0; q < DIM; ++q) {
		const double d = me[q] - c[q];
		d2 += d * d;
	}
	return d2;
}

template <int DIM, int KMAX, int SLICES>
__global__ void __launch_bounds__(kKnnQuery *SLICES)
    k_knn_graph(const double *__restrict__ aos, int64_t N, double radius2, int k, int32_t *__restrict__ nbr /*N*k*/,
                int32_t *__restrict__ deg) {
	__shared__ double s_d[SLICES - 1][KMAX][kKnnQuery];
	__shared__ int s_i[SLICES - 1][KMAX][kKnnQuery];
	if (KMAX == 5) k = 5; // (compile-time list length: see launch_knn_dim)
	const int q = threadIdx.x % kKnnQuery, slice = threadIdx.x / kKnnQuery;
	const int64_t i = (int64_t)blockIdx.x * kKnnQuery + q;
	double me[DIM];
#pragma unroll
	for (int c = 0; c < DIM; ++c) me[c] = (i < N) ? aos[i * DIM + c] : 0.0;
	double bd[KMAX];
	int bi[KMAX];
#pragma unroll
	for (int r = 0; r < KMAX; ++r) {
		bd[r] = DBL_MAX;
		bi[r] = -1;
	}
	double worst = DBL_MAX; // bd[k - 1]
	const int64_t chunk = (N + SLICES - 1) / SLICES;
	const int64_t j_begin = min(N, slice * chunk), j_end = min(N, j_begin + chunk);
	if (i < N) {
		int64_t j = j_begin;
		for (; j + 4 <= j_end; j += 4) { // four independent distance chains, then the (rare) insertions in index order
			double d2[4];
#pragma unroll
			for (int u = 0; u < 4; ++u) d2[u] = knn_dist2<DIM>(me, aos + (j + u) * DIM);
#pragma unroll
			for (int u = 0; u < 4; ++u)
				if (j + u != i && d2[u] <= radius2 && d2[u] < worst) knn_insert<KMAX>(bd, bi, worst, d2[u], (int)(j + u), k);
		}
		for (; j < j_end; ++j) {
			const double d2 = knn_dist2<DIM>(me, aos + j * DIM);
			if (j != i && d2 <= radius2 && d2 < worst) knn_insert<KMAX>(bd, bi, worst, d2, (int)j, k);
		}
	}
	if (slice > 0) {
#pragma unroll
		for (int r = 0; r < KMAX; ++r) {
			s_d[slice - 1][r][q] = bd[r];
			s_i[slice - 1][r][q] = bi[r];
		}
	}
	__syncthreads();
	if (slice == 0 && i < N) {
		for (int s2 = 0; s2 < SLICES - 1; ++s2)
			for (int r = 0; r < k; ++r) { // a slice's list is ascending: once an entry fails, the rest of the list fails too
				const double cd = s_d[s2][r][q];
				const int ci = s_i[s2][r][q];
				if (ci < 0 || !(cd < worst)) break;
				knn_insert<KMAX>(bd, bi, worst, cd, ci, k);
			}
		int cnt = 0;
#pragma unroll
		for (int r = 0; r < KMAX; ++r)
			if (r < k) {
				nbr[i * k + r] = bi[r];
				cnt += bi[r] >= 0;
			}
		deg[i] = cnt;
	}
}

template <int DIM>
static void launch_knn_dim(pxb_ctx *ctx, unsigned grid, double radius, int k, int32_t *nbr, int32_t *deg) {
	const Points &p = ctx->pts;
	if (k == 5) // the driver's default degree (graph_degree()): list length known at compile time
		k_knn_graph<DIM, 5, 8><<<grid, kKnnQuery * 8, 0, ctx->stream>>>(p.aos, p.N, radius * radius, 5, nbr, deg);
	else if (k <= 8)
		k_knn_graph<DIM, 8, 8><<<grid, kKnnQuery * 8, 0, ctx->stream>>>(p.aos, p.N, radius * radius, k, nbr, deg);
	else
		k_knn_graph<DIM, kKnnMax, 4><<<grid, kKnnQuery * 4, 0, ctx->stream>>>(p.aos, p.N, radius * radius, k, nbr, deg);
}

int launch_knn_graph(pxb_ctx *ctx, double radius, int k, int32_t *nbr, int32_t *deg) {
	const Points &p = ctx->pts;
	if (k < 1 || k > kKnnMax) {
		set_error("k must be in [1, %d]", kKnnMax);
		return PXB_ERR_ARGUMENT;
	}
	const unsigned grid = (unsigned)((p.N + kKnnQuery - 1) / kKnnQuery);
	if (p.dim == 4)
		launch_knn_dim<4>(ctx, grid, radius, k, nbr, deg);
	else if (p.dim == 2)
		launch_knn_dim<2>(ctx, grid, radius, k, nbr, deg);
	else
		launch_knn_dim<5>(ctx, grid, radius, k, nbr, deg);
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

// ------------------------------------------------------------------------------------------------
// batched non-minimal homography fit
// ------------------------------------------------------------------------------------------------
constexpr int kFitThreads = 256;

__device__ __forceinline__ double fit_block_sum(double x, double *s_tmp /*8*/) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) x = add(x, __shfl_xor_sync(0xffffffffu, x, o));
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	__syncthreads();
	if (lane == 0) s_tmp[warp] = x;
	__syncthreads();
	double t = 0.0;
#pragma unroll
	for (int w = 0; w < kFitThreads / 32; ++w) t = add(t, s_tmp[w]);
	return t;
}

// NV block sums at once with the topology of fit_block_sum (butterfly inside the warp, then the warps in order): two
// barriers for the whole vector instead of two per value. Every thread receives the sums in v[].
template <int NV> __device__ __forceinline__ void fit_block_sum_vec(double (&v)[NV], double *s_vec /*[warps][NV]*/, double *s_out /*[NV]*/) {
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
		for (int w = 0; w < kFitThreads / 32; ++w) t = add(t, s_vec[w * NV + threadIdx.x]);
		s_out[threadIdx.x] = t;
	}
	__syncthreads();
#pragma unroll
	for (int a = 0; a < NV; ++a) v[a] = s_out[a];
}

// problems: CSR (off[P+1], idx[]) of point indices; weights: NULL, or an array indexed BY ROW OF THE NORMALISED SAMPLE
// (the reference passes weights_[i], i = 0..n-1, once the sample has been gathered into `normalized_points` with a null
// sample pointer -- solver_homography_four_point.h:207-220 with sample_ == nullptr -- i.e. the first n entries of the
// caller's per-point weight array; replicated as is).
// FP64 tensor-core form of the normal equations (MMA = true): the 2n x 8 design matrix is walked in chunks of 4 rows (two
// points); for a chunk Mc (4 x 8) one DMMA m8n8k4 adds Mc^T Mc to the 8 x 8 accumulator -- the A operand (8 x 4, row
// major) and the B operand (4 x 8, column major) of that shape are THE SAME register per lane (lane l holds Mc[l % 4][l / 4])
// -- and a second one adds Mc^T [b | 0] (column 0 = A^T b). Warps take chunks round robin; their accumulators are added
// in warp order. Same sums as the scalar form up to the order of additions (1e-12 relative on the fitted H).
__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
	             : "+d"(d0), "+d"(d1)
	             : "d"(a), "d"(b));
}

template <bool MMA>
__global__ void __launch_bounds__(kFitThreads)
    k_fit_h(const double *__restrict__ aos, const int32_t *__restrict__ off, const int32_t *__restrict__ idx,
            const double *__restrict__ weights, double *__restrict__ H_out, int32_t *__restrict__ ok_out) {
	__shared__ double s_acc[44];
	pdl_launch_dependents();
	pdl_wait();
	const int pb = blockIdx.x;
	const int beg = off[pb], n = off[pb + 1] - beg;
	const int tid = threadIdx.x;
	if (n < 4) { // estimateModelNonminimal: sample_number_ < nonMinimalSampleSize() -> false
		if (tid == 0) ok_out[pb] = 0;
		return;
	}
	// ---- normalizePoints (homography_estimator.h:201-309): mass points, mean distance, sqrt(2)/mean ----
	double sx1 = 0, sy1 = 0, sx2 = 0, sy2 = 0;
	for (int t = tid; t < n; t += kFitThreads) {
		const double *q = aos + 4 * (int64_t)idx[beg + t];
		sx1 = add(sx1, q[0]);
		sy1 = add(sy1, q[1]);
		sx2 = add(sx2, q[2]);
		sy2 = add(sy2, q[3]);
	}
	__shared__ double s_vec[(kFitThreads / 32) * 72]; // 44 sums per warp (scalar form) or a warp's 8 x 8 + 8 accumulators (MMA form)
	double sums4[4] = {sx1, sy1, sx2, sy2};
	fit_block_sum_vec<4>(sums4, s_vec, s_acc);
	const double mx1 = divd(sums4[0], (double)n), my1 = divd(sums4[1], (double)n);
	const double mx2 = divd(sums4[2], (double)n), my2 = divd(sums4[3], (double)n);
	double d1 = 0, d2 = 0;
	for (int t = tid; t < n; t += kFitThreads) {
		const double *q = aos + 4 * (int64_t)idx[beg + t];
		const double dx1 = sub(mx1, q[0]), dy1 = sub(my1, q[1]), dx2 = sub(mx2, q[2]), dy2 = sub(my2, q[3]);
		d1 = add(d1, __dsqrt_rn(add(mul(dx1, dx1), mul(dy1, dy1))));
		d2 = add(d2, __dsqrt_rn(add(mul(dx2, dx2), mul(dy2, dy2))));
	}
	double sums2[2] = {d1, d2};
	fit_block_sum_vec<2>(sums2, s_vec, s_acc);
	const double avg1 = divd(sums2[0], (double)n), avg2 = divd(sums2[1], (double)n);
	const double r1 = divd(1.4142135623730951, avg1), r2 = divd(1.4142135623730951, avg2); // M_SQRT2 / mean distance
	// ---- A^T A (upper triangle, 36) and A^T b (8) of the 2n x 8 system (solver_homography_four_point.h:207-252)
	if (MMA) {
		const int lane = tid & 31, warp = tid >> 5;
		const int krow = lane & 3, jcol = lane >> 2; // this lane's entry of a chunk: design row krow, coefficient jcol
		double g0 = 0.0, g1 = 0.0, b0 = 0.0, b1 = 0.0;
		const int chunks = (n + 1) / 2;
		for (int c = warp; c < chunks; c += kFitThreads / 32) {
			const int t = 2 * c + (krow >> 1);
			double a = 0.0, bv = 0.0; // rows past the end are zero rows
			if (t < n) {
				const double *q = aos + 4 * (int64_t)idx[beg + t];
				const double x1 = mul(sub(q[0], mx1), r1), y1 = mul(sub(q[1], my1), r1);
				const double x2 = mul(sub(q[2], mx2), r2), y2 = mul(sub(q[3], my2), r2);
				const double w = weights ? weights[t] : 1.0;
				const bool second = krow & 1; // equation of y2 (else x2)
				const double tgt = second ? y2 : x2;
				switch (jcol) {
				case 0: a = second ? 0.0 : mul(-w, x1); break;
				case 1: a = second ? 0.0 : mul(-w, y1); break;
				case 2: a = second ? 0.0 : -w; break;
				case 3: a = second ? mul(-w, x1) : 0.0; break;
				case 4: a = second ? mul(-w, y1) : 0.0; break;
				case 5: a = second ? -w : 0.0; break;
				case 6: a = mul(mul(w, tgt), x1); break;
				default: a = mul(mul(w, tgt), y1); break;
				}
				bv = jcol == 0 ? -mul(w, tgt) : 0.0; // B' = [b | 0]: only column 0 carries the right-hand side
			}
			dmma_m8n8k4(g0, g1, a, a);
			dmma_m8n8k4(b0, b1, a, bv);
		}
		// accumulators of the 8 warps, added in warp order: lane holds G[lane / 4][2 (lane % 4) + {0, 1}] and, for lane % 4 == 0,
		// (A^T b)[lane / 4] in b0
		double *s_g = s_vec; // [warps][72]: 64 entries of G + 8 of A^T b
		s_g[warp * 72 + (lane >> 2) * 8 + 2 * (lane & 3)] = g0;
		s_g[warp * 72 + (lane >> 2) * 8 + 2 * (lane & 3) + 1] = g1;
		if ((lane & 3) == 0) s_g[warp * 72 + 64 + (lane >> 2)] = b0;
		__syncthreads();
		if (tid < 44) {
			int r = 0, c = 0, a = 0; // entry tid of the packed layout: upper triangle row by row, then A^T b
			bool found = false;
			if (tid >= 36) {
				r = tid - 36;
			} else {
				for (r = 0; r < 8 && !found; ++r)
					for (c = r; c < 8; ++c, ++a)
						if (a == tid) {
							found = true;
							break;
						}
				--r;
			}
			double v = 0.0;
			for (int w = 0; w < kFitThreads / 32; ++w) v = add(v, tid >= 36 ? s_g[w * 72 + 64 + r] : s_g[w * 72 + r * 8 + c]);
			s_acc[tid] = v;
		}
		__syncthreads();
	} else {
	double acc[44];
#pragma unroll
	for (int a = 0; a < 44; ++a) acc[a] = 0.0;
	for (int t = tid; t < n; t += kFitThreads) {
		const double *q = aos + 4 * (int64_t)idx[beg + t];
		const double x1 = mul(sub(q[0], mx1), r1), y1 = mul(sub(q[1], my1), r1);
		const double x2 = mul(sub(q[2], mx2), r2), y2 = mul(sub(q[3], my2), r2);
		const double w = weights ? weights[t] : 1.0;
		const double mwx1 = mul(-w, x1), mwy1 = mul(-w, y1), wx2 = mul(w, x2), wy2 = mul(w, y2);
		double ra[8] = {mwx1, mwy1, -w, 0, 0, 0, mul(wx2, x1), mul(wx2, y1)};
		double rb[8] = {0, 0, 0, mwx1, mwy1, -w, mul(wy2, x1), mul(wy2, y1)};
		const double ba = -wx2, bb = -wy2;
		int a = 0;
#pragma unroll
		for (int r = 0; r < 8; ++r)
#pragma unroll
			for (int c = r; c < 8; ++c, ++a) acc[a] = add(acc[a], add(mul(ra[r], ra[c]), mul(rb[r], rb[c])));
#pragma unroll
		for (int r = 0; r < 8; ++r) acc[36 + r] = add(acc[36 + r], add(mul(ra[r], ba), mul(rb[r], bb)));
	}
	fit_block_sum_vec<44>(acc, s_vec, s_acc);
	}
	if (tid >= 32) return;
	// ---- solve the 8x8 SPD system: Gaussian elimination with partial pivoting (robust to semi-definite input).
	// Lane r < 8 holds row r of [A^T A | A^T b] in registers; pivot rows travel by shuffle. Every element sees exactly the
	// operations of the sequential elimination (same pivot choice: the first row holding the largest |entry|), so the
	// result is bit-identical to it -- only the rows are processed side by side.
	const int lane = tid;
	double row[9];
#pragma unroll
	for (int c = 0; c < 8; ++c) {
		const int r0 = lane < c ? lane : c, c0 = lane < c ? c : lane; // upper-triangle index of (lane, c)
		row[c] = lane < 8 ? s_acc[r0 * 8 - r0 * (r0 - 1) / 2 + (c0 - r0)] : 0.0;
	}
	row[8] = lane < 8 ? s_acc[36 + lane] : 0.0;
	bool singular = false;
#pragma unroll
	for (int c = 0; c < 8; ++c) {
		if (singular) break; // warp-uniform
		const double mine = fabs(row[c]);
		double best = __shfl_sync(0xffffffffu, mine, c);
		int piv = c;
#pragma unroll
		for (int r = c + 1; r < 8; ++r) {
			const double v = __shfl_sync(0xffffffffu, mine, r);
			if (v > best) {
				best = v;
				piv = r;
			}
		}
		if (!(best > 0.0)) {
			singular = true;
			break;
		}
		const int src = lane == c ? piv : (lane == piv ? c : lane); // swap rows c and piv
		double pc[9];
#pragma unroll
		for (int j = 0; j < 9; ++j) {
			row[j] = __shfl_sync(0xffffffffu, row[j], src);
			pc[j] = __shfl_sync(0xffffffffu, row[j], c);
		}
		if (lane > c && lane < 8) {
			const double f = divd(row[c], pc[c]);
#pragma unroll
			for (int j = 0; j < 9; ++j)
				if (j >= c) row[j] = sub(row[j], mul(f, pc[j]));
		}
	}
	double h[9];
#pragma unroll
	for (int r = 0; r < 9; ++r) h[r] = 0.0;
	if (!singular) {
#pragma unroll
		for (int r = 7; r >= 0; --r) {
			double v = row[8];
#pragma unroll
			for (int j = r + 1; j < 8; ++j) v = sub(v, mul(row[j], h[j]));
			h[r] = __shfl_sync(0xffffffffu, divd(v, row[r]), r);
		}
	}
	if (lane != 0) return;
	h[8] = 1.0;
	bool bad = singular;
	for (int r = 0; r < 8 && !bad; ++r) bad = !(fabs(h[r]) <= DBL_MAX);
	// ---- denormalise: H = T2^-1 * Hn * T1 (homography_estimator.h:169-172); T = [r 0 -r*m; 0 r -r*m; 0 0 1]
	// T2^-1 = [1/r2 0 mx2; 0 1/r2 my2; 0 0 1] (closed form of Eigen's 3x3 cofactor inverse up to rounding)
	const double t1x = mul(-r1, mx1), t1y = mul(-r1, my1);
	const double ir2 = divd(1.0, r2);
	double A[9]; // A = T2^-1 * Hn
	for (int c = 0; c < 3; ++c) {
		A[0 + c] = add(mul(ir2, h[0 + c]), mul(mx2, h[6 + c]));
		A[3 + c] = add(mul(ir2, h[3 + c]), mul(my2, h[6 + c]));
		A[6 + c] = h[6 + c];
	}
	double *out = H_out + 9 * (int64_t)pb;
	for (int r = 0; r < 3; ++r) {
		out[3 * r + 0] = mul(A[3 * r + 0], r1);
		out[3 * r + 1] = mul(A[3 * r + 1], r1);
		out[3 * r + 2] = add(add(mul(A[3 * r + 0], t1x), mul(A[3 * r + 1], t1y)), A[3 * r + 2]);
	}
	ok_out[pb] = bad ? 0 : 1;
}

int launch_fit_h(pxb_ctx *ctx, int P, const int32_t *off, const int32_t *idx, const double *weights, double *H_out,
                 int32_t *ok_out) {
	if (P <= 0) return PXB_OK;
	// PXB_FIT_H_MMA=1 selects the FP64 tensor-core accumulation (measured in profiles/: it does not beat the scalar
	// block reduction on these 8 x 8 problems, so the scalar form stays the default)
	static const bool use_mma = getenv("PXB_FIT_H_MMA") && atoi(getenv("PXB_FIT_H_MMA")) != 0;
	if (use_mma)
		k_fit_h<true><<<(unsigned)P, kFitThreads, 0, ctx->stream>>>(ctx->pts.aos, off, idx, weights, H_out, ok_out);
	else
		PXB_CUDA(launch_pdl(k_fit_h<false>, dim3((unsigned)P), dim3(kFitThreads), 0, ctx->stream, ctx->pts.aos, off, idx, weights, H_out, ok_out));
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

} // namespace pxb

namespace pxb {
int launch_fit_f(pxb_ctx *ctx, int P, const int32_t *off, const int32_t *idx, const double *weights, double *F_out,
                 int32_t *ok_out);
int launch_fit_pnp(pxb_ctx *ctx, int P, const int32_t *off, const int32_t *idx, double *P_out, int32_t *ok_out);
int launch_fit_vp(pxb_ctx *ctx, int P, const int32_t *off, const int32_t *idx, const double *weights_by_point, double *out,
                  int32_t *ok_out);
int launch_fit_line(pxb_ctx *ctx, int P, const int32_t *off, const int32_t *idx, double *out, int32_t *ok_out);
}

extern "C" {
using namespace pxb;

int pxb_knn_graph(pxb_ctx *ctx, double radius, int k, int32_t *nbr_out_host, int32_t *deg_out_host) {
	PXB_CHECK_ARG(ctx && nbr_out_host && deg_out_host, "null argument");
	PXB_CUDA(cudaSetDevice(ctx->device));
	if (ctx->pts.N <= 0) {
		set_error("no points uploaded");
		return PXB_ERR_STATE;
	}
	const int64_t N = ctx->pts.N;
	PXB_TRY(ctx->idx.reserve(sizeof(int32_t) * (size_t)N * (k + 1)));
	int32_t *d_nbr = ctx->idx.as<int32_t>(), *d_deg = d_nbr + (size_t)N * k;
	PXB_TRY(launch_knn_graph(ctx, radius, k, d_nbr, d_deg));
	PXB_CUDA(cudaMemcpyAsync(nbr_out_host, d_nbr, sizeof(int32_t) * (size_t)N * k, cudaMemcpyDeviceToHost, ctx->stream));
	PXB_CUDA(cudaMemcpyAsync(deg_out_host, d_deg, sizeof(int32_t) * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
	PXB_TRY(ctx_wait(ctx));
	return PXB_OK;
}

int pxb_fit_nonminimal(pxb_ctx *ctx, int32_t P, const int32_t *off_host, const int32_t *idx_host,
                       const double *weights_by_row_host, double *H_out_host, int32_t *ok_out_host) {
	PXB_CHECK_ARG(ctx && off_host && idx_host && H_out_host && ok_out_host && P >= 0, "null argument");
	if (P == 0) return PXB_OK;
	PXB_CUDA(cudaSetDevice(ctx->device));
	if (ctx->pts.N <= 0) {
		set_error("no points uploaded");
		return PXB_ERR_STATE;
	}
	const int32_t total = off_host[P];
	for (int32_t t = 0; t < total; ++t)
		if (idx_host[t] < 0 || idx_host[t] >= ctx->pts.N) {
			set_error("point index %d out of range", idx_host[t]);
			return PXB_ERR_ARGUMENT;
		}
	PXB_TRY(ctx->idx.reserve(sizeof(int32_t) * (size_t)(P + 1 + total) + 64));
	int32_t *d_off = ctx->idx.as<int32_t>(), *d_idx = d_off + (P + 1);
	const int ms = model_size(ctx->pts.type);
	PXB_TRY(ctx->models.reserve(sizeof(double) * (size_t)P * ms));
	PXB_TRY(ctx->outA.reserve(sizeof(int32_t) * (size_t)P));
	double *d_w = nullptr;
	if (weights_by_row_host) {
		// H, F: first `total` entries, read by row (the reference's indexing); VP: all N entries, read by point
		const bool by_point = ctx->pts.type == PXB_MODEL_VANISHING_POINT;
		PXB_CHECK_ARG(P == 1 || by_point, "row-indexed weighted fits are issued one problem at a time");
		const size_t count = by_point ? (size_t)ctx->pts.N : (size_t)total;
		PXB_TRY(ctx->pref2.reserve(sizeof(double) * count));
		d_w = ctx->pref2.as<double>();
		PXB_CUDA(cudaMemcpyAsync(d_w, weights_by_row_host, sizeof(double) * count, cudaMemcpyHostToDevice, ctx->stream));
	}
	PXB_CUDA(cudaMemcpyAsync(d_off, off_host, sizeof(int32_t) * (size_t)(P + 1), cudaMemcpyHostToDevice, ctx->stream));
	PXB_CUDA(cudaMemcpyAsync(d_idx, idx_host, sizeof(int32_t) * (size_t)total, cudaMemcpyHostToDevice, ctx->stream));
	switch (ctx->pts.type) {
	case PXB_MODEL_HOMOGRAPHY: PXB_TRY(launch_fit_h(ctx, P, d_off, d_idx, d_w, ctx->models.as<double>(), ctx->outA.as<int32_t>())); break;
	case PXB_MODEL_FUNDAMENTAL: PXB_TRY(launch_fit_f(ctx, P, d_off, d_idx, d_w, ctx->models.as<double>(), ctx->outA.as<int32_t>())); break;
	case PXB_MODEL_PNP: PXB_TRY(launch_fit_pnp(ctx, P, d_off, d_idx, ctx->models.as<double>(), ctx->outA.as<int32_t>())); break;
	case PXB_MODEL_VANISHING_POINT: PXB_TRY(launch_fit_vp(ctx, P, d_off, d_idx, d_w, ctx->models.as<double>(), ctx->outA.as<int32_t>())); break;
	default: PXB_TRY(launch_fit_line(ctx, P, d_off, d_idx, ctx->models.as<double>(), ctx->outA.as<int32_t>())); break;
	}
	PXB_CUDA(cudaMemcpyAsync(H_out_host, ctx->models.ptr, sizeof(double) * (size_t)P * ms, cudaMemcpyDeviceToHost, ctx->stream));
	PXB_CUDA(cudaMemcpyAsync(ok_out_host, ctx->outA.ptr, sizeof(int32_t) * (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
	PXB_TRY(ctx_wait(ctx));
	return PXB_OK;
}

int pxb_fit_homographies(pxb_ctx *ctx, int32_t P, const int32_t *off_host, const int32_t *idx_host,
                         const double *weights_by_row_host, double *H_out_host, int32_t *ok_out_host) {
	if (ctx && ctx->pts.type != PXB_MODEL_HOMOGRAPHY) {
		set_error("pxb_fit_homographies needs homography correspondences uploaded");
		return PXB_ERR_STATE;
	}
	return pxb_fit_nonminimal(ctx, P, off_host, idx_host, weights_by_row_host, H_out_host, ok_out_host);
}
}
