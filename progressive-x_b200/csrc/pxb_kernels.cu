// pxb_kernels.cu -- the data-parallel N-point loops of Progressive-X as sm_100a kernels.
//
//   k_residual_matrix   a1/a2/a3  N x K squared residuals (f64 or f32) + inlier bit matrix      HBM-write bound
//   (a4, the fused compound-aware score, lives in pxb_score.cu)
//   k_preference        a5        preference vector of one model
//   k_tanimoto          a5        dot / squared norms, one block, fixed topology
//   k_compound_max      a5
//   k_pearl_datacost    a9        N x (L+1) data-cost matrix
//   k_segment_sums      a12       per-instance residual sums, one block per instance
//   k_lo_unary          a13       GC-RANSAC LO unary terms
//   k_tukey             a13       Tukey weights
//
// Layout: points are SoA (coordinate-major, stride padded to 64) so that a warp's 32 lanes read 256 contiguous
// bytes per coordinate. Models of the current tile are staged in shared memory and read with warp-uniform
// addresses (LDS broadcast). The residual matrix is hypothesis-major (r2[k*N + i]): lanes map to consecutive
// points, so every warp store covers two full 128-byte lines; stores are streaming (st.global.cs) because the
// matrix is never re-read by the producer.
//
// Summation order: a thread accumulates its points in increasing index order, lanes combine by a
// fixed xor-butterfly, warps combine in warp order, chunks combine in chunk order. The topology depends on N
// only -- never on K, the grid or the device -- so equal inputs always give equal sums.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "pxb_internal.h"
#include "pxb_residuals.cuh"
#include "pxb_screen.cuh"

namespace pxb {


// ------------------------------------------------------------------------------------------------
// AoS -> SoA re-tiling of the uploaded points
// ------------------------------------------------------------------------------------------------
// Bounding box of the finite coordinates -> NormDev (centre + common scale) for the float32 screening copy.
// Two tiny launches: per-block partial boxes over a grid-strided slice, then one warp folds them.
constexpr int kStatBlocks = 64, kStatThreads = 256;

__global__ void __launch_bounds__(kStatThreads)
    k_point_stats(const double *__restrict__ aos, int64_t N, int dim, double *__restrict__ partial /*[blocks][10]*/) {
	__shared__ double s_lo[kStatThreads / 32][5], s_hi[kStatThreads / 32][5];
	double lo[5], hi[5];
	for (int c = 0; c < 5; ++c) lo[c] = 1e300, hi[c] = -1e300;
	for (int64_t i = (int64_t)blockIdx.x * kStatThreads + threadIdx.x; i < N; i += (int64_t)gridDim.x * kStatThreads)
		for (int c = 0; c < dim; ++c) {
			const double v = aos[i * dim + c];
			if (fabs(v) <= 1e300) lo[c] = fmin(lo[c], v), hi[c] = fmax(hi[c], v);
		}
	for (int c = 0; c < 5; ++c)
		for (int o = 16; o > 0; o >>= 1) {
			lo[c] = fmin(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
			hi[c] = fmax(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
		}
	if ((threadIdx.x & 31) == 0)
		for (int c = 0; c < 5; ++c) s_lo[threadIdx.x >> 5][c] = lo[c], s_hi[threadIdx.x >> 5][c] = hi[c];
	__syncthreads();
	if (threadIdx.x < 5) {
		const int c = threadIdx.x;
		double l = 1e300, h = -1e300;
		for (int w = 0; w < kStatThreads / 32; ++w) l = fmin(l, s_lo[w][c]), h = fmax(h, s_hi[w][c]);
		partial[blockIdx.x * 10 + c] = l;
		partial[blockIdx.x * 10 + 5 + c] = h;
	}
}

__global__ void k_point_norm(const double *__restrict__ partial, int blocks, int dim, int type, NormDev *out) {
	if (threadIdx.x != 0) return;
	double L[5], Hh[5];
	for (int c = 0; c < 5; ++c) {
		L[c] = 1e300, Hh[c] = -1e300;
		for (int b = 0; b < blocks; ++b) L[c] = fmin(L[c], partial[b * 10 + c]), Hh[c] = fmax(Hh[c], partial[b * 10 + 5 + c]);
	}
	if (type == PXB_MODEL_VANISHING_POINT) // both end points of a segment live in the same image: one centre
		for (int c = 0; c < 2; ++c) {
			L[c] = L[c + 2] = fmin(L[c], L[c + 2]);
			Hh[c] = Hh[c + 2] = fmax(Hh[c], Hh[c + 2]);
		}
	NormDev nd;
	double s = 0.0;
	for (int c = 0; c < 5; ++c) {
		const double l = L[c], h = Hh[c];
		const bool used = c < dim && !(type == PXB_MODEL_PNP && c < 2) && l <= h;
		nd.c[c] = used ? 0.5 * (l + h) : 0.0;
		if (used) s = fmax(s, 0.5 * (h - l));
	}
	nd.s = (s > 1e-300 && s < 1e300) ? s : 1.0;
	nd.inv_s = 1.0 / nd.s;
	*out = nd;
}

__global__ void k_aos_to_soa(const double *__restrict__ aos, double *__restrict__ soa, float *__restrict__ f32n,
                             float *__restrict__ q, const NormDev *__restrict__ norm, int64_t N, int64_t stride, int dim,
                             int type) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= stride) return;
	double p[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
	for (int c = 0; c < dim; ++c) soa[c * stride + i] = p[c] = (i < N) ? aos[i * dim + c] : 0.0;
	const NormDev nd = *norm;
	float pf[5], qq;
	PXB_DISPATCH_TYPE(type, screen_point<TYPE>(p, nd, pf, qq));
	for (int c = 0; c < dim; ++c) f32n[c * stride + i] = pf[c];
	q[i] = qq;
}

int launch_aos_to_soa(pxb_ctx *ctx) {
	Points &p = ctx->pts;
	const int blocks = (int)std::min<int64_t>(kStatBlocks, (p.N + kStatThreads - 1) / kStatThreads);
	PXB_TRY(ctx->stats.reserve(sizeof(double) * 10 * kStatBlocks));
	k_point_stats<<<blocks, kStatThreads, 0, ctx->stream>>>(p.aos, p.N, p.dim, ctx->stats.as<double>());
	k_point_norm<<<1, 32, 0, ctx->stream>>>(ctx->stats.as<double>(), blocks, p.dim, p.type, p.norm);
	ctx->launches += 2;
	const int grid = (int)((p.stride + kThreads - 1) / kThreads);
	k_aos_to_soa<<<grid, kThreads, 0, ctx->stream>>>(p.aos, p.soa, p.f32n, p.q, p.norm, p.N, p.stride, p.dim, p.type);
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}


// ------------------------------------------------------------------------------------------------
// a1/a2/a3: residual-and-inlier matrix
// ------------------------------------------------------------------------------------------------
// Register tile: each lane owns kRmP points (base + lane + 32 j), the block walks a tile of hypotheses staged in
// shared memory (padded to an even number of doubles so that a model is read with LDS.128 broadcasts).
// Per evaluation: 28 FP64-pipe instructions (H), ~1.5 LDS, 1 STG, 6 range-test instructions, one branch per
// hypothesis for the whole register tile -> the FP64 pipe (2 issue slots per instruction) is the binding unit.
// hypotheses per block: 32 keeps the grid full at RANSAC batch sizes (K ~ 500); large batches take 64 (half the
// per-block prologue: point loads, input screening, barrier)
constexpr int kRmHypsSmall = 32, kRmHypsLarge = 64;
constexpr int64_t kRmLargeBatch = 4096;

template <typename OUT> __device__ __forceinline__ void store_stream(OUT *p, double v);
template <> __device__ __forceinline__ void store_stream<double>(double *p, double v) { __stcs(p, v); }
template <> __device__ __forceinline__ void store_stream<float>(float *p, double v) { __stcs(p, __double2float_rn(v)); }

template <int TYPE, typename OUT, bool HAS_R2, bool HAS_MASK, int kRmP, bool HI_ONLY, bool FULL>
__device__ __forceinline__ void rm_hypothesis_loop(const double (&p)[kRmP][5], const bool (&valid)[kRmP],
                                                   const double *s_models, int nk, int wild, double T2, unsigned hiT,
                                                   OUT *out, uint32_t *mout, int64_t N, int64_t words, int nwords,
                                                   int lane) {
	constexpr int MP = ModelTraits<TYPE>::kPadded;
#pragma unroll 2
	for (int k = 0; k < nk; ++k) {
		double m[12];
		load_model_smem<TYPE>(s_models + k * MP, m);
		double r[kRmP];
		float lo[kRmP];
#pragma unroll
		for (int j = 0; j < kRmP; ++j) r[j] = squared_residual_tile<TYPE>(p[j], m, lo[j]);
		if (__builtin_expect(!(tile_min4(lo) >= __int_as_float(kHiMinPattern)) || wild, 0))
			PXB_RESIDUAL_TILE_EXACT(TYPE, kRmP, p, m, r);
		if (HAS_R2) {
#pragma unroll
			for (int j = 0; j < kRmP; ++j)
				if (FULL || valid[j]) store_stream<OUT>(out + 32 * j, r[j]);
			out += N;
		}
		if (HAS_MASK) {
			uint32_t w[kRmP];
#pragma unroll
			for (int j = 0; j < kRmP; ++j)
				w[j] = __ballot_sync(0xffffffffu, (FULL || valid[j]) && below_threshold<HI_ONLY>(r[j], T2, hiT));
			uint32_t mine = w[0];
#pragma unroll
			for (int j = 1; j < kRmP; ++j) mine = (lane == j) ? w[j] : mine;
			if (lane < (FULL ? kRmP : nwords)) mout[lane] = mine;
			mout += words;
		}
	}
}

template <int TYPE, typename OUT, bool HAS_R2, bool HAS_MASK, int kRmP, int MINB, bool HI_ONLY, int kRmHypsPerBlock>
__global__ void __launch_bounds__(kThreads, MINB)
    k_residual_matrix(const double *__restrict__ soa, int64_t stride, int64_t N, const double *__restrict__ models,
                      int64_t K, double T2, OUT *__restrict__ r2, uint32_t *__restrict__ mask, int64_t words) {
	constexpr int DIM = ModelTraits<TYPE>::kDim, MS = ModelTraits<TYPE>::kSize, MP = ModelTraits<TYPE>::kPadded;
	__shared__ __align__(16) double s_models[kRmHypsPerBlock * MP];

	const int64_t k0 = (int64_t)blockIdx.y * kRmHypsPerBlock;
	const int nk = (int)min((int64_t)kRmHypsPerBlock, K - k0);
	// inputs outside +-2^60 (or NaN/inf) make the whole block take the plain div.rn.f64 loop (see pxb_residuals.cuh)
	int wild = 0;
	for (int t = threadIdx.x; t < nk * MS; t += kThreads) {
		const double v = models[k0 * MS + t];
		s_models[(t / MS) * MP + (t % MS)] = v;
		wild |= !(fabs(v) <= kInputMagnitudeLimit);
	}

	constexpr int kRmPointsPerWarp = 32 * kRmP, kRmPointsPerBlock = (kThreads / 32) * kRmPointsPerWarp;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int64_t base = (int64_t)blockIdx.x * kRmPointsPerBlock + warp * kRmPointsPerWarp;
	double p[kRmP][5];
	bool valid[kRmP];
#pragma unroll
	for (int j = 0; j < kRmP; ++j) {
		const int64_t i = base + lane + 32 * j;
		valid[j] = i < N;
		load_point<DIM>(soa, stride, valid[j] ? i : (N - 1), p[j]);
#pragma unroll
		for (int c = 0; c < DIM; ++c) wild |= !(fabs(p[j][c]) <= kInputMagnitudeLimit);
	}
	wild = __syncthreads_or(wild);
	if (base >= N) return;
	const unsigned hiT = (unsigned)__double2hiint(T2);
	OUT *out = HAS_R2 ? r2 + k0 * N + base + lane : nullptr;
	uint32_t *mout = HAS_MASK ? mask + k0 * words + (base >> 5) : nullptr;
	// Interior warps (all kRmP x 32 points exist) run a loop without any per-point validity predicate: the 64-bit
	// index compares and predicated stores of the ragged variant cost ~4 issue slots per evaluation. The branch is
	// warp-uniform; only the single warp straddling N takes the ragged loop.
	if (base + kRmPointsPerWarp <= N)
		rm_hypothesis_loop<TYPE, OUT, HAS_R2, HAS_MASK, kRmP, HI_ONLY, true>(p, valid, s_models, nk, wild, T2, hiT, out,
		                                                                      mout, N, words, kRmP, lane);
	else
		rm_hypothesis_loop<TYPE, OUT, HAS_R2, HAS_MASK, kRmP, HI_ONLY, false>(
		    p, valid, s_models, nk, wild, T2, hiT, out, mout, N, words, (int)min((int64_t)kRmP, words - (base >> 5)), lane);
}

template <int TYPE, typename OUT, bool HI_ONLY, int HYPS>
static int launch_rm_v(pxb_ctx *ctx, const double *models, int64_t K, double T2, OUT *r2, uint32_t *mask) {
	// points per lane, min blocks per SM. Tuned on B200 (1.023 ms): (5,2) 1.039, (6,2) 1.052 (spills), (8,1) 1.121,
	// (4,3), (2,3), (2,4) within 3 % but never faster
	constexpr int P = 4, MINB = 2;
	const Points &p = ctx->pts;
	const int64_t words = (p.N + 31) / 32;
	const int64_t ppb = (int64_t)(kThreads / 32) * 32 * P;
	const int64_t gx = (p.N + ppb - 1) / ppb;
	int64_t done = 0;
	while (done < K) { // gridDim.y is limited to 65535
		const int64_t kk = std::min<int64_t>(K - done, (int64_t)65535 * HYPS);
		dim3 grid((unsigned)gx, (unsigned)((kk + HYPS - 1) / HYPS));
		const double *mm = models + done * ModelTraits<TYPE>::kSize;
		OUT *rr = r2 ? r2 + done * p.N : nullptr;
		uint32_t *mk = mask ? mask + done * words : nullptr;
		if (rr && mk)
			k_residual_matrix<TYPE, OUT, true, true, P, MINB, HI_ONLY, HYPS><<<grid, kThreads, 0, ctx->stream>>>(p.soa, p.stride, p.N, mm, kk, T2, rr, mk, words);
		else if (rr)
			k_residual_matrix<TYPE, OUT, true, false, P, MINB, false, HYPS><<<grid, kThreads, 0, ctx->stream>>>(p.soa, p.stride, p.N, mm, kk, T2, rr, mk, words);
		else if (mk)
			k_residual_matrix<TYPE, OUT, false, true, P, MINB, HI_ONLY, HYPS><<<grid, kThreads, 0, ctx->stream>>>(p.soa, p.stride, p.N, mm, kk, T2, rr, mk, words);
		ctx->launches++;
		done += kk;
	}
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

template <int TYPE, typename OUT>
static int launch_rm_t(pxb_ctx *ctx, const double *models, int64_t K, double T2, OUT *r2, uint32_t *mask) {
	static const int forced = getenv("PXB_RM_HYPS") ? atoi(getenv("PXB_RM_HYPS")) : 0; // tuning knob (tools/time_kernels.py)
	const bool large = forced ? forced >= kRmHypsLarge : K >= kRmLargeBatch;
	const bool hi = threshold_low_word_is_zero(T2);
	if (large)
		return hi ? launch_rm_v<TYPE, OUT, true, kRmHypsLarge>(ctx, models, K, T2, r2, mask)
		          : launch_rm_v<TYPE, OUT, false, kRmHypsLarge>(ctx, models, K, T2, r2, mask);
	return hi ? launch_rm_v<TYPE, OUT, true, kRmHypsSmall>(ctx, models, K, T2, r2, mask)
	          : launch_rm_v<TYPE, OUT, false, kRmHypsSmall>(ctx, models, K, T2, r2, mask);
}

int launch_residual_matrix(pxb_ctx *ctx, const double *models, int64_t K, double T2, double *r2, float *r2f,
                           uint32_t *mask) {
	if (K <= 0) return PXB_OK;
	// bit matrix only: the float32-screened kernel (pxb_score.cu) -- same bits, and provable outliers skip the float64 path.
	// Single models (inlier lists of one hypothesis) stay here: they are inlier rich and one launch cheaper.
	static const bool no_screen = getenv("PXB_MASK_EXACT") && atoi(getenv("PXB_MASK_EXACT")) != 0; // A/B knob for the tests
	if (!r2 && !r2f && mask && K >= 8 && !no_screen) return launch_inlier_mask(ctx, models, K, T2, mask);
	int rc = PXB_OK;
	if (r2f)
		PXB_DISPATCH_TYPE(ctx->pts.type, rc = (launch_rm_t<TYPE, float>(ctx, models, K, T2, r2f, mask)));
	else
		PXB_DISPATCH_TYPE(ctx->pts.type, rc = (launch_rm_t<TYPE, double>(ctx, models, K, T2, r2, mask)));
	return rc;
}

// ------------------------------------------------------------------------------------------------
// a5: preference vector, tanimoto, compound max
// ------------------------------------------------------------------------------------------------
template <int TYPE>
__global__ void k_preference(const double *__restrict__ soa, int64_t stride, int64_t N, const double *__restrict__ model,
                             double T, double *__restrict__ pref) {
	constexpr int DIM = ModelTraits<TYPE>::kDim, MS = ModelTraits<TYPE>::kSize;
	__shared__ double m[MS];
	if (threadIdx.x < MS) m[threadIdx.x] = model[threadIdx.x];
	__syncthreads();
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= N) return;
	double p[5];
	load_point<DIM>(soa, stride, i, p);
	const double r2 = squared_residual<TYPE>(p, m);
	pref[i] = cv_max(0.0, sub(1.0, divd(r2, T))); // progx_model.h:84-85
}

int launch_preference(pxb_ctx *ctx, const double *model, double T, double *pref) {
	const Points &p = ctx->pts;
	const unsigned grid = (unsigned)((p.N + kThreads - 1) / kThreads);
	PXB_DISPATCH_TYPE(p.type, (k_preference<TYPE><<<grid, kThreads, 0, ctx->stream>>>(p.soa, p.stride, p.N, model, T, pref)));
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

constexpr int kOneBlock = 1024;

// block-wide sum with the fixed topology (strided thread partials -> xor butterfly -> warps in order)
__device__ __forceinline__ double block_sum_1024(double x, double *s_tmp /*32*/) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) x = add(x, __shfl_xor_sync(0xffffffffu, x, o));
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	__syncthreads();
	if (lane == 0) s_tmp[warp] = x;
	__syncthreads();
	double t = 0.0;
	for (int w = 0; w < kOneBlock / 32; ++w) t = add(t, s_tmp[w]);
	return t;
}

__global__ void __launch_bounds__(kOneBlock)
    k_tanimoto(const double *__restrict__ a, const double *__restrict__ b, int64_t N, double *__restrict__ out3) {
	__shared__ double s_tmp[32];
	double d = 0.0, na = 0.0, nb = 0.0;
	for (int64_t i = threadIdx.x; i < N; i += kOneBlock) {
		const double x = a[i], y = b[i];
		d = add(d, mul(x, y));
		na = add(na, mul(x, x));
		nb = add(nb, mul(y, y));
	}
	d = block_sum_1024(d, s_tmp);
	na = block_sum_1024(na, s_tmp);
	nb = block_sum_1024(nb, s_tmp);
	if (threadIdx.x == 0) {
		out3[0] = d;
		out3[1] = na;
		out3[2] = nb;
	}
}

int launch_tanimoto(pxb_ctx *ctx, const double *a, const double *b, int64_t N, double *out3) {
	k_tanimoto<<<1, kOneBlock, 0, ctx->stream>>>(a, b, N, out3);
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

__global__ void k_compound_max(const double *__restrict__ prefs, int64_t L, int64_t N, double *__restrict__ out) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= N) return;
	double m = 0.0; // progressive_x.h:604 setConstant(0)
	for (int64_t k = 0; k < L; ++k) m = cv_max(m, prefs[k * N + i]);
	out[i] = m;
}

int launch_compound_max(pxb_ctx *ctx, const double *prefs, int64_t L, int64_t N, double *out) {
	k_compound_max<<<(unsigned)((N + kThreads - 1) / kThreads), kThreads, 0, ctx->stream>>>(prefs, L, N, out);
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

// ------------------------------------------------------------------------------------------------
// a9: PEARL data-cost matrix, D[i*(L+1) + l]
// ------------------------------------------------------------------------------------------------
constexpr int kMaxLabels = 64; // the reference never exceeds 11 (outer loop cap 10, progressive_x.h:272)

template <int TYPE>
__global__ void k_pearl_datacost(const double *__restrict__ soa, int64_t stride, int64_t N,
                                 const double *__restrict__ models, int L, double T, double one_minus,
                                 double *__restrict__ D) {
	constexpr int DIM = ModelTraits<TYPE>::kDim, MS = ModelTraits<TYPE>::kSize;
	extern __shared__ double s_m[];
	for (int t = threadIdx.x; t < L * MS; t += blockDim.x) s_m[t] = models[t];
	__syncthreads();
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= N) return;
	double p[5];
	load_point<DIM>(soa, stride, i, p);
	double *row = D + i * (L + 1);
	const double two = mul(2.0, one_minus);
	for (int l = 0; l < L; ++l) {
		const double r2 = squared_residual<TYPE>(p, s_m + l * MS);
		// PEARL.h:123-127: r2 > T -> 2(1-lambda); else (1-lambda) * r2 / T  (left-to-right)
		row[l] = (r2 > T) ? two : divd(mul(one_minus, r2), T);
	}
	row[L] = one_minus; // PEARL.h:100-101
}

int launch_pearl_datacost(pxb_ctx *ctx, const double *models, int64_t L, double thr, double lambda, double *D) {
	const Points &p = ctx->pts;
	if (L > kMaxLabels) {
		set_error("too many labels (%lld > %d)", (long long)L, kMaxLabels);
		return PXB_ERR_ARGUMENT;
	}
	const double T = 9.0 / 4.0 * thr * thr; // PEARL.h:51 spelling
	const double one_minus = 1.0 - lambda;
	const unsigned grid = (unsigned)((p.N + kThreads - 1) / kThreads);
	const size_t smem = sizeof(double) * (size_t)std::max<int64_t>(L, 1) * model_size(p.type);
	PXB_DISPATCH_TYPE(p.type, (k_pearl_datacost<TYPE><<<grid, kThreads, smem, ctx->stream>>>(p.soa, p.stride, p.N, models, (int)L, T, one_minus, D)));
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

// ------------------------------------------------------------------------------------------------
// a12: per-instance residual sums (one block per instance, fixed topology)
// ------------------------------------------------------------------------------------------------
// The summation topology is block_sum_1024's over 1024 "virtual" threads (thread v adds the points v, v + 1024, ... of its
// instance in index order; xor butterfly over the lanes; the 32 warp sums in warp order) -- but the virtual threads of an
// instance are spread over kSegSplit blocks of 128, so that L instances occupy 8 L SMs instead of L (N = 10^5, L = 10:
// 99 -> ~15 us per launch). Every block leaves its four warp sums in scratch memory; the block that arrives last at the
// instance's ticket adds the 32 of them up in warp order: the same additions in the same order as one 1024-thread block,
// bit for bit.
constexpr int kSegSplit = 8, kSegThreads = kOneBlock / kSegSplit;
constexpr size_t kSegScratchBytes = kMaxLabels * 32 * (sizeof(double) + sizeof(int)) + kMaxLabels * sizeof(unsigned);

template <int TYPE>
__global__ void __launch_bounds__(kSegThreads)
    k_segment_sums(const double *__restrict__ soa, int64_t stride, int64_t N, const double *__restrict__ models,
                   const int32_t *__restrict__ labels, double *__restrict__ sums, int64_t *__restrict__ counts,
                   double *ws /*[L][32]*/, int *wc /*[L][32]*/, unsigned *ticket /*[L], zero between launches*/) {
	constexpr int DIM = ModelTraits<TYPE>::kDim, MS = ModelTraits<TYPE>::kSize;
	__shared__ double m[MS];
	__shared__ bool s_last;
	const int l = blockIdx.x / kSegSplit, v = (blockIdx.x % kSegSplit) * kSegThreads + threadIdx.x; // virtual thread
	if (threadIdx.x < MS) m[threadIdx.x] = models[l * MS + threadIdx.x];
	__syncthreads();
	double acc = 0.0;
	int cnt = 0;
	// four visits at a time: their labels, then the rows of the matching ones, are in flight together and the four
	// residuals are independent; the additions stay in visit order (a thread's loop used to be one dependent
	// label -> row -> divide -> sqrt -> add chain per visit: ~900 cycles each, 98 visits at N = 10^5)
	constexpr int U = 4;
	for (int64_t i0 = v; i0 < N; i0 += (int64_t)U * kOneBlock) {
		bool mine[U];
		double r[U];
#pragma unroll
		for (int u = 0; u < U; ++u) {
			const int64_t i = i0 + (int64_t)u * kOneBlock;
			mine[u] = i < N && labels[i] == l;
		}
#pragma unroll
		for (int u = 0; u < U; ++u) {
			r[u] = 0.0;
			if (mine[u]) {
				double p[5];
				load_point<DIM>(soa, stride, i0 + (int64_t)u * kOneBlock, p);
				r[u] = __dsqrt_rn(squared_residual<TYPE>(p, m)); // Estimator::residual = sqrt(squaredResidual)
			}
		}
#pragma unroll
		for (int u = 0; u < U; ++u)
			if (mine[u]) {
				acc = add(acc, r[u]);
				cnt++;
			}
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		acc = add(acc, __shfl_xor_sync(0xffffffffu, acc, o));
		cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
	}
	if ((threadIdx.x & 31) == 0) {
		ws[l * 32 + (v >> 5)] = acc;
		wc[l * 32 + (v >> 5)] = cnt;
		__threadfence();
	}
	__syncthreads();
	if (threadIdx.x == 0) s_last = atomicAdd(&ticket[l], 1u) == (unsigned)(kSegSplit - 1);
	__syncthreads();
	if (s_last && threadIdx.x == 0) {
		__threadfence();
		double t = 0.0;
		long long c = 0;
		for (int w = 0; w < 32; ++w) {
			t = add(t, __ldcg(ws + l * 32 + w));
			c += __ldcg(wc + l * 32 + w);
		}
		sums[l] = t;
		counts[l] = c;
		ticket[l] = 0u; // ready for the next launch on this stream
	}
}

int launch_segment_sums(pxb_ctx *ctx, const double *models, int64_t L, const int32_t *labels, double *sums,
                        int64_t *counts) {
	if (L <= 0) return PXB_OK;
	if (L > kMaxLabels) {
		set_error("at most %d instances", kMaxLabels);
		return PXB_ERR_ARGUMENT;
	}
	const Points &p = ctx->pts;
	if (!ctx->seg_scratch.ptr) { // allocated once per context (captured chains keep its address), tickets start at zero
		PXB_TRY(ctx->seg_scratch.reserve(kSegScratchBytes));
		PXB_CUDA(cudaMemsetAsync(ctx->seg_scratch.ptr, 0, kSegScratchBytes, ctx->stream));
	}
	double *ws = ctx->seg_scratch.as<double>();
	int *wc = reinterpret_cast<int *>(ws + kMaxLabels * 32);
	unsigned *ticket = reinterpret_cast<unsigned *>(wc + kMaxLabels * 32);
	PXB_DISPATCH_TYPE(p.type, (k_segment_sums<TYPE><<<(unsigned)(L * kSegSplit), kSegThreads, 0, ctx->stream>>>(
	                               p.soa, p.stride, p.N, models, labels, sums, counts, ws, wc, ticket)));
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

// ------------------------------------------------------------------------------------------------
// a13: GC-RANSAC local-optimisation unary terms and Tukey weights
// ------------------------------------------------------------------------------------------------
template <int TYPE>
__device__ __forceinline__ void lo_unary_terms(const double (&p)[5], const double *m, double T, double one_minus, double &dist,
                                               double &e0, double &e1) {
	const double r2 = squared_residual<TYPE>(p, m);
	// std::clamp(v, 0.0, 1.0): (v < lo) ? lo : (hi < v) ? hi : v   -- NaN passes through
	const double q = divd(r2, T);
	dist = (q < 0.0) ? 0.0 : ((1.0 < q) ? 1.0 : q);
	const double tmp = sub(1.0, dist);
	if (r2 <= T) { // GCRANSAC.h:958-961
		e0 = mul(one_minus, tmp);
		e1 = 0.0;
	} else {
		e0 = 0.0;
		e1 = mul(one_minus, sub(1.0, tmp));
	}
}

template <int TYPE>
__global__ void k_lo_unary(const double *__restrict__ soa, int64_t stride, int64_t N, const double *__restrict__ model,
                           double T, double one_minus, double *__restrict__ d, double *__restrict__ e0,
                           double *__restrict__ e1) {
	constexpr int DIM = ModelTraits<TYPE>::kDim, MS = ModelTraits<TYPE>::kSize;
	__shared__ double m[MS];
	if (threadIdx.x < MS) m[threadIdx.x] = model[threadIdx.x];
	__syncthreads();
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= N) return;
	double p[5];
	load_point<DIM>(soa, stride, i, p);
	double dist, a, b;
	lo_unary_terms<TYPE>(p, m, T, one_minus, dist, a, b);
	d[i] = dist;
	e0[i] = a;
	e1[i] = b;
}

// The cut of GCRANSAC::labeling without a smoothness term (lambda = 0 or no neighbourhood graph) decomposes per node:
// SINK (= inlier) iff the t-link residual cap_source - cap_sink = e1 - e0 is negative (gcr/energy.h:204-208).
template <int TYPE>
__global__ void k_lo_unary_cut(const double *__restrict__ soa, int64_t stride, int64_t N, const double *__restrict__ model,
                               double T, double one_minus, uint8_t *__restrict__ inlier) {
	constexpr int DIM = ModelTraits<TYPE>::kDim, MS = ModelTraits<TYPE>::kSize;
	__shared__ double m[MS];
	pdl_launch_dependents(); // (first kernel of the local-optimisation chain: the ordered compaction may be scheduled at once)
	if (threadIdx.x < MS) m[threadIdx.x] = model[threadIdx.x];
	__syncthreads();
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= N) return;
	double p[5];
	load_point<DIM>(soa, stride, i, p);
	double dist, a, b;
	lo_unary_terms<TYPE>(p, m, T, one_minus, dist, a, b);
	inlier[i] = sub(b, a) < 0.0 ? 1 : 0;
}

int launch_lo_unary_cut(pxb_ctx *ctx, const double *model, double thr, double lambda, uint8_t *inlier) {
	const Points &p = ctx->pts;
	const double T = thr * thr * 9 / 4; // GCRANSAC.h:942 spelling
	const double one_minus = 1.0 - lambda;
	const unsigned grid = (unsigned)((p.N + kThreads - 1) / kThreads);
	PXB_DISPATCH_TYPE(p.type, (k_lo_unary_cut<TYPE><<<grid, kThreads, 0, ctx->stream>>>(p.soa, p.stride, p.N, model, T, one_minus, inlier)));
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

int launch_lo_unary(pxb_ctx *ctx, const double *model, double thr, double lambda, double *d, double *e0, double *e1) {
	const Points &p = ctx->pts;
	const double T = thr * thr * 9 / 4; // GCRANSAC.h:942 spelling
	const double one_minus = 1.0 - lambda;
	const unsigned grid = (unsigned)((p.N + kThreads - 1) / kThreads);
	PXB_DISPATCH_TYPE(p.type, (k_lo_unary<TYPE><<<grid, kThreads, 0, ctx->stream>>>(p.soa, p.stride, p.N, model, T, one_minus, d, e0, e1)));
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

template <int TYPE>
__global__ void k_tukey(const double *__restrict__ soa, int64_t stride, int64_t N, const double *__restrict__ model,
                        double T2, double *__restrict__ w) {
	constexpr int DIM = ModelTraits<TYPE>::kDim, MS = ModelTraits<TYPE>::kSize;
	__shared__ double m[MS];
	if (threadIdx.x < MS) m[threadIdx.x] = model[threadIdx.x];
	__syncthreads();
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= N) return;
	double p[5];
	load_point<DIM>(soa, stride, i, p);
	const double r2 = squared_residual<TYPE>(p, m);
	const double t = cv_max(0.0, sub(1.0, divd(r2, T2))); // GCRANSAC.h:667-668
	w[i] = mul(t, t);
}

int launch_tukey(pxb_ctx *ctx, const double *model, double T2, double *w) {
	const Points &p = ctx->pts;
	const unsigned grid = (unsigned)((p.N + kThreads - 1) / kThreads);
	PXB_DISPATCH_TYPE(p.type, (k_tukey<TYPE><<<grid, kThreads, 0, ctx->stream>>>(p.soa, p.stride, p.N, model, T2, w)));
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

} // namespace pxb

// ------------------------------------------------------------------------------------------------
// self-test: dual_div() against __ddiv_rn on pseudo-random operands (count of mismatching bit patterns)
// ------------------------------------------------------------------------------------------------
namespace pxb {
__device__ __forceinline__ uint64_t splitmix64(uint64_t &x) {
	uint64_t z = (x += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
__global__ void k_selftest_division(uint64_t seed, int iters, int mode, unsigned long long *mismatches) {
	uint64_t s = seed + 0x1234567ull * ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1);
	unsigned long long bad = 0;
	for (int it = 0; it < iters; ++it) {
		double a1, a2, b;
		if (mode == 0) { // arbitrary bit patterns (incl. NaN, inf, denormals)
			a1 = __longlong_as_double((long long)splitmix64(s));
			a2 = __longlong_as_double((long long)splitmix64(s));
			b = __longlong_as_double((long long)splitmix64(s));
		} else { // magnitudes typical for the residual kernels: |x| in [2^-20, 2^20], random sign
			auto gen = [&]() {
				const uint64_t r = splitmix64(s);
				const int e = (int)(r % 41) - 20;
				const double m = 1.0 + (double)((r >> 11) & ((1ull << 52) - 1)) * (1.0 / 4503599627370496.0);
				return ((r >> 63) ? -1.0 : 1.0) * ldexp(m, e);
			};
			a1 = gen();
			a2 = gen();
			b = gen();
		}
		double q1, q2;
		dual_div(a1, a2, b, q1, q2);
		const double r1 = __ddiv_rn(a1, b), r2 = __ddiv_rn(a2, b);
		const bool ok1 = (__double_as_longlong(q1) == __double_as_longlong(r1)) || (q1 != q1 && r1 != r1);
		const bool ok2 = (__double_as_longlong(q2) == __double_as_longlong(r2)) || (q2 != q2 && r2 != r2);
		bad += (ok1 ? 0 : 1) + (ok2 ? 0 : 1);
	}
	if (bad) atomicAdd(mismatches, bad);
}
} // namespace pxb

extern "C" int pxb_selftest_division(pxb_ctx *ctx, uint64_t seed, int64_t n_triples, int mode, int64_t *mismatches) {
	using namespace pxb;
	PXB_CHECK_ARG(ctx && mismatches && n_triples > 0, "null argument");
	PXB_CUDA(cudaSetDevice(ctx->device));
	PXB_TRY(ctx->outA.reserve(sizeof(unsigned long long)));
	PXB_CUDA(cudaMemsetAsync(ctx->outA.ptr, 0, sizeof(unsigned long long), ctx->stream));
	const int threads = 256, blocks = 148 * 8;
	const int iters = (int)((n_triples + (int64_t)threads * blocks - 1) / ((int64_t)threads * blocks));
	k_selftest_division<<<blocks, threads, 0, ctx->stream>>>(seed, iters, mode, ctx->outA.as<unsigned long long>());
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	unsigned long long bad = 0;
	PXB_CUDA(cudaMemcpyAsync(&bad, ctx->outA.ptr, sizeof(bad), cudaMemcpyDeviceToHost, ctx->stream));
	PXB_CUDA(cudaStreamSynchronize(ctx->stream));
	*mismatches = (int64_t)bad;
	return PXB_OK;
}
