// pxb_score.cu -- a4: the fused compound-aware MSAC score (no residual matrix materialised).
//
// Replaces MSACScoringFunctionWithCompoundModel::getScore (px/include/scoring_function_with_compound_model.h:61-125)
// for K hypotheses at once: per hypothesis the inlier count, sum max(0, 1 - r2/T2) and the support shared with the
// compound preference vector, sum min(compound_pref_i, pref_i).
//
// Shape of the kernel (same register tile as k_residual_matrix, so the same FP64-dispatch bound applies):
//   block = 256 threads = 1024 points (lane owns base+lane+32j, j<4, kept in registers) x a tile of 32 hypotheses
//   staged in shared memory. Inliers are rare, so per hypothesis a thread only tests 4 residuals and -- under one
//   branch -- accumulates its own (count, value, shared) triple, which it parks in shared memory [hyp][thread].
//   After every 8 hypotheses the block folds those columns.
//
// Summation topology (a function of N only -- never of K, the grid or the batch split):
//   thread partial over its 4 points in j order -> per lane, over the 8 warps in warp order -> xor butterfly over
//   the 32 lanes -> chunks (blocks of 1024 points) in chunk order (k_score_finalize).
#include "pxb_internal.h"
#include "pxb_residuals.cuh"

namespace pxb {

constexpr int kScP = 4;                                   // points per lane
constexpr int kScChunk = kThreads * kScP;                 // 1024 points per block
constexpr int kScHypsPerBlock = 32; // 128 gains ~1% at K=10k but starves the grid at RANSAC batch sizes
constexpr int kScSub = 8;                                 // hypotheses folded per reduction pass

struct ScorePartial {
	double value, shared;
	long long count;
};

template <int TYPE, bool HAS_CP, bool HI_ONLY>
__global__ void __launch_bounds__(kThreads, 2)
    k_score_partial(const double *__restrict__ soa, int64_t stride, int64_t N, const double *__restrict__ models,
                    int64_t K, double T2, const double *__restrict__ compound_pref, ScorePartial *__restrict__ partials,
                    int nchunks) {
	constexpr int DIM = ModelTraits<TYPE>::kDim, MS = ModelTraits<TYPE>::kSize, MP = ModelTraits<TYPE>::kPadded;
	__shared__ __align__(16) double s_models[kScHypsPerBlock * MP];
	__shared__ double s_v[kScSub][kThreads];
	__shared__ double s_s[HAS_CP ? kScSub : 1][kThreads];
	__shared__ int s_c[kScSub][kThreads];

	const int64_t k0 = (int64_t)blockIdx.y * kScHypsPerBlock;
	const int nk = (int)min((int64_t)kScHypsPerBlock, K - k0);
	int wild = 0;
	for (int t = threadIdx.x; t < nk * MS; t += kThreads) {
		const double mv = models[k0 * MS + t];
		s_models[(t / MS) * MP + (t % MS)] = mv;
		wild |= !(fabs(mv) <= kInputMagnitudeLimit);
	}

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int chunk = blockIdx.x;
	const int64_t base = (int64_t)chunk * kScChunk + warp * (32 * kScP);
	double p[kScP][5], cp[kScP];
	bool valid[kScP];
#pragma unroll
	for (int j = 0; j < kScP; ++j) {
		const int64_t i = base + lane + 32 * j;
		valid[j] = i < N;
		load_point<DIM>(soa, stride, valid[j] ? i : (N - 1), p[j]);
		cp[j] = (HAS_CP && valid[j]) ? __ldg(compound_pref + i) : 0.0;
#pragma unroll
		for (int c = 0; c < DIM; ++c) wild |= !(fabs(p[j][c]) <= kInputMagnitudeLimit);
	}
	wild = __syncthreads_or(wild); // also orders the s_models writes
	const unsigned hiT = (unsigned)__double2hiint(T2);
	// r2 / T2 divides by a launch constant: one Newton reciprocal per thread, then the 3-instruction quotient of
	// ptxas' own fast path (bit-identical inside its domain: 2^-300 <= r2, T2 within 2^+-200 -- checked by the host;
	// anything else takes the plain division)
	const bool fastT = fabs(T2) >= 6.2e-61 && fabs(T2) <= 1.6e60;
	const double rT = rcp_newton(fastT ? T2 : 1.0);

	for (int ks = 0; ks < nk; ks += kScSub) {
		const int nsub = min(kScSub, nk - ks);
#pragma unroll 2
		for (int h = 0; h < nsub; ++h) {
			double m[12];
			load_model_smem<TYPE>(s_models + (ks + h) * MP, m);
			double r[kScP];
			float lo[kScP];
#pragma unroll
			for (int j = 0; j < kScP; ++j) r[j] = squared_residual_tile<TYPE>(p[j], m, lo[j]);
			if (__builtin_expect(!(tile_min4(lo) >= __int_as_float(kHiMinPattern)) || wild, 0))
				PXB_RESIDUAL_TILE_EXACT(TYPE, kScP, p, m, r);
			int c = 0;
			double v = 0.0, s = 0.0;
			bool any = false;
#pragma unroll
			for (int j = 0; j < kScP; ++j) any |= valid[j] && below_threshold<HI_ONLY>(r[j], T2, hiT);
			if (any) { // scoring_function_with_compound_model.h:85-102, points in j order
#pragma unroll
				for (int j = 0; j < kScP; ++j) {
					if (valid[j] && below_threshold<HI_ONLY>(r[j], T2, hiT)) {
						++c;
						const double q = (fastT && __double2hiint(r[j]) >= kHiMinPattern) ? fast_quotient_nocheck(r[j], T2, rT)
						                                                                   : divd(r[j], T2);
						const double sv = cv_max(0.0, sub(1.0, q));
						v = add(v, sv);
						if (HAS_CP) s = add(s, cv_min(cp[j], sv)); // :115-117 (pref is 0 off the inlier set)
					}
				}
			}
			s_v[h][threadIdx.x] = v;
			if (HAS_CP) s_s[h][threadIdx.x] = s;
			s_c[h][threadIdx.x] = c;
		}
		__syncthreads();
		// fold: warp w owns hypothesis ks + w of this pass
		if (warp < nsub) {
			double v = 0.0, s = 0.0;
			int c = 0;
#pragma unroll
			for (int w = 0; w < kThreads / 32; ++w) {
				v = add(v, s_v[warp][w * 32 + lane]);
				if (HAS_CP) s = add(s, s_s[warp][w * 32 + lane]);
				c += s_c[warp][w * 32 + lane];
			}
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) {
				v = add(v, __shfl_xor_sync(0xffffffffu, v, o));
				if (HAS_CP) s = add(s, __shfl_xor_sync(0xffffffffu, s, o));
				c += __shfl_xor_sync(0xffffffffu, c, o);
			}
			if (lane == 0) {
				ScorePartial out;
				out.value = v;
				out.shared = s;
				out.count = c;
				partials[(k0 + ks + warp) * nchunks + chunk] = out;
			}
		}
		__syncthreads();
	}
}

__global__ void k_score_finalize(const ScorePartial *__restrict__ partials, int64_t K, int nchunks,
                                 int64_t *__restrict__ count, double *__restrict__ value, double *__restrict__ shared) {
	const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= K) return;
	double v = 0.0, s = 0.0;
	long long c = 0;
	for (int j = 0; j < nchunks; ++j) {
		const ScorePartial p = partials[k * nchunks + j];
		v = add(v, p.value);
		s = add(s, p.shared);
		c += p.count;
	}
	count[k] = c;
	value[k] = v;
	shared[k] = s;
}

template <int TYPE>
static void launch_partial(pxb_ctx *ctx, dim3 grid, const double *m, int64_t kk, double T2, const double *cp,
                           ScorePartial *pp, int nchunks) {
	const Points &p = ctx->pts;
	const bool hi = threshold_low_word_is_zero(T2);
	if (cp && hi)
		k_score_partial<TYPE, true, true><<<grid, kThreads, 0, ctx->stream>>>(p.soa, p.stride, p.N, m, kk, T2, cp, pp, nchunks);
	else if (cp)
		k_score_partial<TYPE, true, false><<<grid, kThreads, 0, ctx->stream>>>(p.soa, p.stride, p.N, m, kk, T2, cp, pp, nchunks);
	else if (hi)
		k_score_partial<TYPE, false, true><<<grid, kThreads, 0, ctx->stream>>>(p.soa, p.stride, p.N, m, kk, T2, cp, pp, nchunks);
	else
		k_score_partial<TYPE, false, false><<<grid, kThreads, 0, ctx->stream>>>(p.soa, p.stride, p.N, m, kk, T2, cp, pp, nchunks);
}

int launch_score_compound(pxb_ctx *ctx, const double *models, int64_t K, double T2, const double *compound_pref,
                          int64_t *count, double *value_sum, double *shared) {
	if (K <= 0) return PXB_OK;
	const Points &p = ctx->pts;
	const int nchunks = (int)((p.N + kScChunk - 1) / kScChunk);
	PXB_TRY(ctx->partials.reserve(sizeof(ScorePartial) * (size_t)K * nchunks));
	ScorePartial *part = ctx->partials.as<ScorePartial>();
	const int ms = model_size(p.type);
	int64_t done = 0;
	while (done < K) { // gridDim.y is limited to 65535
		const int64_t kk = std::min<int64_t>(K - done, (int64_t)65535 * kScHypsPerBlock);
		dim3 grid((unsigned)nchunks, (unsigned)((kk + kScHypsPerBlock - 1) / kScHypsPerBlock));
		const double *m = models + done * ms;
		ScorePartial *pp = part + done * nchunks;
		switch (p.type) {
		case PXB_MODEL_HOMOGRAPHY: launch_partial<PXB_MODEL_HOMOGRAPHY>(ctx, grid, m, kk, T2, compound_pref, pp, nchunks); break;
		case PXB_MODEL_FUNDAMENTAL: launch_partial<PXB_MODEL_FUNDAMENTAL>(ctx, grid, m, kk, T2, compound_pref, pp, nchunks); break;
		default: launch_partial<PXB_MODEL_PNP>(ctx, grid, m, kk, T2, compound_pref, pp, nchunks); break;
		}
		ctx->launches++;
		done += kk;
	}
	k_score_finalize<<<(unsigned)((K + 127) / 128), 128, 0, ctx->stream>>>(part, K, nchunks, count, value_sum, shared);
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

} // namespace pxb
