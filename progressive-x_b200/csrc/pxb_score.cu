// pxb_score.cu -- a4: the fused compound-aware MSAC score (no residual matrix materialised).
//
// Replaces MSACScoringFunctionWithCompoundModel::getScore (px/include/scoring_function_with_compound_model.h:61-125)
// for K hypotheses at once: per hypothesis the inlier count, sum max(0, 1 - r2/T2) and the support shared with the
// compound preference vector, sum min(compound_pref_i, pref_i). All three only involve the INLIERS of a hypothesis.
//
// Shape of the kernel
//   block = 256 threads = 8 warps x 128 points (lane owns base + lane + 32 j, j < 4) x a tile of 32 hypotheses.
//   1. Screening (hot loop, float32, pxb_screen.cuh): every (point, hypothesis) pair is tested for "certainly not an
//      inlier" in normalised coordinates -- ~16 FP32 instructions instead of ~28 FP64 ones. Pairs that cannot be
//      dismissed (inliers plus a band of up to 2 thresholds around the model, plus anything numerically suspicious)
//      are pushed into a per-warp queue in shared memory, in (hypothesis, ascending point) order.
//   2. Exact path (cold, float64): whenever 32 candidates are queued the warp drains them with all lanes busy: one
//      candidate per lane, the reference's exact residual (plain div.rn.f64), the inlier test r2 < T2, the MSAC terms.
//      Lane h then adds the terms of hypothesis h of the tile SEQUENTIALLY in queue order.
//
// Summation topology (a function of N only -- never of K, the batch split, the tile or the grid):
//   inliers of a 128-point chunk in ascending point order, sequentially (exactly the reference's loop order)
//   -> the 8 chunks of a block in order -> blocks (1024 points) in order (k_score_finalize).
// Counts are exact integers. Skipped pairs contribute exactly nothing in the reference as well (r2 >= T2).
#include <algorithm>

#include "pxb_internal.h"
#include "pxb_screen.cuh"

namespace pxb {

constexpr int kScP = 4;                   // points per lane
constexpr int kScWarps = kThreads / 32;   // 8
constexpr int kScChunk = kThreads * kScP; // 1024 points per block
constexpr int kScHyps = 32;               // hypotheses per tile = lanes (lane h accumulates hypothesis h)
constexpr int kScQueue = 256;             // circular candidate queue per warp (>= 31 + 128)
constexpr unsigned kEmptySlotTag = 0x7fc0e000u; // first float of an unfilled solver slot's screening model

struct ScorePartial {
	double value, shared;
	long long count;
};

struct ScoreAcc {
	double v, s;
	int c;
};

// Shared-memory plan of one block (dynamic): the block's 1024 points in float64 (the exact path gathers from here),
// optionally their compound preferences, the tile's models in float64 and in normalised float32, the per-warp
// candidate queues and the per-warp accumulators.
template <int TYPE, bool HAS_CP> struct ScoreSmem {
	static constexpr int DIM = ModelTraits<TYPE>::kDim, MP = ModelTraits<TYPE>::kPadded, MF = ScreenTraits<TYPE>::kFloats;
	static constexpr size_t kPts = 0;
	static constexpr size_t kCp = kPts + sizeof(double) * DIM * kScChunk;
	static constexpr size_t kModels = kCp + (HAS_CP ? sizeof(double) * kScChunk : 0);
	static constexpr size_t kMf = kModels + sizeof(double) * kScHyps * MP;
	static constexpr size_t kQueue = kMf + sizeof(float) * kScHyps * MF;
	static constexpr size_t kAcc = kQueue + sizeof(unsigned short) * kScWarps * kScQueue;
	static constexpr size_t kRes = kAcc + sizeof(ScoreAcc) * kScWarps * kScHyps; // per warp: 32 values (+32 shared) + 32 segments
	static constexpr size_t kSeg = kRes + sizeof(double) * kScWarps * 32 * (HAS_CP ? 2 : 1);
	static constexpr size_t kEmpty = kSeg + sizeof(unsigned short) * kScWarps * 32; // per hypothesis of the tile: empty slot
	static constexpr size_t kBytes = kEmpty + kScHyps + 16; // slot map of the compacted tile + its size
};

// Drains up to 32 queued candidates of one warp: one candidate per lane, the reference's exact float64 residual, the
// inlier test and the MSAC terms. The queue is sorted by (hypothesis, point), so the candidates of hypothesis h form
// one contiguous segment; lane h then adds the terms of its segment SEQUENTIALLY (ascending point index). Candidates
// that turn out not to be inliers contribute an exact +0.0.
template <int TYPE, bool HAS_CP>
__device__ __noinline__ ScoreAcc score_drain(const double *s_pts, const double *s_cp, const double *s_models, int pbase,
                                             double T2, const unsigned short *queue, int head, int n, double *res_v,
                                             double *res_s, unsigned short *seg, ScoreAcc acc) {
	constexpr int DIM = ModelTraits<TYPE>::kDim, MP = ModelTraits<TYPE>::kPadded;
	const int lane = threadIdx.x & 31;
	int qh = 255;
	bool inl = false;
	double sv = 0.0, sh = 0.0;
	if (lane < n) {
		const unsigned e = queue[(head + lane) & (kScQueue - 1)];
		qh = (int)(e >> 7);
		const int idx = pbase + (int)(e & 127u);
		double p[5];
#pragma unroll
		for (int c = 0; c < DIM; ++c) p[c] = s_pts[c * kScChunk + idx];
		const double *m = s_models + qh * MP;
		bool ok = true;
		double r2 = squared_residual_fast<TYPE>(p, m, ok); // bit-identical to the plain division inside its domain
		if (!ok) r2 = squared_residual<TYPE>(p, m);
		inl = r2 < T2; // scoring_function_with_compound_model.h:85
		if (inl) {
			sv = cv_max(0.0, sub(1.0, divd(r2, T2)));   // :96-99
			if (HAS_CP) sh = cv_min(s_cp[idx], sv);     // :115-117 (pref is 0 off the inlier set)
		}
	}
	const unsigned inliers = __ballot_sync(0xffffffffu, inl);
	if (inliers == 0) return acc;
	// segment of every hypothesis present in this batch: written by the first lane of the segment, read by lane h
	const int prev = __shfl_up_sync(0xffffffffu, qh, 1);
	const bool first = lane < n && (lane == 0 || prev != qh);
	const unsigned firsts = __ballot_sync(0xffffffffu, first);
	seg[lane] = 0;
	res_v[lane] = sv;
	if (HAS_CP) res_s[lane] = sh;
	__syncwarp();
	if (first) {
		const unsigned later = firsts & ~((2u << lane) - 1u);
		const int end = later ? (__ffs(later) - 1) : n;
		seg[qh] = (unsigned short)(lane | ((end - lane) << 8));
	}
	__syncwarp();
	const int start = seg[lane] & 255, len = seg[lane] >> 8;
	const int maxlen = __reduce_max_sync(0xffffffffu, len);
	acc.c += __popc(inliers & (len >= 32 ? 0xffffffffu : (((1u << len) - 1u) << start)));
#pragma unroll 4
	for (int k = 0; k < maxlen; ++k) {
		const int i = (start + k) & 31;
		const double v = res_v[i];
		const double s2 = HAS_CP ? res_s[i] : 0.0;
		if (k < len) {
			acc.v = add(acc.v, v);
			if (HAS_CP) acc.s = add(acc.s, s2);
		}
	}
	__syncwarp();
	return acc;
}

// The hypothesis loop of one warp: screening, queueing, draining.
template <int TYPE, bool HAS_CP, bool FULL>
__device__ __forceinline__ ScoreAcc score_tile(const float (&p)[kScP][5], const float (&zq)[kScP], const bool (&valid)[kScP],
                                               const float *s_mf, int nk, float cT, const double *s_pts, const double *s_cp,
                                               const double *s_models, int pbase, double T2, unsigned short *queue,
                                               double *res_v, double *res_s, unsigned short *seg) {
	constexpr int MF = ScreenTraits<TYPE>::kFloats;
	const int lane = threadIdx.x & 31;
	const unsigned lt = (1u << lane) - 1u;
	int head = 0, tail = 0;
	ScoreAcc acc = {0.0, 0.0, 0};
#pragma unroll 2
	for (int h = 0; h < nk; ++h) {
		float m[MF];
		const float4 *s4 = reinterpret_cast<const float4 *>(s_mf + h * MF);
#pragma unroll
		for (int i = 0; i < MF / 4; ++i) {
			const float4 v = s4[i];
			m[4 * i] = v.x, m[4 * i + 1] = v.y, m[4 * i + 2] = v.z, m[4 * i + 3] = v.w;
		}
		bool cand[kScP];
		bool any = false;
#pragma unroll
		for (int j = 0; j < kScP; ++j) {
			const bool sure = screen_sure_outlier<TYPE>(p[j], m, cT, zq[j]);
			cand[j] = FULL ? !sure : (!sure & valid[j]);
			any |= cand[j];
		}
		if (__any_sync(0xffffffffu, any)) {
#pragma unroll
			for (int j = 0; j < kScP; ++j) {
				const unsigned b = __ballot_sync(0xffffffffu, cand[j]);
				if (b == 0) continue; // warp-uniform
				if (cand[j]) queue[(tail + __popc(b & lt)) & (kScQueue - 1)] = (unsigned short)((h << 7) | (32 * j + lane));
				tail += __popc(b);
			}
			__syncwarp();
			while (tail - head >= 32) {
				acc = score_drain<TYPE, HAS_CP>(s_pts, s_cp, s_models, pbase, T2, queue, head, 32, res_v, res_s, seg, acc);
				head += 32;
			}
		}
	}
	if (tail > head)
		acc = score_drain<TYPE, HAS_CP>(s_pts, s_cp, s_models, pbase, T2, queue, head, tail - head, res_v, res_s, seg, acc);
	return acc;
}

// Per launch, before the score kernel: the normalised float32 copy of every hypothesis (kFloats each) and the two
// launch constants of the test. One thread per hypothesis; keeps all float64 conjugation out of the hot kernel.
template <int TYPE>
__global__ void k_screen_prepare(const double *__restrict__ models, int64_t K, double T2, const NormDev *__restrict__ norm,
                                 float *__restrict__ consts, float *__restrict__ mf) {
	constexpr int MS = ModelTraits<TYPE>::kSize, MF = ScreenTraits<TYPE>::kFloats;
	pdl_launch_dependents();
	pdl_wait();
	const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const NormDev nd = *norm;
	if (k == 0) {
		const ScreenConsts sc = screen_consts<TYPE>(T2, nd);
		consts[0] = sc.cT;
		consts[1] = sc.cE;
	}
	if (k >= K) return;
	float out[MF];
#pragma unroll
	for (int i = 0; i < MF; ++i) out[i] = 0.0f;
	screen_model<TYPE>(models + k * MS, nd, out);
#pragma unroll
	for (int i = 0; i < MF; ++i) mf[k * MF + i] = out[i];
	bool zero = true;
#pragma unroll
	for (int i = 0; i < MS; ++i) zero &= models[k * MS + i] == 0.0;
	// H, F, PnP: an all-zero model gives 0/0 residuals everywhere: the slot is tagged (a NaN payload in its first float) and
	// the kernels leave it out of their tiles. (A zero LINE has residual 0 -- all inliers -- and is scored.)
	if (zero && TYPE <= PXB_MODEL_PNP) mf[k * MF] = __uint_as_float(kEmptySlotTag);
}

template <int TYPE, bool HAS_CP, int PASSES>
__global__ void __launch_bounds__(kThreads, 3)
    k_score_screened(const double *__restrict__ soa, int64_t stride, int64_t N, const float *__restrict__ pf,
                     const float *__restrict__ pq, const float *__restrict__ consts, const float *__restrict__ mfg,
                     const double *__restrict__ models, int64_t K, double T2, const double *__restrict__ compound_pref,
                     ScorePartial *__restrict__ partials, int nchunks, int tile_arg /* hypotheses per pass (PASSES == 1 only) */) {
	using L = ScoreSmem<TYPE, HAS_CP>;
	constexpr int DIM = L::DIM, MS = ModelTraits<TYPE>::kSize, MP = L::MP, MF = L::MF;
	extern __shared__ __align__(16) unsigned char smem[];
	pdl_wait();
	double *s_pts = reinterpret_cast<double *>(smem + L::kPts);
	double *s_cp = reinterpret_cast<double *>(smem + L::kCp);
	double *s_models = reinterpret_cast<double *>(smem + L::kModels);
	float *s_mf = reinterpret_cast<float *>(smem + L::kMf);
	unsigned short *s_queue = reinterpret_cast<unsigned short *>(smem + L::kQueue);
	ScoreAcc *s_acc = reinterpret_cast<ScoreAcc *>(smem + L::kAcc);
	double *s_res = reinterpret_cast<double *>(smem + L::kRes);
	unsigned short *s_seg = reinterpret_cast<unsigned short *>(smem + L::kSeg);
	unsigned char *s_empty = smem + L::kEmpty;

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int chunk = blockIdx.x;
	const int64_t block_base = (int64_t)chunk * kScChunk;
	for (int t = threadIdx.x; t < kScChunk; t += kThreads) { // stride is a multiple of 64 >= N: rows never overrun
		const int64_t i = min(block_base + t, stride - 1);
#pragma unroll
		for (int c = 0; c < DIM; ++c) s_pts[c * kScChunk + t] = __ldg(soa + c * stride + i);
		if (HAS_CP) s_cp[t] = (block_base + t < N) ? __ldg(compound_pref + block_base + t) : 0.0;
	}
	const float cT = __ldg(consts), cE = __ldg(consts + 1);
	const int64_t warp_base = block_base + warp * (32 * kScP);
	float p[kScP][5], zq[kScP];
	bool valid[kScP];
#pragma unroll
	for (int j = 0; j < kScP; ++j) {
		const int64_t i = warp_base + lane + 32 * j;
		valid[j] = i < N;
		const int64_t ii = valid[j] ? i : (N - 1);
#pragma unroll
		for (int c = 0; c < DIM; ++c) p[j][c] = __ldg(pf + c * stride + ii);
		zq[j] = cE * __ldg(pq + ii); // error allowance of this point
	}
	double *res_v = s_res + warp * 32 * (HAS_CP ? 2 : 1), *res_s = res_v + (HAS_CP ? 32 : 0);
	const bool full = warp_base + 32 * kScP <= N; // interior warps run without validity predicates (warp-uniform)

	const int tile = PASSES == 1 ? tile_arg : kScHyps; // big batches always walk full tiles
	for (int pass = 0; pass < PASSES; ++pass) {
		const int64_t k0 = ((int64_t)blockIdx.y * PASSES + pass) * tile;
		if (k0 >= K) break; // block-uniform
		const int nk = (int)min((int64_t)tile, K - k0);
		// Unfilled solution slots of a minimal solver (all-zero models: every residual is NaN, nothing is an inlier; count,
		// score and shared support are exactly zero, as the reference's loop would find) never enter the hypothesis loop:
		// for the families whose solvers return several solutions per sample (F: up to 3, PnP: up to 4) the tile is
		// compacted while it is staged, s_empty[j] = original slot of compacted hypothesis j. The one-solution families
		// keep the straight staging (the extra bookkeeping costs registers in the hot loop: measured -9 % on the H grid).
		constexpr bool COMPACT = TYPE == PXB_MODEL_FUNDAMENTAL || TYPE == PXB_MODEL_PNP;
		int nc = nk;
		if (COMPACT) {
			__syncthreads(); // the previous pass has read s_empty / s_acc
			if (warp == 0) {
				const bool filled = lane < nk && __float_as_uint(__ldg(mfg + (k0 + lane) * MF)) != kEmptySlotTag;
				const unsigned bits = __ballot_sync(0xffffffffu, filled);
				if (filled) s_empty[__popc(bits & ((1u << lane) - 1u))] = (unsigned char)lane;
				if (lane == 0) s_empty[kScHyps] = (unsigned char)__popc(bits);
				if (lane < nk && !filled) partials[(k0 + lane) * nchunks + chunk] = ScorePartial{0.0, 0.0, 0};
			}
			__syncthreads();
			nc = s_empty[kScHyps];
			for (int t = threadIdx.x; t < nc * MS; t += kThreads)
				s_models[(t / MS) * MP + (t % MS)] = models[(k0 + s_empty[t / MS]) * MS + (t % MS)];
			for (int t = threadIdx.x; t < nc * MF; t += kThreads) s_mf[t] = __ldg(mfg + (k0 + s_empty[t / MF]) * MF + (t % MF));
		} else {
			for (int t = threadIdx.x; t < nk * MS; t += kThreads) s_models[(t / MS) * MP + (t % MS)] = models[k0 * MS + t];
			for (int t = threadIdx.x; t < nk * MF; t += kThreads) s_mf[t] = __ldg(mfg + k0 * MF + t);
		}
		__syncthreads();
		ScoreAcc acc;
		if (full)
			acc = score_tile<TYPE, HAS_CP, true>(p, zq, valid, s_mf, nc, cT, s_pts, s_cp, s_models, warp * (32 * kScP), T2,
			                                     s_queue + warp * kScQueue, res_v, res_s, s_seg + warp * 32);
		else
			acc = score_tile<TYPE, HAS_CP, false>(p, zq, valid, s_mf, nc, cT, s_pts, s_cp, s_models, warp * (32 * kScP), T2,
			                                      s_queue + warp * kScQueue, res_v, res_s, s_seg + warp * 32);
		s_acc[warp * kScHyps + lane] = acc;
		__syncthreads();
		if (threadIdx.x < nc) { // the 8 chunks of the block, in order
			ScorePartial out = {0.0, 0.0, 0};
#pragma unroll
			for (int w = 0; w < kScWarps; ++w) {
				const ScoreAcc a = s_acc[w * kScHyps + threadIdx.x];
				out.value = add(out.value, a.v);
				out.shared = add(out.shared, a.s);
				out.count += a.c;
			}
			partials[(k0 + (COMPACT ? (int)s_empty[threadIdx.x] : (int)threadIdx.x)) * nchunks + chunk] = out;
		}
	}
}

__global__ void k_score_finalize(const ScorePartial *__restrict__ partials, int64_t K, int nchunks,
                                 int64_t *__restrict__ count, double *__restrict__ value, double *__restrict__ shared) {
	pdl_wait();
	const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= K) return;
	double v = 0.0, s = 0.0;
	long long c = 0;
	// eight partials in flight per thread, added in chunk order (one load -> add chain per chunk cost 30 us for the 98 chunks
	// of N = 10^5)
	const ScorePartial *row = partials + k * nchunks;
	int j = 0;
	for (; j + 8 <= nchunks; j += 8) {
		ScorePartial p[8];
#pragma unroll
		for (int u = 0; u < 8; ++u) p[u] = row[j + u];
#pragma unroll
		for (int u = 0; u < 8; ++u) {
			v = add(v, p[u].value);
			s = add(s, p[u].shared);
			c += p[u].count;
		}
	}
	for (; j < nchunks; ++j) {
		const ScorePartial p = row[j];
		v = add(v, p.value);
		s = add(s, p.shared);
		c += p.count;
	}
	count[k] = c;
	value[k] = v;
	shared[k] = s;
}

template <int TYPE, bool HAS_CP, int PASSES>
static void launch_screened(pxb_ctx *ctx, int nchunks, const float *consts, const float *mf, const double *m, int64_t kk,
                            double T2, const double *cp, ScorePartial *pp, int tile_hyps) {
	const Points &p = ctx->pts;
	constexpr int kBytes = (int)ScoreSmem<TYPE, HAS_CP>::kBytes;
	// opt in to > 48 KB of dynamic shared memory: once per device and instantiation (every API call counts when eight
	// host threads drive one GPU)
	static bool opted_in[64] = {};
	if (ctx->device < 0 || ctx->device >= 64 || !opted_in[ctx->device]) {
		cudaFuncSetAttribute(k_score_screened<TYPE, HAS_CP, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBytes);
		if (ctx->device >= 0 && ctx->device < 64) opted_in[ctx->device] = true;
	}
	const int64_t tile = (int64_t)tile_hyps * PASSES;
	dim3 grid((unsigned)nchunks, (unsigned)((kk + tile - 1) / tile));
	(void)launch_pdl(k_score_screened<TYPE, HAS_CP, PASSES>, grid, dim3(kThreads), (size_t)kBytes, ctx->stream, p.soa, p.stride, p.N, p.f32n,
	                 p.q, consts, mf, m, kk, T2, cp, pp, nchunks, tile_hyps);
}

// ------------------------------------------------------------------------------------------------
// a1-a3, bit-matrix only: the inlier mask of every (hypothesis, point) pair WITHOUT the r2 matrix
// ------------------------------------------------------------------------------------------------
// Same screening as the score kernel: float32 proves "r2 >= T2" for most pairs (bit 0, nothing else to do); the pairs it
// cannot dismiss are queued per warp and drained 32 at a time through the reference's float64 residual, whose exact
// comparison r2 < T2 sets the bit in a shared-memory tile [hypothesis][word of 32 points]; the tile goes to global memory
// once per block. The mask is bit-identical to the one k_residual_matrix writes (tests/test_gpu_parity.py).
template <int TYPE> struct MaskSmem {
	static constexpr int DIM = ModelTraits<TYPE>::kDim, MP = ModelTraits<TYPE>::kPadded, MF = ScreenTraits<TYPE>::kFloats;
	static constexpr size_t kPts = 0;
	static constexpr size_t kModels = kPts + sizeof(double) * DIM * kScChunk;
	static constexpr size_t kMf = kModels + sizeof(double) * kScHyps * MP;
	static constexpr size_t kQueue = kMf + sizeof(float) * kScHyps * MF;
	static constexpr size_t kMask = kQueue + sizeof(unsigned short) * kScWarps * kScQueue;
	static constexpr size_t kEmpty = kMask + sizeof(uint32_t) * kScHyps * (kScChunk / 32);
	static constexpr size_t kBytes = kEmpty + kScHyps + 16; // slot map of the compacted tile + its size
};

template <int TYPE>
__device__ __forceinline__ void mask_drain(const double *s_pts, const double *s_models, int pbase, double T2,
                                           const unsigned short *queue, int head, int n, uint32_t *s_mask, int word0) {
	constexpr int DIM = ModelTraits<TYPE>::kDim, MP = ModelTraits<TYPE>::kPadded;
	const int lane = threadIdx.x & 31;
	if (lane < n) {
		const unsigned e = queue[(head + lane) & (kScQueue - 1)];
		const int qh = (int)(e >> 7), local = (int)(e & 127u);
		double p[5];
#pragma unroll
		for (int c = 0; c < DIM; ++c) p[c] = s_pts[c * kScChunk + pbase + local];
		const double *m = s_models + qh * MP;
		bool ok = true;
		double r2 = squared_residual_fast<TYPE>(p, m, ok); // bit-identical to the plain division inside its domain
		if (!ok) r2 = squared_residual<TYPE>(p, m);
		if (r2 < T2) atomicOr(&s_mask[qh * (kScChunk / 32) + word0 + (local >> 5)], 1u << (local & 31));
	}
	__syncwarp();
}

template <int TYPE, bool FULL>
__device__ __forceinline__ void mask_tile(const float (&p)[kScP][5], const float (&zq)[kScP], const bool (&valid)[kScP],
                                          const float *s_mf, int nk, float cT, const double *s_pts, const double *s_models,
                                          int pbase, double T2, unsigned short *queue, uint32_t *s_mask, int word0) {
	constexpr int MF = ScreenTraits<TYPE>::kFloats;
	const int lane = threadIdx.x & 31;
	const unsigned lt = (1u << lane) - 1u;
	int head = 0, tail = 0;
#pragma unroll 2
	for (int h = 0; h < nk; ++h) {
		float m[MF];
		const float4 *s4 = reinterpret_cast<const float4 *>(s_mf + h * MF);
#pragma unroll
		for (int i = 0; i < MF / 4; ++i) {
			const float4 v = s4[i];
			m[4 * i] = v.x, m[4 * i + 1] = v.y, m[4 * i + 2] = v.z, m[4 * i + 3] = v.w;
		}
		bool cand[kScP];
		bool any = false;
#pragma unroll
		for (int j = 0; j < kScP; ++j) {
			const bool sure = screen_sure_outlier<TYPE>(p[j], m, cT, zq[j]);
			cand[j] = FULL ? !sure : (!sure & valid[j]);
			any |= cand[j];
		}
		if (__any_sync(0xffffffffu, any)) {
#pragma unroll
			for (int j = 0; j < kScP; ++j) {
				const unsigned b = __ballot_sync(0xffffffffu, cand[j]);
				if (b == 0) continue; // warp-uniform
				if (cand[j]) queue[(tail + __popc(b & lt)) & (kScQueue - 1)] = (unsigned short)((h << 7) | (32 * j + lane));
				tail += __popc(b);
			}
			__syncwarp();
			while (tail - head >= 32) {
				mask_drain<TYPE>(s_pts, s_models, pbase, T2, queue, head, 32, s_mask, word0);
				head += 32;
			}
		}
	}
	if (tail > head) mask_drain<TYPE>(s_pts, s_models, pbase, T2, queue, head, tail - head, s_mask, word0);
}

template <int TYPE>
__global__ void __launch_bounds__(kThreads, 3)
    k_mask_screened(const double *__restrict__ soa, int64_t stride, int64_t N, const float *__restrict__ pf,
                    const float *__restrict__ pq, const float *__restrict__ consts, const float *__restrict__ mfg,
                    const double *__restrict__ models, int64_t K, double T2, uint32_t *__restrict__ mask, int64_t words) {
	using L = MaskSmem<TYPE>;
	constexpr int DIM = L::DIM, MS = ModelTraits<TYPE>::kSize, MP = L::MP, MF = L::MF, WPB = kScChunk / 32;
	extern __shared__ __align__(16) unsigned char smem[];
	double *s_pts = reinterpret_cast<double *>(smem + L::kPts);
	double *s_models = reinterpret_cast<double *>(smem + L::kModels);
	float *s_mf = reinterpret_cast<float *>(smem + L::kMf);
	unsigned short *s_queue = reinterpret_cast<unsigned short *>(smem + L::kQueue);
	uint32_t *s_mask = reinterpret_cast<uint32_t *>(smem + L::kMask);
	unsigned char *s_empty = smem + L::kEmpty;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int chunk = blockIdx.x;
	const int64_t block_base = (int64_t)chunk * kScChunk;
	for (int t = threadIdx.x; t < kScChunk; t += kThreads) {
		const int64_t i = min(block_base + t, stride - 1);
#pragma unroll
		for (int c = 0; c < DIM; ++c) s_pts[c * kScChunk + t] = __ldg(soa + c * stride + i);
	}
	const int64_t k0 = (int64_t)blockIdx.y * kScHyps;
	const int nk = (int)min((int64_t)kScHyps, K - k0);
	if (warp == 0) { // compact the tile: all-zero models set no bit (every residual is NaN) and never enter the loop
		const bool filled = lane < nk && __float_as_uint(__ldg(mfg + (k0 + lane) * MF)) != kEmptySlotTag;
		const unsigned bits = __ballot_sync(0xffffffffu, filled);
		if (filled) s_empty[__popc(bits & ((1u << lane) - 1u))] = (unsigned char)lane;
		if (lane == 0) s_empty[kScHyps] = (unsigned char)__popc(bits);
	}
	for (int t = threadIdx.x; t < kScHyps * WPB; t += kThreads) s_mask[t] = 0u;
	__syncthreads();
	const int nc = s_empty[kScHyps];
	for (int t = threadIdx.x; t < nc * MS; t += kThreads) s_models[(t / MS) * MP + (t % MS)] = models[(k0 + s_empty[t / MS]) * MS + (t % MS)];
	for (int t = threadIdx.x; t < nc * MF; t += kThreads) s_mf[t] = __ldg(mfg + (k0 + s_empty[t / MF]) * MF + (t % MF));
	const float cT = __ldg(consts), cE = __ldg(consts + 1);
	const int64_t warp_base = block_base + warp * (32 * kScP);
	float p[kScP][5], zq[kScP];
	bool valid[kScP];
#pragma unroll
	for (int j = 0; j < kScP; ++j) {
		const int64_t i = warp_base + lane + 32 * j;
		valid[j] = i < N;
		const int64_t ii = valid[j] ? i : (N - 1);
#pragma unroll
		for (int c = 0; c < DIM; ++c) p[j][c] = __ldg(pf + c * stride + ii);
		zq[j] = cE * __ldg(pq + ii);
	}
	__syncthreads();
	if (warp_base + 32 * kScP <= N)
		mask_tile<TYPE, true>(p, zq, valid, s_mf, nc, cT, s_pts, s_models, warp * (32 * kScP), T2, s_queue + warp * kScQueue, s_mask,
		                      warp * kScP);
	else
		mask_tile<TYPE, false>(p, zq, valid, s_mf, nc, cT, s_pts, s_models, warp * (32 * kScP), T2, s_queue + warp * kScQueue, s_mask,
		                       warp * kScP);
	__syncthreads();
	const int64_t w0 = (int64_t)chunk * WPB;
	// rows of the compacted hypotheses go to their original slots; the rows of empty slots are zero
	for (int t = threadIdx.x; t < nk * WPB; t += kThreads) {
		const int h = t / WPB, w = t % WPB;
		if (w0 + w < words) mask[(k0 + h) * words + w0 + w] = 0u;
	}
	__syncthreads();
	for (int t = threadIdx.x; t < nc * WPB; t += kThreads) {
		const int h = t / WPB, w = t % WPB;
		if (w0 + w < words) mask[(k0 + s_empty[h]) * words + w0 + w] = s_mask[h * WPB + w];
	}
}

template <int TYPE> static int launch_mask_t(pxb_ctx *ctx, const double *models, int64_t K, double T2, uint32_t *mask) {
	constexpr int MF = ScreenTraits<TYPE>::kFloats;
	constexpr int kBytes = (int)MaskSmem<TYPE>::kBytes;
	const Points &p = ctx->pts;
	const int nchunks = (int)((p.N + kScChunk - 1) / kScChunk);
	const int64_t words = (p.N + 31) / 32;
	static bool opted_in[64] = {};
	if (ctx->device < 0 || ctx->device >= 64 || !opted_in[ctx->device]) {
		cudaFuncSetAttribute(k_mask_screened<TYPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBytes);
		if (ctx->device >= 0 && ctx->device < 64) opted_in[ctx->device] = true;
	}
	int64_t done = 0;
	while (done < K) { // gridDim.y is limited to 65535
		const int64_t kk = std::min<int64_t>(K - done, (int64_t)65535 * kScHyps);
		const double *m = models + done * ModelTraits<TYPE>::kSize;
		PXB_TRY(ctx->screen.reserve(sizeof(float) * (4 + (size_t)kk * MF)));
		float *consts = ctx->screen.as<float>(), *mf = consts + 4;
		PXB_CUDA(launch_pdl(k_screen_prepare<TYPE>, dim3((unsigned)((kk + 127) / 128)), dim3(128), 0, ctx->stream, m, kk, T2, p.norm, consts, mf));
		dim3 grid((unsigned)nchunks, (unsigned)((kk + kScHyps - 1) / kScHyps));
		k_mask_screened<TYPE><<<grid, kThreads, kBytes, ctx->stream>>>(p.soa, p.stride, p.N, p.f32n, p.q, consts, mf, m, kk, T2,
		                                                              mask + done * words, words);
		ctx->launches += 2;
		done += kk;
	}
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

// inlier bit matrix only (mask[k * words + (i >> 5)], bit i & 31), float32-screened
int launch_inlier_mask(pxb_ctx *ctx, const double *models, int64_t K, double T2, uint32_t *mask) {
	if (K <= 0) return PXB_OK;
	int rc = PXB_OK;
	PXB_DISPATCH_TYPE(ctx->pts.type, rc = launch_mask_t<TYPE>(ctx, models, K, T2, mask));
	return rc;
}

template <int TYPE>
static int launch_partial(pxb_ctx *ctx, int nchunks, const double *m, int64_t kk, double T2, const double *cp, ScorePartial *pp) {
	constexpr int MF = ScreenTraits<TYPE>::kFloats;
	PXB_TRY(ctx->screen.reserve(sizeof(float) * (4 + (size_t)kk * MF)));
	float *consts = ctx->screen.as<float>(), *mf = consts + 4;
	PXB_CUDA(launch_pdl(k_screen_prepare<TYPE>, dim3((unsigned)((kk + 127) / 128)), dim3(128), 0, ctx->stream, m, kk, T2, ctx->pts.norm, consts, mf));
	ctx->launches++;
	// big batches walk 4 tiles of 32 hypotheses per block (the per-block prologue -- staging 1024 points -- is paid once);
	// RANSAC-sized batches keep one tile per block so that the grid still fills the GPU
	const bool big = kk * nchunks >= (int64_t)4 * 32 * 3 * ctx->sm_count * 2;
	// small batches (the <= 50 refits of a local-optimisation step, the single model of a least-squares step): fewer
	// hypotheses per block so that the grid still covers the SMs -- the hypothesis loop of a warp is sequential, and these
	// models are inlier rich (every inlier takes the exact float64 path). The sums do not depend on the tile.
	int tile = kScHyps;
	if (!big) {
		const int64_t want_blocks = 2 * (int64_t)ctx->sm_count;
		tile = (int)std::min<int64_t>(kScHyps, std::max<int64_t>(1, kk * nchunks / want_blocks));
	}
	if (cp) {
		if (big) launch_screened<TYPE, true, 4>(ctx, nchunks, consts, mf, m, kk, T2, cp, pp, kScHyps);
		else launch_screened<TYPE, true, 1>(ctx, nchunks, consts, mf, m, kk, T2, cp, pp, tile);
	} else {
		if (big) launch_screened<TYPE, false, 4>(ctx, nchunks, consts, mf, m, kk, T2, cp, pp, kScHyps);
		else launch_screened<TYPE, false, 1>(ctx, nchunks, consts, mf, m, kk, T2, cp, pp, tile);
	}
	ctx->launches++;
	return PXB_OK;
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
		const int64_t kk = std::min<int64_t>(K - done, (int64_t)65535 * kScHyps);
		const double *m = models + done * ms;
		ScorePartial *pp = part + done * nchunks;
		int rc = PXB_OK;
		PXB_DISPATCH_TYPE(p.type, rc = launch_partial<TYPE>(ctx, nchunks, m, kk, T2, compound_pref, pp));
		PXB_TRY(rc);
		done += kk;
	}
	PXB_CUDA(launch_pdl(k_score_finalize, dim3((unsigned)((K + 127) / 128)), dim3(128), 0, ctx->stream, part, K, nchunks, count, value_sum, shared));
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

} // namespace pxb
