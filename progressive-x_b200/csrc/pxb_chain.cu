// pxb_chain.cu -- device-side glue that keeps the host driver's loops to ONE round trip per iteration: index lists are
// built, sampled and consumed on the device instead of travelling to the host and back.
//
// PEARL iteration (px/include/PEARL.h:319-401, parameterEstimation): labeling -> per-instance point lists -> residual
// sums -> non-minimal refits -> residual sums of the refits is one stream-ordered chain with a single small copy at its
// end (it used to take four round trips: labels down, sums, index lists up + fits, sums again).
// GC-RANSAC local optimisation (gcr/GCRANSAC.h:781-911): labelling -> inlier list -> the <= 50 inner-RANSAC samples ->
// fits -> scores, one copy back (two round trips and an N-byte copy before).
// Least-squares tail (gcr/GCRANSAC.h:561-618, 631-759): inlier list of the current model -> Tukey weights -> weighted
// fit -> score, one copy back (three round trips per iteration before).
//
//   k_label_lists     per-instance point lists from the label array: block l writes the indices of the points with
//                     label l in ascending order to idx[off[l] ...] (the order in which PEARL.h:342-352 collects them),
//                     off[l] = number of points with a label below l
//   k_select_models   cand[l] = ok[l] ? fitted[l] : current[l]   (PEARL.h:381-391 evaluates the refit only where it succeeded)
#include "pxb_internal.h"

namespace pxb {

constexpr int kListThreads = 1024;

__global__ void __launch_bounds__(kListThreads)
    k_label_lists(const int32_t *__restrict__ labels, int64_t N, int L, int32_t *__restrict__ off /*L+1*/,
                  int32_t *__restrict__ idx) {
	__shared__ int s_warp[32];
	__shared__ int s_a, s_b;
	const int l = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	// pass 1: how many points carry a label below l / equal to l
	int below = 0, mine = 0;
	for (int64_t i = tid; i < N; i += kListThreads) {
		const int v = labels[i];
		below += (v >= 0 && v < l);
		mine += (v == l);
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		below += __shfl_xor_sync(0xffffffffu, below, o);
		mine += __shfl_xor_sync(0xffffffffu, mine, o);
	}
	if (tid == 0) s_a = 0, s_b = 0;
	__syncthreads();
	if (lane == 0) {
		atomicAdd(&s_a, below);
		atomicAdd(&s_b, mine);
	}
	__syncthreads();
	const int base = s_a;
	if (tid == 0) {
		off[l] = base;
		if (l == L - 1) off[L] = base + s_b;
	}
	// pass 2: ordered compaction, one chunk of 1024 points per step
	int running = base;
	for (int64_t c0 = 0; c0 < N; c0 += kListThreads) {
		const int64_t i = c0 + tid;
		const bool pred = i < N && labels[i] == l;
		const unsigned b = __ballot_sync(0xffffffffu, pred);
		if (lane == 0) s_warp[warp] = __popc(b);
		__syncthreads();
		if (warp == 0) {
			const int v = s_warp[lane];
			int inc = v;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const int t = __shfl_up_sync(0xffffffffu, inc, o);
				if (lane >= o) inc += t;
			}
			s_warp[lane] = inc - v; // exclusive prefix of the warp counts
			if (lane == 31) s_a = inc; // chunk total
		}
		__syncthreads();
		if (pred) idx[running + s_warp[warp] + __popc(b & ((1u << lane) - 1u))] = (int32_t)i;
		running += s_a;
		__syncthreads();
	}
}

__global__ void k_select_models(const double *__restrict__ current, const double *__restrict__ fitted,
                                const int32_t *__restrict__ ok, int L, int ms, double *__restrict__ cand) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= L * ms) return;
	cand[t] = ok[t / ms] ? fitted[t] : current[t];
}

// ---- ordered compaction of a 0/1 byte array (the LO labelling: SINK = inlier) or of an inlier bit mask -------------------
// One block; indices ascending, the order in which the reference walks the points (GCRANSAC.h:1006-1016,
// scoring_function_with_compound_model.h:85-95). count_out[0] receives the number of entries.
template <bool BITS>
__global__ void __launch_bounds__(kListThreads)
    k_flag_compact(const void *__restrict__ flags, int64_t N, int32_t *__restrict__ idx, int64_t *__restrict__ count_out,
                   int32_t *__restrict__ off2 /* optional: off2[0] = 0, off2[1] = count (a one-problem CSR) */) {
	pdl_launch_dependents();
	pdl_wait();
	__shared__ int s_warp[32];
	__shared__ int s_total;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	int running = 0;
	// per thread and step: one 32-bit word of the bit mask, or eight consecutive flag bytes (one 64-bit load)
	const int64_t items = BITS ? (N + 31) / 32 : (N + 7) / 8;
	for (int64_t c0 = 0; c0 < items; c0 += kListThreads) {
		const int64_t i = c0 + tid;
		int mine;
		unsigned word = 0; // bit j set: entry base + j is flagged
		if (BITS) {
			word = i < items ? reinterpret_cast<const uint32_t *>(flags)[i] : 0u;
		} else if (i < items) {
			const uint8_t *f = reinterpret_cast<const uint8_t *>(flags) + i * 8;
			if (i * 8 + 8 <= N && (reinterpret_cast<uintptr_t>(f) & 7u) == 0) {
				const unsigned long long v = *reinterpret_cast<const unsigned long long *>(f);
#pragma unroll
				for (int j = 0; j < 8; ++j) word |= ((v >> (8 * j)) & 0xffull) ? (1u << j) : 0u;
			} else {
				for (int j = 0; j < 8 && i * 8 + j < N; ++j) word |= f[j] ? (1u << j) : 0u;
			}
		}
		mine = __popc(word);
		int inc = mine; // inclusive scan over the warp, then over the warp totals
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const int t = __shfl_up_sync(0xffffffffu, inc, o);
			if (lane >= o) inc += t;
		}
		if (lane == 31) s_warp[warp] = inc;
		__syncthreads();
		if (warp == 0) {
			const int v = s_warp[lane];
			int w = v;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const int t = __shfl_up_sync(0xffffffffu, w, o);
				if (lane >= o) w += t;
			}
			s_warp[lane] = w - v;
			if (lane == 31) s_total = w;
		}
		__syncthreads();
		int at = running + s_warp[warp] + inc - mine;
		const int64_t base = i * (BITS ? 32 : 8);
		while (word) {
			const int bpos = __ffs(word) - 1;
			idx[at++] = (int32_t)(base + bpos);
			word &= word - 1;
		}
		running += s_total;
		__syncthreads();
	}
	if (tid == 0) {
		count_out[0] = running;
		if (off2) off2[0] = 0, off2[1] = running;
	}
}

// ---- the inner-RANSAC samples of one local-optimisation step (GCRANSAC.h:823-851) --------------------------------------
// count > limit : `trials` samples of `limit` distinct inliers each; trial t draws from its own generator (splitmix64
//                 seeded by lo_substream(seed, event, t)), exactly as the host sampler would: unique_set by rejection
//                 (gcr/uniform_random_generator.h:76-122) over positions of the inlier list;
// m < count     : one problem holding every inlier (all trials would refit the same set);
// otherwise     : nothing (the host loop breaks).
__device__ __forceinline__ uint64_t lo_rng_next(uint64_t &s) {
	uint64_t z = (s += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t lo_substream(uint64_t seed, uint64_t event, uint64_t trial) {
	uint64_t z = seed + 0x9E3779B97F4A7C15ull * (event * 64 + trial + 1);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
constexpr int kLoMaxSample = 64;

// One warp per trial. splitmix64 is counter based (output k = mix(s0 + k * golden)), so a warp computes 32 generator
// outputs at once; they are then CONSUMED in order exactly like the sequential sampler does: an output r >= lim is
// rejected (uniform_int_distribution), a value already in the set is rejected (unique_set redraws that position).
__global__ void __launch_bounds__(32)
    k_lo_sample(const int32_t *__restrict__ inl, const int64_t *__restrict__ count_in, int m, int limit, int trials,
                const uint64_t *__restrict__ seed_event, int32_t *__restrict__ off /*trials + 1*/, int32_t *__restrict__ idx) {
	pdl_launch_dependents();
	pdl_wait();
	const uint64_t seed = seed_event[0], event = seed_event[1]; // on the device: the launch is part of a replayed graph
	const int64_t count = count_in[0];
	const int t = blockIdx.x, lane = threadIdx.x;
	if (count > (int64_t)limit) {
		if (lane == 0) {
			off[t] = t * limit;
			if (t == trials - 1) off[trials] = trials * limit;
		}
		uint64_t s0 = lo_substream(seed, event, (uint64_t)t);
		if (s0 == 0) s0 = 0x9E3779B97F4A7C15ull;
		const uint64_t range = (uint64_t)count; // uniform(max = count - 1)
		const uint64_t lim = UINT64_MAX - (UINT64_MAX % range);
		int32_t acc_lo = -1, acc_hi = -1; // lane l holds accepted[l] and accepted[l + 32]
		int a = 0;
		for (uint64_t k_base = 0; a < limit; k_base += 32) {
			uint64_t z = s0 + (k_base + (uint64_t)lane + 1) * 0x9E3779B97F4A7C15ull;
			z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
			z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
			const uint64_t r = z ^ (z >> 31);
			const int valid = r < lim;
			const int32_t v = (int32_t)(r % range);
			for (int j = 0; j < 32 && a < limit; ++j) {
				const int32_t vj = __shfl_sync(0xffffffffu, v, j);
				const int okj = __shfl_sync(0xffffffffu, valid, j);
				const bool hit = (lane < a && acc_lo == vj) || (lane + 32 < a && acc_hi == vj);
				const bool dup = __any_sync(0xffffffffu, hit);
				if (okj && !dup) {
					if (a < 32) {
						if (lane == a) acc_lo = vj;
					} else if (lane == a - 32) {
						acc_hi = vj;
					}
					++a;
				}
			}
		}
		if (lane < limit) idx[t * limit + lane] = inl[acc_lo];
		if (lane + 32 < limit) idx[t * limit + lane + 32] = inl[acc_hi];
	} else if (count > (int64_t)m) {
		if (lane == 0) {
			off[t] = t == 0 ? 0 : (int32_t)count;
			if (t == trials - 1) off[trials] = (int32_t)count;
		}
		for (int64_t i = (int64_t)t * 32 + lane; i < count; i += (int64_t)trials * 32) idx[i] = inl[i];
	} else if (lane == 0) {
		off[t] = 0;
		if (t == trials - 1) off[trials] = 0;
	}
}

int launch_flag_compact(pxb_ctx *ctx, const uint8_t *flags_dev, int64_t N, int32_t *idx_dev, int64_t *count_dev, int32_t *off2_dev) {
	PXB_CUDA(launch_pdl(k_flag_compact<false>, dim3(1), dim3(kListThreads), 0, ctx->stream, flags_dev, N, idx_dev, count_dev, off2_dev));
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}
int launch_mask_compact(pxb_ctx *ctx, const uint32_t *mask_dev, int64_t N, int32_t *idx_dev, int64_t *count_dev, int32_t *off2_dev) {
	k_flag_compact<true><<<1, kListThreads, 0, ctx->stream>>>(mask_dev, N, idx_dev, count_dev, off2_dev);
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}
int launch_lo_sample(pxb_ctx *ctx, const int32_t *inl_dev, const int64_t *count_dev, int m, int limit, int trials,
                     const uint64_t *seed_event_dev, int32_t *off_dev, int32_t *idx_dev) {
	if (limit > kLoMaxSample || trials < 1) {
		set_error("local optimisation sample of %d points / %d trials exceeds the kernel limits", limit, trials);
		return PXB_ERR_UNSUPPORTED;
	}
	PXB_CUDA(launch_pdl(k_lo_sample, dim3((unsigned)trials), dim3(32), 0, ctx->stream, inl_dev, count_dev, m, limit, trials, seed_event_dev, off_dev, idx_dev));
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

int launch_label_lists(pxb_ctx *ctx, const int32_t *labels_dev, int64_t N, int L, int32_t *off_dev, int32_t *idx_dev) {
	if (L <= 0) return PXB_OK;
	k_label_lists<<<(unsigned)L, kListThreads, 0, ctx->stream>>>(labels_dev, N, L, off_dev, idx_dev);
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

int launch_select_models(pxb_ctx *ctx, const double *current, const double *fitted, const int32_t *ok, int L, int ms,
                         double *cand) {
	if (L <= 0) return PXB_OK;
	const int n = L * ms;
	k_select_models<<<(n + 127) / 128, 128, 0, ctx->stream>>>(current, fitted, ok, L, ms, cand);
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

} // namespace pxb
