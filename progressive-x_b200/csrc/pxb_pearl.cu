// pxb_pearl.cu -- device-side glue of one PEARL iteration (px/include/PEARL.h:319-401, parameterEstimation), so that
// labeling -> per-instance point lists -> residual sums -> non-minimal refits -> residual sums of the refits is ONE
// stream-ordered chain with a single small device-to-host copy at its end (the host driver used to take four round
// trips for it: labels down, sums, index lists up + fits, sums again).
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
