// pxb_label.cu -- the PEARL label sweep (rows a10/a11 of the scope table).
//
//   k_greedy_ufl         a10  GCoptimization::solveGreedy (gcr/GCoptimization.cpp:608-751): greedy uncapacitated
//                             facility location over the dense N x (L+1) data-cost matrix, reached through
//                             solveSpecialCases (:483-555) when there is no smoothness term (lambda == 0, the Python
//                             default) -- a handful of column reductions, no graph cut at all.
//   alpha-expansion      a11  see pxb_expansion.cu
//
// The whole greedy solve is one persistent single-block kernel: L+1 <= 11 labels and at most L+1 rounds of
// column sums over N <= 1e5 sites is a few hundred microseconds of work; a multi-kernel version would be launch
// bound. Column sums use the fixed block topology of pxb_kernels.cu (thread-strided, butterfly, warps in order).
#include "pxb_internal.h"
#include "pxb_residuals.cuh"

namespace pxb {

constexpr int kBlock = 1024;
constexpr int kMaxL = 64;

__device__ __forceinline__ double block_sum(double x, double *s_tmp) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) x = add(x, __shfl_xor_sync(0xffffffffu, x, o));
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	__syncthreads();
	if (lane == 0) s_tmp[warp] = x;
	__syncthreads();
	double t = 0.0;
	for (int w = 0; w < kBlock / 32; ++w) t = add(t, s_tmp[w]);
	return t;
}

__global__ void __launch_bounds__(kBlock)
    k_greedy_ufl(const double *__restrict__ D, int64_t N, int L1, double label_cost,
                 const int32_t *__restrict__ init_labels, int32_t *__restrict__ labels_out, double *__restrict__ cur,
                 int32_t *__restrict__ lab, double *__restrict__ energy_out) {
	__shared__ double s_tmp[32];
	__shared__ double e[kMaxL];
	__shared__ int order[kMaxL];
	__shared__ unsigned char active[kMaxL];
	__shared__ int used[kMaxL];
	__shared__ int s_alpha, s_alpha_prev, s_stop;
	__shared__ double s_estart;

	const int tid = threadIdx.x;
	// ---- estart = compute_energy() of the initial labelling (:614; data + active label costs, no smooth term)
	if (tid < kMaxL) used[tid] = 0;
	__syncthreads();
	double de = 0.0;
	for (int64_t i = tid; i < N; i += kBlock) {
		const int l = init_labels ? init_labels[i] : 0;
		de = add(de, D[i * L1 + l]);
		used[l] = 1; // benign race: all writers store 1
	}
	de = block_sum(de, s_tmp);
	if (tid == 0) {
		double le = 0.0;
		for (int l = L1 - 1; l >= 0; --l)
			if (used[l]) le = add(le, label_cost);
		s_estart = add(add(de, 0.0), le);
	}
	// ---- first label: argmin_l (label_cost + sum_i D[i,l]) with strict <, first wins (:634-650)
	for (int l = 0; l < L1; ++l) {
		double a = 0.0;
		for (int64_t i = tid; i < N; i += kBlock) a = add(a, D[i * L1 + l]);
		a = block_sum(a, s_tmp);
		if (tid == 0) e[l] = add(label_cost, a);
	}
	__syncthreads();
	if (tid == 0) {
		int alpha = 0;
		for (int l = 0; l < L1; ++l)
			if (e[l] < e[alpha]) alpha = l;
		for (int l = 0; l < L1; ++l) {
			order[l] = l;
			active[l] = 0;
		}
		order[alpha] = 0;
		order[0] = alpha;
		active[alpha] = 1;
		s_alpha = alpha;
		s_stop = 0;
	}
	__syncthreads();
	{
		const int alpha = s_alpha;
		for (int64_t i = tid; i < N; i += kBlock) {
			lab[i] = alpha;
			cur[i] = D[i * L1 + alpha];
		}
	}
	__syncthreads();
	// ---- greedy expansion rounds (:667-722)
	for (int alpha_count = 1; alpha_count <= L1; ++alpha_count) {
		if (tid == 0) s_alpha_prev = s_alpha;
		__syncthreads();
		for (int li = alpha_count; li < L1; ++li) {
			const int l = order[li];
			double drop = 0.0;
			for (int64_t i = tid; i < N; i += kBlock) {
				const double delta = sub(D[i * L1 + l], cur[i]);
				if (delta < 0) drop = add(drop, delta);
			}
			drop = block_sum(drop, s_tmp);
			if (tid == 0) {
				double v = e[s_alpha_prev];
				if (!active[l]) v = add(v, label_cost);
				e[l] = add(v, drop);
			}
		}
		__syncthreads();
		if (tid == 0) {
			int alpha = s_alpha;
			int alpha_index = alpha_count - 1;
			for (int li = alpha_count; li < L1; ++li) {
				const int l = order[li];
				if (e[l] < e[alpha]) {
					alpha = l;
					alpha_index = li;
				}
			}
			if (alpha == s_alpha_prev) {
				s_stop = 1;
			} else {
				const int t = order[alpha_count];
				order[alpha_count] = order[alpha_index];
				order[alpha_index] = t;
				active[alpha] = 1;
				s_alpha = alpha;
			}
		}
		__syncthreads();
		if (s_stop) break;
		const int alpha = s_alpha;
		for (int64_t i = tid; i < N; i += kBlock) {
			const double dc_l = D[i * L1 + alpha];
			if (sub(dc_l, cur[i]) < 0) {
				lab[i] = alpha;
				cur[i] = dc_l;
			}
		}
		__syncthreads();
	}
	// ---- accept only if strictly better than the start labelling (:724-741)
	const double efinal = e[s_alpha];
	const bool better = efinal < s_estart;
	for (int64_t i = tid; i < N; i += kBlock) labels_out[i] = better ? lab[i] : (init_labels ? init_labels[i] : 0);
	if (tid == 0) *energy_out = better ? efinal : s_estart;
}

int launch_greedy_label(pxb_ctx *ctx, const double *D, int64_t N, int32_t L1, double label_cost,
                        const int32_t *init_labels, int32_t *labels_out, double *energy_out_dev) {
	if (L1 > kMaxL || L1 < 1) {
		set_error("label count %d outside [1, %d]", L1, kMaxL);
		return PXB_ERR_ARGUMENT;
	}
	PXB_TRY(ctx->outC.reserve(sizeof(double) * (size_t)N));
	PXB_TRY(ctx->outD.reserve(sizeof(int32_t) * (size_t)N));
	k_greedy_ufl<<<1, kBlock, 0, ctx->stream>>>(D, N, L1, label_cost, init_labels, labels_out, ctx->outC.as<double>(),
	                                           ctx->outD.as<int32_t>(), energy_out_dev);
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

} // namespace pxb
