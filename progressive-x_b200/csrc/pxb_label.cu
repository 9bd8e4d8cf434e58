// pxb_label.cu -- the PEARL label sweep (rows a10/a11 of the scope table).
//
//   k_greedy_ufl         a10  GCoptimization::solveGreedy (gcr/GCoptimization.cpp:608-751): greedy uncapacitated
//                             facility location over the dense N x (L+1) data-cost matrix, reached through
//                             solveSpecialCases (:483-555) when there is no smoothness term (lambda == 0, the Python
//                             default) -- a handful of column reductions, no graph cut at all.
//   alpha-expansion      a11  see pxb_expansion.cu
//
// The whole greedy solve is one persistent kernel (a multi-kernel version would be launch bound). k_greedy_ufl_fused
// reads every row of D once per round and accumulates ALL column sums of the round in registers (the first version made
// one pass over N per column: ~L^2/2 passes); for N > 16k it runs on several blocks of a cooperative launch, block
// partials combined in block order after a grid barrier, every block taking the (identical) decision redundantly.
// Column sums use the fixed topology of pxb_kernels.cu (grid-strided thread partials in index order, butterfly, warps in
// order, blocks in order): a function of N only. k_greedy_ufl (one block, one column per pass) remains for L + 1 > 16.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

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

constexpr int kUflMaxL = 16;

// sums of NV per-thread values over the block: butterfly inside the warp, then the warps in order (thread a < NV adds them)
template <int NV> __device__ __forceinline__ void ufl_block_sums(double (&v)[NV], double *s_vec /*[32][NV]*/, double *s_out /*[NV]*/) {
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
		for (int w = 0; w < kBlock / 32; ++w) t = add(t, s_vec[w * NV + threadIdx.x]);
		s_out[threadIdx.x] = t;
	}
	__syncthreads();
}

__global__ void __launch_bounds__(kBlock)
    k_greedy_ufl_fused(const double *__restrict__ D, int64_t N, int L1, double label_cost, const int32_t *__restrict__ init_labels,
                       int32_t *__restrict__ labels_out, double *__restrict__ cur, int32_t *__restrict__ lab,
                       double *__restrict__ energy_out, double *__restrict__ partials /*[2][G][kUflMaxL + 1]*/,
                       unsigned *__restrict__ used_bits /*[G]*/) {
	constexpr int NV = kUflMaxL + 1;
	__shared__ double s_vec[(kBlock / 32) * NV];
	__shared__ double s_col[NV];
	__shared__ double e[kUflMaxL];
	__shared__ int order[kUflMaxL];
	__shared__ unsigned char active[kUflMaxL];
	__shared__ unsigned s_used;
	__shared__ int s_alpha, s_alpha_prev, s_stop;
	__shared__ double s_estart;
	const int tid = threadIdx.x, G = gridDim.x, b = blockIdx.x;
	const int64_t first = (int64_t)b * kBlock + tid, step = (int64_t)G * kBlock;
	int round = 0; // selects the partials buffer: a block can be at most one round ahead of another
	// combines the block sums of all blocks (in block order) into s_col; one grid barrier per round
	auto combine = [&]() {
		if (G > 1) {
			double *buf = partials + (size_t)(round & 1) * G * NV;
			if (tid < NV) buf[b * NV + tid] = s_col[tid];
			__threadfence();
			cooperative_groups::this_grid().sync();
			if (tid < NV) {
				double t = 0.0;
				for (int g = 0; g < G; ++g) t = add(t, buf[g * NV + tid]);
				s_col[tid] = t;
			}
			__syncthreads();
			++round;
		}
	};
	// ---- estart = compute_energy() of the initial labelling (:614) and all column sums (:634-650), one pass
	if (tid == 0) s_used = 0;
	__syncthreads();
	double acc[NV];
#pragma unroll
	for (int a = 0; a < NV; ++a) acc[a] = 0.0;
	unsigned used = 0;
	for (int64_t i = first; i < N; i += step) {
		const double *row = D + i * L1;
		const int l0 = init_labels ? init_labels[i] : 0;
		used |= 1u << l0;
#pragma unroll
		for (int l = 0; l < kUflMaxL; ++l)
			if (l < L1) {
				const double d = row[l];
				acc[l] = add(acc[l], d);
				if (l == l0) acc[kUflMaxL] = add(acc[kUflMaxL], d);
			}
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) used |= __shfl_xor_sync(0xffffffffu, used, o);
	if ((tid & 31) == 0 && used) atomicOr(&s_used, used);
	ufl_block_sums<NV>(acc, s_vec, s_col);
	if (G > 1 && tid == 0) used_bits[b] = s_used;
	combine();
	if (tid == 0) {
		unsigned u = s_used;
		if (G > 1) {
			u = 0;
			for (int g = 0; g < G; ++g) u |= used_bits[g];
		}
		double le = 0.0;
		for (int l = L1 - 1; l >= 0; --l)
			if (u >> l & 1u) le = add(le, label_cost);
		s_estart = add(add(s_col[kUflMaxL], 0.0), le);
		for (int l = 0; l < L1; ++l) e[l] = add(label_cost, s_col[l]);
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
		for (int64_t i = first; i < N; i += step) {
			lab[i] = alpha;
			cur[i] = D[i * L1 + alpha];
		}
	}
	// ---- greedy expansion rounds (:667-722): every candidate's drop in one pass
	for (int alpha_count = 1; alpha_count <= L1; ++alpha_count) {
		if (tid == 0) s_alpha_prev = s_alpha;
		__syncthreads();
		unsigned cand = 0; // labels still outside the solution
		for (int li = alpha_count; li < L1; ++li) cand |= 1u << order[li];
#pragma unroll
		for (int a = 0; a < NV; ++a) acc[a] = 0.0;
		if (cand)
			for (int64_t i = first; i < N; i += step) {
				const double *row = D + i * L1;
				const double c = cur[i];
#pragma unroll
				for (int l = 0; l < kUflMaxL; ++l)
					if (cand >> l & 1u) {
						const double delta = sub(row[l], c);
						if (delta < 0) acc[l] = add(acc[l], delta);
					}
			}
		ufl_block_sums<NV>(acc, s_vec, s_col);
		combine();
		if (tid == 0) {
			for (int li = alpha_count; li < L1; ++li) {
				const int l = order[li];
				double v = e[s_alpha_prev];
				if (!active[l]) v = add(v, label_cost);
				e[l] = add(v, s_col[l]);
			}
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
		if (s_stop) break; // identical in every block: all of them decided on the same sums
		const int alpha = s_alpha;
		for (int64_t i = first; i < N; i += step) {
			const double dc_l = D[i * L1 + alpha];
			if (sub(dc_l, cur[i]) < 0) {
				lab[i] = alpha;
				cur[i] = dc_l;
			}
		}
	}
	// ---- accept only if strictly better than the start labelling (:724-741)
	const double efinal = e[s_alpha];
	const bool better = efinal < s_estart;
	for (int64_t i = first; i < N; i += step) labels_out[i] = better ? lab[i] : (init_labels ? init_labels[i] : 0);
	if (b == 0 && tid == 0) *energy_out = better ? efinal : s_estart;
}

// ---- the same solve on ONE THREAD-BLOCK CLUSTER (N <= 16384: every PEARL sweep of the 5k-10k point problems) -----------
// One 1024-thread block reads the N x (L+1) cost matrix (480 KB at N = 10^4) once or twice per round through a single SM:
// 57-83 us per sweep, the top kernel of a lambda = 0 fit. Here the 1024 threads of that block are VIRTUAL: virtual thread
// v = 128 * rank + tid of an 8-CTA cluster visits the rows v, v + 1024, ... exactly as before, the butterfly runs over the
// same 32 lanes, and the 32 warp sums are added in warp order -- by every CTA redundantly, after each CTA has stored its
// four warp rows into the shared memory of the other seven (st.shared::cluster) and one barrier.cluster. Eight SMs pull
// the matrix (each slice then stays in its SM's L1), every sum is bit-identical to the single-block kernel's, and all
// CTAs take the same decisions. The exchange buffer alternates by reduction parity: a CTA can be at most one reduction
// ahead of the slowest one.
constexpr int kUflCtas = 8, kUflCt = kBlock / kUflCtas, kUflWarpsPerCta = kUflCt / 32;

__device__ __forceinline__ void ufl_cluster_barrier() {
	asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(kUflCtas, 1, 1) __launch_bounds__(kUflCt)
    k_greedy_ufl_cluster(const double *__restrict__ D, int64_t N, int L1, double label_cost, const int32_t *__restrict__ init_labels,
                         int32_t *__restrict__ labels_out, double *__restrict__ cur, int32_t *__restrict__ lab,
                         double *__restrict__ energy_out) {
	constexpr int NV = kUflMaxL + 1, ROW = kUflWarpsPerCta * NV; // doubles one CTA contributes per reduction
	__shared__ double s_vec[2][(kBlock / 32) * NV];
	__shared__ unsigned s_used_all[kUflCtas];
	__shared__ double s_col[NV];
	__shared__ double e[kUflMaxL];
	__shared__ int order[kUflMaxL];
	__shared__ unsigned char active[kUflMaxL];
	__shared__ unsigned s_used;
	__shared__ int s_alpha, s_alpha_prev, s_stop;
	__shared__ double s_estart;
	unsigned rank;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
	const int tid = threadIdx.x, lane = tid & 31;
	const int v = (int)rank * kUflCt + tid, vwarp = v >> 5; // virtual thread / warp of the 1024-thread topology
	int parity = 0;
	// block sums of the cluster: butterfly, this CTA's warp rows to everybody, barrier, the 32 warp sums in warp order
	auto reduce = [&](double(&acc)[NV], bool with_used) {
#pragma unroll
		for (int a = 0; a < NV; ++a)
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) acc[a] = add(acc[a], __shfl_xor_sync(0xffffffffu, acc[a], o));
		double *buf = s_vec[parity];
		if (lane == 0)
#pragma unroll
			for (int a = 0; a < NV; ++a) buf[vwarp * NV + a] = acc[a];
		__syncthreads();
		const unsigned base = (unsigned)__cvta_generic_to_shared(buf + (int)rank * ROW);
		for (int t = tid; t < ROW * kUflCtas; t += kUflCt) {
			const unsigned dst = (unsigned)(t / ROW), k = (unsigned)(t % ROW);
			if (dst == rank) continue;
			unsigned remote;
			asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(base + 8u * k), "r"(dst));
			asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote), "d"(buf[(int)rank * ROW + (int)k]) : "memory");
		}
		if (with_used && tid < kUflCtas) {
			unsigned remote;
			asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"((unsigned)__cvta_generic_to_shared(s_used_all + rank)), "r"((unsigned)tid));
			asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(s_used) : "memory");
		}
		ufl_cluster_barrier();
		if (tid < NV) {
			double t = 0.0;
			for (int w = 0; w < kBlock / 32; ++w) t = add(t, buf[w * NV + tid]);
			s_col[tid] = t;
		}
		__syncthreads();
		parity ^= 1;
	};
	// ---- estart = compute_energy() of the initial labelling (:614) and all column sums (:634-650), one pass
	if (tid == 0) s_used = 0;
	__syncthreads();
	double acc[NV];
#pragma unroll
	for (int a = 0; a < NV; ++a) acc[a] = 0.0;
	unsigned used = 0;
	for (int64_t i = v; i < N; i += kBlock) {
		const double *row = D + i * L1;
		const int l0 = init_labels ? init_labels[i] : 0;
		used |= 1u << l0;
#pragma unroll
		for (int l = 0; l < kUflMaxL; ++l)
			if (l < L1) {
				const double d = row[l];
				acc[l] = add(acc[l], d);
				if (l == l0) acc[kUflMaxL] = add(acc[kUflMaxL], d);
			}
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) used |= __shfl_xor_sync(0xffffffffu, used, o);
	if (lane == 0 && used) atomicOr(&s_used, used);
	__syncthreads();
	reduce(acc, true);
	if (tid == 0) {
		unsigned u = 0;
		for (int g = 0; g < kUflCtas; ++g) u |= s_used_all[g];
		double le = 0.0;
		for (int l = L1 - 1; l >= 0; --l)
			if (u >> l & 1u) le = add(le, label_cost);
		s_estart = add(add(s_col[kUflMaxL], 0.0), le);
		for (int l = 0; l < L1; ++l) e[l] = add(label_cost, s_col[l]);
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
		for (int64_t i = v; i < N; i += kBlock) {
			lab[i] = alpha;
			cur[i] = D[i * L1 + alpha];
		}
	}
	// ---- greedy expansion rounds (:667-722): every candidate's drop in one pass
	for (int alpha_count = 1; alpha_count <= L1; ++alpha_count) {
		if (tid == 0) s_alpha_prev = s_alpha;
		__syncthreads();
		unsigned cand = 0; // labels still outside the solution
		for (int li = alpha_count; li < L1; ++li) cand |= 1u << order[li];
#pragma unroll
		for (int a = 0; a < NV; ++a) acc[a] = 0.0;
		if (cand)
			for (int64_t i = v; i < N; i += kBlock) {
				const double *row = D + i * L1;
				const double c = cur[i];
#pragma unroll
				for (int l = 0; l < kUflMaxL; ++l)
					if (cand >> l & 1u) {
						const double delta = sub(row[l], c);
						if (delta < 0) acc[l] = add(acc[l], delta);
					}
			}
		reduce(acc, false);
		if (tid == 0) {
			for (int li = alpha_count; li < L1; ++li) {
				const int l = order[li];
				double x = e[s_alpha_prev];
				if (!active[l]) x = add(x, label_cost);
				e[l] = add(x, s_col[l]);
			}
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
		if (s_stop) break; // identical in every CTA: all of them decided on the same sums
		const int alpha = s_alpha;
		for (int64_t i = v; i < N; i += kBlock) {
			const double dc_l = D[i * L1 + alpha];
			if (sub(dc_l, cur[i]) < 0) {
				lab[i] = alpha;
				cur[i] = dc_l;
			}
		}
	}
	// ---- accept only if strictly better than the start labelling (:724-741)
	const double efinal = e[s_alpha];
	const bool better = efinal < s_estart;
	for (int64_t i = v; i < N; i += kBlock) labels_out[i] = better ? lab[i] : (init_labels ? init_labels[i] : 0);
	if (rank == 0 && tid == 0) *energy_out = better ? efinal : s_estart;
	ufl_cluster_barrier(); // no CTA may exit while another can still store into its shared memory
}

int launch_greedy_label(pxb_ctx *ctx, const double *D, int64_t N, int32_t L1, double label_cost,
                        const int32_t *init_labels, int32_t *labels_out, double *energy_out_dev) {
	if (L1 > kMaxL || L1 < 1) {
		set_error("label count %d outside [1, %d]", L1, kMaxL);
		return PXB_ERR_ARGUMENT;
	}
	int G = 1;
	if (N > 16384) G = (int)std::min<int64_t>((N + 8191) / 8192, ctx->sm_count); // a function of N (and the part) only
	PXB_TRY(ctx->outC.reserve(sizeof(double) * ((size_t)N + 2 * (size_t)G * (kUflMaxL + 1))));
	PXB_TRY(ctx->outD.reserve(sizeof(int32_t) * ((size_t)N + (size_t)G)));
	double *cur = ctx->outC.as<double>(), *partials = cur + N;
	int32_t *lab = ctx->outD.as<int32_t>();
	unsigned *used_bits = reinterpret_cast<unsigned *>(lab + N);
	if (L1 > kUflMaxL) {
		k_greedy_ufl<<<1, kBlock, 0, ctx->stream>>>(D, N, L1, label_cost, init_labels, labels_out, cur, lab, energy_out_dev);
	} else if (G == 1) {
		static const bool one_block = getenv("PXB_UFL_CLUSTER") && atoi(getenv("PXB_UFL_CLUSTER")) == 0; // A/B: the single-block kernel
		if (one_block)
			k_greedy_ufl_fused<<<1, kBlock, 0, ctx->stream>>>(D, N, (int)L1, label_cost, init_labels, labels_out, cur, lab, energy_out_dev,
			                                                 partials, used_bits);
		else
			k_greedy_ufl_cluster<<<kUflCtas, kUflCt, 0, ctx->stream>>>(D, N, (int)L1, label_cost, init_labels, labels_out, cur, lab,
			                                                          energy_out_dev);
	} else {
		int l1 = (int)L1;
		void *args[] = {(void *)&D, (void *)&N, (void *)&l1, (void *)&label_cost, (void *)&init_labels, (void *)&labels_out,
		                (void *)&cur, (void *)&lab, (void *)&energy_out_dev, (void *)&partials, (void *)&used_bits};
		PXB_CUDA(cudaLaunchCooperativeKernel((void *)k_greedy_ufl_fused, dim3((unsigned)G), dim3(kBlock), args, 0, ctx->stream));
	}
	ctx->launches++;
	PXB_CUDA(cudaGetLastError());
	return PXB_OK;
}

} // namespace pxb
