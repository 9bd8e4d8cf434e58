// pxb_expansion.cu -- a11: the alpha-expansion label sweep of PEARL and the st-cut of the GC-RANSAC local
// optimisation, both on one GPU min-cut engine.
//
//   k_maxflow                 lock-free asynchronous push-relabel (Hong & He) on a CSR residual graph, alternating with
//                             exact backward-BFS global relabelling, one persistent cooperative kernel per cut (no host
//                             round trips). Graphs that fit in shared memory are relabelled by block 0 alone (queue
//                             BFS, four lanes per node, bottom-up levels while the frontier is the larger side, two
//                             block barriers per level, one grid barrier per relabel); larger ones by a grid-wide
//                             level-synchronous BFS. DESIGN.md section 4 has the measurements behind each choice.
//   pxb_lo_graph_cut          GCRANSAC::labeling                       gcr/GCRANSAC.h:964-1018
//   launch_alpha_expansion    GCoptimization::expansion / oneExpansionIteration / alpha_expansion
//                                                                     gcr/GCoptimization.cpp:1003-1086,1239-1318
//
// Parity with the reference's Boykov-Kolmogorov solver (gcr/maxflow.cpp). BK labels a node SINK iff it is in the sink
// tree when the search trees stop growing, i.e. iff the node can still reach the sink in the final residual graph;
// every other node (source tree or free) reads as SOURCE (gcr/graph.h:478-487). That set is the same for every
// maximum flow (it is the sink side of the minimal-sink-side minimum cut), so a different max-flow algorithm followed
// by a backward BFS from the sink reproduces BK's labels. The phase-1 preflow of push-relabel is enough: excess that
// is stranded on the source side never crosses the cut. Terminal capacities are accumulated with the reference's own
// add_tweights arithmetic (gcr/graph.h: tr_cap = (cap_source [+ old tr_cap]) - (cap_sink [- old tr_cap])), so
// structural zeros (equal data costs on both sides) are exact zeros here too. What can differ is the rounding of
// partially used capacities; that only matters on exact ties between cuts (DESIGN.md "Max-flow parity").
//
// A first, fully synchronous (pulse) version was deterministic to the bit but needed ~10^4 grid barriers per cut
// (0.5 s at N = 10^4); the asynchronous phase removes the barriers. Floating-point atomics make the rounding of
// partially used capacities run-dependent, which cannot change the cut except on exact ties (saturating pushes and
// emptied excesses are exact zeros in every order).
//
// Split of work: the binary-energy graph of every expansion move and of every LO cut is assembled ON THE DEVICE over a
// fixed arc skeleton (k_exp_assemble / k_lo_assemble: the reference's add_term1 / add_term2 / add_tweights sequence, one
// thread per node over the node's own neighbour list); the host keeps the move loop and evaluates the labelling energies
// that decide whether a move is kept, in the reference's sequential summation order. pxb_lo_graph_cut (terms given by the
// caller) still assembles on the host.
#include <cooperative_groups.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_set>
#include <vector>

#include "pxb_internal.h"
#include "pxb_maxflow.h"

namespace cg = cooperative_groups;

namespace pxb {


#ifndef PXB_MF_THREADS
#define PXB_MF_THREADS 1024
#endif
constexpr int kMfThreads = PXB_MF_THREADS;
constexpr int kWideDegree = 64;

// predicated loads: issued back to back, no branch; the destination keeps `otherwise` when the predicate is false
__device__ __forceinline__ int ld_nc_s32_if(const int32_t *p, bool pred, int otherwise) {
	int v = otherwise;
	asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q ld.global.nc.s32 %0, [%1];\n\t}" : "+r"(v) : "l"(p), "r"((int)pred));
	return v;
}
__device__ __forceinline__ double ld_cg_f64_if(const double *p, bool pred) {
	double v = 0.0;
	asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q ld.global.cg.f64 %0, [%1];\n\t}" : "+d"(v) : "l"(p), "r"((int)pred));
	return v;
}
__device__ __forceinline__ double ld_volatile_f64_if(const double *p, bool pred) {
	double v = 0.0;
	asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q ld.volatile.global.f64 %0, [%1];\n\t}" : "+d"(v) : "l"(p), "r"((int)pred) : "memory");
	return v;
}
__device__ __forceinline__ int ld_volatile_s32_if(const int32_t *p, bool pred, int otherwise) {
	int v = otherwise;
	asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q ld.volatile.global.s32 %0, [%1];\n\t}" : "+r"(v) : "l"(p), "r"((int)pred) : "memory");
	return v;
}

// Exact distance-to-sink labels by a level-synchronous BACKWARD breadth-first search in "push" form: the nodes of the
// current frontier (height == level) mark every unlabelled in-neighbour v (residual arc v -> u, i.e. cap[rev(a)] > 0 for
// the arc a = u -> v) with level + 1. Each node is expanded exactly once and all of its arcs are examined with
// independent loads (the earlier "pull" form re-scanned every unlabelled node at every level with a serial, early-exit
// arc loop: 16 us per level at N = 10^4; source-side nodes -- never labelled -- paid that at every level).
__device__ void mf_global_relabel(const FlowGraphDev &G, int32_t *h, cg::grid_group &grid, int tid, int nthreads) {
	const int n = G.n;
	volatile int32_t *hv = h;
	const int lane = threadIdx.x & 31;
	// flags[9], [11], [15]: size of the frontier of level L in slot L % 3 -- they choose the direction of every level
	// (below). Like the 'changed' flags they rotate over three slots: during level L the slot of L + 1 is accumulated and
	// the slot of L + 2 (= L - 1, read by everyone before the previous barrier) is cleared.
	const int cnt_slot[3] = {9, 11, 15};
	if (tid == 0) {
		G.flags[0] = 0;
		G.flags[1] = 0;
		G.flags[2] = 0;
		G.flags[9] = 0;
		G.flags[11] = 0;
		G.flags[15] = 0;
	}
	grid.sync();
	int first = 0;
	for (int u = tid; u < n; u += nthreads) {
		const bool at_sink = G.sink_cap[u] > 0.0;
		h[u] = at_sink ? 1 : n;
		first += at_sink;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) first += __shfl_xor_sync(0xffffffffu, first, o);
	if (lane == 0 && first) atomicAdd(&G.flags[cnt_slot[1]], first);
	grid.sync();
	int labelled = 0;
	// flags rotate over three slots: the slot of level L+1 is cleared during level L, while stragglers may still be
	// reading the slot of level L-1 (they are past that level's barrier but not yet past its test)
	for (int level = 1; level < n; ++level) {
		int32_t *changed = &G.flags[level % 3];
		int32_t *count_next = &G.flags[cnt_slot[(level + 1) % 3]];
		const int frontier = *(volatile int32_t *)&G.flags[cnt_slot[level % 3]]; // stable: written before the last barrier
		labelled += frontier;
		bool mine = false;
		int found_here = 0;
		// Direction-optimising step (Beamer et al.), as in the single-block form: when the frontier is larger than what
		// is still unlabelled -- level 1 of every cut: most nodes hang on the sink directly -- every still-unlabelled
		// node looks among its OWN out-arcs for a residual one into the frontier instead of the frontier expanding all of
		// its arcs (the top-down level 1 of a 10^5-node light move cost 0.4 ms: 9.4 10^4 nodes x ~10 arcs x 4 dependent
		// loads to find the ~6000 others). Same labels: a node gets level + 1 iff it has a residual arc into this level
		// and was not labelled before.
		const bool bottom_up = 2 * (long long)frontier > (long long)(n - labelled);
		if (bottom_up) {
			for (int u = tid; u < G.wide_begin; u += nthreads) {
				if (hv[u] != n) continue;
				const int a0 = G.arc_off[u], a1 = G.arc_off[u + 1];
				bool hit = false;
				for (int base = a0; base < a1 && !hit; base += 8) {
					int v[8];
					double c[8];
#pragma unroll
					for (int j = 0; j < 8; ++j) {
						v[j] = ld_nc_s32_if(G.arc_head + base + j, base + j < a1, -1);
						c[j] = ld_volatile_f64_if(G.cap + base + j, base + j < a1);
					}
#pragma unroll
					for (int j = 0; j < 8; ++j)
						hit |= v[j] >= 0 && c[j] > 0.0 && ld_volatile_s32_if(h + (v[j] >= 0 ? v[j] : 0), v[j] >= 0 && c[j] > 0.0, 0) == level;
				}
				if (hit) {
					hv[u] = level + 1;
					mine = true;
					++found_here;
				}
			}
			for (int w = blockIdx.x; w < G.wide_count; w += gridDim.x) { // an unlabelled wide node: its block scans its arcs
				const int u = G.wide_begin + w;
				if (hv[u] != n) continue; // block-uniform
				bool hit = false;
				for (int a = G.arc_off[u] + threadIdx.x; a < G.arc_off[u + 1] && !hit; a += blockDim.x)
					hit = G.cap[a] > 0.0 && hv[G.arc_head[a]] == level;
				if (__syncthreads_or(hit)) {
					if (threadIdx.x == 0) {
						hv[u] = level + 1;
						++found_here;
					}
					mine = true;
				}
			}
		} else {
			for (int u = tid; u < G.wide_begin; u += nthreads) {
				if (hv[u] != level) continue;
				const int a0 = G.arc_off[u], a1 = G.arc_off[u + 1];
				for (int base = a0; base < a1; base += 8) { // chunks of 8 arcs, predicated loads (see mf_process)
					int v[8], r[8];
					bool want[8];
					double c[8];
#pragma unroll
					for (int j = 0; j < 8; ++j) v[j] = ld_nc_s32_if(G.arc_head + base + j, base + j < a1, -1);
#pragma unroll
					for (int j = 0; j < 8; ++j) want[j] = ld_volatile_s32_if(h + (v[j] >= 0 ? v[j] : 0), v[j] >= 0, 0) == n;
#pragma unroll
					for (int j = 0; j < 8; ++j) r[j] = ld_nc_s32_if(G.arc_rev + base + j, want[j], 0);
#pragma unroll
					for (int j = 0; j < 8; ++j) c[j] = ld_volatile_f64_if(G.cap + r[j], want[j]);
#pragma unroll
					for (int j = 0; j < 8; ++j)
						if (want[j] && c[j] > 0.0 && atomicCAS(&h[v[j]], n, level + 1) == n) { // claimed once: exact frontier counts
							mine = true;
							++found_here;
						}
				}
			}
			for (int w = blockIdx.x; w < G.wide_count; w += gridDim.x) { // a wide node's arcs are expanded by its whole block
				const int u = G.wide_begin + w;
				if (hv[u] != level) continue; // block-uniform: h[u] was written before the previous barrier
				for (int a = G.arc_off[u] + threadIdx.x; a < G.arc_off[u + 1]; a += blockDim.x) {
					const int v = G.arc_head[a];
					if (G.cap[G.arc_rev[a]] > 0.0 && hv[v] == n && atomicCAS(&h[v], n, level + 1) == n) {
						mine = true;
						++found_here;
					}
				}
			}
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) found_here += __shfl_xor_sync(0xffffffffu, found_here, o);
		if (lane == 0 && found_here) atomicAdd(count_next, found_here);
		if (mine) *changed = 1;
		if (tid == 0) {
			G.flags[(level + 1) % 3] = 0;
			G.flags[cnt_slot[(level + 2) % 3]] = 0;
			G.flags[8]++; // statistics: BFS levels
		}
		grid.sync();
		if (*changed == 0) break;
	}
	grid.sync();
}

// The same labels computed by ONE block with the heights and the BFS queue in shared memory (graphs up to ~28k nodes:
// every PEARL / LO cut of the configurations this engine targets). A level of the grid-wide form costs a chain of
// dependent L2 round trips plus a grid barrier (~10 us) whatever the size of the frontier, and the heavy expansion moves
// need 20-30 relabels of ~35 levels each; here a level is a few block barriers around a frontier-sized loop (queue
// form: every node is expanded exactly once), and the other blocks wait at one grid barrier per relabel.
// Returns (block-uniformly) whether any node with excess can still reach the sink.
// Appends the items of the lanes with `take` set to the shared queue with ONE atomic per warp (a same-address shared
// atomic per item serialises: at ~200 discoveries per level that alone was half of a level's time).
__device__ __forceinline__ void bfs_enqueue(bool take, int item, int32_t *queue, int *tail) {
	const unsigned mask = __ballot_sync(0xffffffffu, take);
	if (mask == 0) return;
	const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
	int base = 0;
	if (lane == leader) base = atomicAdd(tail, __popc(mask));
	base = __shfl_sync(0xffffffffu, base, leader);
	if (take) queue[base + __popc(mask & ((1u << lane) - 1u))] = item;
}

__device__ bool mf_global_relabel_block(const FlowGraphDev &G, int32_t *h, bool first) {
	extern __shared__ int32_t mf_smem[];
	__shared__ int s_tail, s_nwide, s_wide[32], s_end[2], s_utail;
	const int n = G.n;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	int32_t *hs = mf_smem, *queue = mf_smem + n;
	// block_bfs >= 2: the CSR offsets fit as well (filled once per launch) and save a round trip per expansion;
	// block_bfs == 3: and the list of the nodes that level 1 left unlabelled, for the bottom-up levels (below)
	const int32_t *offs = G.arc_off;
	int32_t *ulist = G.block_bfs == 3 ? mf_smem + 3 * n + 1 : nullptr;
	if (G.block_bfs >= 2) {
		int32_t *so = mf_smem + 2 * n;
		if (first)
			for (int u = threadIdx.x; u <= n; u += blockDim.x) so[u] = G.arc_off[u];
		offs = so;
	}
	const double *capp = G.cap;
	if (threadIdx.x == 0) {
		s_tail = 0;
		s_nwide = 0;
		s_utail = 0;
	}
	__syncthreads();
	// level 1: the nodes with residual capacity to the sink (warp-uniform trip count: bfs_enqueue is a warp collective)
	for (int base = warp * 32; base < n; base += 4 * blockDim.x) {
		double sc[4];
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const int u = base + j * (int)blockDim.x + lane;
			sc[j] = u < n ? __ldcg(&G.sink_cap[u]) : 0.0;
		}
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const int u = base + j * (int)blockDim.x + lane;
			const bool at_sink = u < n && sc[j] > 0.0;
			if (u < n) hs[u] = at_sink ? 1 : n;
			bfs_enqueue(at_sink, u, queue, &s_tail);
			if (ulist) bfs_enqueue(u < n && !at_sink, u, ulist, &s_utail);
		}
	}
	__syncthreads();
	if (threadIdx.x == 0) s_end[1] = s_tail;
	const int nulist = s_utail;
	int begin = 0, level = 1;
	for (;;) {
		__syncthreads(); // the previous level's appends and its end marker are visible
		const int end = s_end[level & 1];
		if (begin == end) break; // block-uniform
		// Direction-optimising step (Beamer et al.): when the frontier is larger than what is still unlabelled -- the
		// first levels of every cut: thousands of nodes hang on the sink directly -- the level is computed BOTTOM-UP:
		// every still-unlabelled node looks among its OWN out-arcs for a residual one into the frontier (capacity and
		// head of an out-arc are addressed directly: one round trip, no reverse-arc indirection) and takes level + 1.
		// Same labels as the top-down expansion (a node gets level + 1 iff it has a residual arc to a node of this
		// level and none to a lower one); a light expansion move spent 60 of its 100 us per relabel expanding the
		// ~8800 level-1 nodes to find the ~1200 others.
		// Work split of both forms: a warp takes 8 nodes at a time, FOUR LANES PER NODE, lane j of a group looks at arcs
		// j, j + 4, j + 8, ... of its node, four of them per pass with predicated loads in flight together. The relabel
		// runs on one SM and is bound by what that SM can issue: one node per thread makes every load instruction touch
		// 32 different lines (load-path bound, ~3.8 us per level), one node per warp pass executes ~125 warp
		// instructions per node with half of the lanes idle (issue bound, ~4 us per level); four lanes per node need
		// ~25 instructions and 12 sectors per node.
		const int grp = lane >> 2, sub = lane & 3;
		const unsigned grp_mask = 0xFu << (grp * 4);
		if (ulist != nullptr && 2 * (end - begin) > (n - end) + nulist / 8) {
			for (int i0 = warp * 8; i0 < nulist; i0 += nwarps * 8) {
				int v = i0 + grp < nulist ? ulist[i0 + grp] : -1;
				if (v >= 0 && hs[v] != n) v = -1; // labelled at an earlier level
				const int a0 = v >= 0 ? offs[v] : 0, deg = v >= 0 ? offs[v + 1] - a0 : 0;
				bool found = false; // group-uniform
				for (int k0 = 0; __any_sync(0xffffffffu, k0 < deg && !found); k0 += 16) { // warp-uniform trip count
					int uu[4];
					double c[4];
#pragma unroll
					for (int t = 0; t < 4; ++t) {
						const int k = k0 + 4 * t + sub;
						const bool in = k < deg && !found;
						uu[t] = ld_nc_s32_if(G.arc_head + a0 + k, in, -1);
						c[t] = ld_cg_f64_if(capp + a0 + k, in);
					}
					bool hit = false;
#pragma unroll
					for (int t = 0; t < 4; ++t) hit |= uu[t] >= 0 && c[t] > 0.0 && hs[uu[t]] == level;
					found |= (__ballot_sync(0xffffffffu, hit) & grp_mask) != 0;
				}
				const bool take = found && sub == 0;
				if (take) hs[v] = level + 1; // only this group handles v in this pass
				bfs_enqueue(take, v, queue, &s_tail);
			}
			__syncthreads();
			if (threadIdx.x == 0) s_end[(level + 1) & 1] = s_tail;
			begin = end;
			++level;
			continue;
		}
		// Top-down expansion of the frontier: heads and reverse arcs of four arcs per lane in one round trip, then --
		// only for heads that are still unlabelled -- the reverse capacities; atomicCAS claims a head; the winners of a
		// pass are appended to the queue with one atomic per warp (a same-address shared atomic per item serialises).
		for (int i0 = begin + warp * 8; i0 < end; i0 += nwarps * 8) {
			int u = i0 + grp < end ? queue[i0 + grp] : -1;
			if (u >= G.wide_begin) { // expanded by the whole block below
				if (sub == 0) s_wide[atomicAdd(&s_nwide, 1)] = u;
				u = -1;
			}
			const int a0 = u >= 0 ? offs[u] : 0, deg = u >= 0 ? offs[u + 1] - a0 : 0;
			for (int k0 = 0; __any_sync(0xffffffffu, k0 < deg); k0 += 16) { // warp-uniform trip count
				int v[4], r[4];
				double c[4];
				bool want[4], won[4];
#pragma unroll
				for (int t = 0; t < 4; ++t) {
					const int k = k0 + 4 * t + sub;
					v[t] = ld_nc_s32_if(G.arc_head + a0 + k, k < deg, -1);
					r[t] = ld_nc_s32_if(G.arc_rev + a0 + k, k < deg, 0);
				}
#pragma unroll
				for (int t = 0; t < 4; ++t) want[t] = v[t] >= 0 && hs[v[t]] == n;
#pragma unroll
				for (int t = 0; t < 4; ++t) c[t] = ld_cg_f64_if(capp + r[t], want[t]);
#pragma unroll
				for (int t = 0; t < 4; ++t) won[t] = want[t] && c[t] > 0.0 && atomicCAS(&hs[v[t]], n, level + 1) == n;
				unsigned wmask[4];
				int before[4], total = 0;
#pragma unroll
				for (int t = 0; t < 4; ++t) {
					wmask[t] = __ballot_sync(0xffffffffu, won[t]);
					before[t] = total;
					total += __popc(wmask[t]);
				}
				if (total > 0) { // warp-uniform
					int base = 0;
					if (lane == 0) base = atomicAdd(&s_tail, total);
					base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
					for (int t = 0; t < 4; ++t)
						if (won[t]) queue[base + before[t] + __popc(wmask[t] & ((1u << lane) - 1u))] = v[t];
				}
			}
		}
		__syncthreads();
		const int nwide = s_nwide; // block-uniform
		if (nwide > 0) {
			for (int w = 0; w < nwide; ++w) {
				const int u = s_wide[w];
				const int a1 = G.arc_off[u + 1];
				for (int base = G.arc_off[u] + warp * 32; base < a1; base += 4 * blockDim.x) { // warp-uniform
					int v[4], r[4];
					double c[4];
					bool want[4];
#pragma unroll
					for (int j = 0; j < 4; ++j) {
						const int a = base + j * (int)blockDim.x + lane;
						v[j] = ld_nc_s32_if(G.arc_head + a, a < a1, -1);
						r[j] = ld_nc_s32_if(G.arc_rev + a, a < a1, 0);
					}
#pragma unroll
					for (int j = 0; j < 4; ++j) want[j] = v[j] >= 0 && hs[v[j]] == n;
#pragma unroll
					for (int j = 0; j < 4; ++j) c[j] = ld_cg_f64_if(capp + r[j], want[j]);
#pragma unroll
					for (int j = 0; j < 4; ++j) {
						const bool won = want[j] && c[j] > 0.0 && atomicCAS(&hs[v[j]], n, level + 1) == n;
						bfs_enqueue(won, v[j], queue, &s_tail);
					}
				}
			}
			__syncthreads();
			if (threadIdx.x == 0) s_nwide = 0;
		}
		// the end marker of the next level alternates between two slots: the other slot is still being read by
		// stragglers of this level's loop test
		if (threadIdx.x == 0) s_end[(level + 1) & 1] = s_tail;
		begin = end;
		++level;
	}
	bool active = false;
	for (int base = threadIdx.x; base < n; base += 4 * blockDim.x) {
		double ex[4];
		int hu[4];
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const int u = base + j * blockDim.x;
			hu[j] = u < n ? hs[u] : n;
			ex[j] = hu[j] < n ? __ldcg(&G.excess[u]) : 0.0;
		}
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const int u = base + j * blockDim.x;
			if (u < n) h[u] = hu[j];
			active |= ex[j] > 0.0;
		}
	}
	if (threadIdx.x == 0) G.flags[8] += level; // statistics: BFS levels
	if (G.debug) {
		const int cnt = __syncthreads_count(active);
		if (threadIdx.x == 0) printf("[mf] relabel: levels=%d reached=%d threads_with_active=%d\n", level, s_tail, cnt);
	}
	return __syncthreads_or(active) != 0;
}

// One asynchronous push-relabel step of node u (Hong & He's lock-free rule: push to the LOWEST residual neighbour if it
// is lower, else lift to one above it). Only the owner thread of u lowers excess[u] / cap[out-arcs of u] and writes
// height[u]; everybody else only adds to them, so the atomics below can never drive a value negative.
__device__ __forceinline__ bool mf_process(const FlowGraphDev &G, volatile int32_t *h, int u) {
	const int n = G.n;
	volatile double *excess = G.excess, *cap = G.cap;
	// everything a visit needs about u itself travels in ONE round trip, also for the (many) nodes that turn out to be
	// inactive: a visit of an active node is a chain of dependent round trips (~370 cycles each), and the length of
	// that chain times the number of cycles is what an asynchronous phase costs
	const double e0 = excess[u];
	const int hu = h[u];
	const double sc = ld_volatile_f64_if(G.sink_cap + u, true);
	const int a0 = ld_nc_s32_if(G.arc_off + u, true, 0), a1 = ld_nc_s32_if(G.arc_off + u + 1, true, 0);
	if (!(e0 > 0.0) || hu >= n) return false;
	double e = e0;
	if (sc > 0.0) { // the sink (height 0) is always the lowest neighbour; what it cannot absorb goes on below
		const double d = fmin(e, sc);
		G.sink_cap[u] = sc - d; // only u's owner writes its sink link
		atomicAdd(&G.excess[u], -d);
		e -= d;
		if (!(e > 0.0)) return true;
	}
	if (a1 - a0 > kWideDegree) {
		// high-degree node (a label-cost auxiliary node): one pass that pushes to EVERY lower residual neighbour, so
		// that its budget is spread in one visit instead of one neighbour per visit
		double rem = e;
		int lowest = 0x7fffffff;
		for (int a = a0; a < a1 && rem > 0.0; ++a) {
			const double c = cap[a];
			if (!(c > 0.0)) continue;
			const int v = G.arc_head[a];
			const int hv = h[v];
			if (hv < hu) {
				const double d = fmin(rem, c);
				atomicAdd(&G.cap[a], -d);
				atomicAdd(&G.cap[G.arc_rev[a]], d);
				atomicAdd(&G.excess[v], d);
				rem -= d;
			} else {
				lowest = min(lowest, hv);
			}
		}
		if (rem < e) atomicAdd(&G.excess[u], rem - e);
		if (rem > 0.0) { // everything lower is saturated: lift above the lowest remaining residual neighbour
			for (int a = a0; a < a1; ++a)
				if (cap[a] > 0.0) lowest = min(lowest, (int)h[G.arc_head[a]]);
			h[u] = lowest == 0x7fffffff ? n : min(max(lowest + 1, hu), n);
		}
		return true;
	}
	// lowest residual neighbour (the first one in arc order among equals). Chunks of 8 arcs with predicated loads: the
	// capacities and heads of a chunk travel together, then the heights of the residual heads -- two dependent round
	// trips per chunk instead of two per arc.
	int best_h = 0x7fffffff, best_a = -1;
	for (int base = a0; base < a1; base += 8) {
		double c[8];
		int v[8], hv[8];
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			c[j] = ld_volatile_f64_if(G.cap + base + j, base + j < a1);
			v[j] = ld_nc_s32_if(G.arc_head + base + j, base + j < a1, 0);
		}
#pragma unroll
		for (int j = 0; j < 8; ++j) hv[j] = ld_volatile_s32_if(G.height[0] + v[j], c[j] > 0.0, 0x7fffffff);
#pragma unroll
		for (int j = 0; j < 8; ++j)
			if (hv[j] < best_h) {
				best_h = hv[j];
				best_a = base + j;
			}
	}
	if (best_a < 0) { // no residual arc at all: the excess is stranded on the source side
		h[u] = n;
		return true;
	}
	if (hu > best_h) {
		const double d = fmin(e, cap[best_a]);
		atomicAdd(&G.cap[best_a], -d);
		atomicAdd(&G.cap[G.arc_rev[best_a]], d);
		atomicAdd(&G.excess[G.arc_head[best_a]], d);
		atomicAdd(&G.excess[u], -d);
	} else {
		h[u] = min(best_h + 1, n);
	}
	return true;
}

// Block-cooperative discharge of one wide node: every thread looks at a strided share of the arcs, the node's excess is
// handed out in arc order by a block-wide exclusive prefix sum of the eligible capacities (lower residual neighbours),
// and the node is lifted when excess remains after every lower neighbour has been saturated. Same rule as the
// single-thread wide branch of mf_process, ~100x fewer dependent memory round trips per visit.
__device__ bool mf_process_wide_block(const FlowGraphDev &G, volatile int32_t *h, int u) {
	__shared__ double s_e, s_wsum[kMfThreads / 32], s_gsum[kMfThreads / 32];
	__shared__ int s_hu, s_wlow[kMfThreads / 32];
	const int n = G.n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	volatile double *excess = G.excess;
	if (threadIdx.x == 0) {
		s_e = excess[u];
		s_hu = h[u];
	}
	__syncthreads();
	const double e = s_e;
	const int hu = s_hu;
	if (!(e > 0.0) || hu >= n) return false; // block-uniform
	const int a0 = G.arc_off[u], a1 = G.arc_off[u + 1];
	double rem = e, given = 0.0;
	int lowest = 0x7fffffff;
	// Two passes over the arcs. Pass 0 hands every lower neighbour only what its own sink link can still absorb (those
	// units leave the graph on the neighbour's next visit instead of trickling on through lambda-sized n-links); pass 1
	// pushes whatever is left up to the arc capacities. Both are ordinary pushes to lower residual neighbours.
	for (int pass = 0; pass < 2; ++pass) {
		for (int base = a0; base < a1; base += kMfThreads) { // block-uniform trip count
			const int a = base + threadIdx.x;
			// two dependent round trips per chunk (capacity + head, then what is needed of the head), predicated loads
			double want = 0.0;
			const bool in = a < a1;
			const double c = ld_volatile_f64_if(G.cap + a, in);
			const int v = ld_nc_s32_if(G.arc_head + a, in, 0);
			const bool residual = c > 0.0;
			const int hv = ld_volatile_s32_if(G.height[0] + v, residual, 0x7fffffff);
			double sink_v = 0.0, excess_v = 0.0;
			if (pass == 0) {
				sink_v = ld_volatile_f64_if(G.sink_cap + v, residual);
				excess_v = ld_volatile_f64_if(G.excess + v, residual);
			}
			if (residual) {
				if (hv < hu)
					want = pass == 0 ? fmin(c, fmax(0.0, sink_v - excess_v)) : c;
				else if (pass == 1)
					lowest = min(lowest, hv);
			}
			double incl = want;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const double t = __shfl_up_sync(0xffffffffu, incl, o);
				if (lane >= o) incl += t;
			}
			if (lane == 31) s_wsum[warp] = incl;
			__syncthreads();
			double before = 0.0, total = 0.0;
#pragma unroll
			for (int w = 0; w < kMfThreads / 32; ++w) {
				if (w < warp) before += s_wsum[w];
				total += s_wsum[w];
			}
			const double excl = before + incl - want;
			const double give = fmin(want, fmax(0.0, rem - excl));
			if (give > 0.0) {
				atomicAdd(&G.cap[a], -give);
				atomicAdd(&G.cap[G.arc_rev[a]], give);
				atomicAdd(&G.excess[v], give);
			}
			given += give;
			rem = fmax(0.0, rem - total);
			__syncthreads(); // s_wsum is reused by the next chunk
			if (!(rem > 0.0)) break; // block-uniform
		}
		if (!(rem > 0.0)) break; // block-uniform
	}
	// what left the node, and the lowest neighbour that is still residual but not lower
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		given += __shfl_xor_sync(0xffffffffu, given, o);
		lowest = min(lowest, __shfl_xor_sync(0xffffffffu, lowest, o));
	}
	if (lane == 0) {
		s_gsum[warp] = given;
		s_wlow[warp] = lowest;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		double g = 0.0;
		int low = 0x7fffffff;
		for (int w = 0; w < kMfThreads / 32; ++w) {
			g += s_gsum[w];
			low = min(low, s_wlow[w]);
		}
		if (g > 0.0) atomicAdd(&G.excess[u], -fmin(g, e));
		if (rem > 0.0) h[u] = low == 0x7fffffff ? n : min(max(low + 1, hu), n);
	}
	__syncthreads();
	return true;
}

constexpr int kAsyncCycles = 64;
constexpr int kMaxRounds = 100000;

__global__ void __launch_bounds__(kMfThreads) k_maxflow(FlowGraphDev G) {
	cg::grid_group grid = cg::this_grid();
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	const int nthreads = gridDim.x * blockDim.x;
	const int n = G.n;
	int32_t *h = G.height[0];
	int round = 0;
	for (; round < kMaxRounds; ++round) {
		// exact distance-to-sink labels; nodes that cannot reach the sink any more get height n and go quiet
		const long long t0 = clock64();
		if (G.block_bfs) {
			if (blockIdx.x == 0) {
				const bool any = mf_global_relabel_block(G, h, round == 0);
				if (threadIdx.x == 0) {
					G.flags[3] = any ? 1 : 0; // re-written only after the barrier that ends this round
					G.flags[14] = 0;          // stop flag of the asynchronous phase
				}
			}
			__threadfence();
			grid.sync();
			if (tid == 0) G.flags[10] += (int)((clock64() - t0) >> 6);
			if (*(volatile int32_t *)&G.flags[3] == 0) break;
		} else {
			mf_global_relabel(G, h, grid, tid, nthreads);
			if (tid == 0) G.flags[10] += (int)((clock64() - t0) >> 6);
			bool active = false;
			for (int u = tid; u < n; u += nthreads) active |= (G.excess[u] > 0.0 && h[u] < n);
			if (tid == 0) {
				G.flags[3 + ((round + 1) % 3)] = 0;
				G.flags[14] = 0; // stop flag of the asynchronous phase
			}
			if (active) G.flags[3 + (round % 3)] = 1;
			grid.sync();
			if (G.flags[3 + (round % 3)] == 0) break;
		}
		const long long t1 = clock64();
		// Asynchronous phase: no grid barriers, every thread keeps discharging its own nodes. It ends for the whole grid
		// at once: when one block has spent its budget of busy cycles (stop flag), or when no block has reported work
		// for quiet_cycles clocks (blocks with work bump a heartbeat counter at every check). A block must not leave
		// while others still push: nodes are spread over the blocks by index, so a unit of flow crossing the graph
		// lands in a different block at almost every hop, and a block that had left stranded it until the next relabel
		// (PXB_MF_LOCAL_EXIT=1 restores that rule for A/B: 9-12 relabel rounds per heavy move instead of 7-9).
		__shared__ int s_ctl;
		bool busy = false;
		int idle_checks = 0, busy_cycles = 0;
		int last_hb = 0;
		long long last_change = clock64();
		if (threadIdx.x == 0) last_hb = *(volatile int32_t *)&G.flags[13];
		const int hard_cap = 64 * G.async_cycles;
		for (int c = 0; c < hard_cap; ++c) {
			for (int w = blockIdx.x; w < G.wide_count; w += gridDim.x) busy |= mf_process_wide_block(G, h, G.wide_begin + w);
			for (int u = tid; u < G.wide_begin; u += nthreads) busy |= mf_process(G, h, u);
			if ((c & 3) != 3) continue;
			const int block_busy = __syncthreads_or(busy);
			busy = false;
			if (G.local_exit) { // every block for itself: quiet for idle_checks x 8 cycles, or the cycle budget
				idle_checks = block_busy ? 0 : idle_checks + 1;
				if (idle_checks >= 2 * G.idle_checks || c + 1 >= G.async_cycles) break; // block-uniform
				continue;
			}
			if (threadIdx.x == 0) {
				int ctl = 0;
				if (block_busy) {
					atomicAdd(&G.flags[13], 1);
					busy_cycles += 4;
					if (busy_cycles >= G.async_cycles) *(volatile int32_t *)&G.flags[14] = 1;
					last_change = clock64();
				} else {
					const int hb = *(volatile int32_t *)&G.flags[13];
					if (hb != last_hb) {
						last_hb = hb;
						last_change = clock64();
					} else if (clock64() - last_change > G.quiet_cycles) {
						ctl = 1;
					}
				}
				if (*(volatile int32_t *)&G.flags[14] != 0) ctl = 1;
				s_ctl = ctl;
			}
			__syncthreads();
			if (s_ctl) break; // block-uniform
		}
		__threadfence();
		grid.sync();
		if (tid == 0) G.flags[12] += (int)((clock64() - t1) >> 6);
	}
	// heights now hold the final reachability: height < n  <=>  the node can reach the sink in the residual graph
	if (tid == 0) {
		G.flags[6] = round + 1;
		G.flags[7] = round < kMaxRounds ? 1 : 0;
	}
}

// Launch geometry of k_maxflow: a cooperative grid of at most 4 blocks per SM (one node per thread is enough), and --
// when heights + queue of the graph fit -- 2n ints of dynamic shared memory for the single-block global relabel.
struct MfLaunch {
	int grid = 1;
	size_t smem = 0;
};
static int mf_launch_config(pxb_ctx *ctx, FlowGraphDev &G, int min_grid, MfLaunch &out) {
	// tuning / A-B switches, read per call (tests flip them): PXB_MF_GRID_BFS=1 forces the grid-wide relabel,
	// PXB_MF_SMEM_KB caps the shared memory the single-block relabel may use (selects its smaller layouts)
	const bool grid_bfs_only = getenv("PXB_MF_GRID_BFS") != nullptr;
	const int async_cycles = getenv("PXB_MF_ASYNC") ? atoi(getenv("PXB_MF_ASYNC")) : kAsyncCycles;
	const int idle_checks = getenv("PXB_MF_IDLE") ? atoi(getenv("PXB_MF_IDLE")) : 4;
	G.debug = (getenv("PXB_MF_STATS") && getenv("PXB_MF_STATS")[0] == '3') ? 1 : 0;
	G.local_exit = getenv("PXB_MF_LOCAL_EXIT") ? 1 : 0;
	G.quiet_cycles = (long long)((getenv("PXB_MF_QUIET_US") ? atof(getenv("PXB_MF_QUIET_US")) : 10.0) * 1965.0);
	G.async_cycles = std::max(8, async_cycles);
	const bool async_given = getenv("PXB_MF_ASYNC") != nullptr;
	G.idle_checks = std::max(1, idle_checks);
	static bool attribute_set[64] = {}; // per device (function attributes belong to the device's context)
	constexpr size_t kMaxSmem = 200 * 1024;
	size_t smem_cap = kMaxSmem;
	if (const char *e = getenv("PXB_MF_SMEM_KB")) smem_cap = std::min(kMaxSmem, (size_t)std::max(0, atoi(e)) * 1024);
	if (ctx->device < 0 || ctx->device >= 64 || !attribute_set[ctx->device]) {
		PXB_CUDA(cudaFuncSetAttribute(k_maxflow, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
		if (ctx->device >= 0 && ctx->device < 64) attribute_set[ctx->device] = true;
	}
	const size_t need = sizeof(int32_t) * 2 * (size_t)G.n, need_offs = sizeof(int32_t) * (3 * (size_t)G.n + 1);
	const size_t need_ulist = need_offs + sizeof(int32_t) * (size_t)G.n; // + the unlabelled list of the bottom-up levels
	G.block_bfs = (grid_bfs_only || G.wide_count > 32) ? 0 : (need_offs <= smem_cap ? 2 : (need <= smem_cap ? 1 : 0));
	if (G.block_bfs == 2 && need_ulist <= smem_cap && !getenv("PXB_MF_TOP_DOWN")) G.block_bfs = 3;
	out.smem = G.block_bfs == 3 ? need_ulist : (G.block_bfs == 2 ? need_offs : (G.block_bfs == 1 ? need : 0));
	// Graphs too large for the single-block relabel pay a grid-wide BFS per round (hundreds of levels of ~6 us on a
	// 10^5-node neighbourhood graph): longer push phases trade rounds for cycles there (C5 at lambda = 0.1: 2.0 s with 64
	// busy cycles per phase, 1.36 s with 256, 1.20 s with 1024, 1.86 s with 4096)
	if (G.block_bfs == 0 && !async_given) G.async_cycles = 1024;
	// occupancy for this shared-memory size: asked once per (thread, device, size) -- a fit makes ~200 cuts
	thread_local struct {
		int device;
		size_t smem;
		int blocks;
	} occ = {-1, 0, 0};
	if (occ.device != ctx->device || occ.smem != out.smem) {
		int b = 0;
		PXB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_maxflow, kMfThreads, out.smem));
		occ = {ctx->device, out.smem, b};
	}
	const int blocks_per_sm = occ.blocks;
	if (blocks_per_sm < 1) {
		set_error("k_maxflow does not fit on an SM with %zu bytes of shared memory", out.smem);
		return PXB_ERR_CUDA;
	}
	const int want = (G.n + kMfThreads - 1) / kMfThreads;
	const int cap = ctx->sm_count * std::min(blocks_per_sm, 4);
	out.grid = std::max(1, std::min(std::max(min_grid, want), cap));
	return PXB_OK;
}

// ---- host-side graph assembly with the reference's Energy/Graph arithmetic -----------------------------------
struct FlowGraphHost {
	int n = 0;
	std::vector<double> tr;                 // terminal capacity, source minus sink (gcr/graph.h add_tweights)
	std::vector<int32_t> tail, head;        // arc pairs: arc 2p = tail->head, 2p+1 = head->tail
	std::vector<double> cap_fwd, cap_rev;
	explicit FlowGraphHost(int n_) : n(n_), tr((size_t)n_, 0.0) {}
	int add_node() {
		tr.push_back(0.0);
		return n++;
	}
	void add_tweights(int i, double cap_source, double cap_sink) { // gcr/graph.h:add_tweights
		const double delta = tr[i];
		if (delta > 0)
			cap_source += delta;
		else
			cap_sink -= delta;
		tr[i] = cap_source - cap_sink;
	}
	void add_edge(int i, int j, double cap, double rev) {
		tail.push_back(i);
		head.push_back(j);
		cap_fwd.push_back(cap);
		cap_rev.push_back(rev);
	}
	void add_term1(int x, double A, double B) { add_tweights(x, B, A); } // gcr/energy.h:204-208
	void add_term2(int x, int y, double A, double B, double C, double D) { // gcr/energy.h:210-256
		add_tweights(x, D, A);
		B -= A;
		C -= D;
		if (B < 0) {
			add_tweights(x, 0, B);
			add_tweights(y, 0, -B);
			add_edge(x, y, 0, B + C);
		} else if (C < 0) {
			add_tweights(x, 0, -C);
			add_tweights(y, 0, C);
			add_edge(x, y, B + C, 0);
		} else {
			add_edge(x, y, B, C);
		}
	}
};

// Solve the cut on the device. segment[i] = 1 iff node i ends on the SINK side (BK rule).
static int solve_min_cut(pxb_ctx *ctx, const FlowGraphHost &g, std::vector<uint8_t> &segment, int n_sites = -1) {
	const auto t_begin = std::chrono::steady_clock::now();
	const int n = g.n;
	const int pairs = (int)g.tail.size();
	const int m = 2 * pairs;
	segment.assign((size_t)n, 0);
	if (n == 0) return PXB_OK;
	// The whole problem is assembled in ONE pinned host arena that mirrors the device arena (capacities, excess, sink
	// capacities, CSR offsets / heads / reverse arcs) and goes up in a single DMA.
	const size_t mm = (size_t)std::max(m, 1);
	const size_t bytes_d = sizeof(double) * (mm + 2 * (size_t)n);
	const size_t bytes_up = bytes_d + sizeof(int32_t) * ((size_t)n + 1 + 2 * mm);
	const size_t bytes_all = bytes_up + sizeof(int32_t) * (2 * (size_t)n + 16) + 64;
	PXB_TRY(ctx->reserve_pinned(bytes_all));
	PXB_TRY(ctx->partials.reserve(bytes_all));
	unsigned char *hp = static_cast<unsigned char *>(ctx->pinned);
	double *cap = reinterpret_cast<double *>(hp), *excess = cap + mm, *sink_cap = excess + n;
	int32_t *off = reinterpret_cast<int32_t *>(sink_cap + n), *head = off + (n + 1), *rev = head + mm;
	// CSR by tail, arcs of a node in insertion order
	std::fill(off, off + n + 1, 0);
	for (int p = 0; p < pairs; ++p) {
		off[g.tail[p] + 1]++;
		off[g.head[p] + 1]++;
	}
	for (int i = 0; i < n; ++i) off[i + 1] += off[i];
	std::vector<int32_t> fill(off, off + n);
	for (int p = 0; p < pairs; ++p) {
		const int a = fill[g.tail[p]]++, b = fill[g.head[p]]++;
		head[a] = g.head[p];
		cap[a] = g.cap_fwd[p];
		rev[a] = b;
		head[b] = g.tail[p];
		cap[b] = g.cap_rev[p];
		rev[b] = a;
	}
	bool any_sink = false;
	for (int i = 0; i < n; ++i) {
		excess[i] = g.tr[i] > 0 ? g.tr[i] : 0.0;
		sink_cap[i] = g.tr[i] < 0 ? -g.tr[i] : 0.0;
		any_sink |= sink_cap[i] > 0;
	}
	if (!any_sink) return PXB_OK; // nothing can reach the sink: everything is SOURCE
	unsigned char *dp = static_cast<unsigned char *>(ctx->partials.ptr);
	double *d_cap = reinterpret_cast<double *>(dp), *d_excess = d_cap + mm, *d_sink = d_excess + n;
	int32_t *d_off = reinterpret_cast<int32_t *>(d_sink + n), *d_head = d_off + (n + 1), *d_rev = d_head + mm;
	int32_t *d_h0 = d_rev + mm, *d_h1 = d_h0 + n, *d_flags = d_h1 + n;
	int32_t *h_h0 = reinterpret_cast<int32_t *>(hp + bytes_up), *h_flags = h_h0 + 2 * (size_t)n;
	cudaStream_t st = ctx->stream;
	PXB_CUDA(cudaMemcpyAsync(dp, hp, bytes_up, cudaMemcpyHostToDevice, st));
	PXB_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int32_t) * 16, st));
	FlowGraphDev G;
	G.n = n;
	G.m = m;
	G.wide_begin = n_sites >= 0 ? n_sites : n;
	G.wide_count = n - G.wide_begin;
	G.arc_off = d_off;
	G.arc_head = d_head;
	G.arc_rev = d_rev;
	G.cap = d_cap;
	G.excess = d_excess;
	G.sink_cap = d_sink;
	G.height[0] = d_h0;
	G.height[1] = d_h1;
	G.flags = d_flags;
	MfLaunch lc;
	PXB_TRY(mf_launch_config(ctx, G, 1, lc));
	const int grid = lc.grid;
	void *args[] = {&G};
	PXB_CUDA(cudaLaunchCooperativeKernel((void *)k_maxflow, dim3(grid), dim3(kMfThreads), args, lc.smem, st));
	ctx->launches++;
	const int32_t *h = h_h0, *flags = h_flags;
	PXB_CUDA(cudaMemcpyAsync(h_h0, d_h0, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
	PXB_CUDA(cudaMemcpyAsync(h_flags, d_flags, sizeof(int32_t) * 16, cudaMemcpyDeviceToHost, st));
	PXB_TRY(ctx_wait(ctx));
	if (flags[7] != 1 || flags[6] == 0) {
		set_error("max-flow did not converge within %d relabel rounds", kMaxRounds);
		return PXB_ERR_CUDA;
	}
	for (int i = 0; i < n; ++i) segment[i] = h[i] < n ? 1 : 0;
	if (getenv("PXB_MF_STATS")) {
		static int calls = 0;
		static double total_ms = 0;
		const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
		total_ms += ms;
		if (++calls % 20 == 0 || ms > 3 || getenv("PXB_MF_STATS")[0] == '2')
			fprintf(stderr, "[pxb maxflow] call %d: n=%d arcs=%d rounds=%d grid=%d  %.2f ms (total %.1f ms) bfs_levels=%d relabel=%.2f ms async+check=%.2f ms\n", calls, n, m,
			        flags[6], grid, ms, total_ms, flags[8], (double)flags[10] * 64 / 1.965e6, (double)flags[12] * 64 / 1.965e6);
	}
	return PXB_OK;
}

// ---- GC-RANSAC local optimisation entirely on the device --------------------------------------------------------
// GCRANSAC::labeling (gcr/GCRANSAC.h:914-1022) for one model: unary terms (k_lo_unary), the pairwise graph over the
// FIRST occurrence of every unordered neighbour pair (the reference's used_edges matrix, :964-1010) and the st-cut.
// The arc skeleton depends on the neighbourhood graph only, so it is built once per graph and cached on the device;
// per cut only capacities change, and those are computed by k_lo_assemble with the reference's own add_term1 /
// add_term2 / add_tweights arithmetic: the terminal capacity of node i is a sequential function of i's own pairs.
struct LoSkeleton {
	uint64_t key = 0;
	int64_t N = 0;
	int pairs = 0;
	DevBuf buf; // arc_off[N+1] arc_head[m] arc_rev[m] pair_j[p] pair_fwd[p] pair_rev[p] first_off[N+1]
	int32_t *arc_off = nullptr, *arc_head = nullptr, *arc_rev = nullptr, *pair_j = nullptr, *pair_fwd = nullptr,
	        *pair_rev = nullptr, *first_off = nullptr;
	std::vector<int32_t> arc_off_host; // for the launch plan of the cluster-resident engine
	McPlan plan;
	uint64_t plan_key = ~0ull;
};

// The launch plan of k_maxflow_cluster depends on the arc skeleton and on three tuning switches only: cached with the
// skeleton, recomputed when a switch changes (tests flip them).
static uint64_t mc_env_key() {
	auto val = [](const char *name) { const char *e = getenv(name); return e ? (uint64_t)(atoi(e) + 1) : 0ull; };
	return val("PXB_MF_CLUSTER") * 1000003ull + val("PXB_MC_CSIZE") * 10007ull + val("PXB_MC_SMEM_KB");
}

// 64-bit-word multiply-xor hash of the CSR arrays (the cache key of the skeleton; ~10 GB/s on the host)
static uint64_t fnv1a(const void *data, size_t bytes, uint64_t h = 1469598103934665603ull) {
	const unsigned char *p = static_cast<const unsigned char *>(data);
	size_t i = 0;
	if (bytes >= 4096) { // large arrays (the data-cost matrix of the labelling memo): four independent lanes, folded at the end
		uint64_t l[4] = {h, h ^ 0x9E3779B97F4A7C15ull, h ^ 0xBF58476D1CE4E5B9ull, h ^ 0x94D049BB133111EBull};
		for (; i + 32 <= bytes; i += 32) {
			uint64_t w[4];
			memcpy(w, p + i, 32);
			for (int k = 0; k < 4; ++k) {
				l[k] = (l[k] ^ w[k]) * 0x9E3779B97F4A7C15ull;
				l[k] ^= l[k] >> 29;
			}
		}
		for (int k = 0; k < 4; ++k) {
			h = (h ^ l[k]) * 0x9E3779B97F4A7C15ull;
			h ^= h >> 29;
		}
	}
	for (; i + 8 <= bytes; i += 8) {
		uint64_t w;
		memcpy(&w, p + i, 8);
		h = (h ^ w) * 0x9E3779B97F4A7C15ull;
		h ^= h >> 29;
	}
	for (; i < bytes; ++i) h = (h ^ p[i]) * 1099511628211ull;
	return h;
}

// cache key of a neighbourhood graph: content hash, or the key the driver registered for exactly these arrays
uint64_t csr_content_key(pxb_ctx *ctx, int64_t N, const int32_t *off, const int32_t *idx) {
	if (ctx->trusted_csr_key != 0 && off == ctx->trusted_csr_off && idx == ctx->trusted_csr_idx) return ctx->trusted_csr_key;
	uint64_t key = fnv1a(off, sizeof(int32_t) * (size_t)(N + 1));
	return fnv1a(idx, sizeof(int32_t) * (size_t)off[N], key) ^ ((uint64_t)ctx->device << 56) ^ (uint64_t)N;
}

void lo_skeleton_free(void *p) {
	if (!p) return;
	LoSkeleton *sk = static_cast<LoSkeleton *>(p);
	sk->buf.release();
	delete sk;
}

// the skeleton lives in the context (one per context: contexts may be driven by different host threads)
static int lo_skeleton(pxb_ctx *ctx, int64_t N, const int32_t *off, const int32_t *idx) {
	if (!ctx->lo_skeleton) ctx->lo_skeleton = new LoSkeleton();
	LoSkeleton &g_lo = *static_cast<LoSkeleton *>(ctx->lo_skeleton);
	const uint64_t key = csr_content_key(ctx, N, off, idx);
	if (g_lo.key == key && g_lo.N == N && g_lo.buf.ptr) return PXB_OK;
	std::unordered_set<uint64_t> used;
	used.reserve((size_t)off[N] * 2);
	std::vector<int32_t> pi, pj, first_off((size_t)N + 1, 0);
	for (int64_t i = 0; i < N; ++i) {
		for (int32_t e = off[i]; e < off[i + 1]; ++e) {
			const int64_t j = idx[e];
			if (j == i || j < 0) continue;
			const uint64_t k2 = (uint64_t)std::min(i, j) * (uint64_t)N + (uint64_t)std::max(i, j);
			if (!used.insert(k2).second) continue;
			pi.push_back((int32_t)i);
			pj.push_back((int32_t)j);
		}
		first_off[i + 1] = (int32_t)pi.size(); // pairs whose first endpoint is i are contiguous
	}
	const int pairs = (int)pi.size(), m = 2 * pairs;
	std::vector<int32_t> aoff((size_t)N + 1, 0), head((size_t)std::max(m, 1)), rev((size_t)std::max(m, 1)), pf((size_t)std::max(pairs, 1)),
	    pr((size_t)std::max(pairs, 1));
	for (int p = 0; p < pairs; ++p) {
		aoff[pi[p] + 1]++;
		aoff[pj[p] + 1]++;
	}
	for (int64_t i = 0; i < N; ++i) aoff[i + 1] += aoff[i];
	std::vector<int32_t> fill(aoff.begin(), aoff.end() - 1);
	for (int p = 0; p < pairs; ++p) { // same arc order as solve_min_cut's CSR
		const int a = fill[pi[p]]++, b = fill[pj[p]]++;
		head[a] = pj[p];
		rev[a] = b;
		head[b] = pi[p];
		rev[b] = a;
		pf[p] = a;
		pr[p] = b;
	}
	const size_t ints = 2 * ((size_t)N + 1) + 2 * (size_t)std::max(m, 1) + 3 * (size_t)std::max(pairs, 1);
	PXB_TRY(g_lo.buf.reserve(sizeof(int32_t) * ints));
	int32_t *b = g_lo.buf.as<int32_t>();
	g_lo.arc_off = b;
	g_lo.arc_head = g_lo.arc_off + (N + 1);
	g_lo.arc_rev = g_lo.arc_head + std::max(m, 1);
	g_lo.pair_j = g_lo.arc_rev + std::max(m, 1);
	g_lo.pair_fwd = g_lo.pair_j + std::max(pairs, 1);
	g_lo.pair_rev = g_lo.pair_fwd + std::max(pairs, 1);
	g_lo.first_off = g_lo.pair_rev + std::max(pairs, 1);
	cudaStream_t st = ctx->stream;
	PXB_CUDA(cudaMemcpyAsync(g_lo.arc_off, aoff.data(), sizeof(int32_t) * aoff.size(), cudaMemcpyHostToDevice, st));
	PXB_CUDA(cudaMemcpyAsync(g_lo.arc_head, head.data(), sizeof(int32_t) * head.size(), cudaMemcpyHostToDevice, st));
	PXB_CUDA(cudaMemcpyAsync(g_lo.arc_rev, rev.data(), sizeof(int32_t) * rev.size(), cudaMemcpyHostToDevice, st));
	PXB_CUDA(cudaMemcpyAsync(g_lo.pair_j, pj.data(), sizeof(int32_t) * (size_t)pairs, cudaMemcpyHostToDevice, st));
	PXB_CUDA(cudaMemcpyAsync(g_lo.pair_fwd, pf.data(), sizeof(int32_t) * (size_t)pairs, cudaMemcpyHostToDevice, st));
	PXB_CUDA(cudaMemcpyAsync(g_lo.pair_rev, pr.data(), sizeof(int32_t) * (size_t)pairs, cudaMemcpyHostToDevice, st));
	PXB_CUDA(cudaMemcpyAsync(g_lo.first_off, first_off.data(), sizeof(int32_t) * first_off.size(), cudaMemcpyHostToDevice, st));
	PXB_TRY(ctx_wait(ctx)); // the host vectors go out of scope
	g_lo.key = key;
	g_lo.N = N;
	g_lo.pairs = pairs;
	g_lo.arc_off_host = aoff;
	g_lo.plan_key = ~0ull;
	return PXB_OK;
}

// One thread per node: terminal capacity with the reference's add_tweights sequence (gcr/graph.h), n-link capacities
// of the pairs this node opens. add_term1(i, e0, e1) = add_tweights(i, e1, e0) (gcr/energy.h:204-208); add_term2(i, j,
// A = e00 lambda, B = lambda, C = lambda, D = 0) = add_tweights(i, 0, A); edge(i -> j) = B - A, edge(j -> i) = C
// (:210-256; B - A >= 0 because d <= 1).
__global__ void k_lo_assemble(int64_t N, double lambda, const double *__restrict__ d, const double *__restrict__ e0,
                              const double *__restrict__ e1, const int32_t *__restrict__ first_off,
                              const int32_t *__restrict__ pair_j, const int32_t *__restrict__ pair_fwd,
                              const int32_t *__restrict__ pair_rev, double *__restrict__ cap, double *__restrict__ excess,
                              double *__restrict__ sink_cap) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= N) return;
	double tr = 0.0;
	auto add_tweights = [&](double cap_source, double cap_sink) {
		const double delta = tr;
		if (delta > 0) cap_source = __dadd_rn(cap_source, delta);
		else cap_sink = __dsub_rn(cap_sink, delta);
		tr = __dsub_rn(cap_source, cap_sink);
	};
	add_tweights(e1[i], e0[i]);
	const double di = d[i];
	for (int32_t p = first_off[i]; p < first_off[i + 1]; ++p) {
		const double energy_sum = __dadd_rn(di, d[pair_j[p]]);
		const double A = __dmul_rn(__dmul_rn(0.5, energy_sum), lambda);
		add_tweights(0.0, A);
		double B = __dsub_rn(lambda, A), C = lambda; // B -= A; C -= D (= 0 * lambda)
		if (B < 0) { // unreachable for d in [0, 1]; kept for exactness of the restatement
			C = __dadd_rn(B, C);
			B = 0.0;
		}
		cap[pair_fwd[p]] = B;
		cap[pair_rev[p]] = C;
	}
	excess[i] = tr > 0 ? tr : 0.0;
	sink_cap[i] = tr < 0 ? -tr : 0.0;
}

__global__ void k_lo_collect(int64_t N, const int32_t *__restrict__ h, int n_nodes, uint8_t *__restrict__ seg) {
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < N) seg[i] = h[i] < n_nodes ? 1 : 0;
}

int launch_lo_unary(pxb_ctx *ctx, const double *model, double thr, double lambda, double *d, double *e0, double *e1);

// model_dev: device pointer to the model; seg_host: N bytes, 1 = inlier (SINK side)
int lo_labeling_enqueue(pxb_ctx *ctx, const double *model_dev, double thr, double lambda, const int32_t *csr_off_host,
                        const int32_t *csr_idx_host, uint8_t **seg_dev_out, int32_t **flags_dev_out) {
	const int64_t N = ctx->pts.N;
	PXB_TRY(lo_skeleton(ctx, N, csr_off_host, csr_idx_host));
	LoSkeleton &g_lo = *static_cast<LoSkeleton *>(ctx->lo_skeleton);
	if (g_lo.plan_key != mc_env_key()) {
		PXB_TRY(mf_cluster_plan(ctx, (int)N, 0, g_lo.arc_off_host.data(), g_lo.plan));
		g_lo.plan_key = mc_env_key();
	}
	const int n = (int)N, m = std::max(2 * g_lo.pairs, 1);
	const size_t bytes = sizeof(double) * ((size_t)m + 5 * (size_t)n) + sizeof(int32_t) * (2 * (size_t)n + 16) + (size_t)n + 256;
	PXB_TRY(ctx->partials.reserve(bytes));
	double *d_cap = ctx->partials.as<double>();
	double *d_excess = d_cap + m, *d_sink = d_excess + n, *d_d = d_sink + n, *d_e0 = d_d + n, *d_e1 = d_e0 + n;
	int32_t *d_h0 = reinterpret_cast<int32_t *>(d_e1 + n), *d_h1 = d_h0 + n, *d_flags = d_h1 + n;
	uint8_t *d_seg = reinterpret_cast<uint8_t *>(d_flags + 16);
	cudaStream_t st = ctx->stream;
	PXB_TRY(launch_lo_unary(ctx, model_dev, thr, lambda, d_d, d_e0, d_e1));
	PXB_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int32_t) * 16, st));
	k_lo_assemble<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, lambda, d_d, d_e0, d_e1, g_lo.first_off, g_lo.pair_j,
	                                                           g_lo.pair_fwd, g_lo.pair_rev, d_cap, d_excess, d_sink);
	ctx->launches++;
	FlowGraphDev G;
	G.n = n;
	G.m = 2 * g_lo.pairs;
	G.wide_begin = n;
	G.wide_count = 0;
	G.arc_off = g_lo.arc_off;
	G.arc_head = g_lo.arc_head;
	G.arc_rev = g_lo.arc_rev;
	G.cap = d_cap;
	G.excess = d_excess;
	G.sink_cap = d_sink;
	G.height[0] = d_h0;
	G.height[1] = d_h1;
	G.flags = d_flags;
	if (g_lo.plan.ok) { // the graph fits into one thread-block cluster (pxb_maxflow_cluster.cu)
		PXB_TRY(mf_cluster_launch(ctx, G, g_lo.plan));
	} else {
		MfLaunch lc;
		PXB_TRY(mf_launch_config(ctx, G, 1, lc));
		const int grid = lc.grid;
		void *args[] = {&G};
		PXB_CUDA(cudaLaunchCooperativeKernel((void *)k_maxflow, dim3(grid), dim3(kMfThreads), args, lc.smem, st));
		ctx->launches++;
	}
	k_lo_collect<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, d_h0, n, d_seg);
	ctx->launches++;
	*seg_dev_out = d_seg;
	*flags_dev_out = d_flags;
	return PXB_OK;
}

// True when lo_labeling_enqueue for this neighbourhood graph is a fixed sequence of plain launches (skeleton cached, cut on
// the cluster-resident engine): the driver may then capture its local-optimisation chain into a CUDA graph.
// *signature identifies what such a graph bakes in (the skeleton's content and buffers, the launch plan).
bool lo_labeling_capturable(pxb_ctx *ctx, const int32_t *csr_off_host, const int32_t *csr_idx_host, uint64_t *signature) {
	if (!ctx->lo_skeleton) return false;
	const LoSkeleton &g_lo = *static_cast<LoSkeleton *>(ctx->lo_skeleton);
	const int64_t N = ctx->pts.N;
	if (!(g_lo.buf.ptr && g_lo.N == N && g_lo.key == csr_content_key(ctx, N, csr_off_host, csr_idx_host) && g_lo.plan.ok &&
	      g_lo.plan_key == mc_env_key()))
		return false;
	const uint64_t parts[8] = {g_lo.key, (uint64_t)reinterpret_cast<uintptr_t>(g_lo.buf.ptr), (uint64_t)g_lo.pairs, (uint64_t)g_lo.plan.csize,
	                           (uint64_t)g_lo.plan.sites_per_cta, (uint64_t)g_lo.plan.arcs_per_cta, (uint64_t)g_lo.plan.smem, g_lo.plan_key};
	*signature = fnv1a(parts, sizeof(parts));
	return true;
}

int lo_labeling_device(pxb_ctx *ctx, const double *model_dev, double thr, double lambda, const int32_t *csr_off_host,
                       const int32_t *csr_idx_host, uint8_t *seg_host) {
	uint8_t *d_seg = nullptr;
	int32_t *d_flags = nullptr;
	PXB_TRY(lo_labeling_enqueue(ctx, model_dev, thr, lambda, csr_off_host, csr_idx_host, &d_seg, &d_flags));
	int32_t flags[16];
	PXB_CUDA(cudaMemcpyAsync(seg_host, d_seg, (size_t)ctx->pts.N, cudaMemcpyDeviceToHost, ctx->stream));
	PXB_CUDA(cudaMemcpyAsync(flags, d_flags, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
	PXB_TRY(ctx_wait(ctx));
	if (flags[7] != 1 || flags[6] == 0) {
		set_error("max-flow did not converge within %d relabel rounds", kMaxRounds);
		return PXB_ERR_CUDA;
	}
	return PXB_OK;
}

// ---- alpha-expansion (host control, device cuts) ----------------------------------------------------------------
namespace {
struct ExpansionProblem {
	const double *D; // N x L1 row-major
	int64_t N;
	int L1;
	double lambda, label_cost;
	// gco neighbour lists: setNeighbors(i, j) for every directed entry adds j to i's list AND i to j's list, each
	// with addFront (GCoptimization.cpp:1683-1708); finalizeNeighbors walks each list from its front.
	const int32_t *goff = nullptr, *gidx = nullptr;
};

// GCoptimization::compute_energy (GCoptimization.cpp:950-984) in the reference's summation order
double compute_energy(const ExpansionProblem &P, const std::vector<int32_t> &lab) {
	double data = 0;
	for (int64_t i = 0; i < P.N; ++i) data += P.D[i * P.L1 + lab[i]];
	double smooth = 0;
	for (int64_t i = 0; i < P.N; ++i)
		for (int32_t e = P.goff[i]; e < P.goff[i + 1]; ++e) {
			const int32_t nb = P.gidx[e];
			if (nb < i) smooth += 1.0 * (lab[i] != lab[nb] ? P.lambda : 0);
		}
	std::vector<char> used((size_t)P.L1, 0);
	for (int64_t i = 0; i < P.N; ++i) used[lab[i]] = 1;
	double lc = 0;
	// m_labelcostsAll is built by prepending (GCoptimization.cpp:894-925): iteration runs from the last label down
	for (int l = P.L1 - 1; l >= 0; --l)
		if (used[l]) lc += P.label_cost;
	return data + smooth + lc;
}

// The same value without walking every edge for every candidate labelling. The smooth term adds either lambda or 0 per
// neighbour pair, so its sequentially accumulated value depends only on HOW MANY pairs disagree (adding 0.0 changes
// nothing): smooth = lambda added k times, tabulated once. k is maintained incrementally from the sites that switch.
// The data term is re-summed sequentially over all sites (N additions), the label term over the used labels.
struct EnergyCache {
	// lambda_times[k] = ((lambda + lambda) + ...) k terms, sequentially rounded. A function of lambda and the edge count
	// only: kept with the skeleton across the dozen labellings of a fit (30 000 dependent additions per call otherwise)
	std::vector<double> own_table;
	const std::vector<double> *table = nullptr;
	int64_t pairs = 0;                // disagreeing neighbour pairs of the current labelling
	static int64_t count_pairs(const ExpansionProblem &P, const std::vector<int32_t> &lab) {
		int64_t k = 0;
		for (int64_t i = 0; i < P.N; ++i)
			for (int32_t e = P.goff[i]; e < P.goff[i + 1]; ++e) {
				const int32_t nb = P.gidx[e];
				if (nb < i && lab[i] != lab[nb]) ++k;
			}
		return k;
	}
	static void fill_table(const ExpansionProblem &P, std::vector<double> &t) {
		const size_t total = (size_t)P.goff[P.N] / 2 + 1;
		t.resize(total + 1);
		double acc = 0;
		for (size_t k = 0; k <= total; ++k) {
			t[k] = acc;
			acc += 1.0 * P.lambda;
		}
	}
	// cached: a table filled for the same lambda and edge count (the caller keeps it), or null; uniform: every site carries
	// the same label (the all-zero start of a first labelling: no pair disagrees)
	void init(const ExpansionProblem &P, const std::vector<int32_t> &lab, const std::vector<double> *cached, bool uniform) {
		if (cached) {
			table = cached;
		} else {
			fill_table(P, own_table);
			table = &own_table;
		}
		pairs = uniform ? 0 : count_pairs(P, lab);
	}
	// pairs of `cand`, which differs from `lab` exactly on `switched` (all of which take the label alpha)
	int64_t pairs_after(const ExpansionProblem &P, const std::vector<int32_t> &lab, const std::vector<int32_t> &cand,
	                    const std::vector<int32_t> &switched) const {
		int64_t k = pairs;
		for (int32_t s2 : switched)
			for (int32_t e = P.goff[s2]; e < P.goff[s2 + 1]; ++e) {
				const int32_t nb = P.gidx[e];
				// every entry of a site's list is one neighbour pair; a pair between two switched sites appears in both
				// lists and is taken from the larger index only
				if (cand[nb] != lab[nb] && nb > s2) continue;
				k += (int64_t)(cand[s2] != cand[nb]) - (int64_t)(lab[s2] != lab[nb]);
			}
		return k;
	}
	double energy(const ExpansionProblem &P, const std::vector<int32_t> &lab, int64_t k) const {
		double data = 0;
		for (int64_t i = 0; i < P.N; ++i) data += P.D[i * P.L1 + lab[i]];
		std::vector<char> used((size_t)P.L1, 0);
		for (int64_t i = 0; i < P.N; ++i) used[lab[i]] = 1;
		double lc = 0;
		for (int l = P.L1 - 1; l >= 0; --l)
			if (used[l]) lc += P.label_cost;
		return data + (*table)[(size_t)k] + lc;
	}
};
} // namespace

// ---- device-side assembly of one expansion move ------------------------------------------------------------------
// The binary energy of "expand alpha" lives on a FIXED arc skeleton: node s < N is site s, node N + l is the label-cost
// auxiliary node of label l. Site s owns one arc per entry of its gco neighbour list (both directions of an undirected
// pair are entries, so the reverse arc is simply the mirrored entry) plus one arc to the auxiliary node of its current
// label; auxiliary node l owns one arc per site currently labelled l. Only capacities change from move to move, and the
// site <-> auxiliary wiring changes when a move is accepted. Terminal capacities follow the reference's add_term1 /
// add_term2 / add_tweights arithmetic (gcr/energy.h:204-256, gcr/graph.h): the terminal capacity of a site is a
// sequential function of its OWN neighbour list (setupDataCostsExpansion :327-333, setupSmoothCostsExpansion :337-402
// with Potts * lambda: a neighbour that keeps alpha adds add_tweights(lambda, 0); an active neighbour with a smaller
// index adds add_tweights(e11, 0) and the arc pair (lambda, lambda - e11), e11 = lambda iff the two labels differ;
// setupLabelCostsExpansion :1131-1195 adds the arc pair (0, label_cost) to the auxiliary node, whose own terminal
// capacity is label_cost). Sites that already carry alpha are inert (all capacities zero).
__global__ void k_exp_assemble(int64_t N, int L1, int alpha, double lambda, double label_cost, const double *__restrict__ D,
                               const int32_t *__restrict__ lab, const int32_t *__restrict__ goff,
                               const int32_t *__restrict__ gidx, const int32_t *__restrict__ label_count,
                               const int32_t *__restrict__ arc_rev, double *__restrict__ cap, double *__restrict__ excess,
                               double *__restrict__ sink_cap, int32_t *__restrict__ flags, const int32_t *__restrict__ stop) {
	if (stop && *reinterpret_cast<const volatile int32_t *>(stop) != 0) return; // an earlier move of the batch changed the labelling
	const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s < 16) flags[s] = 0; // counters and status words of the max-flow kernel that follows
	if (s >= N + L1) return;
	if (s >= N) { // auxiliary node of label l: add_term1(aux, 0, label_cost) = add_tweights(aux, label_cost, 0)
		const int l = (int)(s - N);
		const bool live = label_cost > 0 && l != alpha && label_count[l] > 0;
		excess[s] = live ? label_cost : 0.0;
		sink_cap[s] = 0.0;
		return; // its arc capacities are written by the sites
	}
	const int ls = lab[s];
	const bool active = ls != alpha;
	const int64_t a0 = goff[s] + s; // arcs of site s: its list entries, then the auxiliary arc
	const int e0 = goff[s], e1 = goff[s + 1];
	double tr = 0.0;
	auto add_tweights = [&](double cap_source, double cap_sink) {
		const double delta = tr;
		if (delta > 0) cap_source = __dadd_rn(cap_source, delta);
		else cap_sink = __dsub_rn(cap_sink, delta);
		tr = __dsub_rn(cap_source, cap_sink);
	};
	if (active) add_tweights(D[s * L1 + ls], D[s * L1 + alpha]); // add_term1(v, D(alpha), D(current))
	double out_cap = 0.0;
	for (int e = e0; e < e1; ++e) {
		const int nb = gidx[e];
		const int lnb = lab[nb];
		double c = 0.0;
		if (active && lambda > 0) {
			if (lnb == alpha) {
				add_tweights(lambda, 0.0); // add_term1(v, e0 = 0, e1 = lambda)
			} else {
				const double e11 = (ls != lnb) ? lambda : 0.0;
				if (nb < s) {
					add_tweights(e11, 0.0); // add_term2(v, w, 0, lambda, lambda, e11): add_tweights(v, D = e11, A = 0)
					c = lambda;             // edge v -> w : B = e01 - e00
				} else {
					c = __dsub_rn(lambda, e11); // edge w -> v seen from w's side: C = e10 - e11
				}
			}
		}
		cap[a0 + (e - e0)] = c;
		out_cap = __dadd_rn(out_cap, c);
	}
	const int64_t aux_arc = a0 + (e1 - e0);
	double to_aux = 0.0, from_aux = 0.0;
	if (active && label_cost > 0) {
		// add_term2(v, aux, 0, 0, label_cost, 0): add_tweights(v, 0, 0) leaves tr as it is; arcs (v -> aux) = 0 and
		// (aux -> v) = label_cost, clamped to what v can pass on (its sink link plus its outgoing n-links): every maximum
		// flow and the set of nodes that reach the sink are unchanged, but the auxiliary node spreads its budget at once.
		add_tweights(0.0, 0.0);
		const double bound = __dadd_rn(tr < 0 ? -tr : 0.0, out_cap);
		from_aux = fmin(label_cost, bound);
	}
	cap[aux_arc] = to_aux;
	cap[arc_rev[aux_arc]] = from_aux;
	excess[s] = (active && tr > 0) ? tr : 0.0;
	sink_cap[s] = (active && tr < 0) ? -tr : 0.0;
}

// Rewires the site <-> auxiliary arcs after the labelling changed: site s points at node N + lab[s]; the auxiliary node
// of label l lists the sites labelled l in index order (rank = number of earlier sites with the same label).
__global__ void k_exp_wire_aux(int64_t N, int L1, int64_t E, const int32_t *__restrict__ lab, const int32_t *__restrict__ rank,
                               const int32_t *__restrict__ label_off, const int32_t *__restrict__ goff,
                               int32_t *__restrict__ arc_off, int32_t *__restrict__ arc_head, int32_t *__restrict__ arc_rev) {
	const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s < N) {
		const int l = lab[s];
		const int64_t site_arc = (int64_t)goff[s + 1] + s;                 // last arc of site s
		const int64_t aux_arc = E + N + label_off[l] + rank[s];            // its mirror in the auxiliary node's list
		arc_head[site_arc] = (int32_t)(N + l);
		arc_rev[site_arc] = (int32_t)aux_arc;
		arc_head[aux_arc] = (int32_t)s;
		arc_rev[aux_arc] = (int32_t)site_arc;
	}
	if (s <= L1) arc_off[N + s] = (int32_t)(E + N + label_off[s]); // label_off[L1] = N: arc_off[N + L1] = E + 2N
}

// The part of the expansion graph that depends on the neighbourhood graph only -- gco's adjacency lists and the
// site-to-site arcs with their mirrors -- is built once per graph and kept in the context (host copy for the energy
// bookkeeping, device copy for the assembly kernels); a fit calls PEARL::labeling a dozen times on the same graph.
struct ExpSkeleton {
	uint64_t key = 0;
	int64_t N = 0, E = 0;
	std::vector<int32_t> goff, gidx; // host: gco neighbour lists (also read by compute_energy / EnergyCache)
	std::vector<int32_t> arc_off_host; // site arc offsets (list entries + one auxiliary arc per site)
	McPlan plan;                       // launch plan of the cluster-resident engine (per label count: n_aux <= 16 only)
	uint64_t plan_key = ~0ull;
	// The last labelling solved on this graph. PEARL's final iteration re-labels with models refitted on an unchanged
	// labelling: data costs and initial labels are bit-identical to the previous call, and so is the result of this
	// deterministic-up-to-exact-ties function; it is handed back instead of re-running L + 1 expansion moves.
	uint64_t memo_key = 0;
	std::vector<int32_t> memo_labels;
	double memo_energy = 0;
	// EnergyCache's table of sequentially accumulated multiples of lambda (valid for lambda_table_of on this graph)
	std::vector<double> lambda_table;
	double lambda_table_of = -1.0;
	DevBuf buf;                      // arc_off[N+1] head[E+N] rev[E+N] goff[N+1] gidx[E]
	int32_t *arc_off = nullptr, *head = nullptr, *rev = nullptr, *d_goff = nullptr, *d_gidx = nullptr;
};

void exp_skeleton_free(void *p) {
	if (!p) return;
	ExpSkeleton *sk = static_cast<ExpSkeleton *>(p);
	sk->buf.release();
	delete sk;
}

static int exp_skeleton(pxb_ctx *ctx, int64_t N, const int32_t *off, const int32_t *idx, ExpSkeleton *&out) {
	if (!ctx->exp_skeleton) ctx->exp_skeleton = new ExpSkeleton();
	ExpSkeleton &sk = *static_cast<ExpSkeleton *>(ctx->exp_skeleton);
	out = &sk;
	const uint64_t key = csr_content_key(ctx, N, off, idx);
	if (sk.key == key && sk.N == N && sk.buf.ptr) return PXB_OK;
	std::vector<int32_t> grev; // mirrored entry of every neighbour-list entry
	// gco adjacency: per site a list built with addFront, so the last inserted neighbour comes first
	sk.goff.assign((size_t)N + 1, 0);
	for (int64_t i = 0; i < N; ++i)
		for (int32_t e = off[i]; e < off[i + 1]; ++e) {
			const int32_t j = idx[e];
			if (j == i) continue; // PEARL.h:535
			sk.goff[i + 1]++;
			sk.goff[j + 1]++;
		}
	for (int64_t i = 0; i < N; ++i) sk.goff[i + 1] += sk.goff[i];
	const int64_t E = sk.goff[N];
	sk.gidx.assign((size_t)std::max<int64_t>(E, 1), 0);
	grev.resize((size_t)std::max<int64_t>(E, 1));
	std::vector<int32_t> next(sk.goff.begin() + 1, sk.goff.end()); // fill every list from its back: reversed push order
	for (int64_t i = 0; i < N; ++i)
		for (int32_t e = off[i]; e < off[i + 1]; ++e) {
			const int32_t j = idx[e];
			if (j == i) continue;
			const int32_t pi = --next[i], pj = --next[j];
			sk.gidx[pi] = j;
			sk.gidx[pj] = (int32_t)i;
			grev[pi] = pj;
			grev[pj] = pi;
		}
	// arcs of site s: its neighbour list shifted by s (one slot per earlier site for that site's auxiliary arc)
	const size_t ms = (size_t)(E + N);
	std::vector<int32_t> arc_off((size_t)N + 1, 0), head(ms, 0), rev(ms, 0);
	for (int64_t s2 = 0; s2 <= N; ++s2) arc_off[s2] = (int32_t)(sk.goff[s2] + s2);
	for (int64_t s2 = 0; s2 < N; ++s2)
		for (int32_t e = sk.goff[s2]; e < sk.goff[s2 + 1]; ++e) {
			const int32_t nb = sk.gidx[e];
			head[(size_t)e + s2] = nb;
			rev[(size_t)e + s2] = grev[e] + nb; // the mirrored entry lives in nb's arc block, shifted by nb aux slots
		}
	const size_t ints = 2 * ((size_t)N + 1) + 2 * ms + (size_t)std::max<int64_t>(E, 1);
	PXB_TRY(sk.buf.reserve(sizeof(int32_t) * ints));
	sk.arc_off = sk.buf.as<int32_t>();
	sk.head = sk.arc_off + (N + 1);
	sk.rev = sk.head + ms;
	sk.d_goff = sk.rev + ms;
	sk.d_gidx = sk.d_goff + (N + 1);
	cudaStream_t st = ctx->stream;
	PXB_CUDA(cudaMemcpyAsync(sk.arc_off, arc_off.data(), sizeof(int32_t) * arc_off.size(), cudaMemcpyHostToDevice, st));
	PXB_CUDA(cudaMemcpyAsync(sk.head, head.data(), sizeof(int32_t) * ms, cudaMemcpyHostToDevice, st));
	PXB_CUDA(cudaMemcpyAsync(sk.rev, rev.data(), sizeof(int32_t) * ms, cudaMemcpyHostToDevice, st));
	PXB_CUDA(cudaMemcpyAsync(sk.d_goff, sk.goff.data(), sizeof(int32_t) * (size_t)(N + 1), cudaMemcpyHostToDevice, st));
	if (E > 0) PXB_CUDA(cudaMemcpyAsync(sk.d_gidx, sk.gidx.data(), sizeof(int32_t) * (size_t)E, cudaMemcpyHostToDevice, st));
	PXB_TRY(ctx_wait(ctx)); // the host vectors go out of scope
	sk.key = key;
	sk.N = N;
	sk.E = E;
	sk.arc_off_host = arc_off;
	sk.plan_key = ~0ull;
	sk.lambda_table_of = -1.0;
	return PXB_OK;
}

// Closes one move of a speculative batch (one block): how many sites would switch to alpha (SOURCE side: cannot reach the
// sink, GCoptimization.cpp:451-469), the max-flow status words of the move, and -- if anything switches or the cut did not
// converge -- the stop flag that turns the remaining launches of the batch into no-ops. res: kMoveRes ints per move.
constexpr int kMoveRes = 12;
__global__ void __launch_bounds__(1024)
    k_exp_close_move(int64_t N, int n_nodes, int alpha, const int32_t *__restrict__ lab, const int32_t *__restrict__ h,
                     const int32_t *__restrict__ goff, const int32_t *__restrict__ gidx, const int32_t *__restrict__ flags,
                     int32_t *__restrict__ res, int32_t *__restrict__ stop) {
	if (*reinterpret_cast<volatile int32_t *>(stop) != 0) return;
	__shared__ int s_cnt, s_delta;
	if (threadIdx.x == 0) s_cnt = 0, s_delta = 0;
	__syncthreads();
	// switched sites, and by how much the number of disagreeing neighbour pairs changes if they take alpha (an integer:
	// EnergyCache::pairs_after, every entry of a site's list is one pair, a pair of two switched sites counts at the larger
	// index only)
	int c = 0, delta = 0;
	for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
		const int li = lab[i];
		if (li == alpha || h[i] < n_nodes) continue;
		++c;
		for (int32_t e = goff[i]; e < goff[i + 1]; ++e) {
			const int32_t nb = gidx[e];
			const int ln = lab[nb];
			const bool nb_switched = ln != alpha && !(h[nb] < n_nodes);
			if (nb_switched && nb > i) continue;
			delta += (int)(alpha != (nb_switched ? alpha : ln)) - (int)(li != ln);
		}
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		c += __shfl_xor_sync(0xffffffffu, c, o);
		delta += __shfl_xor_sync(0xffffffffu, delta, o);
	}
	if ((threadIdx.x & 31) == 0 && c) {
		atomicAdd(&s_cnt, c);
		atomicAdd(&s_delta, delta);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		const bool converged = flags[7] == 1 && flags[6] != 0;
		res[0] = 1; // executed
		res[1] = s_cnt;
		res[2] = flags[6];
		res[3] = flags[7];
		res[4] = flags[8];
		res[5] = flags[10];
		res[6] = flags[12];
		res[7] = s_delta;
		res[8] = flags[5];  // statistics of the cluster engine's push phases: cycles, site visits, auxiliary steps, votes
		res[9] = flags[15];
		res[10] = flags[9];
		res[11] = flags[11];
		if (s_cnt > 0 || !converged) *stop = 1;
	}
}

int launch_alpha_expansion(pxb_ctx *ctx, const double *D_dev, int64_t N, int32_t L1, double lambda, double label_cost,
                           const int32_t *csr_off_host, const int32_t *csr_idx_host, const int32_t *init_labels_dev,
                           int32_t *labels_out_dev, double *energy_out_host) {
	// The data costs and the labelling are mirrored on the host: the labelling energies that decide whether a move is
	// kept are evaluated there in the reference's sequential summation order (compute_energy).
	const auto t_call = std::chrono::steady_clock::now();
	double ms_setup = 0, ms_cut = 0, ms_energy = 0, ms_push = 0;
	auto since = [](std::chrono::steady_clock::time_point t0) {
		return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	};
	// one pinned block for everything that crosses PCIe in this call: heights + flags and the results of a batch of moves
	// (per wait), the data costs and the initial labels (once; a pageable destination made the 480 KB copy a staged,
	// synchronous one)
	const size_t n_nodes_all = (size_t)N + (size_t)L1, res_ints_all = 2 + (size_t)kMoveRes * (size_t)L1;
	const size_t pin_ints = (n_nodes_all + 16 + res_ints_all + 1) & ~size_t(1), D_count = (size_t)N * (size_t)L1;
	PXB_TRY(ctx->reserve_pinned(sizeof(int32_t) * pin_ints + sizeof(double) * D_count + sizeof(int32_t) * (size_t)N));
	double *const D = reinterpret_cast<double *>(static_cast<int32_t *>(ctx->pinned) + pin_ints);
	int32_t *const lab_pinned = reinterpret_cast<int32_t *>(D + D_count);
	std::vector<int32_t> lab((size_t)N, 0);
	cudaStream_t st = ctx->stream;
	PXB_CUDA(cudaMemcpyAsync(D, D_dev, sizeof(double) * D_count, cudaMemcpyDeviceToHost, st));
	if (init_labels_dev)
		PXB_CUDA(cudaMemcpyAsync(lab_pinned, init_labels_dev, sizeof(int32_t) * (size_t)N, cudaMemcpyDeviceToHost, st));
	ExpSkeleton *skp = nullptr;
	PXB_TRY(exp_skeleton(ctx, N, csr_off_host, csr_idx_host, skp)); // overlaps the downloads on a cache hit
	ExpSkeleton &sk = *skp;
	PXB_TRY(ctx_wait(ctx));
	if (init_labels_dev) std::memcpy(lab.data(), lab_pinned, sizeof(int32_t) * (size_t)N);
	uint64_t memo_key = 0;
	if (ctx->label_memo && !getenv("PXB_NO_LABEL_MEMO")) {
		memo_key = fnv1a(D, sizeof(double) * D_count, sk.key ^ 0x5851F42D4C957F2Dull);
		memo_key = fnv1a(lab.data(), sizeof(int32_t) * lab.size(), memo_key);
		const double par[2] = {lambda, label_cost};
		const int64_t dims[2] = {N, L1};
		memo_key = fnv1a(par, sizeof(par), memo_key);
		memo_key = fnv1a(dims, sizeof(dims), memo_key) | 1;
		if (memo_key == sk.memo_key && (int64_t)sk.memo_labels.size() == N) {
			*energy_out_host = sk.memo_energy;
			PXB_CUDA(cudaMemcpyAsync(labels_out_dev, sk.memo_labels.data(), sizeof(int32_t) * (size_t)N, cudaMemcpyHostToDevice, st));
			PXB_TRY(ctx_wait(ctx));
			if (getenv("PXB_MF_STATS")) fprintf(stderr, "[pxb expansion] labelling: identical to the previous call (memo)\n");
			return PXB_OK;
		}
	}

	ExpansionProblem P;
	P.D = D;
	P.N = N;
	P.L1 = L1;
	P.lambda = lambda;
	P.label_cost = label_cost;
	P.goff = sk.goff.data();
	P.gidx = sk.gidx.data();
	const int64_t E = sk.E;
	const int n = (int)(N + L1);
	const int64_t m = E + 2 * N;

	// ---- device arena: per-call copy of the arcs (the auxiliary part is rewired with the labelling) + per-move state ----
	const size_t n_int = (size_t)(n + 1) + 2 * (size_t)m + 2 * (size_t)N + 2 * (size_t)(L1 + 1) + 2 * (size_t)n + 32 +
	                     (size_t)kMoveRes * (size_t)(L1 + 1) + 8;
	const size_t bytes = sizeof(double) * ((size_t)m + 2 * (size_t)n) + sizeof(int32_t) * n_int + 256;
	PXB_TRY(ctx->partials.reserve(bytes));
	double *d_cap = ctx->partials.as<double>(), *d_excess = d_cap + m, *d_sink = d_excess + n;
	int32_t *d_arc_off = reinterpret_cast<int32_t *>(d_sink + n), *d_head = d_arc_off + (n + 1), *d_rev = d_head + m;
	int32_t *d_lab = d_rev + m, *d_rank = d_lab + N;
	int32_t *d_label_off = d_rank + N, *d_label_count = d_label_off + (L1 + 1);
	// heights and kernel flags are adjacent: they come back in one copy after every move
	int32_t *d_h0 = d_label_count + (L1 + 1), *d_flags = d_h0 + n, *d_h1 = d_flags + 16;
	// results of a speculative batch of moves: [stop | pad | L1 x kMoveRes] (one copy back per batch, after the heights)
	int32_t *d_stop = d_h1 + n + 8, *d_res = d_stop + 2;
	const int32_t *d_goff = sk.d_goff, *d_gidx = sk.d_gidx;
	PXB_CUDA(cudaMemcpyAsync(d_arc_off, sk.arc_off, sizeof(int32_t) * (size_t)(N + 1), cudaMemcpyDeviceToDevice, st));
	PXB_CUDA(cudaMemcpyAsync(d_head, sk.head, sizeof(int32_t) * (size_t)(E + N), cudaMemcpyDeviceToDevice, st));
	PXB_CUDA(cudaMemcpyAsync(d_rev, sk.rev, sizeof(int32_t) * (size_t)(E + N), cudaMemcpyDeviceToDevice, st));
	std::vector<int32_t> rank((size_t)N), label_off((size_t)L1 + 1), label_count((size_t)L1 + 1);
	auto push_labelling = [&]() -> int { // labels, ranks and the auxiliary wiring follow the host labelling
		std::fill(label_count.begin(), label_count.end(), 0);
		for (int64_t i = 0; i < N; ++i) rank[i] = label_count[lab[i]]++;
		label_off[0] = 0;
		for (int l = 0; l < L1; ++l) label_off[l + 1] = label_off[l] + label_count[l];
		PXB_CUDA(cudaMemcpyAsync(d_lab, lab.data(), sizeof(int32_t) * (size_t)N, cudaMemcpyHostToDevice, st));
		PXB_CUDA(cudaMemcpyAsync(d_rank, rank.data(), sizeof(int32_t) * (size_t)N, cudaMemcpyHostToDevice, st));
		PXB_CUDA(cudaMemcpyAsync(d_label_off, label_off.data(), sizeof(int32_t) * (size_t)(L1 + 1), cudaMemcpyHostToDevice, st));
		PXB_CUDA(cudaMemcpyAsync(d_label_count, label_count.data(), sizeof(int32_t) * (size_t)(L1 + 1), cudaMemcpyHostToDevice, st));
		k_exp_wire_aux<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, L1, E, d_lab, d_rank, d_label_off, d_goff, d_arc_off, d_head, d_rev);
		ctx->launches++;
		// (no wait: the four sources are pageable vectors, and a host-to-device copy from pageable memory returns only after
		// the source has been copied into the driver's staging buffer -- they may be rewritten at once)
		return PXB_OK;
	};
	PXB_TRY(push_labelling());
	ms_setup = since(t_call);

	FlowGraphDev G;
	G.n = n;
	G.m = (int)m;
	G.wide_begin = (int)N;
	G.wide_count = L1;
	G.arc_off = d_arc_off;
	G.arc_head = d_head;
	G.arc_rev = d_rev;
	G.cap = d_cap;
	G.excess = d_excess;
	G.sink_cap = d_sink;
	G.height[0] = d_h0;
	G.height[1] = d_h1;
	G.flags = d_flags;
	MfLaunch lc;
	PXB_TRY(mf_launch_config(ctx, G, std::min(L1, ctx->sm_count), lc));
	const int grid = lc.grid;
	// graphs that fit run on the cluster-resident engine (pxb_maxflow_cluster.cu)
	if (sk.plan_key != mc_env_key() + (uint64_t)L1 * 0x9E3779B97F4A7C15ull) {
		PXB_TRY(mf_cluster_plan(ctx, (int)N, L1, sk.arc_off_host.data(), sk.plan));
		sk.plan_key = mc_env_key() + (uint64_t)L1 * 0x9E3779B97F4A7C15ull;
	}
	const bool use_cluster = sk.plan.ok;
	const size_t res_ints = res_ints_all; // (the pinned block was reserved at the top of the call)
	int32_t *h_host = static_cast<int32_t *>(ctx->pinned), *flags_host = h_host + n, *res_host = flags_host + 16 + 2;
	// Speculative batches (cluster-resident engine only). Most moves change nothing -- every labelling ends with a full
	// cycle of them -- but the host used to wait for each one to learn that. The moves of a cycle are enqueued back to back;
	// k_exp_close_move counts on the device how many sites a move would switch and, when any does, raises the stop flag that
	// turns the remaining launches of the batch into no-ops: the heights then still belong to that move, and the host
	// continues exactly as the one-move-per-wait loop would (candidate labelling, energies, accept or not, next label).
	// The batch length adapts: 1 after a move that switched something (the first sweep from the all-zero labelling),
	// the rest of the cycle after one that did not. PXB_EXP_SPEC=0 restores one wait per move.
	const bool speculate = use_cluster && !(getenv("PXB_EXP_SPEC") && atoi(getenv("PXB_EXP_SPEC")) == 0);
	int spec = init_labels_dev ? L1 : 1;

	EnergyCache ec;
	if (sk.lambda_table_of != lambda || sk.lambda_table.size() != (size_t)sk.goff[N] / 2 + 2) {
		EnergyCache::fill_table(P, sk.lambda_table);
		sk.lambda_table_of = lambda;
	}
	ec.init(P, lab, &sk.lambda_table, init_labels_dev == nullptr);
	double new_energy = ec.energy(P, lab, ec.pairs), old_energy; // always the energy of `lab`
	std::vector<int32_t> cand, switched;
	std::vector<int> batch;
	const bool stats = getenv("PXB_MF_STATS") != nullptr, check_energy = getenv("PXB_CHECK_ENERGY") != nullptr;
	for (int cycle = 1; cycle <= 1000; ++cycle) { // GCoptimization.cpp:1062-1077
		old_energy = new_energy;
		int alpha = 0;
		while (alpha < L1) { // oneExpansionIteration, fixed label order 0..L
			batch.clear();
			int a = alpha;
			for (; a < L1 && (int)batch.size() < (speculate ? spec : 1); ++a)
				if (label_count[a] != (int32_t)N) batch.push_back(a); // (no site to move: alpha_expansion returns at size == 0)
			if (batch.empty()) break;
			const auto t_move = std::chrono::steady_clock::now();
			if (speculate) PXB_CUDA(cudaMemsetAsync(d_stop, 0, sizeof(int32_t) * res_ints, st));
			for (int al : batch) {
				k_exp_assemble<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(N, L1, al, lambda, label_cost, D_dev, d_lab, d_goff, d_gidx,
				                                                          d_label_count, d_rev, d_cap, d_excess, d_sink, d_flags,
				                                                          speculate ? d_stop : nullptr);
				ctx->launches++;
				if (use_cluster) {
					G.stop = speculate ? d_stop : nullptr;
					PXB_TRY(mf_cluster_launch(ctx, G, sk.plan));
				} else {
					void *args[] = {&G};
					PXB_CUDA(cudaLaunchCooperativeKernel((void *)k_maxflow, dim3(grid), dim3(kMfThreads), args, lc.smem, st));
					ctx->launches++;
				}
				if (speculate) {
					k_exp_close_move<<<1, 1024, 0, st>>>(N, n, al, d_lab, d_h0, d_goff, d_gidx, d_flags, d_res + (size_t)kMoveRes * al, d_stop);
					ctx->launches++;
				}
			}
			PXB_CUDA(cudaMemcpyAsync(h_host, d_h0, sizeof(int32_t) * ((size_t)n + 16), cudaMemcpyDeviceToHost, st)); // + flags
			if (speculate) PXB_CUDA(cudaMemcpyAsync(res_host - 2, d_stop, sizeof(int32_t) * res_ints, cudaMemcpyDeviceToHost, st));
			PXB_TRY(ctx_wait(ctx));
			ms_cut += since(t_move);
			int next_alpha = a;
			bool any_switch = false;
			for (int al : batch) {
				const int32_t *mv = res_host + (size_t)kMoveRes * al; // executed, switched, flags 6 / 7 / 8 / 10 / 12
				if (speculate && !mv[0]) break; // (behind the move that stopped the batch)
				const int32_t rounds = speculate ? mv[2] : flags_host[6], status = speculate ? mv[3] : flags_host[7];
				if (status != 1 || rounds == 0) {
					set_error("max-flow did not converge within %d relabel rounds", kMaxRounds);
					return PXB_ERR_CUDA;
				}
				const auto t_en = std::chrono::steady_clock::now();
				if (stats && getenv("PXB_MF_STATS")[0] == '2')
					fprintf(stderr, "[pxb expansion] alpha=%d active=%d rounds=%d bfs_levels=%d relabel=%.2f ms async=%.2f ms (batch of %d; "
					                "%d push cycles: visits %.2f, auxiliary steps %.2f, votes %.2f ms)\n",
					        al, (int)(N - label_count[al]), rounds, speculate ? mv[4] : flags_host[8],
					        (double)(speculate ? mv[5] : flags_host[10]) * 64 / 1.965e6,
					        (double)(speculate ? mv[6] : flags_host[12]) * 64 / 1.965e6, (int)batch.size(), speculate ? mv[8] : flags_host[5],
					        (double)(speculate ? mv[9] : flags_host[15]) * 64 / 1.965e6, (double)(speculate ? mv[10] : flags_host[9]) * 64 / 1.965e6,
					        (double)(speculate ? mv[11] : flags_host[11]) * 64 / 1.965e6);
				if (speculate && mv[1] == 0) continue; // nothing switches
				// candidate labelling: SOURCE side (cannot reach the sink) takes alpha (:451-469)
				cand = lab;
				switched.clear();
				for (int64_t i = 0; i < N; ++i)
					if (lab[i] != al && !(h_host[i] < n)) {
						cand[i] = al;
						switched.push_back((int32_t)i);
					}
				if (switched.empty()) {
					ms_energy += since(t_en);
					continue;
				}
				any_switch = true;
				next_alpha = al + 1;
				// the reference applies the move iff afterExpansionEnergy < m_beforeExpansionEnergy (:1286); both are the
				// energies of the two labellings, evaluated here directly (in the reference's summation order)
				// (speculative batches: the pair count came back with the move; PXB_CHECK_ENERGY compares the energy below
				// with the full edge walk)
				const int64_t k_after = speculate ? ec.pairs + (int64_t)mv[7] : ec.pairs_after(P, lab, cand, switched);
				const double before = new_energy, after = ec.energy(P, cand, k_after);
				if (check_energy && after != compute_energy(P, cand)) { // PXB_CHECK_ENERGY=1: the full edge walk must agree bit for bit
					set_error("incremental labelling energy %.17g differs from the full evaluation %.17g", after, compute_energy(P, cand));
					return PXB_ERR_STATE;
				}
				ms_energy += since(t_en);
				if (after < before) {
					const auto t_p = std::chrono::steady_clock::now();
					lab.swap(cand);
					new_energy = after;
					ec.pairs = k_after;
					PXB_TRY(push_labelling());
					ms_push += since(t_p);
				}
				break; // (speculative batch: the launches behind this move did nothing; otherwise the batch had one move)
			}
			spec = any_switch ? 1 : L1;
			alpha = next_alpha;
		}
		if (new_energy == old_energy) break;
	}
	*energy_out_host = new_energy;
	PXB_CUDA(cudaMemcpyAsync(labels_out_dev, lab.data(), sizeof(int32_t) * (size_t)N, cudaMemcpyHostToDevice, st));
	PXB_TRY(ctx_wait(ctx));
	if (memo_key) {
		sk.memo_key = memo_key;
		sk.memo_labels = lab;
		sk.memo_energy = new_energy;
	}
	if (stats)
		fprintf(stderr, "[pxb expansion] labelling: N=%lld L1=%d total %.2f ms = setup %.2f + cuts %.2f + energies %.2f + rewiring %.2f\n",
		        (long long)N, L1, since(t_call), ms_setup, ms_cut, ms_energy, ms_push);
	return PXB_OK;
}

} // namespace pxb

using namespace pxb;

// gcr/GCRANSAC.h:964-1018
extern "C" int pxb_lo_graph_cut(pxb_ctx *ctx, const double *e0, const double *e1, const double *d, int64_t N,
                                double lambda, const int32_t *csr_off, const int32_t *csr_idx, uint8_t *inlier_out) {
	PXB_CHECK_ARG(ctx && e0 && e1 && d && inlier_out && N > 0, "null argument");
	PXB_CUDA(cudaSetDevice(ctx->device));
	FlowGraphHost g((int)N);
	for (int64_t i = 0; i < N; ++i) g.add_term1((int)i, e0[i], e1[i]);
	if (lambda > 0 && csr_off && csr_idx) {
		std::unordered_set<uint64_t> used; // the reference's N x N used_edges matrix (:964), same first-come semantics
		used.reserve((size_t)csr_off[N] * 2);
		const double e11 = 0;
		for (int64_t i = 0; i < N; ++i) {
			const double energy1 = d[i];
			for (int32_t e = csr_off[i]; e < csr_off[i + 1]; ++e) {
				const int64_t j = csr_idx[e];
				if (j == i || j < 0) continue;
				const uint64_t key = (uint64_t)std::min(i, j) * (uint64_t)N + (uint64_t)std::max(i, j);
				if (!used.insert(key).second) continue;
				const double energy2 = d[j];
				const double energy_sum = energy1 + energy2;
				const double e00 = 0.5 * energy_sum;
				g.add_term2((int)i, (int)j, e00 * lambda, lambda, lambda, e11 * lambda);
			}
		}
	}
	std::vector<uint8_t> seg;
	PXB_TRY(solve_min_cut(ctx, g, seg));
	std::memcpy(inlier_out, seg.data(), (size_t)N);
	return PXB_OK;
}

// GCRANSAC::labeling as a whole (gcr/GCRANSAC.h:914-1022): unary terms, pairwise graph and st-cut without leaving the
// device. Same result as pxb_lo_unary_terms + pxb_lo_graph_cut.
extern "C" int pxb_lo_labeling(pxb_ctx *ctx, const double *model_host, double thr, double lambda, const int32_t *csr_off,
                               const int32_t *csr_idx, uint8_t *inlier_out) {
	PXB_CHECK_ARG(ctx && model_host && csr_off && csr_idx && inlier_out, "null argument");
	if (ctx->pts.N <= 0) {
		set_error("no points uploaded");
		return PXB_ERR_STATE;
	}
	PXB_CUDA(cudaSetDevice(ctx->device));
	const int ms = model_size(ctx->pts.type);
	PXB_TRY(ctx->models.reserve(sizeof(double) * ms));
	PXB_CUDA(cudaMemcpyAsync(ctx->models.ptr, model_host, sizeof(double) * ms, cudaMemcpyHostToDevice, ctx->stream));
	return lo_labeling_device(ctx, ctx->models.as<double>(), thr, lambda, csr_off, csr_idx, inlier_out);
}
