// pxb_maxflow_cluster.cu -- the min-cut engine for graphs that fit into ONE thread-block cluster (sm_100a).
//
//   k_maxflow_cluster   push-relabel with exact global relabelling, the whole residual graph resident in the
//                       distributed shared memory of a cluster of 8 or 16 CTAs (up to ~16k nodes / ~200k arcs: every
//                       PEARL expansion move and every GC-RANSAC local-optimisation cut of the 5k-10k point problems).
//
// Same role and same results as k_maxflow (pxb_expansion.cu: reference gcr/maxflow.cpp, Graph::maxflow, with BK's
// labelling rule gcr/graph.h:478-487 -- a node is SINK iff it still reaches the sink in the final residual graph).
// k_maxflow keeps the graph in global memory: a push visit is a chain of dependent L2 round trips (~370 cycles each),
// a BFS level of its single-block relabel ~3 us, and every phase ends at a grid barrier. Here
//   * sites are dealt to the CTAs in contiguous blocks (one site per thread); a CTA keeps the capacities, heads and
//     mirror slots of ITS sites' arcs, their excess / sink links and -- REPLICATED in every CTA -- the 16-bit heights of
//     all nodes: every load of a push visit or of a BFS level is a local shared-memory load (~38 cycles);
//   * what crosses CTAs travels one way only: a push adds to the mirror arc and to the excess of the head in the
//     owner's shared memory, a lifted or newly labelled node stores its height into all replicas; phases are separated
//     by barrier.cluster (~600 cycles with 1024-thread CTAs) instead of grid barriers;
//   * capacities stay FLOAT64, like the reference's. A 64-bit fixed-point version (remote integer adds are a single
//     fire-and-forget instruction, remote float64 adds compile to compare-and-swap loops) was written first and is
//     exact and deterministic -- but it is NOT the reference's arithmetic: PEARL's graphs are full of structural ties
//     between multiples of lambda (an outlier with three alpha neighbours carries fl(3 lambda) of excess against three
//     arcs of lambda each), which float64 subtraction chains break one way (0.9 - 0.3 - 0.3 - 0.3 = +5.6e-17: all arcs
//     saturate, the site leaves the sink side) and exact arithmetic the other way (fl(0.9) < 3 fl(0.3)): alpha-expansion
//     then walks to a different local minimum (tests/test_gpu_graphcut.py caught it at lambda = 0.3). So remote adds are
//     two compare-and-swap loops run side by side (~3 DSMEM round trips per push);
//   * the label-cost auxiliary nodes (thousands of arcs) have no arc list here: the capacities of the pair
//     (site <-> auxiliary node) live with the site, sites PULL from the auxiliary node's excess (one fetch-add per CTA
//     and cycle, handed out in site order by a block scan) and push back to it like to any other neighbour.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "pxb_maxflow.h"

namespace pxb {

constexpr int kMcThreads = 1024;
constexpr int kMcMaxAux = 16;
constexpr int kMcMaxRounds = 100000;
constexpr unsigned kInf = 0xFFFFu;

struct McParams {
	int n, n_sites, n_aux;
	int B, amax, npad, auxbase;
	const int32_t *arc_off, *arc_head, *arc_rev;
	const double *cap, *excess, *sink_cap;
	int32_t *height, *flags;
	const int32_t *stop;
	int max_cycles, check_every, debug, bfs_cap, cap_rounds, cap_cycles, aux_mask;
};

// ---- distributed shared memory primitives -------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
	unsigned r;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
	return r;
}
__device__ __forceinline__ void st_cluster_u16(unsigned addr, unsigned v) {
	asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(unsigned addr, unsigned v) {
	asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(unsigned addr, uint4 v) {
	asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ double ld_cluster_f64(unsigned addr) {
	double v;
	asm volatile("ld.volatile.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
	return v;
}
__device__ __forceinline__ unsigned long long cas_cluster(unsigned addr, unsigned long long expect, unsigned long long desired) {
	unsigned long long old;
	asm volatile("atom.relaxed.cluster.shared::cluster.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(addr), "l"(expect), "l"(desired) : "memory");
	return old;
}
// *a += d and *b += d in (possibly) other CTAs' shared memory: two compare-and-swap loops whose round trips overlap
__device__ __forceinline__ void cluster_add2(unsigned a, unsigned b, double d) {
	unsigned long long oa = (unsigned long long)__double_as_longlong(ld_cluster_f64(a));
	unsigned long long ob = (unsigned long long)__double_as_longlong(ld_cluster_f64(b));
	bool da = false, db = false;
	do {
		unsigned long long ra = oa, rb = ob;
		if (!da) ra = cas_cluster(a, oa, (unsigned long long)__double_as_longlong(__longlong_as_double((long long)oa) + d));
		if (!db) rb = cas_cluster(b, ob, (unsigned long long)__double_as_longlong(__longlong_as_double((long long)ob) + d));
		da |= ra == oa;
		db |= rb == ob;
		oa = ra;
		ob = rb;
	} while (!(da && db));
}
__device__ __forceinline__ void cluster_add1(unsigned a, double d) {
	unsigned long long oa = (unsigned long long)__double_as_longlong(ld_cluster_f64(a));
	for (;;) {
		const unsigned long long ra = cas_cluster(a, oa, (unsigned long long)__double_as_longlong(__longlong_as_double((long long)oa) + d));
		if (ra == oa) break;
		oa = ra;
	}
}
// takes up to `want` from *a (>= 0 always): returns what was taken
__device__ __forceinline__ double cluster_take(unsigned a, double want) {
	unsigned long long oa = (unsigned long long)__double_as_longlong(ld_cluster_f64(a));
	for (;;) {
		const double have = __longlong_as_double((long long)oa);
		const double g = fmin(fmax(have, 0.0), want);
		if (!(g > 0.0)) return 0.0;
		const unsigned long long ra = cas_cluster(a, oa, (unsigned long long)__double_as_longlong(have - g));
		if (ra == oa) return g;
		oa = ra;
	}
}
__device__ __forceinline__ void cluster_sync_all() {
	asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void local_add(double *p, double v) { atomicAdd(p, v); }

// Cluster-wide OR of a per-thread predicate: block vote, one 4-byte store per CTA into EVERY CTA's slot, cluster barrier.
// The slots rotate over three words: slot e % 3 is written before and read after the barrier of epoch e, and cleared
// (locally) before the barrier of epoch e - 1, when its last readers (epoch e - 3) are two barriers behind.
__device__ __forceinline__ bool cluster_or(bool pred, int &epoch, volatile int *slots, unsigned nranks) {
	const int cur = epoch % 3, nxt = (epoch + 1) % 3;
	const int any = __syncthreads_or(pred);
	if (any && threadIdx.x < nranks) st_cluster_u32(mapa(smem_addr((const void *)(slots + cur)), threadIdx.x), 1u);
	if (threadIdx.x == 0) slots[nxt] = 0;
	cluster_sync_all();
	++epoch;
	return slots[cur] != 0;
}

// Block-wide exclusive prefix sum (and total) of one value per thread, in thread order.
__device__ __forceinline__ void block_scan(double v, double &excl, double &total, double *s_warp) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	double incl = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const double t = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl += t;
	}
	if (lane == 31) s_warp[warp] = incl;
	__syncthreads();
	double before = 0, all = 0;
#pragma unroll 8
	for (int w = 0; w < kMcThreads / 32; ++w) {
		const double s = s_warp[w];
		if (w < warp) before += s;
		all += s;
	}
	excl = before + incl - v;
	total = all;
	__syncthreads(); // s_warp may be reused at once
}

// Lowest residual neighbour of a site (the first one in arc order among equals; best_a < 0: no residual arc at all).
// Chunks of 8 arcs: capacities and heads of a chunk are loaded together, then the heights of the heads -- two
// dependent shared-memory round trips per chunk instead of three per arc (a hub of the neighbourhood graph has 50+ arcs
// and everybody waits for it at the next barrier).
__device__ __forceinline__ void lowest_neighbour(const volatile double *vcap, const unsigned short *head, const volatile unsigned short *vh,
                                                 int a0, int a1, unsigned &best_h, int &best_a) {
	best_h = 0x7fffffffu;
	best_a = -1;
	for (int base = a0; base < a1; base += 8) {
		double c[8];
		unsigned hd[8], hv[8];
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const int a = min(base + j, a1 - 1);
			c[j] = vcap[a];
			hd[j] = head[a];
		}
#pragma unroll
		for (int j = 0; j < 8; ++j) hv[j] = vh[hd[j]];
#pragma unroll
		for (int j = 0; j < 8; ++j)
			if (base + j < a1 && c[j] > 0.0 && hv[j] < best_h) {
				best_h = hv[j];
				best_a = base + j;
			}
	}
}

// Has the site a residual out-arc into `level`?
__device__ __forceinline__ bool reaches_level(const volatile double *vcap, const unsigned short *head, const volatile unsigned short *vh,
                                              int a0, int a1, unsigned level) {
	for (int base = a0; base < a1; base += 8) {
		double c[8];
		unsigned hd[8], hv[8];
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const int a = min(base + j, a1 - 1);
			c[j] = vcap[a];
			hd[j] = head[a];
		}
#pragma unroll
		for (int j = 0; j < 8; ++j) hv[j] = vh[hd[j]];
		bool hit = false;
#pragma unroll
		for (int j = 0; j < 8; ++j) hit |= c[j] > 0.0 && hv[j] == level; // (a clamped duplicate repeats a real arc: harmless)
		if (hit) return true;
	}
	return false;
}

__global__ void __launch_bounds__(kMcThreads, 1) k_maxflow_cluster(McParams P) {
	extern __shared__ __align__(16) unsigned char mc_smem[];
	__shared__ int s_base[kMcMaxAux + 2];
	__shared__ int s_slots[3];
	__shared__ double s_warp[kMcThreads / 32];
	__shared__ double s_auxe[2][kMcMaxAux];
	__shared__ double s_grant;
	if (P.stop && *reinterpret_cast<const volatile int32_t *>(P.stop) != 0) return; // (written by an earlier kernel: uniform)
	unsigned rank, nranks;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
	asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(nranks));
	const int tid = threadIdx.x, t = tid;
	const int N = P.n_sites, naux = P.n_aux, n = P.n, B = P.B, amax = P.amax, auxbase = P.auxbase;
	// shared arrays (every block a multiple of 16 bytes)
	double *cap = reinterpret_cast<double *>(mc_smem);
	double *exc = cap + amax, *snk = exc + B, *cas = snk + B, *auxe = cas + B;
	unsigned short *head = reinterpret_cast<unsigned short *>(auxe + kMcMaxAux);
	unsigned short *rev = head + amax, *off = rev + amax, *lab = off + (B + 8), *hrep = lab + B;
	volatile double *vcap = cap, *vexc = exc;
	volatile unsigned short *vh = hrep;

	// ---- load: this CTA's block of the global graph -----------------------------------------------------------------
	const int s0 = (int)rank * B;
	const int nb = max(0, min(B, N - s0)); // sites of this CTA
	if (tid <= (int)nranks) s_base[tid] = P.arc_off[min(tid * B, N)];
	if (tid < 3) s_slots[tid] = 0;
	__syncthreads();
	const int abase = s_base[rank], na = s_base[rank + 1] - abase;
	for (int i = tid; i < na; i += kMcThreads) {
		const int v = P.arc_head[abase + i];
		cap[i] = P.cap[abase + i];
		if (v < N) {
			head[i] = (unsigned short)v;
			rev[i] = (unsigned short)(P.arc_rev[abase + i] - s_base[v / B]);
		} else {
			head[i] = (unsigned short)(auxbase + (v - N));
			rev[i] = 0;
		}
	}
	if (t < nb) {
		const int a1g = P.arc_off[s0 + t + 1];
		off[t] = (unsigned short)(P.arc_off[s0 + t] - abase);
		if (t == nb - 1) off[nb] = (unsigned short)(a1g - abase);
		exc[t] = P.excess[s0 + t];
		snk[t] = P.sink_cap[s0 + t];
		if (naux > 0) { // the site's last arc goes to the auxiliary node of its label; the mirror capacity lives here too
			lab[t] = (unsigned short)(P.arc_head[a1g - 1] - N);
			cas[t] = P.cap[P.arc_rev[a1g - 1]];
		} else {
			lab[t] = 0;
			cas[t] = 0.0;
		}
	}
	if (rank == 0 && tid < kMcMaxAux) auxe[tid] = tid < naux ? P.excess[N + tid] : 0.0;
	const int a0 = t < nb ? off[t] : 0; // (own write: visible to this thread)
	__syncthreads();
	const int a1 = t < nb ? off[t + 1] : 0;
	const unsigned my_h_addr = smem_addr(hrep + s0 + (t < B ? t : 0));
	// Only the CTAs that own a neighbour of this site ever read its height (arcs come in mirrored pairs: who has an arc
	// to the site is a head of one of the site's arcs): a height update goes to those replicas only -- 8.6 of 16 on
	// average for 12 arcs dealt at random, every one of them a DSMEM store transaction of its own.
	unsigned readers = 1u << rank;
	for (int a = a0; a < a1; ++a) {
		const unsigned v = head[a];
		if (v < (unsigned)auxbase) readers |= 1u << (v / (unsigned)B);
	}
	unsigned hmine = kInf;
	int epoch = 0, rounds = 0, levels_total = 0;
	long long clk_relabel = 0, clk_push = 0, acc_scan = 0, acc_or = 0, acc_aux = 0, acc_pvote = 0, acc_visit = 0;
	int push_cycles = 0;
	__shared__ unsigned long long s_scan_max, s_or_min, s_aux_max, s_pvote_min, s_visit_max;
	if (tid == 0) s_scan_max = 0ull, s_or_min = ~0ull, s_aux_max = 0ull, s_pvote_min = ~0ull, s_visit_max = 0ull;
	cluster_sync_all();

	for (; rounds < kMcMaxRounds; ++rounds) {
		// ---- exact distance-to-sink labels: level-synchronous BOTTOM-UP breadth-first search. Every unlabelled site
		// looks among its own out-arcs for a residual one into the current level (all local loads); a newly labelled
		// node stores its height into every replica; one cluster barrier per level.
		const long long c0 = clock64();
		hmine = (t < nb && snk[t] > 0.0) ? 1u : kInf;
		if (t < B) hrep[s0 + t] = (unsigned short)hmine; // this CTA's block of the replica (padding slots get kInf)
		if (tid < kMcMaxAux) hrep[auxbase + tid] = (unsigned short)kInf;
		__syncthreads();
		{ // this CTA's block of level-1 heights goes to every other replica in 16-byte pieces
			const int vecs = B / 8;
			const uint4 *src = reinterpret_cast<const uint4 *>(hrep + s0);
			const unsigned base_addr = smem_addr(hrep + s0);
			for (int i = tid; i < vecs * (int)nranks; i += kMcThreads) {
				const unsigned r = i / vecs, k = i - r * vecs;
				if (r != rank) st_cluster_v4(mapa(base_addr + 16u * k, r), src[k]);
			}
		}
		cluster_sync_all();
		unsigned level = 1;
		for (;; ++level) {
			if (rounds < P.cap_rounds && level > (unsigned)P.bfs_cap) {
				// Capped first relabel: whatever the search has not reached within bfs_cap levels gets bfs_cap + 2 -- a VALID
				// labelling (an unlabelled site has no residual arc into a level <= bfs_cap, or it would have been labelled),
				// just not an exact one. The round cannot end the cut (every later relabel is exact and complete).
				if (t < nb && hmine == kInf) {
					hmine = (unsigned)P.bfs_cap + 2u;
#pragma unroll
					for (unsigned r = 0; r < 16; ++r)
						if ((readers >> r) & 1u) st_cluster_u16(mapa(my_h_addr, r), hmine);
				}
				// (an auxiliary node is labelled by a frontier site with a residual arc from it: unlabelled ones follow the
				// same argument; every CTA updates its own replica, they all hold the same state)
				if (tid < naux && vh[auxbase + tid] == kInf) hrep[auxbase + tid] = (unsigned short)(P.bfs_cap + 2);
				cluster_sync_all();
				break;
			}
			bool found = false;
			const long long lv0 = clock64();
			if (t < nb) {
				if (hmine == kInf) {
					found = reaches_level(vcap, head, vh, a0, a1, level);
					if (found) {
						hmine = level + 1;
#pragma unroll
						for (unsigned r = 0; r < 16; ++r)
							if ((readers >> r) & 1u) st_cluster_u16(mapa(my_h_addr, r), hmine);
					}
				} else if (hmine == level && naux > 0 && cas[t] > 0.0 && vh[auxbase + lab[t]] == kInf) {
					// a frontier site with a residual arc FROM its auxiliary node labels that node
					const unsigned addr = smem_addr(hrep + auxbase + lab[t]);
#pragma unroll
					for (unsigned r = 0; r < 16; ++r)
						if (r < nranks) st_cluster_u16(mapa(addr, r), level + 1);
					found = true;
				}
			}
			const long long lv1 = clock64();
			const bool more = cluster_or(found, epoch, s_slots, nranks);
			acc_scan += lv1 - lv0;
			acc_or += clock64() - lv1;
			if (!more) break;
		}
		levels_total += (int)level;
		bool active = t < nb && hmine != kInf && vexc[t] > 0.0;
		if (rank == 0 && tid < naux) active |= auxe[tid] > 0.0 && vh[auxbase + tid] != kInf;
		const bool any_active = cluster_or(active, epoch, s_slots, nranks);
		clk_relabel += clock64() - c0;
		if (P.debug && rank == 0 && tid == 0) printf("[mc] round %d: levels=%u any_active=%d\n", rounds, level, (int)any_active);
		if (!any_active && !(rounds < P.cap_rounds)) break; // (a capped relabel cannot end the cut)

		// ---- push phase: every thread keeps discharging its site (Hong & He's lock-free rule: push to the lowest
		// residual neighbour if it is lower, else lift to one above it); a cluster-wide vote every check_every cycles
		const long long c1 = clock64();
		bool busy = false;
		const int cycles_now = rounds < P.cap_rounds ? P.cap_cycles : P.max_cycles; // (cluster-uniform)
		for (int cyc = 0; cyc < cycles_now; ++cyc) {
			const bool aux_cycle = naux > 0 && (cyc & P.aux_mask) == 0;
			const long long pv0 = clock64();
			++push_cycles;
			double aux_seen = 0.0;
			if (aux_cycle && tid < naux) aux_seen = ld_cluster_f64(mapa(smem_addr(auxe + tid), 0)); // in flight during the visit
			if (t < nb && hmine != kInf) {
				double e = vexc[t];
				if (e > 0.0) {
					busy = true;
					const double sc = snk[t];
					if (sc > 0.0) { // the sink is always the lowest neighbour; what it cannot absorb goes on below
						const double d = fmin(e, sc);
						snk[t] = sc - d;
						local_add(exc + t, -d);
						e -= d;
					}
					if (e > 0.0) {
						unsigned best_h;
						int best_a;
						lowest_neighbour(vcap, head, vh, a0, a1, best_h, best_a);
						if (best_a < 0 || best_h == kInf) { // nothing residual that still reaches the sink: stranded
							hmine = kInf;
#pragma unroll
							for (unsigned r = 0; r < 16; ++r)
								if ((readers >> r) & 1u) st_cluster_u16(mapa(my_h_addr, r), kInf);
						} else if (hmine > best_h) {
							const double d = fmin(e, vcap[best_a]);
							const int v = head[best_a];
							local_add(cap + best_a, -d);
							local_add(exc + t, -d);
							if (v >= auxbase) { // back to the auxiliary node: both capacities of the pair are this site's
								cas[t] += d;
								cluster_add1(mapa(smem_addr(auxe + (v - auxbase)), 0), d);
							} else {
								const unsigned rv = (unsigned)v / (unsigned)B;
								const int tv = v - (int)rv * B;
								cluster_add2(mapa(smem_addr(cap + rev[best_a]), rv), mapa(smem_addr(exc + tv), rv), d);
							}
						} else {
							unsigned nh = best_h + 1;
							if (nh >= (unsigned)n) nh = kInf;
							hmine = nh;
#pragma unroll
							for (unsigned r = 0; r < 16; ++r)
								if ((readers >> r) & 1u) st_cluster_u16(mapa(my_h_addr, r), nh);
						}
					}
				}
			}
			const long long pv1 = clock64();
			acc_visit += pv1 - pv0;
			if (aux_cycle) {
				double *seen = s_auxe[(cyc / (P.aux_mask + 1)) & 1];
				if (tid < naux) seen[tid] = aux_seen;
				__syncthreads();
				for (int l = 0; l < naux; ++l) {
					const unsigned ha = vh[auxbase + l];
					if (!(seen[l] > 0.0) || ha == kInf) continue; // block-uniform
					// Sites pull from the auxiliary node of their label. First only what the site's own sink link can
					// still absorb (those units leave the graph at once instead of trickling on through lambda-sized
					// n-links); when no site wants that, up to the arc capacities.
					const bool elig = t < nb && lab[t] == l && cas[t] > 0.0 && hmine < ha;
					double want = 0.0, excl, total;
					if (elig) want = fmin(cas[t], fmax(snk[t] - vexc[t], 0.0));
					block_scan(want, excl, total, s_warp);
					if (!(total > 0.0)) {
						want = elig ? cas[t] : 0.0;
						block_scan(want, excl, total, s_warp);
						if (!(total > 0.0)) continue;
					}
					if (tid == 0) s_grant = cluster_take(mapa(smem_addr(auxe + l), 0), total);
					__syncthreads();
					const double g = s_grant;
					const double give = fmin(want, fmax(g - excl, 0.0));
					if (give > 0.0) {
						cas[t] -= give;
						local_add(cap + (a1 - 1), give);
						local_add(exc + t, give);
					}
					busy |= g > 0.0;
					__syncthreads(); // s_grant is reused
				}
			}
			const long long pv2 = clock64();
			acc_aux += pv2 - pv1;
			if ((cyc + 1) % P.check_every == 0) {
				const bool more = cluster_or(busy, epoch, s_slots, nranks);
				acc_pvote += clock64() - pv2;
				if (!more) break;
				busy = false;
			}
		}
		clk_push += clock64() - c1;
	}
	// heights now hold the final reachability: height < n  <=>  the node can reach the sink in the residual graph
	if (t < nb) P.height[s0 + t] = hmine == kInf ? n : (int)hmine;
	if (rank == 0 && tid < naux) P.height[N + tid] = vh[auxbase + tid] == kInf ? n : (int)vh[auxbase + tid];
	atomicMax(&s_scan_max, (unsigned long long)acc_scan);
	atomicMin(&s_or_min, (unsigned long long)acc_or);
	atomicMax(&s_aux_max, (unsigned long long)acc_aux);
	atomicMin(&s_pvote_min, (unsigned long long)acc_pvote);
	atomicMax(&s_visit_max, (unsigned long long)acc_visit);
	__syncthreads();
	if (rank == 0 && tid == 0) {
		P.flags[13] = (int)(s_scan_max >> 6); // slowest thread's time in the level scans / fastest thread's time in the votes
		P.flags[14] = (int)(s_or_min >> 6);
		P.flags[9] = (int)(s_aux_max >> 6);    // push phases: slowest thread's time in the auxiliary-node steps,
		P.flags[11] = (int)(s_pvote_min >> 6); // fastest thread's time in the votes,
		P.flags[15] = (int)(s_visit_max >> 6); // slowest thread's time in the site visits
		P.flags[5] = push_cycles;
		P.flags[6] = rounds + 1;
		P.flags[7] = rounds < kMcMaxRounds ? 1 : 0;
		P.flags[8] = levels_total;
		P.flags[10] = (int)(clk_relabel >> 6);
		P.flags[12] = (int)(clk_push >> 6);
	}
	cluster_sync_all(); // no CTA may exit while others can still address its shared memory
}

// ---- host side ------------------------------------------------------------------------------------------------------
static size_t mc_smem_bytes(int B, int amax, int npad) {
	return sizeof(double) * ((size_t)amax + 3 * (size_t)B + kMcMaxAux) +
	       sizeof(unsigned short) * (2 * (size_t)amax + (size_t)(B + 8) + (size_t)B + (size_t)npad);
}

int mf_cluster_plan(pxb_ctx *ctx, int n_sites, int n_aux, const int32_t *arc_off_host, McPlan &plan) {
	plan = McPlan();
	if (const char *e = getenv("PXB_MF_CLUSTER"))
		if (atoi(e) == 0) return PXB_OK;
	if (n_sites < 1 || n_aux > kMcMaxAux || n_sites + n_aux >= 65000) return PXB_OK;
	size_t smem_cap = 200 * 1024;
	if (const char *e = getenv("PXB_MC_SMEM_KB")) smem_cap = std::min((size_t)220 * 1024, (size_t)std::max(0, atoi(e)) * 1024);
	const int forced = getenv("PXB_MC_CSIZE") ? atoi(getenv("PXB_MC_CSIZE")) : 0;
	static bool attribute_set[64] = {};
	if (ctx->device < 0 || ctx->device >= 64 || !attribute_set[ctx->device]) {
		PXB_CUDA(cudaFuncSetAttribute(k_maxflow_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
		PXB_CUDA(cudaFuncSetAttribute(k_maxflow_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
		if (ctx->device >= 0 && ctx->device < 64) attribute_set[ctx->device] = true;
	}
	for (int csize : {8, 16}) {
		if (forced && csize != forced) continue;
		const int B = ((n_sites + csize - 1) / csize + 7) / 8 * 8;
		if (B > kMcThreads) continue;
		int amax = 8, maxdeg = 0;
		for (int r = 0; r < csize; ++r) {
			const int lo = std::min(r * B, n_sites), hi = std::min((r + 1) * B, n_sites);
			amax = std::max(amax, arc_off_host[hi] - arc_off_host[lo]);
		}
		for (int s = 0; s < n_sites; ++s) maxdeg = std::max(maxdeg, arc_off_host[s + 1] - arc_off_host[s]);
		amax = (amax + 7) / 8 * 8;
		if (amax > 65528) continue;
		const int npad = csize * B + kMcMaxAux;
		const size_t smem = mc_smem_bytes(B, amax, npad);
		if (smem > smem_cap) continue;
		// can one such cluster be resident at all? (asked once per geometry and thread)
		thread_local struct {
			int device, csize;
			size_t smem;
			int clusters;
		} occ = {-1, 0, 0, 0};
		if (occ.device != ctx->device || occ.csize != csize || occ.smem != smem) {
			cudaLaunchConfig_t cfg = {};
			cfg.gridDim = dim3(csize);
			cfg.blockDim = dim3(kMcThreads);
			cfg.dynamicSmemBytes = smem;
			cudaLaunchAttribute at[1];
			at[0].id = cudaLaunchAttributeClusterDimension;
			at[0].val.clusterDim.x = csize;
			at[0].val.clusterDim.y = 1;
			at[0].val.clusterDim.z = 1;
			cfg.attrs = at;
			cfg.numAttrs = 1;
			int nc = 0;
			if (cudaOccupancyMaxActiveClusters(&nc, k_maxflow_cluster, &cfg) != cudaSuccess) {
				(void)cudaGetLastError();
				nc = 0;
			}
			occ = {ctx->device, csize, smem, nc};
		}
		if (occ.clusters < 1) continue;
		plan.ok = true;
		plan.csize = csize;
		plan.sites_per_cta = B;
		plan.arcs_per_cta = amax;
		plan.max_degree = maxdeg;
		plan.smem = smem;
		return PXB_OK;
	}
	return PXB_OK;
}

int mf_cluster_launch(pxb_ctx *ctx, const FlowGraphDev &G, const McPlan &plan) {
	if (!plan.ok) {
		set_error("cluster max-flow: no plan for this graph");
		return PXB_ERR_STATE;
	}
	McParams P;
	P.n = G.n;
	P.n_sites = G.wide_begin;
	P.n_aux = G.wide_count;
	P.B = plan.sites_per_cta;
	P.amax = plan.arcs_per_cta;
	P.auxbase = plan.csize * plan.sites_per_cta;
	P.npad = P.auxbase + kMcMaxAux;
	P.arc_off = G.arc_off;
	P.arc_head = G.arc_head;
	P.arc_rev = G.arc_rev;
	P.cap = G.cap;
	P.excess = G.excess;
	P.sink_cap = G.sink_cap;
	P.height = G.height[0];
	P.flags = G.flags;
	P.stop = G.stop;
	P.check_every = getenv("PXB_MC_CHECK") ? std::max(1, atoi(getenv("PXB_MC_CHECK"))) : 8;
	const int cycles = getenv("PXB_MC_CYCLES") ? std::max(1, atoi(getenv("PXB_MC_CYCLES"))) : 64;
	P.max_cycles = (cycles + P.check_every - 1) / P.check_every * P.check_every;
	P.debug = (getenv("PXB_MF_STATS") && getenv("PXB_MF_STATS")[0] == '3') ? 1 : 0;
	// Capped first relabel (see the kernel): on by default for the local-optimisation cut (4 levels: 59-103 BFS levels per cut
	// -> 6, the cut 0.20 -> 0.10 ms), off for expansion moves (their 5-9 rounds want exact labels: capping more than the
	// first one costs rounds, capping only the first one is noise). PXB_MC_BFS_CAP / PXB_MC_BFS_CAP_AUX / PXB_MC_CAP_ROUNDS.
	if (P.n_aux == 0)
		P.bfs_cap = getenv("PXB_MC_BFS_CAP") ? std::max(0, atoi(getenv("PXB_MC_BFS_CAP"))) : 4;
	else
		P.bfs_cap = getenv("PXB_MC_BFS_CAP_AUX") ? std::max(0, atoi(getenv("PXB_MC_BFS_CAP_AUX"))) : 0;
	P.aux_mask = 3; // auxiliary-node step every 4 push cycles (PXB_MC_AUX_EVERY: 1, 2, 4, 8)
	if (getenv("PXB_MC_AUX_EVERY")) {
		const int e = atoi(getenv("PXB_MC_AUX_EVERY"));
		if (e == 1 || e == 2 || e == 4 || e == 8) P.aux_mask = e - 1;
	}
	P.cap_rounds = P.bfs_cap > 0 ? 1 : 0;
	// a capped round only has to drain what sits within a few hops of the sink (8 cycles do on every scene measured: the
	// exact relabel behind it finds nothing active); whatever it leaves is picked up by an ordinary round
	P.cap_cycles = std::min(P.max_cycles, P.check_every);
	if (getenv("PXB_MC_CAP_CYCLES")) P.cap_cycles = (std::max(1, atoi(getenv("PXB_MC_CAP_CYCLES"))) + P.check_every - 1) / P.check_every * P.check_every;
	if (P.bfs_cap > 0 && P.n_aux > 0 && getenv("PXB_MC_CAP_ROUNDS")) P.cap_rounds = std::max(1, atoi(getenv("PXB_MC_CAP_ROUNDS")));
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(plan.csize);
	cfg.blockDim = dim3(kMcThreads);
	cfg.dynamicSmemBytes = plan.smem;
	cfg.stream = ctx->stream;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeClusterDimension;
	at[0].val.clusterDim.x = plan.csize;
	at[0].val.clusterDim.y = 1;
	at[0].val.clusterDim.z = 1;
	cfg.attrs = at;
	cfg.numAttrs = 1;
	PXB_CUDA(cudaLaunchKernelEx(&cfg, k_maxflow_cluster, P));
	ctx->launches++;
	return PXB_OK;
}

} // namespace pxb
