// pxb_expansion.cu -- a11: the alpha-expansion label sweep of PEARL and the st-cut of the GC-RANSAC local
// optimisation, both on one GPU min-cut engine.
//
//   k_maxflow                 lock-free asynchronous push-relabel (Hong & He) on a CSR residual graph, alternating with
//                             exact backward-BFS global relabelling, one persistent cooperative kernel per cut (grid
//                             barriers only around the BFS levels, no host round trips).
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
// Round-1 split of work: the host builds the binary-energy graph of each move (integer bookkeeping plus two or three
// flops per edge, exactly the add_term1/add_term2 sequence of the reference) and evaluates labelling energies in the
// reference's summation order; the device solves the cuts. Moving graph construction to the device is listed as the
// next step in DESIGN.md.
#include <cooperative_groups.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_set>
#include <vector>

#include "pxb_internal.h"

namespace cg = cooperative_groups;

namespace pxb {

struct FlowGraphDev {
	int n, m;
	const int32_t *arc_off, *arc_head, *arc_rev;
	double *cap, *pushed, *excess, *sink_cap;
	int32_t *height[2];
	int32_t *flags; // [0..2] BFS 'changed' (level mod 3), [3..5] 'active' (pulse mod 3), [6] pulses, [7] status
};

constexpr int kMfThreads = 256;
constexpr int kWideDegree = 64;

__device__ void mf_global_relabel(const FlowGraphDev &G, int32_t *h, cg::grid_group &grid, int tid, int nthreads) {
	const int n = G.n;
	for (int u = tid; u < n; u += nthreads) h[u] = (G.sink_cap[u] > 0.0) ? 1 : n;
	if (tid == 0) {
		G.flags[0] = 0;
		G.flags[1] = 0;
		G.flags[2] = 0;
	}
	grid.sync();
	// flags rotate over three slots: the slot of level L+1 is cleared during level L, while stragglers may still be
	// reading the slot of level L-1 (they are past that level's barrier but not yet past its test)
	for (int level = 1; level < n; ++level) {
		int32_t *changed = &G.flags[level % 3];
		bool mine = false;
		for (int u = tid; u < n; u += nthreads) {
			if (h[u] != n) continue;
			for (int a = G.arc_off[u]; a < G.arc_off[u + 1]; ++a)
				if (G.cap[a] > 0.0 && h[G.arc_head[a]] == level) {
					h[u] = level + 1;
					mine = true;
					break;
				}
		}
		if (mine) *changed = 1;
		if (tid == 0) G.flags[(level + 1) % 3] = 0;
		grid.sync();
		if (*changed == 0) break;
	}
	grid.sync();
}

// One asynchronous push-relabel step of node u (Hong & He's lock-free rule: push to the LOWEST residual neighbour if it
// is lower, else lift to one above it). Only the owner thread of u lowers excess[u] / cap[out-arcs of u] and writes
// height[u]; everybody else only adds to them, so the atomics below can never drive a value negative.
__device__ __forceinline__ void mf_process(const FlowGraphDev &G, volatile int32_t *h, int u) {
	const int n = G.n;
	volatile double *excess = G.excess, *cap = G.cap;
	const double e = excess[u];
	const int hu = h[u];
	if (!(e > 0.0) || hu >= n) return;
	if (G.sink_cap[u] > 0.0) { // the sink (height 0) is always the lowest neighbour
		const double d = fmin(e, G.sink_cap[u]);
		G.sink_cap[u] -= d;
		atomicAdd(&G.excess[u], -d);
		return;
	}
	const int a0 = G.arc_off[u], a1 = G.arc_off[u + 1];
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
		return;
	}
	int best_h = 0x7fffffff, best_a = -1;
	for (int a = a0; a < a1; ++a)
		if (cap[a] > 0.0) {
			const int hv = h[G.arc_head[a]];
			if (hv < best_h) {
				best_h = hv;
				best_a = a;
			}
		}
	if (best_a < 0) { // no residual arc at all: the excess is stranded on the source side
		h[u] = n;
		return;
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
}

constexpr int kAsyncCycles = 192;
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
		mf_global_relabel(G, h, grid, tid, nthreads);
		bool active = false;
		for (int u = tid; u < n; u += nthreads) active |= (G.excess[u] > 0.0 && h[u] < n);
		if (tid == 0) G.flags[3 + ((round + 1) % 3)] = 0;
		if (active) G.flags[3 + (round % 3)] = 1;
		grid.sync();
		if (G.flags[3 + (round % 3)] == 0) break;
		// asynchronous phase: no barriers, every thread keeps discharging its own nodes
		for (int c = 0; c < kAsyncCycles; ++c)
			for (int u = tid; u < n; u += nthreads) mf_process(G, h, u);
		__threadfence();
		grid.sync();
	}
	// heights now hold the final reachability: height < n  <=>  the node can reach the sink in the residual graph
	if (tid == 0) {
		G.flags[6] = round + 1;
		G.flags[7] = round < kMaxRounds ? 1 : 0;
	}
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
static int solve_min_cut(pxb_ctx *ctx, const FlowGraphHost &g, std::vector<uint8_t> &segment) {
	const auto t_begin = std::chrono::steady_clock::now();
	const int n = g.n;
	const int pairs = (int)g.tail.size();
	const int m = 2 * pairs;
	segment.assign((size_t)n, 0);
	if (n == 0) return PXB_OK;
	// CSR by tail, arcs of a node in insertion order
	std::vector<int32_t> off((size_t)n + 1, 0), head((size_t)std::max(m, 1)), rev((size_t)std::max(m, 1));
	std::vector<double> cap((size_t)std::max(m, 1));
	for (int p = 0; p < pairs; ++p) {
		off[g.tail[p] + 1]++;
		off[g.head[p] + 1]++;
	}
	for (int i = 0; i < n; ++i) off[i + 1] += off[i];
	std::vector<int32_t> fill(off.begin(), off.end() - 1);
	for (int p = 0; p < pairs; ++p) {
		const int a = fill[g.tail[p]]++, b = fill[g.head[p]]++;
		head[a] = g.head[p];
		cap[a] = g.cap_fwd[p];
		rev[a] = b;
		head[b] = g.tail[p];
		cap[b] = g.cap_rev[p];
		rev[b] = a;
	}
	std::vector<double> excess((size_t)n), sink_cap((size_t)n);
	bool any_source = false, any_sink = false;
	for (int i = 0; i < n; ++i) {
		excess[i] = g.tr[i] > 0 ? g.tr[i] : 0.0;
		sink_cap[i] = g.tr[i] < 0 ? -g.tr[i] : 0.0;
		any_source |= excess[i] > 0;
		any_sink |= sink_cap[i] > 0;
	}
	if (!any_sink) return PXB_OK; // nothing can reach the sink: everything is SOURCE
	// device buffers (one arena)
	const size_t bytes_i = sizeof(int32_t) * ((size_t)n + 1 + 2 * (size_t)std::max(m, 1) + 2 * (size_t)n + 16);
	const size_t bytes_d = sizeof(double) * (2 * (size_t)std::max(m, 1) + 2 * (size_t)n);
	PXB_TRY(ctx->partials.reserve(bytes_d + bytes_i + 256));
	double *d_cap = ctx->partials.as<double>();
	double *d_pushed = d_cap + std::max(m, 1);
	double *d_excess = d_pushed + std::max(m, 1);
	double *d_sink = d_excess + n;
	int32_t *d_off = reinterpret_cast<int32_t *>(d_sink + n);
	int32_t *d_head = d_off + (n + 1);
	int32_t *d_rev = d_head + std::max(m, 1);
	int32_t *d_h0 = d_rev + std::max(m, 1);
	int32_t *d_h1 = d_h0 + n;
	int32_t *d_flags = d_h1 + n;
	cudaStream_t st = ctx->stream;
	PXB_CUDA(cudaMemcpyAsync(d_cap, cap.data(), sizeof(double) * cap.size(), cudaMemcpyHostToDevice, st));
	PXB_CUDA(cudaMemsetAsync(d_pushed, 0, sizeof(double) * (size_t)std::max(m, 1), st));
	PXB_CUDA(cudaMemcpyAsync(d_excess, excess.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
	PXB_CUDA(cudaMemcpyAsync(d_sink, sink_cap.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
	PXB_CUDA(cudaMemcpyAsync(d_off, off.data(), sizeof(int32_t) * off.size(), cudaMemcpyHostToDevice, st));
	PXB_CUDA(cudaMemcpyAsync(d_head, head.data(), sizeof(int32_t) * head.size(), cudaMemcpyHostToDevice, st));
	PXB_CUDA(cudaMemcpyAsync(d_rev, rev.data(), sizeof(int32_t) * rev.size(), cudaMemcpyHostToDevice, st));
	PXB_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int32_t) * 16, st));
	FlowGraphDev G;
	G.n = n;
	G.m = m;
	G.arc_off = d_off;
	G.arc_head = d_head;
	G.arc_rev = d_rev;
	G.cap = d_cap;
	G.pushed = d_pushed;
	G.excess = d_excess;
	G.sink_cap = d_sink;
	G.height[0] = d_h0;
	G.height[1] = d_h1;
	G.flags = d_flags;
	(void)any_source;
	int blocks_per_sm = 0;
	PXB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_maxflow, kMfThreads, 0));
	const int want = (n + kMfThreads - 1) / kMfThreads;
	const int grid = std::max(1, std::min(want, ctx->sm_count * std::max(1, std::min(blocks_per_sm, 4))));
	void *args[] = {&G};
	PXB_CUDA(cudaLaunchCooperativeKernel((void *)k_maxflow, dim3(grid), dim3(kMfThreads), args, 0, st));
	ctx->launches++;
	std::vector<int32_t> h((size_t)n);
	int32_t flags[16];
	PXB_CUDA(cudaMemcpyAsync(h.data(), d_h0, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
	PXB_CUDA(cudaMemcpyAsync(flags, d_flags, sizeof(flags), cudaMemcpyDeviceToHost, st));
	PXB_CUDA(cudaStreamSynchronize(st));
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
		if (++calls % 20 == 0 || ms > 50)
			fprintf(stderr, "[pxb maxflow] call %d: n=%d arcs=%d rounds=%d grid=%d  %.2f ms (total %.1f ms)\n", calls, n, m,
			        flags[6], grid, ms, total_ms);
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
	std::vector<int32_t> goff, gidx;
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
} // namespace

int launch_alpha_expansion(pxb_ctx *ctx, const double *D_dev, int64_t N, int32_t L1, double lambda, double label_cost,
                           const int32_t *csr_off_dev, const int32_t *csr_idx_dev, int64_t n_dir_edges,
                           const int32_t *init_labels_dev, int32_t *labels_out_dev, double *energy_out_host) {
	// The problem description comes back to the host for graph assembly (see the header comment).
	std::vector<double> D((size_t)N * L1);
	std::vector<int32_t> off((size_t)N + 1), idx((size_t)std::max<int64_t>(n_dir_edges, 1)), lab((size_t)N, 0);
	cudaStream_t st = ctx->stream;
	PXB_CUDA(cudaMemcpyAsync(D.data(), D_dev, sizeof(double) * D.size(), cudaMemcpyDeviceToHost, st));
	PXB_CUDA(cudaMemcpyAsync(off.data(), csr_off_dev, sizeof(int32_t) * off.size(), cudaMemcpyDeviceToHost, st));
	if (n_dir_edges > 0)
		PXB_CUDA(cudaMemcpyAsync(idx.data(), csr_idx_dev, sizeof(int32_t) * (size_t)n_dir_edges, cudaMemcpyDeviceToHost, st));
	if (init_labels_dev)
		PXB_CUDA(cudaMemcpyAsync(lab.data(), init_labels_dev, sizeof(int32_t) * (size_t)N, cudaMemcpyDeviceToHost, st));
	PXB_CUDA(cudaStreamSynchronize(st));

	ExpansionProblem P;
	P.D = D.data();
	P.N = N;
	P.L1 = L1;
	P.lambda = lambda;
	P.label_cost = label_cost;
	{ // gco adjacency: per site a list built with addFront, so the last inserted neighbour comes first
		std::vector<std::vector<int32_t>> lists((size_t)N);
		for (int64_t i = 0; i < N; ++i)
			for (int32_t e = off[i]; e < off[i + 1]; ++e) {
				const int32_t j = idx[e];
				if (j == i) continue; // PEARL.h:535
				lists[i].push_back(j);
				lists[j].push_back((int32_t)i);
			}
		P.goff.assign((size_t)N + 1, 0);
		for (int64_t i = 0; i < N; ++i) P.goff[i + 1] = P.goff[i] + (int32_t)lists[i].size();
		P.gidx.resize((size_t)P.goff[N]);
		for (int64_t i = 0; i < N; ++i)
			std::copy(lists[i].rbegin(), lists[i].rend(), P.gidx.begin() + P.goff[i]);
	}

	double new_energy = compute_energy(P, lab), old_energy;
	std::vector<int32_t> active, lookup((size_t)N, -1);
	std::vector<uint8_t> seg;
	for (int cycle = 1; cycle <= 1000; ++cycle) { // GCoptimization.cpp:1062-1077
		old_energy = new_energy;
		for (int alpha = 0; alpha < L1; ++alpha) { // oneExpansionIteration, fixed label order 0..L
			active.clear();
			for (int64_t i = 0; i < N; ++i)
				if (lab[i] != alpha) active.push_back((int32_t)i);
			const int size = (int)active.size();
			if (size == 0) continue;
			for (int v = 0; v < size; ++v) lookup[active[v]] = v;
			FlowGraphHost g(size);
			// setupDataCostsExpansion (:327-333): add_term1(i, D(site, alpha), D(site, current))
			for (int v = 0; v < size; ++v) {
				const int64_t s = active[v];
				g.add_term1(v, D[s * L1 + alpha], D[s * L1 + lab[s]]);
			}
			// setupSmoothCostsExpansion (:337-402), Potts * lambda, weight 1
			if (lambda > 0)
				for (int v = size - 1; v >= 0; --v) {
					const int64_t s = active[v];
					for (int32_t e = P.goff[s]; e < P.goff[s + 1]; ++e) {
						const int32_t nb = P.gidx[e];
						if (lookup[nb] == -1) { // neighbour keeps alpha
							const double e0 = (alpha != lab[nb]) ? lambda : 0, e1 = (lab[s] != lab[nb]) ? lambda : 0;
							g.add_term1(v, e0 * 1.0, e1 * 1.0);
						} else if (nb < s) {
							const double e00 = 0, e01 = (alpha != lab[nb]) ? lambda : 0, e10 = (lab[s] != alpha) ? lambda : 0,
							             e11 = (lab[s] != lab[nb]) ? lambda : 0;
							g.add_term2(v, lookup[nb], e00 * 1.0, e01 * 1.0, e10 * 1.0, e11 * 1.0);
						}
					}
				}
			// setupLabelCostsExpansion (:1131-1195): one auxiliary node per non-alpha label present among the active sites
			if (label_cost > 0) {
				std::vector<int> aux((size_t)L1, -1);
				for (int v = 0; v < size; ++v) {
					const int l = lab[active[v]];
					if (aux[l] < 0) {
						aux[l] = g.add_node();
						g.add_term1(aux[l], 0, label_cost);
					}
					g.add_term2(v, aux[l], 0, 0, label_cost, 0);
				}
			}
			// An auxiliary arc aux -> site can never carry more than the site can pass on (its own sink link plus its
			// outgoing n-links). Clamping it to that bound leaves every maximum flow -- and the set of nodes that can
			// reach the sink -- unchanged, but lets the auxiliary node spread its budget in one discharge.
			if (label_cost > 0 && g.n > size) {
				std::vector<double> out_cap((size_t)g.n, 0.0);
				for (size_t pidx = 0; pidx < g.tail.size(); ++pidx) {
					out_cap[g.tail[pidx]] += g.cap_fwd[pidx];
					out_cap[g.head[pidx]] += g.cap_rev[pidx];
				}
				for (size_t pidx = 0; pidx < g.tail.size(); ++pidx)
					if (g.head[pidx] >= size) { // site -> aux pair: cap_fwd = 0, cap_rev = label cost (aux -> site)
						const int v = g.tail[pidx];
						const double bound = (g.tr[v] < 0 ? -g.tr[v] : 0.0) + out_cap[v];
						g.cap_rev[pidx] = std::min(g.cap_rev[pidx], bound);
					}
			}
			PXB_TRY(solve_min_cut(ctx, g, seg));
			// candidate labelling: SOURCE side (get_var == 0) takes alpha (:451-469)
			bool any_switch = false;
			std::vector<int32_t> cand = lab;
			for (int v = 0; v < size; ++v)
				if (!seg[v]) {
					cand[active[v]] = alpha;
					any_switch = true;
				}
			for (int v = 0; v < size; ++v) lookup[active[v]] = -1;
			if (!any_switch) continue;
			// the reference applies the move iff afterExpansionEnergy < m_beforeExpansionEnergy (:1286); both are the
			// energies of the two labellings, evaluated here directly
			const double before = compute_energy(P, lab), after = compute_energy(P, cand);
			if (after < before) lab.swap(cand);
		}
		new_energy = compute_energy(P, lab);
		if (new_energy == old_energy) break;
	}
	*energy_out_host = new_energy;
	PXB_CUDA(cudaMemcpyAsync(labels_out_dev, lab.data(), sizeof(int32_t) * (size_t)N, cudaMemcpyHostToDevice, st));
	PXB_CUDA(cudaStreamSynchronize(st));
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
