// pxb_maxflow.h -- the residual graph both max-flow engines of libpxb200.so work on (not part of the ABI).
//   k_maxflow          (pxb_expansion.cu)        cooperative grid, float64 capacities in global memory: any size
//   k_maxflow_cluster  (pxb_maxflow_cluster.cu)  one thread-block cluster, the whole residual graph resident in the
//                                                cluster's (distributed) shared memory
#pragma once
#include <cstdint>

#include "pxb_internal.h"

namespace pxb {

struct FlowGraphDev {
	int n, m;
	int wide_begin, wide_count; // nodes [wide_begin, wide_begin + wide_count) are label-cost auxiliary nodes: thousands
	                            // of arcs each, handled by a whole thread block instead of one owner thread
	const int32_t *arc_off, *arc_head, *arc_rev;
	double *cap, *excess, *sink_cap;
	int32_t *height[2];
	int32_t *flags; // [0..2] BFS 'changed' (level mod 3), [3..5] 'active' (pulse mod 3), [6] pulses, [7] status
	const int32_t *stop = nullptr; // cluster engine only: when set and *stop != 0 the launch does nothing (a speculative
	                               // batch of expansion moves was cut short by an earlier move, pxb_expansion.cu)
	int async_cycles, idle_checks; // tuning knobs of the asynchronous phase (PXB_MF_ASYNC, PXB_MF_IDLE)
	int local_exit;                // PXB_MF_LOCAL_EXIT=1: blocks leave the phase on their own (A/B)
	long long quiet_cycles;        // PXB_MF_QUIET_US: grid-wide silence that ends the phase
	int debug;      // PXB_MF_STATS=3: block 0 prints the number of active nodes after every relabel
	int block_bfs;  // 1/2/3: the launch carries 2n / 3n+1 / 4n+1 ints of dynamic shared memory and block 0 runs the global
	                // relabel alone (2: CSR offsets in shared memory, 3: and bottom-up levels)
};

// Launch plan of the cluster-resident engine for one arc skeleton (computed once per skeleton on the host).
struct McPlan {
	bool ok = false;   // the graph fits: n < 65535 nodes, <= 1024 sites and < 65536 arc slots per CTA, shared memory
	int csize = 0;     // CTAs in the cluster (8 or 16)
	int sites_per_cta = 0;
	int arcs_per_cta = 0; // largest arc block of a CTA (sizes the shared arrays)
	int max_degree = 0;
	size_t smem = 0;
};

// arc_off_host: CSR offsets of the site nodes (n_sites + 1 entries, auxiliary arcs included in a site's block);
// n_aux: label-cost auxiliary nodes behind the sites (0 for the local-optimisation cut)
int mf_cluster_plan(pxb_ctx *ctx, int n_sites, int n_aux, const int32_t *arc_off_host, McPlan &plan);

// Enqueues the cut on ctx->stream. Results as k_maxflow leaves them: height[0][u] < n iff u reaches the sink, flags[6] =
// relabel rounds, flags[7] = 1.
int mf_cluster_launch(pxb_ctx *ctx, const FlowGraphDev &G, const McPlan &plan);

} // namespace pxb
