// pxb_internal.h -- shared internals of libpxb200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "pxb200.h"

namespace pxb {

struct NormDev;
void set_error(const char *fmt, ...);

#define PXB_CUDA(call)                                                                                  \
	do {                                                                                                \
		cudaError_t err__ = (call);                                                                     \
		if (err__ != cudaSuccess) {                                                                     \
			pxb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
			return PXB_ERR_CUDA;                                                                        \
		}                                                                                               \
	} while (0)

#define PXB_CHECK_ARG(cond, msg)                    \
	do {                                            \
		if (!(cond)) {                              \
			pxb::set_error("bad argument: %s", msg); \
			return PXB_ERR_ARGUMENT;                \
		}                                           \
	} while (0)

#define PXB_TRY(expr)              \
	do {                           \
		int rc__ = (expr);         \
		if (rc__ != PXB_OK) return rc__; \
	} while (0)

constexpr int kMaxDim = 5;

inline int point_dim(int t) { return t == PXB_MODEL_PNP ? 5 : (t == PXB_MODEL_LINE2D ? 2 : 4); }
inline int model_size(int t) { return t == PXB_MODEL_PNP ? 12 : (t >= PXB_MODEL_VANISHING_POINT ? 3 : 9); }
inline int sample_size(int t) {
	switch (t) {
	case PXB_MODEL_HOMOGRAPHY: return 4;
	case PXB_MODEL_FUNDAMENTAL: return 7;
	case PXB_MODEL_PNP: return 3;
	default: return 2;
	}
}
inline int max_solutions(int t) { return t == PXB_MODEL_FUNDAMENTAL ? 3 : (t == PXB_MODEL_PNP ? 4 : 1); }

// A growable device scratch buffer (never shrinks; freed with the context).
struct DevBuf {
	void *ptr = nullptr;
	size_t cap = 0;
	int reserve(size_t bytes);
	void release();
	template <class T> T *as() { return reinterpret_cast<T *>(ptr); }
};

// Points live on the device in SoA form: coordinate c of point i at soa[c * stride + i], stride a multiple of 64
// so that every coordinate row starts 512-byte aligned and warps read 256 contiguous bytes.
struct Points {
	int type = -1;
	int dim = 0;
	int64_t N = 0;
	int64_t stride = 0;
	double *soa = nullptr; // [dim][stride]
	double *aos = nullptr; // [N][dim] as uploaded (used by the solvers' gathers)
	// float32 screening copy (pxb_screen.cuh): normalised coordinates [dim][stride], per-point error scale [stride],
	// and the normalisation itself (device-resident, written by k_point_stats)
	float *f32n = nullptr;
	float *q = nullptr;
	struct NormDev *norm = nullptr;
};

} // namespace pxb

struct pxb_ctx {
	int device = 0;
	int sm_count = 0;
	cudaStream_t stream = nullptr;
	int64_t launches = 0;
	pxb::Points pts;
	// scratch
	pxb::DevBuf models, pref, pref2, outA, outB, outC, outD, idx, mask, partials, staging, screen, stats, cpref;
	void *pinned = nullptr;
	size_t pinned_cap = 0;
	void *lo_skeleton = nullptr; // pxb_expansion.cu: cached arc skeleton of the last neighbourhood graph
	// set by the host driver for the lifetime of its neighbourhood graph: the CSR arrays at these addresses are not
	// modified, so the skeleton caches may skip re-hashing them (external ABI callers always get the content hash)
	const int32_t *trusted_csr_off = nullptr, *trusted_csr_idx = nullptr;
	uint64_t trusted_csr_key = 0;
	void *exp_skeleton = nullptr; // pxb_expansion.cu: cached gco adjacency + static arcs of the alpha-expansion graph
	// set by the host driver around PEARL's labelling calls: an alpha-expansion whose data costs and initial labels are
	// bit-identical to the previous call of the same run hands back that call's result (the standalone operator
	// pxb_pearl_label always computes)
	bool label_memo = false;
	// Pinned staging arena of the host-pointer entry points: small H2D payloads are copied here first and small D2H
	// results land here and are handed to the caller's (pageable) buffers after the stream synchronises. Pageable
	// cudaMemcpyAsync calls are synchronous, take the driver's big lock and serialise concurrent contexts; pinned ones
	// are queued DMA descriptors.
	unsigned char *stage = nullptr;
	size_t stage_cap = 0, stage_used = 0;
	struct PendingCopy {
		void *dst;
		const void *src;
		size_t bytes;
	};
	std::vector<PendingCopy> pending;
	int reserve_pinned(size_t bytes);
	// hypothesis-block sharding (pxb_ctx_set_shard, pxb_nccl.cu): an ncclComm_t owned by the caller; when set, the task-level
	// find* entry points are collective over its ranks
	// cooperative waiting (pxb_batch.cu): when set, every wait on this context's stream polls cudaStreamQuery and calls
	// yield_fn between polls instead of blocking the host thread -- the batch driver runs several problems per host thread
	void (*yield_fn)(void *) = nullptr;
	void *yield_arg = nullptr;
	void *shard_comm = nullptr;
	int shard_world = 1, shard_rank = 0;
	pxb::DevBuf shard_msg, shard_rec;
	// ProgressiveX::getMutableSettings / getStatistics (pxb_ctx_set_settings, pxb_ctx_get_statistics)
	pxb_multi_model_settings engine_settings;
	bool has_engine_settings = false;
	pxb_multi_model_statistics last_statistics = {};
	cudaEvent_t timing_events[5 * PXB_MAX_ROUNDS + 2] = {}; // five marks per round + start / end of the run
	bool timing_events_ready = false;
	// Replayable device chains (pxb_driver.cu run_chain): the loops of the driver issue the same fixed sequence of copies
	// and kernels again and again; the second time a sequence with the same signature is seen it is captured into a CUDA
	// graph and from then on replayed with ONE launch. Inputs travel through a fixed pinned slot (chain_in -> chain_par on
	// the device), results come back into another (chain_out).
	struct ChainGraph {
		uint64_t key = 0;
		cudaGraphExec_t exec = nullptr;
		int state = 0; // 0: seen once (not captured yet), 1: captured, 2: cannot be captured
		int launches = 0;
		uint64_t last_use = 0;
	};
	std::vector<ChainGraph> chain_graphs;
	uint64_t chain_tick = 0;
	unsigned char *chain_in = nullptr, *chain_out = nullptr; // pinned, kChainInBytes / kChainOutBytes
	pxb::DevBuf chain_par;
	pxb::DevBuf labels, pack; // PEARL labels (kept on the device between iterations) and the packed per-call results
	pxb::DevBuf seg_scratch;  // warp partials + tickets of k_segment_sums (fixed size, allocated once)
	pxb::DevBuf pref_rows;    // preference vectors of the accepted instances, one row of N per instance (pxb_driver.cu)
};

namespace pxb {
// ---- programmatic dependent launch (sm_90+) ------------------------------------------------------------------------------
// The driver's chains are sequences of 4-15 us kernels; between two kernel nodes of a captured graph the GPU idles for
// the launch latency of the second one. A kernel launched with launch_pdl may be SCHEDULED while its predecessor in the
// stream still runs: it must call pdl_wait() before it touches anything a predecessor wrote (the wait returns when the
// preceding grid -- and, transitively, everything before it -- has completed and its writes are visible), and a small
// predecessor may call pdl_launch_dependents() at its top so that the successor's blocks are resident, parked at their
// wait, by the time it finishes. Both instructions do nothing in a kernel that was launched the ordinary way, and an edge
// whose upstream node is a copy or a memset is an ordinary dependency. PXB_PDL=0 launches everything the ordinary way.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
inline bool pdl_enabled() {
	static const bool on = !(getenv("PXB_PDL") && atoi(getenv("PXB_PDL")) == 0);
	return on;
}
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = grid;
	cfg.blockDim = block;
	cfg.dynamicSmemBytes = smem;
	cfg.stream = st;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
	cfg.attrs = at;
	cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
constexpr size_t kChainInBytes = size_t(256) << 10, kChainOutBytes = size_t(1) << 20;
void lo_skeleton_free(void *p);
void exp_skeleton_free(void *p);
uint64_t csr_content_key(pxb_ctx *ctx, int64_t N, const int32_t *off, const int32_t *idx);
// staged transfers of the host-pointer entry points (pxb_api.cu): small payloads go through the context's pinned arena;
// api_sync synchronises the stream and delivers the staged D2H results
int api_h2d(pxb_ctx *ctx, void *dst, const void *src, size_t bytes);
int api_d2h(pxb_ctx *ctx, void *dst, const void *src, size_t bytes);
int api_sync(pxb_ctx *ctx);
// waits for everything queued on ctx->stream (blocking, or yielding to the batch scheduler when the context has one)
int ctx_wait(pxb_ctx *ctx);
// NCCL exchange steps of the sharded driver (pxb_nccl.cu): device buffers, asynchronous on ctx->stream
int shard_broadcast(pxb_ctx *ctx, void *buf_dev, size_t bytes, int root);
int shard_allgather(pxb_ctx *ctx, void *recv_dev, size_t bytes_per_rank);
int pearl_label_device(pxb_ctx *ctx, const double *D_dev, int64_t N, int32_t L1, double lambda, double label_cost,
                       const int32_t *csr_off_host, const int32_t *csr_idx_host, const int32_t *init_labels_host,
                       int32_t *labels_out_host, double *energy_out);
// the same labelling with the labels staying on the device: init_labels_dev (may be null) and labels_out_dev are device
// arrays of N int32; the energy arrives either on the device (*energy_dev_out != null: greedy path, asynchronous) or in
// *energy_host (alpha-expansion: its host move loop has synchronised). n_dir < 0: count the smooth edges here.
int pearl_label_enqueue(pxb_ctx *ctx, const double *D_dev, int64_t N, int32_t L1, double lambda, double label_cost,
                        const int32_t *csr_off_host, const int32_t *csr_idx_host, int64_t n_dir, const int32_t *init_labels_dev,
                        int32_t *labels_out_dev, double *energy_host, double **energy_dev_out);
int launch_flag_compact(pxb_ctx *ctx, const uint8_t *flags_dev, int64_t N, int32_t *idx_dev, int64_t *count_dev, int32_t *off2_dev);
int launch_mask_compact(pxb_ctx *ctx, const uint32_t *mask_dev, int64_t N, int32_t *idx_dev, int64_t *count_dev, int32_t *off2_dev);
// seed_event_dev: two uint64 on the device (generator seed of the proposal's LO sampler, index of this LO labelling)
int launch_lo_sample(pxb_ctx *ctx, const int32_t *inl_dev, const int64_t *count_dev, int m, int limit, int trials,
                     const uint64_t *seed_event_dev, int32_t *off_dev, int32_t *idx_dev);
// GCRANSAC::labeling with the 0/1 result left on the device (*seg_dev_out, N bytes) and the max-flow status words in
// *flags_dev_out (16 int32; converged iff flags[7] == 1 && flags[6] != 0): asynchronous
int lo_labeling_enqueue(pxb_ctx *ctx, const double *model_dev, double thr, double lambda, const int32_t *csr_off_host,
                        const int32_t *csr_idx_host, uint8_t **seg_dev_out, int32_t **flags_dev_out);
bool lo_labeling_capturable(pxb_ctx *ctx, const int32_t *csr_off_host, const int32_t *csr_idx_host, uint64_t *signature);
int launch_label_lists(pxb_ctx *ctx, const int32_t *labels_dev, int64_t N, int L, int32_t *off_dev, int32_t *idx_dev);
int launch_select_models(pxb_ctx *ctx, const double *current, const double *fitted, const int32_t *ok, int L, int ms,
                         double *cand);
// kernel launchers (device pointers, asynchronous on ctx->stream)
int launch_residual_matrix(pxb_ctx *ctx, const double *models, int64_t K, double T2, double *r2, float *r2f,
                           uint32_t *mask);
// inlier bit matrix only, float32-screened (pxb_score.cu): bit-identical to the mask launch_residual_matrix writes
int launch_inlier_mask(pxb_ctx *ctx, const double *models, int64_t K, double T2, uint32_t *mask);
int launch_score_compound(pxb_ctx *ctx, const double *models, int64_t K, double T2, const double *compound_pref,
                          int64_t *count, double *value_sum, double *shared);
int launch_preference(pxb_ctx *ctx, const double *model, double T, double *pref);
int launch_tanimoto(pxb_ctx *ctx, const double *a, const double *b, int64_t N, double *out3 /*dot,na,nb*/);
int launch_compound_max(pxb_ctx *ctx, const double *prefs, int64_t L, int64_t N, double *out);
int launch_pearl_datacost(pxb_ctx *ctx, const double *models, int64_t L, double thr, double lambda, double *D);
int launch_segment_sums(pxb_ctx *ctx, const double *models, int64_t L, const int32_t *labels, double *sums,
                        int64_t *counts);
int launch_lo_unary(pxb_ctx *ctx, const double *model, double thr, double lambda, double *d, double *e0, double *e1);
int launch_lo_unary_cut(pxb_ctx *ctx, const double *model, double thr, double lambda, uint8_t *inlier);
int launch_tukey(pxb_ctx *ctx, const double *model, double T2, double *w);
int launch_solve_plane_parallax(pxb_ctx *ctx, const int64_t *samples, int64_t K, const double *H_dev, double *models_out,
                                int32_t *n_models, uint8_t *sample_valid, uint8_t *model_valid);
int launch_solve_minimal(pxb_ctx *ctx, const int64_t *samples, int64_t K, double *models_out, int32_t *n_models,
                         uint8_t *sample_valid, uint8_t *model_valid);
int launch_greedy_label(pxb_ctx *ctx, const double *D, int64_t N, int32_t L1, double label_cost,
                        const int32_t *init_labels, int32_t *labels_out, double *energy_out_dev);
int launch_alpha_expansion(pxb_ctx *ctx, const double *D_dev, int64_t N, int32_t L1, double lambda, double label_cost,
                           const int32_t *csr_off_host, const int32_t *csr_idx_host, const int32_t *init_labels_dev,
                           int32_t *labels_out_dev, double *energy_out_host);
} // namespace pxb
