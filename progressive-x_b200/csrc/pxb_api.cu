// pxb_api.cu -- the extern "C" boundary of libpxb200.so: context, memory helpers and the host-pointer wrappers
// around the kernel launchers. See include/pxb200.h for the contract of every entry point.
#include <algorithm>
#include <cstring>

#include "pxb_internal.h"

namespace pxb {

static thread_local std::string g_last_error;

void set_error(const char *fmt, ...) {
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_last_error = buf;
}

int DevBuf::reserve(size_t bytes) {
	if (bytes <= cap) return PXB_OK;
	if (ptr) PXB_CUDA(cudaFree(ptr));
	ptr = nullptr;
	cap = 0;
	size_t want = std::max<size_t>(bytes, 256);
	want = (want + 255) & ~size_t(255);
	PXB_CUDA(cudaMalloc(&ptr, want));
	cap = want;
	return PXB_OK;
}
void DevBuf::release() {
	if (ptr) cudaFree(ptr);
	ptr = nullptr;
	cap = 0;
}

int launch_aos_to_soa(pxb_ctx *ctx);

// drops whatever a previous, failed call may have left in the staging arena
static void begin_call(pxb_ctx *ctx) {
	ctx->pending.clear();
	ctx->stage_used = 0;
}

static int require_points(pxb_ctx *ctx) {
	if (!ctx) {
		set_error("null context");
		return PXB_ERR_ARGUMENT;
	}
	begin_call(ctx);
	if (ctx->pts.N <= 0 || !ctx->pts.soa) {
		set_error("no points uploaded: call pxb_upload_points first");
		return PXB_ERR_STATE;
	}
	PXB_CUDA(cudaSetDevice(ctx->device)); // entry points may be called from any host thread
	return PXB_OK;
}

constexpr size_t kStageBytes = size_t(4) << 20, kStageMaxItem = size_t(256) << 10; // larger payloads go direct

// page-locked (cudaMallocHost / cudaHostRegister) caller memory is DMA-able as it is: no staging copy for it
static bool caller_memory_is_pinned(const void *p) {
	cudaPointerAttributes attr;
	if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
		(void)cudaGetLastError();
		return false;
	}
	return attr.type == cudaMemoryTypeHost;
}

static unsigned char *stage_take(pxb_ctx *ctx, size_t bytes, const void *caller_ptr) {
	if (!ctx->stage || bytes > kStageMaxItem) return nullptr;
	if (bytes >= (size_t(16) << 10) && caller_memory_is_pinned(caller_ptr)) return nullptr;
	const size_t at = (ctx->stage_used + 255) & ~size_t(255);
	if (at + bytes > ctx->stage_cap) return nullptr;
	ctx->stage_used = at + bytes;
	return ctx->stage + at;
}

int api_h2d(pxb_ctx *ctx, void *dst, const void *src, size_t bytes) {
	if (bytes == 0) return PXB_OK;
	if (unsigned char *st = stage_take(ctx, bytes, src)) {
		std::memcpy(st, src, bytes);
		src = st;
	}
	PXB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
	return PXB_OK;
}
int api_d2h(pxb_ctx *ctx, void *dst, const void *src, size_t bytes) {
	if (bytes == 0) return PXB_OK;
	if (unsigned char *st = stage_take(ctx, bytes, dst)) {
		PXB_CUDA(cudaMemcpyAsync(st, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
		ctx->pending.push_back({dst, st, bytes}); // delivered by sync()
		return PXB_OK;
	}
	PXB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	return PXB_OK;
}
int ctx_wait(pxb_ctx *ctx) {
	if (ctx->yield_fn) {
		for (;;) {
			const cudaError_t q = cudaStreamQuery(ctx->stream);
			if (q == cudaSuccess) return PXB_OK;
			if (q != cudaErrorNotReady) {
				set_error("cudaStreamQuery failed: %s", cudaGetErrorString(q));
				return PXB_ERR_CUDA;
			}
			ctx->yield_fn(ctx->yield_arg);
		}
	}
	PXB_CUDA(cudaStreamSynchronize(ctx->stream));
	return PXB_OK;
}

int api_sync(pxb_ctx *ctx) {
	cudaError_t err = cudaSuccess;
	if (ctx_wait(ctx) != PXB_OK) err = cudaErrorUnknown;
	if (err == cudaSuccess)
		for (const auto &c : ctx->pending) std::memcpy(c.dst, c.src, c.bytes);
	ctx->pending.clear();
	ctx->stage_used = 0;
	return err == cudaSuccess ? PXB_OK : PXB_ERR_CUDA;
}

// One PEARL::labeling call on a data-cost matrix that already sits on the device (D_dev, N x L1), labels in and out on
// the device. Shared by pxb_pearl_label and the host driver (which keeps matrix and labels on the device).
int pearl_label_enqueue(pxb_ctx *ctx, const double *D_dev, int64_t N, int32_t L1, double lambda, double label_cost,
                        const int32_t *csr_off_host, const int32_t *csr_idx_host, int64_t n_dir, const int32_t *init_labels_dev,
                        int32_t *labels_out_dev, double *energy_host, double **energy_dev_out) {
	*energy_dev_out = nullptr;
	// Count the undirected edges setNeighbors would insert (self loops are skipped, PEARL.h:535).
	if (n_dir < 0) {
		n_dir = 0;
		if (lambda > 0.0 && csr_off_host && csr_idx_host)
			for (int64_t i = 0; i < N; ++i)
				for (int32_t e = csr_off_host[i]; e < csr_off_host[i + 1]; ++e)
					if (csr_idx_host[e] != i) ++n_dir;
	}
	if (n_dir == 0) {
		// solveSpecialCases: data costs + per-label costs, no smooth term -> solveGreedy (GCoptimization.cpp:542-552).
		// With label_cost == 0 the reference takes the per-site argmin branch (:499-517); the greedy solver with a
		// zero label cost is not the same algorithm, so that case is handled explicitly.
		if (!(label_cost > 0.0)) {
			set_error("pxb_pearl_label: label_cost must be > 0 (PEARL always sets it to minimum_inlier_number)");
			return PXB_ERR_UNSUPPORTED;
		}
		PXB_TRY(ctx->outB.reserve(sizeof(double)));
		PXB_TRY(launch_greedy_label(ctx, D_dev, N, L1, label_cost, init_labels_dev, labels_out_dev, ctx->outB.as<double>()));
		*energy_dev_out = ctx->outB.as<double>();
		return PXB_OK;
	}
	return launch_alpha_expansion(ctx, D_dev, N, L1, lambda, label_cost, csr_off_host, csr_idx_host, init_labels_dev,
	                              labels_out_dev, energy_host);
}

int pearl_label_device(pxb_ctx *ctx, const double *D_dev, int64_t N, int32_t L1, double lambda, double label_cost,
                       const int32_t *csr_off_host, const int32_t *csr_idx_host, const int32_t *init_labels_host,
                       int32_t *labels_out_host, double *energy_out) {
	// GCoptimization::setLabel range-checks labels (GCoptimization.cpp:929-934 throws GCException)
	if (init_labels_host)
		for (int64_t i = 0; i < N; ++i)
			if (init_labels_host[i] < 0 || init_labels_host[i] >= L1) {
				set_error("init label %d of site %lld outside [0, %d)", init_labels_host[i], (long long)i, L1);
				return PXB_ERR_ARGUMENT;
			}
	PXB_TRY(ctx->labels.reserve(sizeof(int32_t) * (size_t)N * 2 + 64));
	int32_t *lab_out = ctx->labels.as<int32_t>();
	int32_t *lab_in = nullptr;
	if (init_labels_host) {
		lab_in = lab_out + N;
		PXB_TRY(api_h2d(ctx, lab_in, init_labels_host, sizeof(int32_t) * (size_t)N));
	}
	double *energy_dev = nullptr;
	PXB_TRY(pearl_label_enqueue(ctx, D_dev, N, L1, lambda, label_cost, csr_off_host, csr_idx_host, -1, lab_in, lab_out, energy_out,
	                            &energy_dev));
	PXB_TRY(api_d2h(ctx, labels_out_host, lab_out, sizeof(int32_t) * (size_t)N));
	if (energy_dev) PXB_TRY(api_d2h(ctx, energy_out, energy_dev, sizeof(double)));
	return api_sync(ctx);
}

static int h2d(pxb_ctx *ctx, void *dst, const void *src, size_t bytes) { return api_h2d(ctx, dst, src, bytes); }
static int d2h(pxb_ctx *ctx, void *dst, const void *src, size_t bytes) { return api_d2h(ctx, dst, src, bytes); }
static int sync(pxb_ctx *ctx) { return api_sync(ctx); }

} // namespace pxb

using namespace pxb;

int pxb_ctx::reserve_pinned(size_t bytes) {
	if (bytes <= pinned_cap) return PXB_OK;
	if (pinned) cudaFreeHost(pinned);
	pinned = nullptr;
	pinned_cap = 0;
	PXB_CUDA(cudaMallocHost(&pinned, bytes));
	pinned_cap = bytes;
	return PXB_OK;
}

extern "C" {

const char *pxb_last_error(void) { return g_last_error.c_str(); }
const char *pxb_version(void) { return "pxb200 0.1 (sm_100a)"; }

int pxb_ctx_create(int device, pxb_ctx **out) {
	PXB_CHECK_ARG(out != nullptr, "out is null");
	*out = nullptr;
	int count = 0;
	cudaError_t err = cudaGetDeviceCount(&count);
	if (err != cudaSuccess || count == 0) {
		set_error("no CUDA device available (%s); libpxb200 has no CPU fallback",
		          err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0");
		return PXB_ERR_NO_DEVICE;
	}
	if (device < 0 || device >= count) {
		set_error("device %d out of range (0..%d)", device, count - 1);
		return PXB_ERR_ARGUMENT;
	}
	cudaDeviceProp prop;
	PXB_CUDA(cudaGetDeviceProperties(&prop, device));
	if (prop.major != 10) {
		set_error("device %d is sm_%d%d; libpxb200 is built for sm_100a only", device, prop.major, prop.minor);
		return PXB_ERR_NO_DEVICE;
	}
	PXB_CUDA(cudaSetDevice(device));
	pxb_ctx *ctx = new pxb_ctx();
	ctx->device = device;
	ctx->sm_count = prop.multiProcessorCount;
	err = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
	if (err != cudaSuccess) {
		set_error("cudaStreamCreate failed: %s", cudaGetErrorString(err));
		delete ctx;
		return PXB_ERR_CUDA;
	}
	if (cudaMallocHost(reinterpret_cast<void **>(&ctx->stage), kStageBytes) == cudaSuccess)
		ctx->stage_cap = kStageBytes;
	else
		(void)cudaGetLastError(); // no staging arena: transfers fall back to direct copies
	// fixed pinned slots of the replayable device chains (without them the driver simply never captures a graph)
	if (cudaMallocHost(reinterpret_cast<void **>(&ctx->chain_in), kChainInBytes) != cudaSuccess ||
	    cudaMallocHost(reinterpret_cast<void **>(&ctx->chain_out), kChainOutBytes) != cudaSuccess ||
	    ctx->chain_par.reserve(kChainInBytes) != PXB_OK) {
		(void)cudaGetLastError();
		if (ctx->chain_in) cudaFreeHost(ctx->chain_in);
		if (ctx->chain_out) cudaFreeHost(ctx->chain_out);
		ctx->chain_in = ctx->chain_out = nullptr;
	}
	*out = ctx;
	return PXB_OK;
}

void pxb_ctx_destroy(pxb_ctx *ctx) {
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	if (ctx->pts.soa) cudaFree(ctx->pts.soa);
	if (ctx->pts.aos) cudaFree(ctx->pts.aos);
	if (ctx->pts.f32n) cudaFree(ctx->pts.f32n);
	if (ctx->pts.q) cudaFree(ctx->pts.q);
	if (ctx->pts.norm) cudaFree(ctx->pts.norm);
	DevBuf *bufs[] = {&ctx->models, &ctx->pref, &ctx->pref2, &ctx->outA, &ctx->outB, &ctx->outC,
	                  &ctx->outD,   &ctx->idx,  &ctx->mask,  &ctx->partials, &ctx->staging, &ctx->screen, &ctx->stats, &ctx->cpref,
	                  &ctx->shard_msg, &ctx->shard_rec, &ctx->labels, &ctx->pack, &ctx->chain_par, &ctx->pref_rows, &ctx->seg_scratch};
	for (DevBuf *b : bufs) b->release();
	if (ctx->pinned) cudaFreeHost(ctx->pinned);
	if (ctx->stage) cudaFreeHost(ctx->stage);
	for (auto &g : ctx->chain_graphs)
		if (g.exec) cudaGraphExecDestroy(g.exec);
	if (ctx->chain_in) cudaFreeHost(ctx->chain_in);
	if (ctx->chain_out) cudaFreeHost(ctx->chain_out);
	if (ctx->timing_events_ready)
		for (cudaEvent_t e : ctx->timing_events) cudaEventDestroy(e);
	lo_skeleton_free(ctx->lo_skeleton);
	exp_skeleton_free(ctx->exp_skeleton);
	cudaStreamDestroy(ctx->stream);
	delete ctx;
}

void *pxb_ctx_stream(pxb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int pxb_settings_default(pxb_multi_model_settings *out) { // MultiModelSettings() (progressive_x.h:60-75), gcr/settings.h:66-86
	PXB_CHECK_ARG(out != nullptr, "null argument");
	out->minimum_number_of_inliers = 20;
	out->max_proposal_number_without_change = 10;
	out->cell_number_in_neighborhood_graph = 8;
	out->maximum_model_number = SIZE_MAX;
	out->maximum_tanimoto_similarity = 0.5;
	out->confidence = 0.95;
	out->inlier_outlier_threshold = 2.0;
	out->spatial_coherence_weight = 0.14;
	out->max_iteration_number = 5000;
	out->min_iteration_number = 20;
	out->min_iteration_number_before_lo = 20;
	out->max_local_optimization_number = 50;
	out->max_graph_cut_number = 10;
	out->max_least_squares_iterations = 10;
	out->max_unsuccessful_model_generations = 100;
	out->scoring_exponent = 2;
	return PXB_OK;
}

int pxb_ctx_set_settings(pxb_ctx *ctx, const pxb_multi_model_settings *settings) {
	PXB_CHECK_ARG(ctx != nullptr, "null context");
	ctx->has_engine_settings = settings != nullptr;
	if (settings) {
		PXB_CHECK_ARG(settings->max_local_optimization_number >= 1 && settings->max_local_optimization_number <= 512,
		              "max_local_optimization_number in [1, 512]");
		ctx->engine_settings = *settings;
	}
	return PXB_OK;
}

int pxb_ctx_get_statistics(pxb_ctx *ctx, pxb_multi_model_statistics *out) {
	PXB_CHECK_ARG(ctx && out, "null argument");
	*out = ctx->last_statistics;
	return PXB_OK;
}

int pxb_sync(pxb_ctx *ctx) {
	PXB_CHECK_ARG(ctx != nullptr, "null context");
	PXB_CUDA(cudaSetDevice(ctx->device));
	return sync(ctx);
}

int64_t pxb_launch_count(pxb_ctx *ctx) { return ctx ? ctx->launches : 0; }

int pxb_dev_alloc(pxb_ctx *ctx, size_t bytes, void **dev_ptr) {
	PXB_CHECK_ARG(ctx && dev_ptr, "null argument");
	PXB_CUDA(cudaSetDevice(ctx->device));
	PXB_CUDA(cudaMalloc(dev_ptr, std::max<size_t>(bytes, 1)));
	return PXB_OK;
}
int pxb_dev_free(pxb_ctx *ctx, void *dev_ptr) {
	PXB_CHECK_ARG(ctx != nullptr, "null context");
	if (dev_ptr) PXB_CUDA(cudaFree(dev_ptr));
	return PXB_OK;
}
int pxb_memcpy_h2d(pxb_ctx *ctx, void *dev_dst, const void *host_src, size_t bytes) {
	PXB_CHECK_ARG(ctx && dev_dst && host_src, "null argument");
	PXB_CUDA(cudaSetDevice(ctx->device));
	begin_call(ctx);
	PXB_TRY(h2d(ctx, dev_dst, host_src, bytes));
	return sync(ctx);
}
int pxb_memcpy_d2h(pxb_ctx *ctx, void *host_dst, const void *dev_src, size_t bytes) {
	PXB_CHECK_ARG(ctx && host_dst && dev_src, "null argument");
	PXB_CUDA(cudaSetDevice(ctx->device));
	begin_call(ctx);
	PXB_TRY(d2h(ctx, host_dst, dev_src, bytes));
	return sync(ctx);
}
int pxb_host_alloc_pinned(size_t bytes, void **host_ptr) {
	PXB_CHECK_ARG(host_ptr != nullptr, "null argument");
	PXB_CUDA(cudaMallocHost(host_ptr, std::max<size_t>(bytes, 1)));
	return PXB_OK;
}
int pxb_host_free_pinned(void *host_ptr) {
	if (host_ptr) PXB_CUDA(cudaFreeHost(host_ptr));
	return PXB_OK;
}

// ---- data ------------------------------------------------------------------------------------------------
int pxb_upload_points(pxb_ctx *ctx, int model_type, const double *pts_host, int64_t N) {
	PXB_CHECK_ARG(ctx != nullptr, "null context");
	begin_call(ctx);
	PXB_CHECK_ARG(model_type >= 0 && model_type <= PXB_MODEL_LINE2D, "unknown model type");
	PXB_CHECK_ARG(pts_host != nullptr && N > 0, "points must be a non-empty [N, dim] array");
	PXB_CUDA(cudaSetDevice(ctx->device));
	Points &p = ctx->pts;
	const int dim = point_dim(model_type);
	const int64_t stride = ((N + 63) / 64) * 64;
	if (p.soa == nullptr || p.stride != stride || p.dim != dim) {
		PXB_CUDA(cudaStreamSynchronize(ctx->stream));
		if (p.soa) PXB_CUDA(cudaFree(p.soa));
		if (p.aos) PXB_CUDA(cudaFree(p.aos));
		if (p.f32n) PXB_CUDA(cudaFree(p.f32n));
		if (p.q) PXB_CUDA(cudaFree(p.q));
		p.soa = p.aos = nullptr;
		p.f32n = p.q = nullptr;
		PXB_CUDA(cudaMalloc(&p.soa, sizeof(double) * (size_t)stride * dim));
		PXB_CUDA(cudaMalloc(&p.aos, sizeof(double) * (size_t)stride * dim));
		PXB_CUDA(cudaMalloc(&p.f32n, sizeof(float) * (size_t)stride * dim));
		PXB_CUDA(cudaMalloc(&p.q, sizeof(float) * (size_t)stride));
		if (!p.norm) PXB_CUDA(cudaMalloc(&p.norm, 64));
	}
	p.type = model_type;
	p.dim = dim;
	p.N = N;
	p.stride = stride;
	PXB_TRY(h2d(ctx, p.aos, pts_host, sizeof(double) * (size_t)N * dim));
	PXB_TRY(launch_aos_to_soa(ctx));
	return sync(ctx);
}

int64_t pxb_point_count(pxb_ctx *ctx) { return ctx ? ctx->pts.N : 0; }

// ---- a1/a2/a3 --------------------------------------------------------------------------------------------
int pxb_residual_matrix_dev(pxb_ctx *ctx, const double *models_dev, int64_t K, double T2, double *r2_dev,
                            uint32_t *mask_dev) {
	PXB_TRY(require_points(ctx));
	PXB_CHECK_ARG(models_dev != nullptr && K >= 0, "models");
	return launch_residual_matrix(ctx, models_dev, K, T2, r2_dev, nullptr, mask_dev);
}

int pxb_residual_matrix(pxb_ctx *ctx, const double *models_host, int64_t K, double T2, double *r2_host,
                        uint32_t *mask_host) {
	PXB_TRY(require_points(ctx));
	PXB_CHECK_ARG(models_host != nullptr && K >= 0, "models");
	if (K == 0) return PXB_OK;
	const int64_t N = ctx->pts.N, words = (N + 31) / 32;
	const int ms = model_size(ctx->pts.type);
	PXB_TRY(ctx->models.reserve(sizeof(double) * (size_t)K * ms));
	PXB_TRY(h2d(ctx, ctx->models.ptr, models_host, sizeof(double) * (size_t)K * ms));
	// the matrix can be large: process it in hypothesis slabs through a bounded device staging buffer
	const size_t row_bytes = (r2_host ? sizeof(double) * (size_t)N : 0) + (mask_host ? sizeof(uint32_t) * (size_t)words : 0);
	const int64_t slab = std::max<int64_t>(1, std::min<int64_t>(K, (int64_t)((size_t(1) << 30) / std::max<size_t>(row_bytes, 1))));
	if (r2_host) PXB_TRY(ctx->staging.reserve(sizeof(double) * (size_t)slab * N));
	if (mask_host) PXB_TRY(ctx->mask.reserve(sizeof(uint32_t) * (size_t)slab * words));
	for (int64_t k0 = 0; k0 < K; k0 += slab) {
		const int64_t kk = std::min(slab, K - k0);
		PXB_TRY(launch_residual_matrix(ctx, ctx->models.as<double>() + k0 * ms, kk, T2,
		                               r2_host ? ctx->staging.as<double>() : nullptr, nullptr,
		                               mask_host ? ctx->mask.as<uint32_t>() : nullptr));
		if (r2_host) PXB_TRY(d2h(ctx, r2_host + k0 * N, ctx->staging.ptr, sizeof(double) * (size_t)kk * N));
		if (mask_host) PXB_TRY(d2h(ctx, mask_host + k0 * words, ctx->mask.ptr, sizeof(uint32_t) * (size_t)kk * words));
	}
	return sync(ctx);
}

// ---- a4 --------------------------------------------------------------------------------------------------
int pxb_score_compound_dev(pxb_ctx *ctx, const double *models_dev, int64_t K, double T2,
                           const double *compound_pref_dev, int64_t *count_dev, double *value_sum_dev,
                           double *shared_dev) {
	PXB_TRY(require_points(ctx));
	PXB_CHECK_ARG(models_dev && count_dev && value_sum_dev && shared_dev && K >= 0, "null argument");
	return launch_score_compound(ctx, models_dev, K, T2, compound_pref_dev, count_dev, value_sum_dev, shared_dev);
}

int pxb_score_compound(pxb_ctx *ctx, const double *models_host, int64_t K, double T2,
                       const double *compound_pref_host, int64_t *count_host, double *value_sum_host,
                       double *shared_host) {
	PXB_TRY(require_points(ctx));
	PXB_CHECK_ARG(models_host && count_host && value_sum_host && shared_host && K >= 0, "null argument");
	if (K == 0) return PXB_OK;
	const int64_t N = ctx->pts.N;
	const int ms = model_size(ctx->pts.type);
	PXB_TRY(ctx->models.reserve(sizeof(double) * (size_t)K * ms));
	PXB_TRY(h2d(ctx, ctx->models.ptr, models_host, sizeof(double) * (size_t)K * ms));
	const double *cp = nullptr;
	if (compound_pref_host) {
		PXB_TRY(ctx->pref.reserve(sizeof(double) * (size_t)N));
		PXB_TRY(h2d(ctx, ctx->pref.ptr, compound_pref_host, sizeof(double) * (size_t)N));
		cp = ctx->pref.as<double>();
	}
	PXB_TRY(ctx->outA.reserve(sizeof(int64_t) * (size_t)K * 3));
	int64_t *cnt = ctx->outA.as<int64_t>();
	double *val = reinterpret_cast<double *>(cnt + K);
	double *shr = val + K;
	PXB_TRY(launch_score_compound(ctx, ctx->models.as<double>(), K, T2, cp, cnt, val, shr));
	PXB_TRY(d2h(ctx, count_host, cnt, sizeof(int64_t) * (size_t)K));
	PXB_TRY(d2h(ctx, value_sum_host, val, sizeof(double) * (size_t)K));
	PXB_TRY(d2h(ctx, shared_host, shr, sizeof(double) * (size_t)K));
	return sync(ctx);
}

int pxb_inliers(pxb_ctx *ctx, const double *model_host, double T2, int64_t *inliers_host, int64_t *n_inliers) {
	PXB_TRY(require_points(ctx));
	PXB_CHECK_ARG(model_host && inliers_host && n_inliers, "null argument");
	const int64_t N = ctx->pts.N, words = (N + 31) / 32;
	const int ms = model_size(ctx->pts.type);
	PXB_TRY(ctx->models.reserve(sizeof(double) * ms));
	PXB_TRY(h2d(ctx, ctx->models.ptr, model_host, sizeof(double) * ms));
	PXB_TRY(ctx->mask.reserve(sizeof(uint32_t) * (size_t)words));
	PXB_TRY(launch_residual_matrix(ctx, ctx->models.as<double>(), 1, T2, nullptr, nullptr, ctx->mask.as<uint32_t>()));
	PXB_TRY(ctx->reserve_pinned(sizeof(uint32_t) * (size_t)words));
	PXB_TRY(d2h(ctx, ctx->pinned, ctx->mask.ptr, sizeof(uint32_t) * (size_t)words));
	PXB_TRY(sync(ctx));
	// bit matrix -> ascending index list (format conversion of the kernel's output, no arithmetic)
	const uint32_t *w = reinterpret_cast<const uint32_t *>(ctx->pinned);
	int64_t n = 0;
	for (int64_t j = 0; j < words; ++j) {
		uint32_t bits = w[j];
		while (bits) {
			const int b = __builtin_ctz(bits);
			inliers_host[n++] = j * 32 + b;
			bits &= bits - 1;
		}
	}
	*n_inliers = n;
	return PXB_OK;
}

// ---- a5 --------------------------------------------------------------------------------------------------
int pxb_preference_vector(pxb_ctx *ctx, const double *model_host, double T, double *pref_host) {
	PXB_TRY(require_points(ctx));
	PXB_CHECK_ARG(model_host && pref_host, "null argument");
	const int64_t N = ctx->pts.N;
	const int ms = model_size(ctx->pts.type);
	PXB_TRY(ctx->models.reserve(sizeof(double) * ms));
	PXB_TRY(h2d(ctx, ctx->models.ptr, model_host, sizeof(double) * ms));
	PXB_TRY(ctx->pref2.reserve(sizeof(double) * (size_t)N));
	PXB_TRY(launch_preference(ctx, ctx->models.as<double>(), T, ctx->pref2.as<double>()));
	PXB_TRY(d2h(ctx, pref_host, ctx->pref2.ptr, sizeof(double) * (size_t)N));
	return sync(ctx);
}

int pxb_tanimoto(pxb_ctx *ctx, const double *a_host, const double *b_host, int64_t N, double *similarity) {
	PXB_CHECK_ARG(ctx && a_host && b_host && similarity && N > 0, "null argument");
	PXB_CUDA(cudaSetDevice(ctx->device));
	begin_call(ctx);
	PXB_TRY(ctx->pref.reserve(sizeof(double) * (size_t)N));
	PXB_TRY(ctx->pref2.reserve(sizeof(double) * (size_t)N));
	PXB_TRY(ctx->outA.reserve(sizeof(double) * 3));
	PXB_TRY(h2d(ctx, ctx->pref.ptr, a_host, sizeof(double) * (size_t)N));
	PXB_TRY(h2d(ctx, ctx->pref2.ptr, b_host, sizeof(double) * (size_t)N));
	PXB_TRY(launch_tanimoto(ctx, ctx->pref.as<double>(), ctx->pref2.as<double>(), N, ctx->outA.as<double>()));
	double out3[3];
	PXB_TRY(d2h(ctx, out3, ctx->outA.ptr, sizeof(out3)));
	PXB_TRY(sync(ctx));
	// progressive_x.h:584-585
	*similarity = out3[0] / (out3[1] + out3[2] - out3[0]);
	return PXB_OK;
}

int pxb_compound_max(pxb_ctx *ctx, const double *prefs_host, int64_t L, int64_t N, double *out_host) {
	PXB_CHECK_ARG(ctx && prefs_host && out_host && L >= 0 && N > 0, "null argument");
	PXB_CUDA(cudaSetDevice(ctx->device));
	begin_call(ctx);
	PXB_TRY(ctx->staging.reserve(sizeof(double) * (size_t)std::max<int64_t>(L, 1) * N));
	PXB_TRY(ctx->pref.reserve(sizeof(double) * (size_t)N));
	PXB_TRY(h2d(ctx, ctx->staging.ptr, prefs_host, sizeof(double) * (size_t)L * N));
	PXB_TRY(launch_compound_max(ctx, ctx->staging.as<double>(), L, N, ctx->pref.as<double>()));
	PXB_TRY(d2h(ctx, out_host, ctx->pref.ptr, sizeof(double) * (size_t)N));
	return sync(ctx);
}

// ---- a6/a7/a8 --------------------------------------------------------------------------------------------
int pxb_solve_minimal(pxb_ctx *ctx, const int64_t *samples_host, int64_t K, double *models_out_host,
                      int32_t *n_models_host, uint8_t *sample_valid_host, uint8_t *model_valid_host) {
	PXB_TRY(require_points(ctx));
	PXB_CHECK_ARG(samples_host && models_out_host && n_models_host && K >= 0, "null argument");
	if (K == 0) return PXB_OK;
	const int t = ctx->pts.type;
	const int m = sample_size(t), ms = model_size(t), mx = max_solutions(t);
	const int64_t N = ctx->pts.N;
	for (int64_t i = 0; i < K * m; ++i)
		if (samples_host[i] < 0 || samples_host[i] >= N) {
			set_error("sample index %lld out of range [0, %lld)", (long long)samples_host[i], (long long)N);
			return PXB_ERR_ARGUMENT;
		}
	PXB_TRY(ctx->idx.reserve(sizeof(int64_t) * (size_t)K * m));
	PXB_TRY(ctx->models.reserve(sizeof(double) * (size_t)K * mx * ms));
	PXB_TRY(ctx->outA.reserve(sizeof(int32_t) * (size_t)K + 2 * (size_t)K + 64));
	int32_t *nm = ctx->outA.as<int32_t>();
	uint8_t *sv = reinterpret_cast<uint8_t *>(nm + K);
	uint8_t *mv = sv + K;
	PXB_TRY(h2d(ctx, ctx->idx.ptr, samples_host, sizeof(int64_t) * (size_t)K * m));
	PXB_CUDA(cudaMemsetAsync(ctx->models.ptr, 0, sizeof(double) * (size_t)K * mx * ms, ctx->stream));
	PXB_TRY(launch_solve_minimal(ctx, ctx->idx.as<int64_t>(), K, ctx->models.as<double>(), nm, sv, mv));
	PXB_TRY(d2h(ctx, models_out_host, ctx->models.ptr, sizeof(double) * (size_t)K * mx * ms));
	PXB_TRY(d2h(ctx, n_models_host, nm, sizeof(int32_t) * (size_t)K));
	if (sample_valid_host) PXB_TRY(d2h(ctx, sample_valid_host, sv, (size_t)K));
	if (model_valid_host) PXB_TRY(d2h(ctx, model_valid_host, mv, (size_t)K));
	return sync(ctx);
}

int pxb_solve_plane_parallax(pxb_ctx *ctx, const int64_t *samples_host, int64_t K, const double *H_host,
                             double *models_out_host, int32_t *n_models_host) {
	PXB_TRY(require_points(ctx));
	PXB_CHECK_ARG(samples_host && H_host && models_out_host && n_models_host && K >= 0, "null argument");
	PXB_CHECK_ARG(ctx->pts.type == PXB_MODEL_FUNDAMENTAL, "plane-and-parallax needs two-view correspondences (PXB_MODEL_FUNDAMENTAL)");
	if (K == 0) return PXB_OK;
	const int64_t N = ctx->pts.N;
	for (int64_t i = 0; i < K * 2; ++i)
		if (samples_host[i] < 0 || samples_host[i] >= N) {
			set_error("sample index %lld out of range [0, %lld)", (long long)samples_host[i], (long long)N);
			return PXB_ERR_ARGUMENT;
		}
	PXB_TRY(ctx->idx.reserve(sizeof(int64_t) * (size_t)K * 2));
	PXB_TRY(ctx->models.reserve(sizeof(double) * (size_t)K * 9));
	PXB_TRY(ctx->outA.reserve(sizeof(int32_t) * (size_t)K));
	PXB_TRY(ctx->outC.reserve(sizeof(double) * 9));
	PXB_TRY(h2d(ctx, ctx->idx.ptr, samples_host, sizeof(int64_t) * (size_t)K * 2));
	PXB_TRY(h2d(ctx, ctx->outC.ptr, H_host, sizeof(double) * 9));
	PXB_CUDA(cudaMemsetAsync(ctx->models.ptr, 0, sizeof(double) * (size_t)K * 9, ctx->stream));
	PXB_TRY(launch_solve_plane_parallax(ctx, ctx->idx.as<int64_t>(), K, ctx->outC.as<double>(), ctx->models.as<double>(),
	                                    ctx->outA.as<int32_t>(), nullptr, nullptr));
	PXB_TRY(d2h(ctx, models_out_host, ctx->models.ptr, sizeof(double) * (size_t)K * 9));
	PXB_TRY(d2h(ctx, n_models_host, ctx->outA.ptr, sizeof(int32_t) * (size_t)K));
	return sync(ctx);
}

// ---- a9/a10/a11/a12 ----------------------------------------------------------------------------------------
int pxb_pearl_datacost(pxb_ctx *ctx, const double *models_host, int64_t L, double thr, double lambda,
                       double *D_host) {
	PXB_TRY(require_points(ctx));
	PXB_CHECK_ARG(D_host && L >= 0 && (L == 0 || models_host), "null argument");
	const int64_t N = ctx->pts.N;
	const int ms = model_size(ctx->pts.type);
	PXB_TRY(ctx->models.reserve(sizeof(double) * (size_t)std::max<int64_t>(L, 1) * ms));
	PXB_TRY(h2d(ctx, ctx->models.ptr, models_host, sizeof(double) * (size_t)L * ms));
	PXB_TRY(ctx->staging.reserve(sizeof(double) * (size_t)N * (L + 1)));
	PXB_TRY(launch_pearl_datacost(ctx, ctx->models.as<double>(), L, thr, lambda, ctx->staging.as<double>()));
	PXB_TRY(d2h(ctx, D_host, ctx->staging.ptr, sizeof(double) * (size_t)N * (L + 1)));
	return sync(ctx);
}

int pxb_pearl_label(pxb_ctx *ctx, const double *D_host, int64_t N, int32_t L1, double lambda, double label_cost,
                    const int32_t *csr_off_host, const int32_t *csr_idx_host, const int32_t *init_labels_host,
                    int32_t *labels_out_host, double *energy_out) {
	PXB_CHECK_ARG(ctx && D_host && labels_out_host && energy_out && N > 0 && L1 >= 1, "null argument");
	PXB_CUDA(cudaSetDevice(ctx->device));
	begin_call(ctx);
	PXB_TRY(ctx->staging.reserve(sizeof(double) * (size_t)N * L1));
	PXB_TRY(h2d(ctx, ctx->staging.ptr, D_host, sizeof(double) * (size_t)N * L1));
	return pearl_label_device(ctx, ctx->staging.as<double>(), N, L1, lambda, label_cost, csr_off_host, csr_idx_host,
	                          init_labels_host, labels_out_host, energy_out);
}

int pxb_segment_residual_sums(pxb_ctx *ctx, const double *models_host, int64_t L, const int32_t *labels_host,
                              double *sums_host, int64_t *counts_host) {
	PXB_TRY(require_points(ctx));
	PXB_CHECK_ARG(models_host && labels_host && sums_host && counts_host && L >= 0, "null argument");
	if (L == 0) return PXB_OK;
	const int64_t N = ctx->pts.N;
	const int ms = model_size(ctx->pts.type);
	PXB_TRY(ctx->models.reserve(sizeof(double) * (size_t)L * ms));
	PXB_TRY(h2d(ctx, ctx->models.ptr, models_host, sizeof(double) * (size_t)L * ms));
	PXB_TRY(ctx->outA.reserve(sizeof(int32_t) * (size_t)N));
	PXB_TRY(h2d(ctx, ctx->outA.ptr, labels_host, sizeof(int32_t) * (size_t)N));
	PXB_TRY(ctx->outB.reserve(sizeof(double) * (size_t)L * 2));
	double *sums = ctx->outB.as<double>();
	int64_t *counts = reinterpret_cast<int64_t *>(sums + L);
	PXB_TRY(launch_segment_sums(ctx, ctx->models.as<double>(), L, ctx->outA.as<int32_t>(), sums, counts));
	PXB_TRY(d2h(ctx, sums_host, sums, sizeof(double) * (size_t)L));
	PXB_TRY(d2h(ctx, counts_host, counts, sizeof(int64_t) * (size_t)L));
	return sync(ctx);
}

// ---- a13 -------------------------------------------------------------------------------------------------
int pxb_lo_unary_terms(pxb_ctx *ctx, const double *model_host, double thr, double lambda, double *d_host,
                       double *e0_host, double *e1_host) {
	PXB_TRY(require_points(ctx));
	PXB_CHECK_ARG(model_host && d_host && e0_host && e1_host, "null argument");
	const int64_t N = ctx->pts.N;
	const int ms = model_size(ctx->pts.type);
	PXB_TRY(ctx->models.reserve(sizeof(double) * ms));
	PXB_TRY(h2d(ctx, ctx->models.ptr, model_host, sizeof(double) * ms));
	PXB_TRY(ctx->staging.reserve(sizeof(double) * (size_t)N * 3));
	double *d = ctx->staging.as<double>(), *e0 = d + N, *e1 = e0 + N;
	PXB_TRY(launch_lo_unary(ctx, ctx->models.as<double>(), thr, lambda, d, e0, e1));
	PXB_TRY(d2h(ctx, d_host, d, sizeof(double) * (size_t)N));
	PXB_TRY(d2h(ctx, e0_host, e0, sizeof(double) * (size_t)N));
	PXB_TRY(d2h(ctx, e1_host, e1, sizeof(double) * (size_t)N));
	return sync(ctx);
}

int pxb_tukey_weights(pxb_ctx *ctx, const double *model_host, double T2, double *weights_host) {
	PXB_TRY(require_points(ctx));
	PXB_CHECK_ARG(model_host && weights_host, "null argument");
	const int64_t N = ctx->pts.N;
	const int ms = model_size(ctx->pts.type);
	PXB_TRY(ctx->models.reserve(sizeof(double) * ms));
	PXB_TRY(h2d(ctx, ctx->models.ptr, model_host, sizeof(double) * ms));
	PXB_TRY(ctx->pref2.reserve(sizeof(double) * (size_t)N));
	PXB_TRY(launch_tukey(ctx, ctx->models.as<double>(), T2, ctx->pref2.as<double>()));
	PXB_TRY(d2h(ctx, weights_host, ctx->pref2.ptr, sizeof(double) * (size_t)N));
	return sync(ctx);
}

} // extern "C"
