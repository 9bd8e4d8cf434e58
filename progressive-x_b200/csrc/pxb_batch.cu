// pxb_batch.cu -- many independent problems on one GPU (BASELINE config C4) without one host thread per problem.
//
// One fit is a chain of ~150 short kernel sequences separated by host decisions (the block replay of GCRANSAC::run, the
// LO / IRLS / PEARL loops), so a single problem leaves the GPU idle most of the time and a host thread that blocks in
// cudaStreamSynchronize does nothing for tens of microseconds per decision. Here every problem runs as a FIBER (ucontext)
// with its own pxb_ctx (stream + scratch); wherever the driver would block on its stream it polls cudaStreamQuery and
// yields to the scheduler of its host thread, which resumes the next fiber. One host thread thus keeps `in_flight`
// problems moving: the host work of one overlaps the kernel / copy latency of the others, and the kernels of different
// problems overlap on the device. Results do not depend on the interleaving: every problem only touches its own context
// and its own seed. The driver code is unchanged -- the only hook is pxb::ctx_wait (pxb_api.cu).
#include <ucontext.h>

#include <atomic>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "pxb_internal.h"

namespace pxb {
namespace {

struct BatchJob { // shared by all host threads of one call
	int64_t n_pairs;
	const double *const *corr;
	const int64_t *n_points;
	int64_t *const *labeling_out;
	double *const *models_out;
	int64_t max_models_out;
	int32_t *n_models_out;
	size_t w1, h1, w2, h2;
	double lambda, threshold, confidence, radius, max_tanimoto;
	size_t max_iters, min_points;
	int max_models;
	size_t sampler_id;
	double scoring_exponent;
	uint64_t seed;
	int per_pair_seed;
	std::atomic<int64_t> next{0};
	std::atomic<int> first_error{0};
	std::mutex error_mutex;
	std::string error_text;
};

struct Fiber {
	ucontext_t uc;
	ucontext_t *sched = nullptr;
	std::vector<unsigned char> stack;
	pxb_ctx *ctx = nullptr;
	BatchJob *job = nullptr;
	bool done = false;
};

void fiber_yield(void *arg) {
	Fiber *f = static_cast<Fiber *>(arg);
	swapcontext(&f->uc, f->sched);
}

void fiber_main(unsigned lo, unsigned hi) {
	Fiber *f = reinterpret_cast<Fiber *>(((uint64_t)hi << 32) | (uint64_t)lo);
	BatchJob *j = f->job;
	for (;;) {
		const int64_t p = j->next.fetch_add(1);
		if (p >= j->n_pairs || j->first_error.load() != 0) break;
		const uint64_t seed = j->per_pair_seed ? j->seed + (uint64_t)p : j->seed;
		const int rc = pxb_find_homographies(f->ctx, j->corr[p], j->n_points[p], j->labeling_out[p], j->models_out[p],
		                                     j->max_models_out, j->w1, j->h1, j->w2, j->h2, j->lambda, j->threshold, j->confidence,
		                                     j->radius, j->max_tanimoto, j->max_iters, j->min_points, j->max_models, j->sampler_id,
		                                     j->scoring_exponent, 0, seed);
		if (rc < 0) {
			std::lock_guard<std::mutex> g(j->error_mutex);
			if (j->first_error.load() == 0) {
				j->first_error.store(rc);
				j->error_text = pxb_last_error();
			}
			break;
		}
		j->n_models_out[p] = rc;
	}
	f->done = true;
	swapcontext(&f->uc, f->sched); // never resumed
}

// the contexts of the batch driver are kept per device across calls (stream, pinned arena and scratch buffers are reused)
std::mutex g_pool_mutex;
std::vector<std::vector<pxb_ctx *>> g_pool; // [device][slot]

int take_contexts(int device, int count, std::vector<pxb_ctx *> &out) {
	std::lock_guard<std::mutex> g(g_pool_mutex);
	if ((int)g_pool.size() <= device) g_pool.resize(device + 1);
	auto &pool = g_pool[device];
	while ((int)pool.size() < count) {
		pxb_ctx *c = nullptr;
		PXB_TRY(pxb_ctx_create(device, &c));
		pool.push_back(c);
	}
	out.assign(pool.begin(), pool.begin() + count);
	return PXB_OK;
}

std::mutex g_call_mutex; // one batch call at a time per process: the pooled contexts are not re-entrant

void run_thread(BatchJob *job, pxb_ctx **ctxs, int n_fibers, int device) {
	cudaSetDevice(device);
	ucontext_t sched;
	std::vector<Fiber> fibers((size_t)n_fibers);
	constexpr size_t kStack = size_t(1) << 20;
	for (int i = 0; i < n_fibers; ++i) {
		Fiber &f = fibers[i];
		f.stack.resize(kStack);
		f.sched = &sched;
		f.ctx = ctxs[i];
		f.job = job;
		f.ctx->yield_fn = fiber_yield;
		f.ctx->yield_arg = &f;
		getcontext(&f.uc);
		f.uc.uc_stack.ss_sp = f.stack.data();
		f.uc.uc_stack.ss_size = f.stack.size();
		f.uc.uc_link = &sched;
		const uint64_t a = reinterpret_cast<uint64_t>(&f);
		makecontext(&f.uc, reinterpret_cast<void (*)()>(fiber_main), 2, (unsigned)(a & 0xffffffffu), (unsigned)(a >> 32));
	}
	for (;;) { // round robin: a fiber runs until it has to wait for its stream (or finishes)
		bool any = false;
		for (int i = 0; i < n_fibers; ++i) {
			if (fibers[i].done) continue;
			any = true;
			swapcontext(&sched, &fibers[i].uc);
		}
		if (!any) break;
	}
	for (int i = 0; i < n_fibers; ++i) {
		ctxs[i]->yield_fn = nullptr;
		ctxs[i]->yield_arg = nullptr;
	}
}

} // namespace
} // namespace pxb

using namespace pxb;

extern "C" {

int pxb_find_homographies_batch(int device, int64_t n_pairs, const double *const *correspondences, const int64_t *n_points,
                                int64_t *const *labeling_out, double *const *models_out, int64_t max_models_out,
                                int32_t *n_models_out, size_t source_image_width, size_t source_image_height,
                                size_t destination_image_width, size_t destination_image_height,
                                double spatial_coherence_weight, double threshold, double confidence,
                                double neighborhood_ball_radius, double maximum_tanimoto_similarity, size_t max_iters,
                                size_t minimum_point_number, int maximum_model_number, size_t sampler_id,
                                double scoring_exponent, uint64_t seed, int per_pair_seed, int host_threads, int in_flight) {
	PXB_CHECK_ARG(n_pairs >= 0 && correspondences && n_points && labeling_out && models_out && n_models_out, "null argument");
	PXB_CHECK_ARG(host_threads >= 1 && host_threads <= 64 && in_flight >= 1 && in_flight <= 64, "host_threads / in_flight in [1, 64]");
	if (n_pairs == 0) return PXB_OK;
	std::lock_guard<std::mutex> call_guard(g_call_mutex);
	int threads = (int)std::min<int64_t>(host_threads, n_pairs);
	int fibers = (int)std::min<int64_t>(in_flight, (n_pairs + threads - 1) / threads);
	std::vector<pxb_ctx *> ctxs;
	PXB_TRY(take_contexts(device, threads * fibers, ctxs));
	BatchJob job;
	job.n_pairs = n_pairs;
	job.corr = correspondences;
	job.n_points = n_points;
	job.labeling_out = labeling_out;
	job.models_out = models_out;
	job.max_models_out = max_models_out;
	job.n_models_out = n_models_out;
	job.w1 = source_image_width, job.h1 = source_image_height, job.w2 = destination_image_width, job.h2 = destination_image_height;
	job.lambda = spatial_coherence_weight, job.threshold = threshold, job.confidence = confidence;
	job.radius = neighborhood_ball_radius, job.max_tanimoto = maximum_tanimoto_similarity;
	job.max_iters = max_iters, job.min_points = minimum_point_number, job.max_models = maximum_model_number;
	job.sampler_id = sampler_id, job.scoring_exponent = scoring_exponent;
	job.seed = seed, job.per_pair_seed = per_pair_seed;
	if (threads == 1) {
		run_thread(&job, ctxs.data(), fibers, device);
	} else {
		std::vector<std::thread> pool;
		for (int t = 0; t < threads; ++t) pool.emplace_back(run_thread, &job, ctxs.data() + (size_t)t * fibers, fibers, device);
		for (auto &t : pool) t.join();
	}
	if (job.first_error.load() != 0) {
		set_error("%s", job.error_text.c_str());
		return job.first_error.load();
	}
	return PXB_OK;
}

void pxb_batch_release(void) {
	std::lock_guard<std::mutex> call_guard(g_call_mutex);
	std::lock_guard<std::mutex> g(g_pool_mutex);
	for (auto &pool : g_pool) {
		for (pxb_ctx *c : pool) pxb_ctx_destroy(c);
		pool.clear();
	}
}

} // extern "C"
