// pxb_driver.cu -- host side of the drop-in: the Progressive-X anytime loop, the GC-RANSAC proposal engine and
// the PEARL optimiser, re-stated as a block-replay driver over the CUDA operators of this library.
//
// What stays on the host is exactly what the reference keeps sequential (control flow, RNG state machines, the
// <= 10 outer proposals); every N-point loop runs on the device through the operators of include/pxb200.h.
//
//   ProgressiveX::run            px/include/progressive_x.h:251-489      -> Driver::run
//   isPutativeModelValid         px/include/progressive_x.h:565-591      -> Driver::putative_model_valid
//   updateCompoundModel          px/include/progressive_x.h:597-624      -> pxb_compound_max over the stored prefs
//   GCRANSAC::run                gcr/GCRANSAC.h:203-628                  -> Driver::propose (block replay)
//   graphCutLocalOptimization    gcr/GCRANSAC.h:781-911                  -> Driver::local_optimization
//   iteratedLeastSquaresFitting  gcr/GCRANSAC.h:631-759                  -> Driver::irls
//   PEARL::run / labeling / parameterEstimation / rejectInstances
//                                px/include/PEARL.h:275-555              -> Driver::pearl
//
// Block replay. The reference draws one minimal sample, solves, scores N points, compares, repeats. Here the main
// sampler is run ahead for a block of B samples; one launch solves them, one launch scores every model against
// all N points, and the host then replays the reference's sequential bookkeeping (iteration counting incl. failed
// generations, so-far-the-best updates with the early-exit rule of getScore, adaptive max_iteration, LO trigger)
// over the pre-evaluated block in sample order. The decisions are those the sequential loop would take on the same
// sample stream, because a hypothesis' (count, value, shared) does not depend on the state of the loop.
//
// Deviations from the reference (documented in DESIGN.md):
//   * RNG: own seedable generator (the reference seeds std::mt19937 from std::random_device and cannot be replayed);
//   * neighbourhood graph: exact kNN-in-radius on the GPU instead of randomised FLANN;
//   * the reference's two-buffer inlier ping-pong (GCRANSAC.h:244-252,:546-552) is replaced by "the inlier list of the
//     current best model" -- the reference's version depends on list-size coincidences;
//   * samplers: uniform (0), PROSAC (1), Progressive NAPSAC (2, over four grid layers) and NAPSAC (3; id 2 for lines)
//     follow the reference's state machines on the seedable generator;
//   * non-minimal fits ("next"): H through the 8x8 normal equations (same least-squares solution as
//     the reference's QR); F as normalised 8-point + rank-2 projection and PnP as normalised DLT + 10 LM steps, i.e.
//     without PoseLib's bundle adjustment / OpenCV's EPnP (pxb_fit_fp.cu); DEGENSAC is not applied to F.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <numeric>
#include <vector>

#include "pxb_internal.h"

namespace pxb {

int launch_knn_graph(pxb_ctx *ctx, double radius, int k, int32_t *nbr, int32_t *deg);
int launch_fit_h(pxb_ctx *ctx, int P, const int32_t *off, const int32_t *idx, const double *weights, double *H_out,
                 int32_t *ok_out);
int launch_fit_f(pxb_ctx *ctx, int P, const int32_t *off, const int32_t *idx, const double *weights, double *F_out,
                 int32_t *ok_out);
int launch_fit_pnp(pxb_ctx *ctx, int P, const int32_t *off, const int32_t *idx, double *P_out, int32_t *ok_out);
int launch_f_sym_count(pxb_ctx *ctx, const double *model, double T2, double Tsym2, long long *out2);
int launch_fit_vp(pxb_ctx *ctx, int P, const int32_t *off, const int32_t *idx, const double *weights_by_point, double *out,
                  int32_t *ok_out);
int launch_fit_line(pxb_ctx *ctx, int P, const int32_t *off, const int32_t *idx, double *out, int32_t *ok_out);

namespace {

// Neighbours kept per point. The reference's FlannNeighborhoodGraph asks OpenCV's FLANN for radiusMatch with
// SearchParams(checks = 6) (gcr/neighborhood/flann_neighborhood_graph.h:100-139): at most 6 approximate matches come back
// whatever the radius, and the first one is dropped -- <= 5 neighbours per point (2-4.6 on average on the AdelaideRMF
// scenes, probed with cv2 here). The deterministic stand-in keeps the 5 nearest points inside the radius; with more
// neighbours the Potts term of PEARL (lambda per cut edge) outweighs the data term and every instance is swallowed by the
// outlier label at the reference's own lambda = 0.5 (adelaideF.ipynb).
inline int graph_degree() {
	if (const char *e = getenv("PXB_GRAPH_K")) {
		const int k = atoi(e);
		if (k >= 1 && k <= 16) return k;
	}
	return 5;
}

// ---- RNG -----------------------------------------------------------------------------------------------------
struct Rng {
	uint64_t s;
	explicit Rng(uint64_t seed) : s(seed ? seed : 0x9E3779B97F4A7C15ull) {}
	uint64_t next() {
		uint64_t z = (s += 0x9E3779B97F4A7C15ull);
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
		z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
		return z ^ (z >> 31);
	}
	// uniform integer in [0, max] (inclusive, like std::uniform_int_distribution(0, max)), rejection sampled
	size_t uniform(size_t max) {
		const uint64_t range = (uint64_t)max + 1;
		if (range == 0) return (size_t)next();
		const uint64_t limit = UINT64_MAX - (UINT64_MAX % range);
		uint64_t r;
		do r = next(); while (r >= limit);
		return (size_t)(r % range);
	}
	// gcr/uniform_random_generator.h:76-122: unique set by rejection, optional value to skip
	void unique_set(size_t *out, size_t n, size_t max, bool has_skip = false, size_t skip = 0) {
		for (size_t i = 0; i < n; i++) {
			out[i] = uniform(max);
			if (has_skip && out[i] == skip) {
				i--;
				continue;
			}
			for (int j = (int)i - 1; j >= 0; j--)
				if (out[i] == out[j]) {
					i--;
					break;
				}
		}
	}
};

struct Graph { // directed neighbour lists, CSR
	std::vector<int32_t> off, idx;
	int64_t degree(int64_t i) const { return off[i + 1] - off[i]; }
	const int32_t *nbrs(int64_t i) const { return idx.data() + off[i]; }
};

// ---- samplers (gcr/samplers/sampler.h:45-86 contract: sample(pool, subset, m) -> bool) -------------------------
struct Sampler {
	virtual ~Sampler() {}
	virtual bool sample(const std::vector<size_t> &pool, size_t *subset, size_t m) = 0;
};
struct UniformSampler : Sampler { // gcr/samplers/uniform_sampler.h:118-134
	Rng rng;
	explicit UniformSampler(uint64_t seed) : rng(seed) {}
	bool sample(const std::vector<size_t> &pool, size_t *subset, size_t m) override {
		if (m > pool.size()) return false;
		rng.unique_set(subset, m, pool.size() - 1);
		for (size_t i = 0; i < m; ++i) subset[i] = pool[subset[i]];
		return true;
	}
};
struct NapsacSampler : Sampler { // gcr/samplers/napsac_sampler.h:102-151
	Rng rng;
	const Graph *g;
	size_t maximum_iterations = 100;
	NapsacSampler(uint64_t seed, const Graph *g_) : rng(seed), g(g_) {}
	bool sample(const std::vector<size_t> &pool, size_t *subset, size_t m) override {
		if (m > pool.size()) return false;
		size_t attempts = 0;
		while (attempts++ < maximum_iterations) {
			rng.unique_set(subset, 1, pool.size() - 1);
			const int64_t deg = g->degree((int64_t)subset[0]);
			const int32_t *nb = g->nbrs((int64_t)subset[0]);
			if ((size_t)deg < m) continue;
			if ((size_t)deg == m) {
				for (size_t i = 0; i < m; ++i) subset[i] = (size_t)nb[i];
				break;
			}
			// :139-142 passes the centre *point index* as the neighbour-list *index* to skip (reference quirk, kept)
			rng.unique_set(subset + 1, m - 1, (size_t)deg - 1, true, subset[0]);
			for (size_t i = 1; i < m; ++i) subset[i] = (size_t)nb[subset[i]];
			break;
		}
		return attempts < maximum_iterations;
	}
};

// gcr/samplers/prosac_sampler.h: samples are drawn from a pool that grows along the (quality-ordered) point list; the
// newest point of the pool is always part of the sample. reset() state per proposal (progressive_x.h:290-291).
struct ProsacSampler : Sampler {
	Rng rng;
	size_t sample_size, point_number, convergence, kth = 1, subset_size, gen_max;
	std::vector<size_t> growth;
	ProsacSampler(uint64_t seed, size_t m, size_t N, size_t convergence_iterations = 100000)
	    : rng(seed), sample_size(m), point_number(N), convergence(convergence_iterations), subset_size(m), gen_max(m - 1) {
		growth.assign(N, 0); // :initialize
		double T_n = (double)convergence;
		for (size_t i = 0; i < m; i++) T_n *= static_cast<double>(m - i) / (double)(N - i);
		size_t T_n_prime = 1;
		for (size_t i = 0; i < N; ++i) {
			if (i + 1 <= m) {
				growth[i] = T_n_prime;
				continue;
			}
			const double Tn_plus1 = static_cast<double>(i + 1) * T_n / (double)(i + 1 - m);
			growth[i] = T_n_prime + (unsigned int)std::ceil(Tn_plus1 - T_n);
			T_n = Tn_plus1;
			T_n_prime = growth[i];
		}
	}
	void increment() { // incrementIterationNumber
		++kth;
		if (kth > convergence) {
			gen_max = point_number - 1;
		} else if (kth > growth[subset_size - 1]) {
			++subset_size;
			if (subset_size > point_number) subset_size = point_number;
			gen_max = subset_size - 2;
		}
	}
	void set_sample_number(size_t k) { // setSampleNumber
		kth = k;
		if (kth > convergence) {
			gen_max = point_number - 1;
		} else {
			while (kth > growth[subset_size - 1] && subset_size != point_number) {
				++subset_size;
				if (subset_size > point_number) subset_size = point_number;
				gen_max = subset_size - 2;
			}
		}
	}
	bool sample(const std::vector<size_t> &, size_t *subset, size_t m) override {
		if (m != sample_size) { // "PROSAC is not yet implemented to change the sample size"
			increment();
			return false;
		}
		if (kth > convergence) {
			rng.unique_set(subset, m, gen_max);
			return true;
		}
		rng.unique_set(subset, m - 1, gen_max);
		subset[m - 1] = subset_size - 1; // the last index is the point at the end of the current pool
		increment();
		return true;
	}
};

// gcr/neighborhood/grid_neighborhood_graph.h: points hashed into a regular grid of `cells` cells per axis; the
// neighbours of a point are the points of its cell (itself included) in row order. Coordinates outside [0, size) are
// clamped to the border cells (the reference casts a negative floor() to size_t).
struct GridLayer {
	std::vector<std::vector<size_t>> cell_points;
	std::vector<size_t> cell_of_point;
	void build(const double *rows, size_t N, int dim, const double *sizes, size_t cells) {
		std::vector<std::pair<size_t, size_t>> key((size_t)N);
		for (size_t i = 0; i < N; ++i) {
			size_t index = 0, offset = 1;
			for (int d = 0; d < dim; ++d) {
				const double cs = sizes[d] / (double)cells;
				double f = std::floor(rows[i * dim + d] / cs);
				if (!(f >= 0)) f = 0;
				if (f > (double)(cells - 1)) f = (double)(cells - 1);
				index += offset * (size_t)f;
				offset *= cells;
			}
			key[i] = {index, i};
		}
		std::vector<std::pair<size_t, size_t>> sorted = key;
		std::sort(sorted.begin(), sorted.end());
		cell_of_point.assign(N, 0);
		cell_points.clear();
		for (size_t t = 0; t < N; ++t) {
			if (t == 0 || sorted[t].first != sorted[t - 1].first) cell_points.emplace_back();
			cell_points.back().push_back(sorted[t].second); // row order inside a cell (pairs sort by index, then row)
			cell_of_point[sorted[t].second] = cell_points.size() - 1;
		}
	}
	const std::vector<size_t> &neighbors(size_t i) const { return cell_points[cell_of_point[i]]; }
};

// gcr/samplers/progressive_napsac_sampler.h: local samples around a PROSAC-chosen centre from the finest grid layer that
// holds enough points, blending into global PROSAC sampling. Layers {16, 8, 4, 2}, sampler length 0.5
// (progressivex_python.cpp:227-235). The grid layers depend on the data only and are shared by all proposals.
struct ProgressiveNapsacSampler : Sampler {
	Rng rng;
	const std::vector<GridLayer> *layers;
	ProsacSampler one_point, prosac;
	size_t sample_size, point_number, kth = 0, max_local_iterations;
	std::vector<size_t> current_layer, hits, subset_size_of, growth;
	ProgressiveNapsacSampler(uint64_t seed, size_t m, size_t N, const std::vector<GridLayer> *layers_, double sampler_length)
	    : rng(seed), layers(layers_), one_point(seed ^ 0x5851F42D4C957F2Dull, 1, N, N), prosac(seed ^ 0x14057B7EF767814Full, m, N, N),
	      sample_size(m), point_number(N), current_layer(N, 0), hits(N, 0), subset_size_of(N, m) {
		max_local_iterations = static_cast<size_t>(sampler_length * (double)N);
		growth.assign(N, 0);
		const size_t local = m - 1;
		double T_n = (double)max_local_iterations;
		for (size_t i = 0; i < local; ++i) T_n *= static_cast<double>(local - i) / (double)(N - i);
		unsigned int T_n_prime = 1;
		for (size_t i = 0; i < N; ++i) {
			if (i + 1 <= local) {
				growth[i] = T_n_prime;
				continue;
			}
			const double Tn_plus1 = static_cast<double>(i + 1) * T_n / (double)(i + 1 - local);
			growth[i] = T_n_prime + static_cast<size_t>(std::ceil(Tn_plus1 - T_n));
			T_n = Tn_plus1;
			T_n_prime = (unsigned int)growth[i];
		}
	}
	bool sample(const std::vector<size_t> &pool, size_t *subset, size_t m) override {
		++kth;
		if (m != sample_size) return false;
		if (m > pool.size()) return false;
		if (kth > max_local_iterations) { // fully blended into global sampling
			prosac.set_sample_number(kth);
			return prosac.sample(pool, subset, m);
		}
		if (!one_point.sample(pool, subset, 1)) return false;
		const size_t centre = subset[0];
		const size_t h = ++hits[centre];
		size_t &ss = subset_size_of[centre];
		while (h > growth[ss - 1] && ss < point_number) ss = std::min(ss + 1, point_number);
		size_t &layer = current_layer[centre];
		bool last = false;
		while (true) { // the finest grid whose cell holds enough points
			if (layer >= layers->size()) {
				last = true;
				break;
			}
			if ((*layers)[layer].neighbors(centre).size() < ss) {
				++layer;
				continue;
			}
			break;
		}
		if (last) {
			prosac.set_sample_number(kth);
			const bool ok = prosac.sample(pool, subset, m);
			subset[m - 1] = centre;
			return ok;
		}
		const std::vector<size_t> &nb = (*layers)[layer].neighbors(centre);
		subset[m - 1] = centre;
		subset[m - 2] = nb[ss - 1];
		rng.unique_set(subset, m - 2, ss - 2, true, centre); // (:..., neighbour index compared with a point index: kept)
		for (size_t i = 0; i + 2 < m; ++i) {
			subset[i] = nb[subset[i]];
			++hits[subset[i]];
		}
		++hits[subset[m - 2]];
		return true;
	}
};

// Samples drawn ahead per refill of the block replay (GCRANSAC::run's main loop). 1024 covers the Python default
// max_iters = 1000 in one solve + score round trip; the result does not depend on it.
constexpr size_t kReplayBlock = 1024;

struct Score { // gcr/scoring_function.h:48-73
	int64_t inliers = 0;
	double value = 0.0;
};

struct Settings {
	int type = PXB_MODEL_HOMOGRAPHY;
	double threshold = 2.0, confidence = 0.95, lambda = 0.14, max_tanimoto = 0.5;
	size_t min_inliers = 20, max_iters = 5000, max_models = std::numeric_limits<size_t>::max();
	size_t max_proposals_without_change = 10; // progressive_x.h:63
	int exponent = 2;                         // scoring_function_with_compound_model.h:20 (int!)
	size_t sampler_id = 0;
	bool napsac = false;                      // main sampler = NapsacSampler (H/F: id 3, lines: id 2)
	bool progressive_napsac = false;          // main sampler = ProgressiveNapsacSampler (H/F: id 2)
	double sizes[4] = {0, 0, 0, 0};           // image sizes: cell sizes of the P-NAPSAC grid layers
	std::vector<double> point_weights;        // MultiModelSettings::point_weights (progressive_x.h:36), VP only
	// FundamentalMatrixEstimator(minimum_inlier_ratio_in_validity_check = 0.5, use_degensac = true)
	double sym_epipolar_ratio = 0.5;
	bool use_degensac = true;
	// DEGENSAC's nested run: FundamentalMatrixPlaneParallaxSolver with a fixed homography as the minimal solver
	bool plane_parallax = false;
	double pp_H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
	const double *rows_host = nullptr; // the caller's [N, dim] rows (DEGENSAC's seven-point test runs on the host)
	bool do_logging = false;
	bool allow_shard = true; // nested (DEGENSAC) runs never shard
	uint64_t seed = 1;
	// gcransac::utils::Settings defaults (gcr/settings.h:66-86) as overridden by progressive_x.h:64-71
	size_t min_iteration_number = 20, min_iteration_number_before_lo = 20, max_local_optimization_number = 50,
	       max_graph_cut_number = 10, max_least_squares_iterations = 10, max_unsuccessful_model_generations = 100;
};

// PXB_PROFILE=1: host wall time per phase of one find* call, printed to stderr when the driver finishes (tuning aid).
struct PhaseTimes {
	bool on = getenv("PXB_PROFILE") != nullptr;
	std::vector<std::pair<const char *, double>> acc;
	std::vector<long> calls;
	void add(const char *name, double ms) {
		for (size_t i = 0; i < acc.size(); ++i)
			if (acc[i].first == name) {
				acc[i].second += ms;
				calls[i]++;
				return;
			}
		acc.emplace_back(name, ms);
		calls.push_back(1);
	}
	void print() const {
		if (!on) return;
		double tot = 0;
		for (auto &a : acc) tot += a.second;
		fprintf(stderr, "[pxb profile] total %.2f ms\n", tot);
		for (size_t i = 0; i < acc.size(); ++i)
			fprintf(stderr, "[pxb profile]   %-22s %8.2f ms  %6ld calls\n", acc[i].first, acc[i].second, calls[i]);
	}
};
struct Scoped {
	PhaseTimes &pt;
	const char *name;
	std::chrono::steady_clock::time_point t0;
	Scoped(PhaseTimes &p, const char *n) : pt(p), name(n), t0(std::chrono::steady_clock::now()) {}
	~Scoped() {
		if (pt.on) pt.add(name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
	}
};

struct Instance {
	std::vector<double> model;
	// (the preference vector as of acceptance -- never refreshed: progressive_x.h:597-624 -- is row k of ctx->pref_rows)
};

class Driver {
  public:
	Driver(pxb_ctx *ctx, const Settings &s) : ctx_(ctx), s_(s) {
		if (ctx->has_engine_settings && s.allow_shard) { // getMutableSettings(): the knobs no argument list carries
			const pxb_multi_model_settings &e = ctx->engine_settings;
			s_.max_proposals_without_change = e.max_proposal_number_without_change;
			s_.min_iteration_number = e.min_iteration_number;
			s_.min_iteration_number_before_lo = e.min_iteration_number_before_lo;
			s_.max_local_optimization_number = e.max_local_optimization_number;
			s_.max_graph_cut_number = e.max_graph_cut_number;
			s_.max_least_squares_iterations = e.max_least_squares_iterations;
			s_.max_unsuccessful_model_generations = e.max_unsuccessful_model_generations;
		}
		N_ = ctx->pts.N;
		ms_ = model_size(s.type);
		m_ = s.plane_parallax ? 2 : sample_size(s.type);
		maxsol_ = s.plane_parallax ? 1 : max_solutions(s.type);
	}
	~Driver() {
		if (ctx_->trusted_csr_off == graph_.off.data()) { // the arrays die with this driver
			ctx_->trusted_csr_key = 0;
			ctx_->trusted_csr_off = ctx_->trusted_csr_idx = nullptr;
		}
	}
	int build_graph(double radius, int k);
	void build_grid_layers(const double *rows) { // ProgressiveNapsacSampler<4>(.., {16, 8, 4, 2}, .., sizes, 0.5)
		const size_t cells[4] = {16, 8, 4, 2};
		grid_layers_.resize(4);
		for (int l = 0; l < 4; ++l) grid_layers_[l].build(rows, (size_t)N_, 4, s_.sizes, cells[l]);
	}
	int run(int setup_status = PXB_OK);
	int run_local();
	bool sharded() const { return ctx_->shard_comm != nullptr && s_.allow_shard && !s_.plane_parallax; }
	bool is_worker() const { return sharded() && ctx_->shard_rank != 0; }
	const std::vector<Instance> &instances() const { return models_; }
	const std::vector<int64_t> &labeling() const { return labeling_; }

	PhaseTimes prof_;

  private:
	pxb_ctx *ctx_;
	Settings s_;
	int64_t N_;
	int ms_, m_, maxsol_;
	Graph graph_;
	std::vector<GridLayer> grid_layers_;
	std::vector<Instance> models_;
	std::vector<int64_t> labeling_;
	size_t pearl_outliers_ = 0;
	// GC-RANSAC statistics of the last proposal
	size_t iteration_number_ = 0, graph_cut_number_ = 0, lo_number_ = 0;
	size_t degensac_degenerate_ = 0, degensac_updates_ = 0; // H-degenerate samples seen / models replaced (logging)
	std::vector<int64_t> proposal_inliers_;
	// one packed device-to-host copy of (count, value, shared) per scoring call, unpacked after the sync
	struct PendingScores {
		std::vector<int64_t> *cnt;
		std::vector<double> *val, *shr;
		int64_t K;
	} pending_scores_ = {nullptr, nullptr, nullptr, 0};
	std::vector<int64_t> score_raw_;
	std::vector<uint8_t> flags_raw_;

	// --- operators (thin wrappers; every N-point loop is a kernel) ---
	Score finish_score(int64_t count, double value_sum, double shared, int64_t best_inliers) const {
		Score sc;
		// scoring_function_with_compound_model.h:105-106: zero score when the candidate cannot reach the best
		if ((uint64_t)(count + 1) < (uint64_t)best_inliers) return sc;
		sc.inliers = count;
		sc.value = value_sum;
		if (!models_.empty()) sc.value -= std::pow(shared, s_.exponent); // :110-121
		return sc;
	}
	// The compound preference vector and the instances' preference vectors never leave the device: ctx->cpref [N] is the
	// element-wise maximum of rows 0 .. |models_|-1 of ctx->pref_rows (updateCompoundModel, progressive_x.h:597-624).
	const double *compound_dev() const { return models_.empty() ? nullptr : ctx_->cpref.as<double>(); }
	double *pref_row(size_t k) const { return ctx_->pref_rows.as<double>() + k * (size_t)N_; }
	int reserve_round_buffers() { // once per run: all-zero compound vector (progressive_x.h:262), the preference table
		PXB_TRY(ctx_->cpref.reserve(sizeof(double) * (size_t)N_));
		PXB_TRY(ctx_->pref2.reserve(sizeof(double) * (size_t)N_));
		PXB_TRY(ctx_->pref_rows.reserve(sizeof(double) * (size_t)N_ * (PXB_MAX_ROUNDS + 1)));
		PXB_CUDA(cudaMemsetAsync(ctx_->cpref.ptr, 0, sizeof(double) * (size_t)N_, ctx_->stream));
		return PXB_OK;
	}
	int publish_compound() { // stream ordered, no wait: the next chain that scores against the vector comes after it
		PXB_TRY(launch_compound_max(ctx_, ctx_->pref_rows.as<double>(), (int64_t)models_.size(), N_, ctx_->cpref.as<double>()));
		if (sharded()) { // the other ranks score against the same compound preference vector
			ShardMsg h{};
			h.op = kShardCompound;
			PXB_TRY(shard_send(h, nullptr, 0));
			PXB_TRY(shard_broadcast(ctx_, ctx_->cpref.ptr, sizeof(double) * (size_t)N_, 0));
		}
		return PXB_OK;
	}
	// PEARL rejected instance l of S: the rows behind it move down (rare; rows do not overlap)
	int erase_pref_row(size_t l, size_t S) {
		for (size_t k = l + 1; k < S; ++k)
			PXB_CUDA(cudaMemcpyAsync(pref_row(k - 1), pref_row(k), sizeof(double) * (size_t)N_, cudaMemcpyDeviceToDevice, ctx_->stream));
		return PXB_OK;
	}
	// scores K models that already sit in ctx->models (device) and brings (count, value, shared) back in ONE copy (the
	// three arrays are contiguous on the device): no sync here, sync_scores() waits and unpacks
	int score_device_models(int64_t K, double T2, std::vector<int64_t> &cnt, std::vector<double> &val, std::vector<double> &shr) {
		cnt.resize(K);
		val.resize(K);
		shr.resize(K);
		pending_scores_ = {nullptr, nullptr, nullptr, 0};
		if (K == 0) return PXB_OK;
		PXB_TRY(ctx_->outB.reserve(sizeof(int64_t) * (size_t)K * 3));
		int64_t *d_cnt = ctx_->outB.as<int64_t>();
		double *d_val = reinterpret_cast<double *>(d_cnt + K), *d_shr = d_val + K;
		PXB_TRY(launch_score_compound(ctx_, ctx_->models.as<double>(), K, T2, compound_dev(), d_cnt, d_val, d_shr));
		score_raw_.resize((size_t)K * 3);
		PXB_TRY(api_d2h(ctx_, score_raw_.data(), d_cnt, sizeof(int64_t) * (size_t)K * 3));
		pending_scores_ = {&cnt, &val, &shr, K};
		return PXB_OK;
	}
	int sync_scores() {
		PXB_TRY(api_sync(ctx_));
		if (pending_scores_.K > 0) {
			const int64_t K = pending_scores_.K;
			std::memcpy(pending_scores_.cnt->data(), score_raw_.data(), sizeof(int64_t) * (size_t)K);
			std::memcpy(pending_scores_.val->data(), score_raw_.data() + K, sizeof(double) * (size_t)K);
			std::memcpy(pending_scores_.shr->data(), score_raw_.data() + 2 * K, sizeof(double) * (size_t)K);
			pending_scores_.K = 0;
		}
		return PXB_OK;
	}
	// ---- hypothesis-block sharding over NCCL (pxb_ctx_set_shard) ------------------------------------------
	// Rank 0 (the coordinator) runs the whole control flow; the other ranks serve "solve and score your slice of this block"
	// requests. A hypothesis' models / flags / (count, value, shared) do not depend on which GPU or batch evaluated them
	// (fixed reduction topology, DESIGN.md "Summation order"), so the sharded run takes exactly the decisions of the
	// single-GPU run for the same seed. Everything that is not a block refill (LO, IRLS, PEARL: N x <= 50 work) stays on
	// the coordinator -- "replicas only" is not even needed for it.
	enum ShardOp : int32_t { kShardBlock = 1, kShardCompound = 2, kShardDone = 3 };
	struct ShardMsg { // 64-byte header in front of the sample indices of a block
		int32_t op, has_compound;
		int64_t want;
		double T2;
		int64_t result;
		int64_t pad[4];
	};
	struct ShardRecord { // one rank's slice of a block inside the all-gather buffer (fields 16-byte aligned)
		size_t models, cnt, val, shr, n, sv, mv, bytes;
	};
	ShardRecord shard_record(size_t S) const {
		auto up16 = [](size_t v) { return (v + 15) & ~size_t(15); };
		ShardRecord r;
		size_t at = 0;
		r.models = at, at += up16(sizeof(double) * S * maxsol_ * ms_);
		r.cnt = at, at += up16(sizeof(int64_t) * S * maxsol_);
		r.val = at, at += up16(sizeof(double) * S * maxsol_);
		r.shr = at, at += up16(sizeof(double) * S * maxsol_);
		r.n = at, at += up16(sizeof(int32_t) * S);
		r.sv = at, at += up16(S);
		r.mv = at, at += up16(S);
		r.bytes = at;
		return r;
	}
	size_t shard_block_cap() const { // largest block a refill may ask for: identical on every rank
		size_t B = kReplayBlock * (size_t)ctx_->shard_world;
		if (const char *e = getenv("PXB_BLOCK_SIZE")) {
			const long v = atol(e);
			if (v >= 1) B = (size_t)v;
		}
		return B;
	}
	size_t shard_msg_bytes() const { return sizeof(ShardMsg) + sizeof(int64_t) * shard_block_cap() * (size_t)m_; }
	// solves and scores this rank's slice of the block whose samples sit behind the header in ctx->shard_msg, then gathers
	int shard_compute_and_gather(size_t want, double T2, bool has_compound) {
		const size_t G = (size_t)ctx_->shard_world, r = (size_t)ctx_->shard_rank;
		const size_t S = (want + G - 1) / G, lo = std::min(want, r * S), ks = std::min(want, lo + S) - lo;
		const ShardRecord rec = shard_record(S);
		PXB_TRY(ctx_->shard_rec.reserve(rec.bytes * G));
		char *mine = ctx_->shard_rec.as<char>() + rec.bytes * r;
		PXB_CUDA(cudaMemsetAsync(mine, 0, rec.bytes, ctx_->stream));
		if (ks > 0) {
			const int64_t *smp = reinterpret_cast<const int64_t *>(ctx_->shard_msg.as<char>() + sizeof(ShardMsg)) + lo * m_;
			double *models = reinterpret_cast<double *>(mine + rec.models);
			PXB_TRY(launch_solve_minimal(ctx_, smp, (int64_t)ks, models, reinterpret_cast<int32_t *>(mine + rec.n),
			                             reinterpret_cast<uint8_t *>(mine + rec.sv), reinterpret_cast<uint8_t *>(mine + rec.mv)));
			PXB_TRY(launch_score_compound(ctx_, models, (int64_t)(ks * maxsol_), T2, has_compound ? ctx_->cpref.as<double>() : nullptr,
			                              reinterpret_cast<int64_t *>(mine + rec.cnt), reinterpret_cast<double *>(mine + rec.val),
			                              reinterpret_cast<double *>(mine + rec.shr)));
		}
		return shard_allgather(ctx_, ctx_->shard_rec.ptr, rec.bytes);
	}
	int shard_send(const ShardMsg &h, const int64_t *samples, size_t n_samples) { // coordinator: header (+ samples) to all ranks
		const size_t bytes = shard_msg_bytes();
		PXB_TRY(ctx_->shard_msg.reserve(bytes));
		shard_host_.resize(bytes);
		std::memcpy(shard_host_.data(), &h, sizeof(h));
		if (n_samples) std::memcpy(shard_host_.data() + sizeof(h), samples, sizeof(int64_t) * n_samples);
		PXB_TRY(api_h2d(ctx_, ctx_->shard_msg.ptr, shard_host_.data(), sizeof(h) + sizeof(int64_t) * n_samples));
		return shard_broadcast(ctx_, ctx_->shard_msg.ptr, bytes, 0);
	}
	int solve_and_score_sharded(const std::vector<int64_t> &samples, size_t want, double T2, std::vector<double> &models,
	                            std::vector<int32_t> &n, std::vector<uint8_t> &sv, std::vector<uint8_t> &mv,
	                            std::vector<int64_t> &cnt, std::vector<double> &val, std::vector<double> &shr) {
		Scoped t(prof_, "refill(sharded solve+score)");
		if (want > shard_block_cap()) {
			set_error("sharded block of %zu samples exceeds the agreed capacity %zu", want, shard_block_cap());
			return PXB_ERR_STATE;
		}
		ShardMsg h{};
		h.op = kShardBlock;
		h.has_compound = models_.empty() ? 0 : 1;
		h.want = (int64_t)want;
		h.T2 = T2;
		PXB_TRY(shard_send(h, samples.data(), want * (size_t)m_));
		PXB_TRY(shard_compute_and_gather(want, T2, h.has_compound != 0));
		const size_t G = (size_t)ctx_->shard_world, S = (want + G - 1) / G;
		const ShardRecord rec = shard_record(S);
		PXB_TRY(ctx_->reserve_pinned(rec.bytes * G));
		PXB_CUDA(cudaMemcpyAsync(ctx_->pinned, ctx_->shard_rec.ptr, rec.bytes * G, cudaMemcpyDeviceToHost, ctx_->stream));
		PXB_TRY(api_sync(ctx_));
		const size_t KS = want * maxsol_;
		cnt.resize(KS);
		val.resize(KS);
		shr.resize(KS);
		for (size_t r = 0; r < G; ++r) { // slices are contiguous in sample order: rank r holds samples [r S, r S + ks)
			const size_t lo = std::min(want, r * S), ks = std::min(want, lo + S) - lo;
			if (!ks) continue;
			const char *src = reinterpret_cast<const char *>(ctx_->pinned) + rec.bytes * r;
			std::memcpy(models.data() + lo * maxsol_ * ms_, src + rec.models, sizeof(double) * ks * maxsol_ * ms_);
			std::memcpy(cnt.data() + lo * maxsol_, src + rec.cnt, sizeof(int64_t) * ks * maxsol_);
			std::memcpy(val.data() + lo * maxsol_, src + rec.val, sizeof(double) * ks * maxsol_);
			std::memcpy(shr.data() + lo * maxsol_, src + rec.shr, sizeof(double) * ks * maxsol_);
			std::memcpy(n.data() + lo, src + rec.n, sizeof(int32_t) * ks);
			std::memcpy(sv.data() + lo, src + rec.sv, ks);
			std::memcpy(mv.data() + lo, src + rec.mv, ks);
		}
		++shard_blocks_;
		return PXB_OK;
	}
	int shard_worker(); // ranks > 0: serve requests until the coordinator sends the result
	int shard_finish(int result); // coordinator: result (model count or a negative status) + payload to all ranks
	std::vector<unsigned char> shard_host_;
	size_t shard_blocks_ = 0;

	int score_models(const double *models, int64_t K, double T2, std::vector<int64_t> &cnt, std::vector<double> &val,
	                 std::vector<double> &shr) {
		cnt.resize(K);
		val.resize(K);
		shr.resize(K);
		if (K == 0) return PXB_OK;
		Scoped t(prof_, "score_models");
		PXB_TRY(ctx_->models.reserve(sizeof(double) * (size_t)K * ms_));
		PXB_TRY(api_h2d(ctx_, ctx_->models.ptr, models, sizeof(double) * (size_t)K * ms_));
		PXB_TRY(score_device_models(K, T2, cnt, val, shr));
		return sync_scores();
	}
	// minimal solves of a block of samples + their scores in ONE round trip (GCRANSAC.h:296-447, block form)
	int solve_and_score(const std::vector<int64_t> &samples, size_t want, double T2, std::vector<double> &models,
	                    std::vector<int32_t> &n, std::vector<uint8_t> &sv, std::vector<uint8_t> &mv, std::vector<int64_t> &cnt,
	                    std::vector<double> &val, std::vector<double> &shr) {
		if (sharded()) return solve_and_score_sharded(samples, want, T2, models, n, sv, mv, cnt, val, shr);
		Scoped t(prof_, "refill(solve+score)");
		const size_t K = want;
		// models, scores and flags of the block sit in one device record (same layout as a sharded slice): ONE copy back.
		// inputs: sample indices [K m] | plane-and-parallax homography [9]
		const ShardRecord rec = shard_record(K);
		const size_t b_smp = sizeof(int64_t) * K * m_, in_bytes = b_smp + sizeof(double) * 9;
		PXB_TRY(ctx_->chain_par.reserve(in_bytes));
		PXB_TRY(ctx_->shard_rec.reserve(rec.bytes));
		const bool slots = chain_slots(in_bytes, rec.bytes);
		chain_tmp_in_.resize(in_bytes);
		unsigned char *hin = slots ? ctx_->chain_in : chain_tmp_in_.data();
		std::memcpy(hin, samples.data(), b_smp);
		std::memcpy(hin + b_smp, s_.pp_H, sizeof(double) * 9);
		pack_host_.resize(rec.bytes);
		unsigned char *hout = slots ? ctx_->chain_out : pack_host_.data();
		auto enqueue = [&]() -> int {
			char *par = ctx_->chain_par.as<char>();
			const int64_t *d_smp = reinterpret_cast<const int64_t *>(par);
			const double *d_ppH = reinterpret_cast<const double *>(par + b_smp);
			char *rc_dev = ctx_->shard_rec.as<char>();
			double *d_models = reinterpret_cast<double *>(rc_dev + rec.models);
			int32_t *d_n = reinterpret_cast<int32_t *>(rc_dev + rec.n);
			uint8_t *d_sv = reinterpret_cast<uint8_t *>(rc_dev + rec.sv), *d_mv = reinterpret_cast<uint8_t *>(rc_dev + rec.mv);
			if (slots)
				PXB_CUDA(cudaMemcpyAsync(par, hin, in_bytes, cudaMemcpyHostToDevice, ctx_->stream));
			else
				PXB_TRY(api_h2d(ctx_, par, hin, in_bytes));
			PXB_CUDA(cudaMemsetAsync(rc_dev, 0, rec.bytes, ctx_->stream));
			if (s_.plane_parallax)
				PXB_TRY(launch_solve_plane_parallax(ctx_, d_smp, (int64_t)K, d_ppH, d_models, d_n, d_sv, d_mv));
			else
				PXB_TRY(launch_solve_minimal(ctx_, d_smp, (int64_t)K, d_models, d_n, d_sv, d_mv));
			PXB_TRY(launch_score_compound(ctx_, d_models, (int64_t)(K * maxsol_), T2, compound_dev(),
			                              reinterpret_cast<int64_t *>(rc_dev + rec.cnt), reinterpret_cast<double *>(rc_dev + rec.val),
			                              reinterpret_cast<double *>(rc_dev + rec.shr)));
			if (slots)
				PXB_CUDA(cudaMemcpyAsync(hout, rc_dev, rec.bytes, cudaMemcpyDeviceToHost, ctx_->stream));
			else
				PXB_TRY(api_d2h(ctx_, hout, rc_dev, rec.bytes));
			return PXB_OK;
		};
		if (slots)
			PXB_TRY(run_chain(chain_key(kChainRefill, {buffers_key(), (uint64_t)K, (uint64_t)s_.plane_parallax, bits(T2)}), enqueue));
		else
			PXB_TRY(enqueue());
		PXB_TRY(api_sync(ctx_));
		const unsigned char *h = hout;
		const size_t KS = K * maxsol_;
		cnt.resize(KS);
		val.resize(KS);
		shr.resize(KS);
		std::memcpy(models.data(), h + rec.models, sizeof(double) * KS * ms_);
		std::memcpy(cnt.data(), h + rec.cnt, sizeof(int64_t) * KS);
		std::memcpy(val.data(), h + rec.val, sizeof(double) * KS);
		std::memcpy(shr.data(), h + rec.shr, sizeof(double) * KS);
		std::memcpy(n.data(), h + rec.n, sizeof(int32_t) * K);
		std::memcpy(sv.data(), h + rec.sv, K);
		std::memcpy(mv.data(), h + rec.mv, K);
		return PXB_OK;
	}
	int inliers_of(const double *model, double T2, std::vector<int64_t> &out) {
		out.resize(N_);
		int64_t n = 0;
		Scoped t(prof_, "inliers_of");
		PXB_TRY(pxb_inliers(ctx_, model, T2, out.data(), &n));
		out.resize(n);
		return PXB_OK;
	}
	int launch_fit_family(int P, const int32_t *d_off, const int32_t *d_idx, const double *d_w, double *models_dev, int32_t *ok_dev);
	std::vector<unsigned char> pack_host_, chain_tmp_in_;
	// ---- replayable device chains (CUDA graphs; see pxb_ctx::ChainGraph) ---------------------------------------------
	enum ChainKind : uint64_t { kChainRefill = 1, kChainLo = 2, kChainTail = 3, kChainPearl = 4, kChainFinish = 5 };
	static uint64_t mix_key(std::initializer_list<uint64_t> parts) {
		uint64_t h = 0x9E3779B97F4A7C15ull;
		for (uint64_t v : parts) {
			h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
			h *= 0xBF58476D1CE4E5B9ull;
			h ^= h >> 29;
		}
		return h | 1;
	}
	static uint64_t bits(double v) {
		uint64_t u;
		std::memcpy(&u, &v, sizeof(u));
		return u;
	}
	static uint64_t bits(const void *p) { return (uint64_t)reinterpret_cast<uintptr_t>(p); }
	// the buffers every chain touches: a reallocation of any of them changes the signature (and forces a new capture)
	uint64_t buffers_key() const {
		const pxb::Points &p = ctx_->pts;
		return mix_key({bits(p.soa), bits(p.aos), bits(p.f32n), bits(p.q), bits(p.norm), (uint64_t)p.N, (uint64_t)p.type,
		                bits(ctx_->idx.ptr), bits(ctx_->pack.ptr), bits(ctx_->screen.ptr), bits(ctx_->partials.ptr),
		                bits(ctx_->outA.ptr), bits(ctx_->mask.ptr), bits(ctx_->pref2.ptr), bits(ctx_->shard_rec.ptr),
		                bits(ctx_->staging.ptr), bits(ctx_->labels.ptr), bits(ctx_->outB.ptr), bits(ctx_->outC.ptr),
		                bits(ctx_->outD.ptr), bits(ctx_->cpref.ptr), bits(ctx_->chain_par.ptr), bits(compound_dev())});
	}
	bool chain_slots(size_t in_bytes, size_t out_bytes) const {
		return ctx_->chain_in && in_bytes <= kChainInBytes && out_bytes <= kChainOutBytes;
	}
	// Runs `enqueue` (a fixed sequence of asynchronous copies and launches on the context's stream that only depends on
	// `key`): plainly the first time a key is seen (buffers get allocated), captured into a CUDA graph the second time,
	// replayed with one cudaGraphLaunch from then on. PXB_GRAPHS=0 disables capturing.
	// PXB_PROFILE=2: the GPU time of every chain (CUDA events around its launch, waited for at once -- the caller waits
	// right afterwards anyway) goes into the phase table beside the host's wall time for the same call
	template <class Enqueue> int run_chain(uint64_t key, Enqueue &&enqueue) {
		static const bool gpu_timed = getenv("PXB_PROFILE") && atoi(getenv("PXB_PROFILE")) == 2;
		if (!gpu_timed) return run_chain_untimed(key, enqueue);
		static thread_local cudaEvent_t ev[2] = {nullptr, nullptr};
		if (!ev[0]) {
			PXB_CUDA(cudaEventCreate(&ev[0]));
			PXB_CUDA(cudaEventCreate(&ev[1]));
		}
		PXB_CUDA(cudaEventRecord(ev[0], ctx_->stream));
		PXB_TRY(run_chain_untimed(key, enqueue));
		PXB_CUDA(cudaEventRecord(ev[1], ctx_->stream));
		PXB_CUDA(cudaEventSynchronize(ev[1]));
		float ms = 0;
		cudaEventElapsedTime(&ms, ev[0], ev[1]);
		static const char *names[] = {"gpu: ?", "gpu: refill chain", "gpu: lo_step chain", "gpu: tail_step chain", "gpu: pearl chain",
		                              "gpu: finish chain"};
		prof_.add(names[chain_kind_ <= 5 ? chain_kind_ : 0], ms);
		return PXB_OK;
	}
	uint64_t chain_kind_ = 0; // kind of the chain keyed last (names the row of the PXB_PROFILE=2 table)
	uint64_t chain_key(ChainKind kind, std::initializer_list<uint64_t> parts) {
		chain_kind_ = kind;
		uint64_t h = mix_key(parts);
		return mix_key({(uint64_t)kind, h});
	}
	template <class Enqueue> int run_chain_untimed(uint64_t key, Enqueue &&enqueue) {
		static const bool enabled = !(getenv("PXB_GRAPHS") && atoi(getenv("PXB_GRAPHS")) == 0);
		if (!enabled) return enqueue();
		auto &cache = ctx_->chain_graphs;
		pxb_ctx::ChainGraph *e = nullptr;
		for (auto &g : cache)
			if (g.key == key) e = &g;
		const uint64_t now = ++ctx_->chain_tick;
		if (e && e->state == 1) {
			e->last_use = now;
			ctx_->launches += e->launches;
			PXB_CUDA(cudaGraphLaunch(e->exec, ctx_->stream));
			return PXB_OK;
		}
		if (e && e->state == 2) return enqueue();
		if (!e) { // first sight: run it plainly (this is also where scratch buffers grow), remember the signature
			if (cache.size() >= 48) { // evict the least recently used entry
				size_t victim = 0;
				for (size_t i = 1; i < cache.size(); ++i)
					if (cache[i].last_use < cache[victim].last_use) victim = i;
				if (cache[victim].exec) cudaGraphExecDestroy(cache[victim].exec);
				cache.erase(cache.begin() + victim);
			}
			pxb_ctx::ChainGraph g;
			g.key = key;
			g.last_use = now;
			cache.push_back(g);
			return enqueue();
		}
		// second sight: capture
		e->last_use = now;
		const int64_t launches0 = ctx_->launches;
		if (cudaStreamBeginCapture(ctx_->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
			(void)cudaGetLastError();
			e->state = 2;
			return enqueue();
		}
		const int rc = enqueue();
		cudaGraph_t graph = nullptr;
		const cudaError_t ce = cudaStreamEndCapture(ctx_->stream, &graph);
		if (rc != PXB_OK || ce != cudaSuccess || !graph) {
			(void)cudaGetLastError();
			if (graph) cudaGraphDestroy(graph);
			e->state = 2;
			ctx_->launches = launches0;
			return enqueue(); // nothing was executed during the failed capture
		}
		cudaGraphExec_t exec = nullptr;
		const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
		cudaGraphDestroy(graph);
		if (ie != cudaSuccess || !exec) {
			(void)cudaGetLastError();
			e->state = 2;
			ctx_->launches = launches0;
			return enqueue();
		}
		e->exec = exec;
		e->state = 1;
		e->launches = (int)(ctx_->launches - launches0);
		PXB_CUDA(cudaGraphLaunch(e->exec, ctx_->stream));
		return PXB_OK;
	}
	int64_t smooth_edges_ = -1; // directed non-loop entries of the neighbour lists (counted once per graph)
	// Estimator::nonMinimalSampleSize(): H four-point 4, F bundle-adjustment solver 7, PnP bundle adjustment 4,
	// vanishing point / 2D line: the minimal solver doubles as the non-minimal one, 2
	size_t non_minimal_sample_size() const {
		return s_.type == PXB_MODEL_FUNDAMENTAL ? 7 : (s_.type >= PXB_MODEL_VANISHING_POINT ? 2 : 4);
	}
	// Estimator::isWeightingApplicable(); LinearModelSolver<2> accepts weights and never reads them
	bool weighting_applicable() const { return s_.type != PXB_MODEL_PNP && s_.type != PXB_MODEL_LINE2D; }
	// H/F solvers read weights_[row of the gathered sample]; the VP solver reads weights_[point index] (:203)
	bool weights_by_point() const { return s_.type == PXB_MODEL_VANISHING_POINT; }
	int model_is_valid(std::vector<double> &model, size_t slot_valid, const int64_t *sample, uint64_t nested_seed, bool &valid,
	                   bool &updated);
	int apply_degensac(std::vector<double> &model, const int64_t *sample, uint64_t nested_seed, bool &valid, bool &updated);
	size_t iteration_number_for(size_t inliers, double log_probability) const;
	int propose(uint64_t round_seed, std::vector<double> &model_out, bool &found);
	struct LoStep { // results of one local-optimisation step
		int64_t inliers = 0;
		std::vector<double> fitted, val, shr;
		std::vector<int64_t> cnt;
		std::vector<int32_t> ok;
	};
	int lo_step(const double *model, uint64_t lo_seed, uint64_t event, double T2, LoStep &out);
	uint64_t lo_events_ = 0; // LO labellings of the current proposal (selects the sampling substreams)
	int local_optimization(uint64_t lo_seed, std::vector<double> &best_model, Score &best_score, double T2);
	struct TailStep { // results of one least-squares step
		int64_t inliers = 0, cnt = 0;
		std::vector<double> fitted;
		double val = 0, shr = 0;
		int32_t ok = 0;
	};
	int tail_step(const double *model, bool weighted, double T2, TailStep &out);
	int proposal_finish(const double *model, std::vector<int64_t> &inliers, double &tanimoto);
	bool defer_inliers_ = false; // the outer loop reads the proposal's inliers in proposal_finish (one round trip for both)
	int pearl();
	size_t predicted_unseen_inliers(size_t iterations, size_t compound_inliers) const;
};

int Driver::build_graph(double radius, int k) {
	PXB_CUDA(cudaSetDevice(ctx_->device));
	std::vector<int32_t> nbr((size_t)N_ * k), deg((size_t)N_);
	PXB_TRY(ctx_->idx.reserve(sizeof(int32_t) * (size_t)N_ * (k + 1)));
	int32_t *d_nbr = ctx_->idx.as<int32_t>(), *d_deg = d_nbr + (size_t)N_ * k;
	PXB_TRY(launch_knn_graph(ctx_, radius, k, d_nbr, d_deg));
	PXB_CUDA(cudaMemcpyAsync(nbr.data(), d_nbr, sizeof(int32_t) * nbr.size(), cudaMemcpyDeviceToHost, ctx_->stream));
	PXB_CUDA(cudaMemcpyAsync(deg.data(), d_deg, sizeof(int32_t) * deg.size(), cudaMemcpyDeviceToHost, ctx_->stream));
	PXB_TRY(ctx_wait(ctx_));
	graph_.off.assign((size_t)N_ + 1, 0);
	for (int64_t i = 0; i < N_; ++i) graph_.off[i + 1] = graph_.off[i] + deg[i];
	graph_.idx.resize((size_t)graph_.off[N_]);
	for (int64_t i = 0; i < N_; ++i)
		for (int t = 0; t < deg[i]; ++t) graph_.idx[graph_.off[i] + t] = nbr[(size_t)i * k + t];
	// the graph is fixed from here on: register its arrays so that the skeleton caches need not re-hash them per cut
	ctx_->trusted_csr_key = 0;
	ctx_->trusted_csr_key = csr_content_key(ctx_, N_, graph_.off.data(), graph_.idx.data()) | 1;
	ctx_->trusted_csr_off = graph_.off.data();
	ctx_->trusted_csr_idx = graph_.idx.data();
	return PXB_OK;
}

// the non-minimal solver of the estimator family on device-resident CSR index lists (asynchronous)
int Driver::launch_fit_family(int P, const int32_t *d_off, const int32_t *d_idx, const double *d_w, double *models_dev,
                              int32_t *ok_dev) {
	switch (s_.type) {
	case PXB_MODEL_HOMOGRAPHY: return launch_fit_h(ctx_, P, d_off, d_idx, d_w, models_dev, ok_dev);
	case PXB_MODEL_FUNDAMENTAL: return launch_fit_f(ctx_, P, d_off, d_idx, d_w, models_dev, ok_dev);
	case PXB_MODEL_PNP: // PerspectiveNPointEstimator::isWeightingApplicable() is false: weights never reach the solver
		return launch_fit_pnp(ctx_, P, d_off, d_idx, models_dev, ok_dev);
	case PXB_MODEL_VANISHING_POINT: return launch_fit_vp(ctx_, P, d_off, d_idx, d_w, models_dev, ok_dev);
	default: return launch_fit_line(ctx_, P, d_off, d_idx, models_dev, ok_dev);
	}
}

// Estimator::isValidModel as called from GCRANSAC::run (:441-447) with threshold_ = truncated_threshold.
//   H   determinant test (already evaluated by the solver kernel)          homography_estimator.h:326-342
//   F   at least max(7, half) of the Sampson inliers must also be inliers under the symmetric epipolar distance
//       (fundamental_estimator.h:268-325), then DEGENSAC (:341-572, Driver::apply_degensac)
//   PnP always true                                                         perspective_n_point_estimator.h:210-218
int Driver::model_is_valid(std::vector<double> &model, size_t slot_valid, const int64_t *sample, uint64_t nested_seed,
                           bool &valid, bool &updated) {
	valid = slot_valid != 0;
	updated = false;
	if (s_.type != PXB_MODEL_FUNDAMENTAL || !valid) return PXB_OK;
	const double tt = 3.0 / 2.0 * s_.threshold, T2 = tt * tt;
	PXB_TRY(ctx_->models.reserve(sizeof(double) * 9));
	PXB_TRY(ctx_->outB.reserve(sizeof(long long) * 2));
	PXB_CUDA(cudaMemcpyAsync(ctx_->models.ptr, model.data(), sizeof(double) * 9, cudaMemcpyHostToDevice, ctx_->stream));
	PXB_TRY(launch_f_sym_count(ctx_, ctx_->models.as<double>(), T2, tt * tt, ctx_->outB.as<long long>()));
	long long c[2];
	PXB_CUDA(cudaMemcpyAsync(c, ctx_->outB.ptr, sizeof(c), cudaMemcpyDeviceToHost, ctx_->stream));
	PXB_TRY(ctx_wait(ctx_));
	const size_t minimum = std::max<size_t>(7, (size_t)((double)c[0] * s_.sym_epipolar_ratio)); // fundamental_estimator.h:303-304
	valid = (size_t)c[1] >= minimum;
	if (!valid) return PXB_OK;
	// :326-334: seven-point models of the outer estimator go through DEGENSAC
	if (s_.use_degensac && !s_.plane_parallax && s_.rows_host && sample) PXB_TRY(apply_degensac(model, sample, nested_seed, valid, updated));
	return PXB_OK;
}

namespace {
// eigenvector of the smallest eigenvalue of a symmetric 3x3 (cyclic Jacobi)
void smallest_eigenvector3(double A[3][3], double v[3]) {
	double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
	for (int sweep = 0; sweep < 60; ++sweep) {
		if (std::fabs(A[0][1]) + std::fabs(A[0][2]) + std::fabs(A[1][2]) == 0.0) break;
		for (int p = 0; p < 2; ++p)
			for (int q = p + 1; q < 3; ++q) {
				if (A[p][q] == 0.0) continue;
				const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
				const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
				const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
				for (int k = 0; k < 3; ++k) {
					const double akp = A[k][p], akq = A[k][q];
					A[k][p] = c * akp - sn * akq;
					A[k][q] = sn * akp + c * akq;
				}
				for (int k = 0; k < 3; ++k) {
					const double apk = A[p][k], aqk = A[q][k];
					A[p][k] = c * apk - sn * aqk;
					A[q][k] = sn * apk + c * aqk;
				}
				for (int k = 0; k < 3; ++k) {
					const double vkp = V[k][p], vkq = V[k][q];
					V[k][p] = c * vkp - sn * vkq;
					V[k][q] = sn * vkp + c * vkq;
				}
			}
	}
	int b = 0;
	for (int i = 1; i < 3; ++i)
		if (A[i][i] < A[b][b]) b = i;
	for (int i = 0; i < 3; ++i) v[i] = V[i][b];
}
inline void cross3(const double *a, const double *b, double *o) {
	o[0] = a[1] * b[2] - a[2] * b[1];
	o[1] = a[2] * b[0] - a[0] * b[2];
	o[2] = a[0] * b[1] - a[1] * b[0];
}
inline double h_transfer_error2(const double *H, const double *q) { // homography_estimator.h:181-199
	const double t1 = H[0] * q[0] + H[1] * q[1] + H[2], t2 = H[3] * q[0] + H[4] * q[1] + H[5], t3 = H[6] * q[0] + H[7] * q[1] + H[8];
	const double d1 = q[2] - (t1 / t3), d2 = q[3] - (t2 / t3);
	return d1 * d1 + d2 * d2;
}
// fundamental_estimator.h:341-476 (see pxb_h_degenerate_sample in include/pxb200.h)
bool h_degenerate_sample(const double *rows, const int64_t *sample, const double *F, double *Hbest) {
	static const int triplets[15] = {0, 1, 2, 3, 4, 5, 0, 1, 6, 3, 4, 6, 2, 5, 6};
	// epipole = third left singular vector of F (null vector of F^T), scaled to z = 1
	double FFt[3][3];
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) FFt[i][j] = F[3 * i] * F[3 * j] + F[3 * i + 1] * F[3 * j + 1] + F[3 * i + 2] * F[3 * j + 2];
	double e[3];
	smallest_eigenvector3(FFt, e);
	for (int i = 0; i < 3; ++i) e[i] /= e[2];
	e[2] = 1.0;
	const double ex[9] = {0, -e[2], e[1], e[2], 0, -e[0], -e[1], e[0], 0};
	double A[9];
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c) A[3 * r + c] = ex[3 * r] * F[c] + ex[3 * r + 1] * F[3 + c] + ex[3 * r + 2] * F[6 + c];
	const double sq_h_thr = 2.0 * 2.0; // homography_threshold_ = 2.0 (:93)
	bool degenerate = false;
	for (int t = 0; t < 5 && !degenerate; ++t) {
		const int64_t pid[3] = {sample[triplets[3 * t]], sample[triplets[3 * t + 1]], sample[triplets[3 * t + 2]]};
		double M[9], b[3];
		for (int k = 0; k < 3; ++k) {
			const double *q = rows + 4 * pid[k];
			const double x1[3] = {q[0], q[1], 1.0}, x2[3] = {q[2], q[3], 1.0};
			double Ax1[3], c1[3], c2[3];
			for (int r = 0; r < 3; ++r) Ax1[r] = A[3 * r] * x1[0] + A[3 * r + 1] * x1[1] + A[3 * r + 2] * x1[2];
			cross3(x2, Ax1, c1);
			cross3(x2, e, c2);
			b[k] = (c1[0] * c2[0] + c1[1] * c2[1] + c1[2] * c2[2]) / (c2[0] * c2[0] + c2[1] * c2[1] + c2[2] * c2[2]);
			M[3 * k] = x1[0], M[3 * k + 1] = x1[1], M[3 * k + 2] = x1[2];
		}
		// M^-1 b by cofactors
		const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
		const double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
		const double inv[9] = {c00 / det, (M[2] * M[7] - M[1] * M[8]) / det, (M[1] * M[5] - M[2] * M[4]) / det,
		                       c01 / det, (M[0] * M[8] - M[2] * M[6]) / det, (M[2] * M[3] - M[0] * M[5]) / det,
		                       c02 / det, (M[1] * M[6] - M[0] * M[7]) / det, (M[0] * M[4] - M[1] * M[3]) / det};
		double mb[3];
		for (int r = 0; r < 3; ++r) mb[r] = inv[3 * r] * b[0] + inv[3 * r + 1] * b[1] + inv[3 * r + 2] * b[2];
		double Hm[9];
		for (int r = 0; r < 3; ++r)
			for (int c = 0; c < 3; ++c) Hm[3 * r + c] = A[3 * r + c] - e[r] * mb[c];
		size_t inlier_number = 3;
		for (int i = 0; i < 7; ++i) {
			const int64_t idx = sample[i];
			if (idx == pid[0] || idx == pid[1] || idx == pid[2]) continue;
			if (h_transfer_error2(Hm, rows + 4 * idx) < sq_h_thr) ++inlier_number;
		}
		if (inlier_number >= 5) {
			std::memcpy(Hbest, Hm, sizeof(Hm));
			degenerate = true;
		}
	}
	return degenerate;
}
} // namespace

// FundamentalMatrixEstimator::applyDegensac (gcr/estimators/fundamental_estimator.h:341-572). The seven-point sample is
// H-degenerate when one of five point triplets induces (together with F) a homography that at least five of the seven
// points follow (2 px). Then: the homography is refitted to the F-inliers that follow it, a nested GC-RANSAC with the
// plane-and-parallax solver (that homography + two off-plane correspondences) runs on all points, and its model replaces
// F when it has more inliers. Differences from the reference: the nested run is capped at 5000 iterations (the
// reference's cap is SIZE_MAX) and draws from the seedable generator.
int Driver::apply_degensac(std::vector<double> &model, const int64_t *sample, uint64_t nested_seed, bool &valid, bool &updated) {
	const double *rows = s_.rows_host;
	const double sq_h_thr = 2.0 * 2.0; // homography_threshold_ = 2.0 (:93)
	double Hbest[9];
	const bool degenerate = h_degenerate_sample(rows, sample, model.data(), Hbest);
	if (!degenerate) return PXB_OK;
	++degensac_degenerate_;
	// the F-inliers that follow the homography
	const double tt = 3.0 / 2.0 * s_.threshold, T2 = tt * tt;
	std::vector<int64_t> f_inliers, h_inliers;
	PXB_TRY(inliers_of(model.data(), T2, f_inliers));
	for (int64_t i : f_inliers)
		if (h_transfer_error2(Hbest, rows + 4 * i) < sq_h_thr) h_inliers.push_back(i);
	if (h_inliers.size() < 4) { // homography_estimator.nonMinimalSampleSize()
		valid = false;
		return PXB_OK;
	}
	// non-minimal homography on them (the H fit kernel reads the same [x1 y1 x2 y2] rows)
	std::vector<int32_t> off = {0, (int32_t)h_inliers.size()}, idx(h_inliers.begin(), h_inliers.end());
	PXB_TRY(ctx_->idx.reserve(sizeof(int32_t) * (off.size() + idx.size()) + 64));
	int32_t *d_off = ctx_->idx.as<int32_t>(), *d_idx = d_off + off.size();
	PXB_TRY(ctx_->models.reserve(sizeof(double) * 9));
	PXB_TRY(ctx_->outA.reserve(sizeof(int32_t)));
	PXB_TRY(api_h2d(ctx_, d_off, off.data(), sizeof(int32_t) * off.size()));
	PXB_TRY(api_h2d(ctx_, d_idx, idx.data(), sizeof(int32_t) * idx.size()));
	PXB_TRY(launch_fit_h(ctx_, 1, d_off, d_idx, nullptr, ctx_->models.as<double>(), ctx_->outA.as<int32_t>()));
	double Hfit[9];
	int32_t ok = 0;
	PXB_TRY(api_d2h(ctx_, Hfit, ctx_->models.ptr, sizeof(Hfit)));
	PXB_TRY(api_d2h(ctx_, &ok, ctx_->outA.ptr, sizeof(ok)));
	PXB_TRY(api_sync(ctx_));
	if (!ok) {
		valid = false;
		return PXB_OK;
	}
	// nested GC-RANSAC: plane-and-parallax minimal solver, eight-point non-minimal solver, plain MSAC scoring
	Settings ns = s_;
	ns.plane_parallax = true;
	std::memcpy(ns.pp_H, Hfit, sizeof(Hfit));
	ns.threshold = tt;          // gcransac.settings.threshold = threshold_ (the outer truncated threshold)
	ns.lambda = 0.0;            // spatial_coherence_weight = 0
	ns.confidence = 0.99;
	ns.max_iters = 5000;
	ns.max_local_optimization_number = 10; // gcr/settings.h:72
	ns.sampler_id = 0;
	ns.napsac = ns.progressive_napsac = false;
	ns.use_degensac = false;    // FundamentalMatrixEstimator<PlaneParallax, EightPoint>(0.0, false)
	ns.sym_epipolar_ratio = 0.0;
	ns.do_logging = false;
	ns.allow_shard = false;
	Driver nested(ctx_, ns);
	std::vector<double> model2;
	bool found = false;
	PXB_TRY(nested.propose(nested_seed, model2, found));
	if (found && nested.proposal_inliers_.size() > f_inliers.size()) {
		model = model2;
		updated = true;
		++degensac_updates_;
	}
	return PXB_OK;
}

// gcr/GCRANSAC.h:158-173
size_t Driver::iteration_number_for(size_t inliers, double log_probability) const {
	const double q = std::pow(static_cast<double>(inliers) / (double)N_, (double)m_);
	const double log2 = std::log(1 - q);
	if (std::fabs(log2) < std::numeric_limits<double>::epsilon()) return std::numeric_limits<size_t>::max();
	const double iter = log_probability / log2;
	return static_cast<size_t>(iter) + 1;
}

// One step of graphCutLocalOptimization (gcr/GCRANSAC.h:806-905) as ONE stream-ordered chain: GCRANSAC::labeling of
// `model` (unary decision for lambda = 0, st-cut otherwise) -> ordered inlier list -> the inner-RANSAC samples (drawn on
// the device, one splitmix64 generator per trial seeded by lo_substream(lo_seed, event, trial) -- k_lo_sample in
// pxb_chain.cu; oracle/px_sequential.py derives the same substreams) -> non-minimal fits -> scores; one packed copy back.
int Driver::lo_step(const double *model, uint64_t lo_seed, uint64_t event, double T2, LoStep &out) {
	Scoped t(prof_, "lo_step");
	const int trials = (int)s_.max_local_optimization_number, limit = 7 * m_; // estimator.inlierLimit()
	const bool cut = s_.lambda > 0 && !graph_.idx.empty(); // st-cut (else the cut decomposes per node: k_lo_unary_cut)
	// inputs: model | seed, event; lists: inliers [N] | off [trials + 1] | idx [max(N, trials * limit)]
	const size_t in_bytes = sizeof(double) * ms_ + 2 * sizeof(uint64_t);
	const size_t n_idx = std::max((size_t)N_, (size_t)trials * limit);
	// packed results: count | max-flow flags [16] | fitted [trials ms] | count / value / shared [trials] | ok [trials]
	const size_t bM = sizeof(double) * (size_t)trials * ms_, bT = sizeof(double) * (size_t)trials;
	const size_t o_flags = 8, o_fit = o_flags + 64, o_cnt = o_fit + bM, o_val = o_cnt + bT, o_shr = o_val + bT, o_ok = o_shr + bT,
	             pack_bytes = o_ok + sizeof(int32_t) * (size_t)trials;
	PXB_TRY(ctx_->chain_par.reserve(in_bytes));
	PXB_TRY(ctx_->outA.reserve((size_t)N_));
	PXB_TRY(ctx_->idx.reserve(sizeof(int32_t) * ((size_t)N_ + trials + 1 + n_idx) + 64));
	PXB_TRY(ctx_->pack.reserve(pack_bytes));
	const bool slots = chain_slots(in_bytes, pack_bytes);
	chain_tmp_in_.resize(in_bytes);
	unsigned char *hin = slots ? ctx_->chain_in : chain_tmp_in_.data();
	std::memcpy(hin, model, sizeof(double) * ms_);
	const uint64_t se[2] = {lo_seed, event};
	std::memcpy(hin + sizeof(double) * ms_, se, sizeof(se));
	pack_host_.resize(pack_bytes);
	unsigned char *hout = slots ? ctx_->chain_out : pack_host_.data();
	auto enqueue = [&]() -> int {
		char *par = ctx_->chain_par.as<char>();
		const double *d_model = reinterpret_cast<const double *>(par);
		const uint64_t *d_se = reinterpret_cast<const uint64_t *>(par + sizeof(double) * ms_);
		int32_t *d_inl = ctx_->idx.as<int32_t>(), *d_off = d_inl + N_, *d_idx = d_off + trials + 1;
		char *pk = ctx_->pack.as<char>();
		if (slots)
			PXB_CUDA(cudaMemcpyAsync(par, hin, in_bytes, cudaMemcpyHostToDevice, ctx_->stream));
		else
			PXB_TRY(api_h2d(ctx_, par, hin, in_bytes));
		uint8_t *d_seg = nullptr;
		int32_t *d_mf_flags = nullptr;
		PXB_CUDA(cudaMemsetAsync(pk, 0, pack_bytes, ctx_->stream)); // (before the labelling: kernel -> kernel edges behind it)
		if (!cut) {
			d_seg = ctx_->outA.as<uint8_t>();
			PXB_TRY(launch_lo_unary_cut(ctx_, d_model, s_.threshold, s_.lambda, d_seg));
		} else {
			PXB_TRY(lo_labeling_enqueue(ctx_, d_model, s_.threshold, s_.lambda, graph_.off.data(), graph_.idx.data(), &d_seg, &d_mf_flags));
		}
		PXB_TRY(launch_flag_compact(ctx_, d_seg, N_, d_inl, reinterpret_cast<int64_t *>(pk), nullptr));
		// (the labelling's scratch is shared with the score kernel's partial sums: take its status words now)
		if (d_mf_flags) PXB_CUDA(cudaMemcpyAsync(pk + o_flags, d_mf_flags, 64, cudaMemcpyDeviceToDevice, ctx_->stream));
		PXB_TRY(launch_lo_sample(ctx_, d_inl, reinterpret_cast<int64_t *>(pk), m_, limit, trials, d_se, d_off, d_idx));
		PXB_TRY(launch_fit_family(trials, d_off, d_idx, nullptr, reinterpret_cast<double *>(pk + o_fit), reinterpret_cast<int32_t *>(pk + o_ok)));
		PXB_TRY(launch_score_compound(ctx_, reinterpret_cast<double *>(pk + o_fit), trials, T2, compound_dev(),
		                              reinterpret_cast<int64_t *>(pk + o_cnt), reinterpret_cast<double *>(pk + o_val),
		                              reinterpret_cast<double *>(pk + o_shr)));
		if (slots)
			PXB_CUDA(cudaMemcpyAsync(hout, pk, pack_bytes, cudaMemcpyDeviceToHost, ctx_->stream));
		else
			PXB_TRY(api_d2h(ctx_, hout, pk, pack_bytes));
		return PXB_OK;
	};
	// (the cooperative k_maxflow over a lazily built skeleton is left out of the graphs; the cluster-resident engine over a
	// cached skeleton is a plain launch)
	uint64_t cut_signature = 0;
	if (slots && (!cut || lo_labeling_capturable(ctx_, graph_.off.data(), graph_.idx.data(), &cut_signature)))
		PXB_TRY(run_chain(chain_key(kChainLo, {buffers_key(), (uint64_t)trials, (uint64_t)limit, bits(T2), bits(s_.threshold), bits(s_.lambda), cut_signature}), enqueue));
	else
		PXB_TRY(enqueue());
	PXB_TRY(api_sync(ctx_));
	const unsigned char *h = hout;
	if (cut) {
		const int32_t *f = reinterpret_cast<const int32_t *>(h + o_flags);
		if (f[7] != 1 || f[6] == 0) {
			set_error("max-flow of the local optimisation did not converge");
			return PXB_ERR_CUDA;
		}
		if (const char *e = getenv("PXB_MF_STATS"))
			if (e[0] == '2')
				fprintf(stderr, "[pxb lo cut] rounds=%d bfs_levels=%d relabel=%.3f ms push=%.3f ms (level scans %.3f ms, votes %.3f ms)\n", f[6], f[8],
				        (double)f[10] * 64 / 1.965e6, (double)f[12] * 64 / 1.965e6, (double)f[13] * 64 / 1.965e6, (double)f[14] * 64 / 1.965e6);
	}
	std::memcpy(&out.inliers, h, sizeof(int64_t));
	out.fitted.assign(reinterpret_cast<const double *>(h + o_fit), reinterpret_cast<const double *>(h + o_fit) + (size_t)trials * ms_);
	out.cnt.assign(reinterpret_cast<const int64_t *>(h + o_cnt), reinterpret_cast<const int64_t *>(h + o_cnt) + trials);
	out.val.assign(reinterpret_cast<const double *>(h + o_val), reinterpret_cast<const double *>(h + o_val) + trials);
	out.shr.assign(reinterpret_cast<const double *>(h + o_shr), reinterpret_cast<const double *>(h + o_shr) + trials);
	out.ok.assign(reinterpret_cast<const int32_t *>(h + o_ok), reinterpret_cast<const int32_t *>(h + o_ok) + trials);
	return PXB_OK;
}

// gcr/GCRANSAC.h:781-911. The <= 50 inner-RANSAC trials of one graph cut are independent given the cut: they are drawn,
// fitted and scored in one device chain (lo_step) and the max_score bookkeeping is replayed in trial order.
int Driver::local_optimization(uint64_t lo_seed, std::vector<double> &best_model, Score &best_score, double T2) {
	const size_t inlier_limit = 7 * (size_t)m_; // estimator.inlierLimit()
	Score max_score = best_score;
	std::vector<double> lo_model = best_model;
	LoStep st;
	++lo_number_; // GCRANSAC.h:806 -- on top of the caller's increment (:486 / :535): the reference's statistic counts a run twice
	while (++graph_cut_number_ < s_.max_graph_cut_number) {
		bool updated = false;
		PXB_TRY(lo_step(lo_model.data(), lo_seed, lo_events_++, T2, st));
		size_t n_trials;
		if (inlier_limit < (size_t)st.inliers)
			n_trials = s_.max_local_optimization_number; // samples of inlier_limit points (:823-851)
		else if ((size_t)m_ < (size_t)st.inliers)
			n_trials = 1; // every trial refits the same set: one evaluation is equivalent
		else
			break;
		for (size_t t = 0; t < n_trials; ++t) {
			if (!st.ok[t]) continue; // estimateModelNonminimal failed -> `continue` (:851-855)
			const Score sc = finish_score(st.cnt[t], st.val[t], st.shr[t], max_score.inliers);
			if (max_score.value < sc.value) {
				updated = true;
				max_score = sc;
				lo_model.assign(st.fitted.begin() + t * ms_, st.fitted.begin() + (t + 1) * ms_);
			}
		}
		if (!updated) break;
	}
	if (best_score.value < max_score.value) {
		best_score = max_score;
		best_model = lo_model;
	}
	return PXB_OK;
}

// One least-squares step of the tail of GCRANSAC::run as ONE stream-ordered chain: ordered inlier list of `model`
// (r2 < T2, the list getScore would return) -> [Tukey bisquare weights of `model`, GCRANSAC.h:658-669] -> one non-minimal
// fit on all of them -> score of the fit; one packed copy back. The Tukey weight of a point is zero exactly when the
// point is not an inlier, so the per-point weight array of the reference (:686-688: zero for non-inliers) is the kernel's
// output as it is; H / F solvers read it by ROW of the gathered sample, the vanishing-point solver by point (k_fit_*).
int Driver::tail_step(const double *model, bool weighted, double T2, TailStep &out) {
	Scoped t(prof_, "tail_step");
	const int64_t words = (N_ + 31) / 32;
	const size_t in_bytes = sizeof(double) * ms_;
	// packed results: count | fitted [ms] | count / value / shared of the fit | ok
	const size_t o_fit = 8, o_cnt = o_fit + sizeof(double) * ms_, o_val = o_cnt + 8, o_shr = o_val + 8, o_ok = o_shr + 8,
	             pack_bytes = o_ok + 8;
	PXB_TRY(ctx_->chain_par.reserve(in_bytes));
	PXB_TRY(ctx_->mask.reserve(sizeof(uint32_t) * (size_t)words));
	PXB_TRY(ctx_->idx.reserve(sizeof(int32_t) * ((size_t)N_ + 2) + 64));
	PXB_TRY(ctx_->pack.reserve(pack_bytes));
	if (weighted) PXB_TRY(ctx_->pref2.reserve(sizeof(double) * (size_t)N_));
	const bool slots = chain_slots(in_bytes, pack_bytes);
	chain_tmp_in_.resize(in_bytes);
	unsigned char *hin = slots ? ctx_->chain_in : chain_tmp_in_.data();
	std::memcpy(hin, model, in_bytes);
	pack_host_.resize(pack_bytes);
	unsigned char *hout = slots ? ctx_->chain_out : pack_host_.data();
	auto enqueue = [&]() -> int {
		double *d_model = ctx_->chain_par.as<double>();
		int32_t *d_off = ctx_->idx.as<int32_t>(), *d_inl = d_off + 2;
		char *pk = ctx_->pack.as<char>();
		if (slots)
			PXB_CUDA(cudaMemcpyAsync(d_model, hin, in_bytes, cudaMemcpyHostToDevice, ctx_->stream));
		else
			PXB_TRY(api_h2d(ctx_, d_model, hin, in_bytes));
		PXB_TRY(launch_residual_matrix(ctx_, d_model, 1, T2, nullptr, nullptr, ctx_->mask.as<uint32_t>()));
		PXB_CUDA(cudaMemsetAsync(pk, 0, pack_bytes, ctx_->stream));
		PXB_TRY(launch_mask_compact(ctx_, ctx_->mask.as<uint32_t>(), N_, d_inl, reinterpret_cast<int64_t *>(pk), d_off));
		double *d_w = nullptr;
		if (weighted) {
			d_w = ctx_->pref2.as<double>();
			PXB_TRY(launch_tukey(ctx_, d_model, T2, d_w));
		}
		PXB_TRY(launch_fit_family(1, d_off, d_inl, d_w, reinterpret_cast<double *>(pk + o_fit), reinterpret_cast<int32_t *>(pk + o_ok)));
		PXB_TRY(launch_score_compound(ctx_, reinterpret_cast<double *>(pk + o_fit), 1, T2, compound_dev(),
		                              reinterpret_cast<int64_t *>(pk + o_cnt), reinterpret_cast<double *>(pk + o_val),
		                              reinterpret_cast<double *>(pk + o_shr)));
		if (slots)
			PXB_CUDA(cudaMemcpyAsync(hout, pk, pack_bytes, cudaMemcpyDeviceToHost, ctx_->stream));
		else
			PXB_TRY(api_d2h(ctx_, hout, pk, pack_bytes));
		return PXB_OK;
	};
	if (slots)
		PXB_TRY(run_chain(chain_key(kChainTail, {buffers_key(), (uint64_t)weighted, bits(T2)}), enqueue));
	else
		PXB_TRY(enqueue());
	PXB_TRY(api_sync(ctx_));
	const unsigned char *h = hout;
	std::memcpy(&out.inliers, h, 8);
	out.fitted.assign(reinterpret_cast<const double *>(h + o_fit), reinterpret_cast<const double *>(h + o_fit) + ms_);
	std::memcpy(&out.cnt, h + o_cnt, 8);
	std::memcpy(&out.val, h + o_val, 8);
	std::memcpy(&out.shr, h + o_shr, 8);
	std::memcpy(&out.ok, h + o_ok, 4);
	return PXB_OK;
}

// gcr/GCRANSAC.h:203-628
int Driver::propose(uint64_t round_seed, std::vector<double> &model_out, bool &found) {
	found = false;
	iteration_number_ = graph_cut_number_ = lo_number_ = 0;
	proposal_inliers_.clear();
	const double log_probability = std::log(1.0 - s_.confidence);
	size_t max_iteration = iteration_number_for(1, log_probability);
	const double truncated_threshold = 3.0 / 2.0 * s_.threshold;
	const double T2 = truncated_threshold * truncated_threshold; // :254-255 spelling
	std::unique_ptr<Sampler> main_sampler;
	if (s_.napsac && !graph_.idx.empty())
		main_sampler.reset(new NapsacSampler(round_seed * 2 + 1, &graph_));
	else if (s_.progressive_napsac && !grid_layers_.empty() && (size_t)N_ > (size_t)m_)
		main_sampler.reset(new ProgressiveNapsacSampler(round_seed * 2 + 1, (size_t)m_, (size_t)N_, &grid_layers_, 0.5));
	else if (s_.sampler_id == 1 && (size_t)N_ > (size_t)m_)
		main_sampler.reset(new ProsacSampler(round_seed * 2 + 1, (size_t)m_, (size_t)N_));
	else
		main_sampler.reset(new UniformSampler(round_seed * 2 + 1));
	const uint64_t lo_seed = round_seed * 2 + 2;
	lo_events_ = 0;
	std::vector<size_t> pool(N_);
	std::iota(pool.begin(), pool.end(), 0);

	Score best_score;
	std::vector<double> best_model;

	// ---- block state ----
	// Block size of the replay. Any value gives the same result for the same seed (the sample stream does not depend
	// on it); PXB_BLOCK_SIZE=1 *is* the reference's sequential loop and is what tests/test_gpu_e2e.py compares against.
	size_t B = kReplayBlock, Bmin = 32;
	if (sharded()) B *= (size_t)ctx_->shard_world; // every rank still sees up to kReplayBlock samples per refill
	if (const char *e = getenv("PXB_BLOCK_SIZE")) {
		const long v = atol(e);
		if (v >= 1) B = Bmin = (size_t)v;
	}
	std::vector<int64_t> samples;       // B x m
	std::vector<uint8_t> sampled_ok;    // sampler success per slot
	std::vector<double> blk_models;     // B x maxsol x ms
	std::vector<int32_t> blk_n;
	std::vector<uint8_t> blk_sv, blk_mv;
	std::vector<int64_t> blk_cnt;       // per (slot, solution)
	std::vector<double> blk_val, blk_shr;
	size_t cursor = 0, filled = 0;
	auto refill = [&](size_t want) -> int {
		want = std::max<size_t>(std::min(want, B), std::min(Bmin, B));
		samples.assign(want * m_, 0);
		sampled_ok.assign(want, 0);
		std::vector<size_t> sub(m_);
		for (size_t b = 0; b < want; ++b) {
			sampled_ok[b] = main_sampler->sample(pool, sub.data(), (size_t)m_) ? 1 : 0;
			for (int j = 0; j < m_; ++j) samples[b * m_ + j] = sampled_ok[b] ? (int64_t)sub[j] : (int64_t)j;
		}
		blk_models.assign(want * maxsol_ * ms_, 0.0);
		blk_n.assign(want, 0);
		blk_sv.assign(want, 0);
		blk_mv.assign(want, 0);
		PXB_TRY(solve_and_score(samples, want, T2, blk_models, blk_n, blk_sv, blk_mv, blk_cnt, blk_val, blk_shr));
		cursor = 0;
		filled = want;
		return PXB_OK;
	};

	while (s_.min_iteration_number > iteration_number_ ||
	       iteration_number_ < std::min(max_iteration, s_.max_iters)) {
		bool do_local_optimization = false;
		++iteration_number_;
		// :296-339 select a sample that yields at least one model (<= 100 attempts)
		int unsuccessful = -1;
		size_t slot = SIZE_MAX;
		while (++unsuccessful < (int)s_.max_unsuccessful_model_generations) {
			if (cursor >= filled) {
				const size_t cap = std::min(max_iteration, s_.max_iters);
				const size_t remaining = cap > iteration_number_ ? cap - iteration_number_ + 1 : 1;
				PXB_TRY(refill(std::max(remaining, s_.min_iteration_number)));
			}
			const size_t b = cursor++;
			if (!sampled_ok[b]) continue;        // sampler failure (:300-310)
			if (!blk_sv[b]) continue;            // isValidSample (:314-323)
			if (blk_n[b] > 0) {                  // estimateModel (:326-330)
				slot = b;
				break;
			}
		}
		iteration_number_ += (size_t)unsuccessful; // :341
		if (slot != SIZE_MAX) {
			for (int j = 0; j < blk_n[slot]; ++j) {
				const size_t q = slot * maxsol_ + j;
				Score sc = finish_score(blk_cnt[q], blk_val[q], blk_shr[q], best_score.inliers);
				// :441-447: better score AND Estimator::isValidModel (which may replace the model: DEGENSAC)
				bool model_ok = false, model_updated = false;
				std::vector<double> candidate(blk_models.begin() + q * ms_, blk_models.begin() + (q + 1) * ms_);
				if (best_score.value < sc.value)
					PXB_TRY(model_is_valid(candidate, blk_mv[slot], samples.data() + slot * m_,
					                       round_seed * 7919ull + iteration_number_ * 31ull + (uint64_t)j, model_ok, model_updated));
				if (best_score.value < sc.value && model_ok) {
					if (model_updated) { // :450-457 re-score the replaced model
						std::vector<int64_t> c1;
						std::vector<double> v1, s1;
						PXB_TRY(score_models(candidate.data(), 1, T2, c1, v1, s1));
						sc = finish_score(c1[0], v1[0], s1[0], best_score.inliers);
					}
					best_model = candidate;
					best_score = sc;
					do_local_optimization = iteration_number_ > s_.min_iteration_number_before_lo &&
					                        (size_t)best_score.inliers > (size_t)m_; // :464-465
					max_iteration = iteration_number_for((size_t)best_score.inliers, log_probability);
				}
			}
		}
		if (do_local_optimization) { // :482-503
			++lo_number_;
			PXB_TRY(local_optimization(lo_seed, best_model, best_score, T2));
			max_iteration = iteration_number_for((size_t)best_score.inliers, log_probability);
		}
	}
	if ((size_t)best_score.inliers <= (size_t)m_) return PXB_OK; // :522-528 no model found

	if (lo_number_ == 0) { // :531-544 final LO if it never ran
		++lo_number_;
		PXB_TRY(local_optimization(lo_seed, best_model, best_score, T2));
	}
	// :546-559 the inliers of the best model; :561-590 iterated least squares (GCRANSAC.h:631-759, the single-model branch);
	// :592-618 else one least-squares fit on all inliers. Every step is one device chain (tail_step): the inlier LIST of
	// a model is only needed on the host once, at the very end.
	const bool weighted = weighting_applicable();
	TailStep first, ts;
	PXB_TRY(tail_step(best_model.data(), weighted, T2, first));
	best_score.inliers = first.inliers;
	bool refit_applied = false;
	{
		std::vector<double> model = best_model;
		size_t current = (size_t)first.inliers, iterations = 0;
		bool success = false;
		Score irls_score;
		if (current > (size_t)m_) {
			ts = first;
			bool have_step = true; // `ts` holds the step of `model`
			while (++iterations < s_.max_least_squares_iterations) {
				if (!have_step) PXB_TRY(tail_step(model.data(), weighted, T2, ts));
				have_step = false;
				if (!ts.ok) break;
				const Score sc = finish_score(ts.cnt, ts.val, ts.shr, 0);
				if ((size_t)sc.inliers < (size_t)m_) break;
				if ((size_t)sc.inliers <= current) break;
				model = ts.fitted;
				irls_score = sc;               // = the score getScore gives the polished model (:575-583)
				current = (size_t)sc.inliers;  // = |inliers of the new model| (same r2 < T2 predicate)
			}
			success = iterations > 1;
		}
		if (success && best_score.value < irls_score.value) {
			refit_applied = true;
			best_model = model;
		}
	}
	if (!refit_applied) { // :592-618 one (unweighted) least-squares fit on all inliers of the best model
		if (weighted)
			PXB_TRY(tail_step(best_model.data(), false, T2, ts));
		else
			ts = first; // the first step already was that fit
		if (ts.ok) {
			const Score sc = finish_score(ts.cnt, ts.val, ts.shr, 0);
			if (best_score.value < sc.value) best_model = ts.fitted;
		}
	}
	if (!defer_inliers_) {
		std::vector<int64_t> best_inliers;
		PXB_TRY(inliers_of(best_model.data(), T2, best_inliers));
		proposal_inliers_ = best_inliers; // statistics.inliers (:621)
	}
	model_out = best_model;
	found = true;
	return PXB_OK;
}

// The end of a proposal as ONE stream-ordered chain (one round trip instead of four, nothing but 24 bytes and the inlier
// bit mask crosses PCIe): the inliers of the proposed model (r2 < T2: statistics.inliers, GCRANSAC.h:546-559 / :621) ->
// its preference vector (progx_model.h:84-85, threshold spelled as progressive_x.h:523) into the scratch row ->
// the three sums of the Tanimoto similarity against the compound vector (progressive_x.h:565-591).
int Driver::proposal_finish(const double *model, std::vector<int64_t> &inliers, double &tanimoto) {
	Scoped t(prof_, "proposal_finish");
	const double truncated_threshold = 3.0 / 2.0 * s_.threshold;
	const double T2 = truncated_threshold * truncated_threshold; // GCRANSAC.h:254-255 spelling
	const double T = 9.0 / 4.0 * s_.threshold * s_.threshold;    // progressive_x.h:523 spelling
	const int64_t words = (N_ + 31) / 32;
	const size_t in_bytes = sizeof(double) * ms_;
	const size_t o_sums = ((sizeof(uint32_t) * (size_t)words + 7) / 8) * 8, pack_bytes = o_sums + 3 * sizeof(double);
	PXB_TRY(ctx_->chain_par.reserve(in_bytes));
	PXB_TRY(ctx_->pack.reserve(pack_bytes));
	const bool slots = chain_slots(in_bytes, pack_bytes);
	chain_tmp_in_.resize(in_bytes);
	unsigned char *hin = slots ? ctx_->chain_in : chain_tmp_in_.data();
	std::memcpy(hin, model, in_bytes);
	pack_host_.resize(pack_bytes);
	unsigned char *hout = slots ? ctx_->chain_out : pack_host_.data();
	auto enqueue = [&]() -> int {
		double *d_model = ctx_->chain_par.as<double>();
		char *pk = ctx_->pack.as<char>();
		if (slots)
			PXB_CUDA(cudaMemcpyAsync(d_model, hin, in_bytes, cudaMemcpyHostToDevice, ctx_->stream));
		else
			PXB_TRY(api_h2d(ctx_, d_model, hin, in_bytes));
		PXB_TRY(launch_residual_matrix(ctx_, d_model, 1, T2, nullptr, nullptr, reinterpret_cast<uint32_t *>(pk)));
		PXB_TRY(launch_preference(ctx_, d_model, T, ctx_->pref2.as<double>()));
		PXB_TRY(launch_tanimoto(ctx_, ctx_->pref2.as<double>(), ctx_->cpref.as<double>(), N_, reinterpret_cast<double *>(pk + o_sums)));
		if (slots)
			PXB_CUDA(cudaMemcpyAsync(hout, pk, pack_bytes, cudaMemcpyDeviceToHost, ctx_->stream));
		else
			PXB_TRY(api_d2h(ctx_, hout, pk, pack_bytes));
		return PXB_OK;
	};
	if (slots)
		PXB_TRY(run_chain(chain_key(kChainFinish, {buffers_key(), bits(ctx_->pref2.ptr), bits(T2), bits(T)}), enqueue));
	else
		PXB_TRY(enqueue());
	PXB_TRY(api_sync(ctx_));
	// bit mask -> ascending index list (format conversion of the kernel's output, no arithmetic)
	const uint32_t *w = reinterpret_cast<const uint32_t *>(hout);
	inliers.clear();
	for (int64_t j = 0; j < words; ++j) {
		uint32_t bits32 = w[j];
		while (bits32) {
			inliers.push_back(j * 32 + __builtin_ctz(bits32));
			bits32 &= bits32 - 1;
		}
	}
	double sums[3];
	std::memcpy(sums, hout + o_sums, sizeof(sums));
	tanimoto = sums[0] / (sums[1] + sums[2] - sums[0]); // progressive_x.h:584-585
	return PXB_OK;
}

// px/include/PEARL.h:405-472 (run), :476-555 (labeling), :319-401 (parameterEstimation), :275-315 (rejectInstances)
int Driver::pearl() {
	size_t iteration_number = 0;
	double energy = std::numeric_limits<double>::max(), previous_energy = -1.0;
	bool model_rejected = false, convergence = false;
	bool have_labels = false;
	const double label_cost = (double)s_.min_inliers; // model_complexity_weight(minimum_inlier_number_) (:147)
	const bool smooth = s_.lambda > 0.0 && !graph_.idx.empty();
	if (smooth && smooth_edges_ < 0) { // the undirected edges setNeighbors would insert (self loops skipped, PEARL.h:535)
		smooth_edges_ = 0;
		for (int64_t i = 0; i < N_; ++i)
			for (int32_t e = graph_.off[i]; e < graph_.off[i + 1]; ++e)
				if (graph_.idx[e] != i) ++smooth_edges_;
	}
	// The labels stay on the device for the whole run ([current | previous]); one iteration -- data costs, label sweep,
	// per-instance point lists, residual sums, refits, residual sums of the refits -- is one stream-ordered chain with
	// one packed copy back (PEARL.h:476-555 + :319-401).
	PXB_TRY(ctx_->labels.reserve(sizeof(int32_t) * (size_t)N_ * 2 + 64));
	int32_t *lab_cur = ctx_->labels.as<int32_t>(), *lab_prev = lab_cur + N_;
	// PEARL.h:373-380 passes settings.point_weights (only findVanishingPoints_ sets them; read by point there)
	double *d_w = nullptr;
	if (weights_by_point() && s_.point_weights.size() == (size_t)N_) {
		PXB_TRY(ctx_->pref2.reserve(sizeof(double) * (size_t)N_));
		d_w = ctx_->pref2.as<double>();
		PXB_TRY(api_h2d(ctx_, d_w, s_.point_weights.data(), sizeof(double) * (size_t)N_));
	}
	std::vector<int64_t> counts;
	while (!convergence && iteration_number++ < 100) {
		const bool init_with_previous = iteration_number > 1 && !model_rejected;
		// ---- labeling ----
		const int64_t L = (int64_t)models_.size();
		if (L == 0) break;
		Scoped tl(prof_, "pearl iteration");
		std::vector<double> flat((size_t)L * ms_);
		for (int64_t l = 0; l < L; ++l) std::copy(models_[l].model.begin(), models_[l].model.end(), flat.begin() + l * ms_);
		const bool init_prev = init_with_previous && have_labels; // no instance was rejected: the previous labels are valid
		// packed results: energy | before[L] | after[L] | counts[L] | counts2[L] | fitted[L ms] | cand[L ms] | ok[L]
		const size_t bL = sizeof(double) * (size_t)L, bM = sizeof(double) * (size_t)L * ms_;
		const size_t o_before = 8, o_after = o_before + bL, o_cnt = o_after + bL, o_cnt2 = o_cnt + bL, o_fit = o_cnt2 + bL,
		             o_cand = o_fit + bM, o_ok = o_cand + bM, pack_bytes = o_ok + sizeof(int32_t) * (size_t)L;
		PXB_TRY(ctx_->chain_par.reserve(bM));
		PXB_TRY(ctx_->staging.reserve(sizeof(double) * (size_t)N_ * (L + 1)));
		PXB_TRY(ctx_->pack.reserve(pack_bytes));
		PXB_TRY(ctx_->idx.reserve(sizeof(int32_t) * ((size_t)N_ + (size_t)L + 1) + 64));
		const bool slots = chain_slots(bM, pack_bytes);
		chain_tmp_in_.resize(bM);
		unsigned char *hin = slots ? ctx_->chain_in : chain_tmp_in_.data();
		std::memcpy(hin, flat.data(), bM);
		pack_host_.resize(pack_bytes);
		unsigned char *hout = slots ? ctx_->chain_out : pack_host_.data();
		double *energy_dev = nullptr;
		auto enqueue = [&]() -> int {
			double *d_models = ctx_->chain_par.as<double>();
			char *pk = ctx_->pack.as<char>();
			if (slots)
				PXB_CUDA(cudaMemcpyAsync(d_models, hin, bM, cudaMemcpyHostToDevice, ctx_->stream));
			else
				PXB_TRY(api_h2d(ctx_, d_models, hin, bM));
			PXB_TRY(launch_pearl_datacost(ctx_, d_models, L, s_.threshold, s_.lambda, ctx_->staging.as<double>()));
			const int32_t *init = nullptr;
			if (init_prev) {
				PXB_CUDA(cudaMemcpyAsync(lab_prev, lab_cur, sizeof(int32_t) * (size_t)N_, cudaMemcpyDeviceToDevice, ctx_->stream));
				init = lab_prev;
			}
			ctx_->label_memo = true;
			const int rc_label = pearl_label_enqueue(ctx_, ctx_->staging.as<double>(), N_, (int32_t)(L + 1), s_.lambda, label_cost,
			                                         smooth ? graph_.off.data() : nullptr, smooth ? graph_.idx.data() : nullptr,
			                                         smooth ? smooth_edges_ : 0, init, lab_cur, &energy, &energy_dev);
			ctx_->label_memo = false;
			PXB_TRY(rc_label);
			if (energy_dev) PXB_CUDA(cudaMemcpyAsync(pk, energy_dev, sizeof(double), cudaMemcpyDeviceToDevice, ctx_->stream));
			// ---- parameterEstimation ----
			int32_t *d_off = ctx_->idx.as<int32_t>(), *d_idx = d_off + (L + 1);
			PXB_TRY(launch_label_lists(ctx_, lab_cur, N_, (int)L, d_off, d_idx));
			PXB_TRY(launch_segment_sums(ctx_, d_models, L, lab_cur, reinterpret_cast<double *>(pk + o_before),
			                            reinterpret_cast<int64_t *>(pk + o_cnt)));
			// instances with fewer points than nonMinimalSampleSize() are not refitted (:363-365): the solvers report !ok for them
			PXB_TRY(launch_fit_family((int)L, d_off, d_idx, d_w, reinterpret_cast<double *>(pk + o_fit), reinterpret_cast<int32_t *>(pk + o_ok)));
			PXB_TRY(launch_select_models(ctx_, d_models, reinterpret_cast<double *>(pk + o_fit), reinterpret_cast<int32_t *>(pk + o_ok),
			                             (int)L, ms_, reinterpret_cast<double *>(pk + o_cand)));
			PXB_TRY(launch_segment_sums(ctx_, reinterpret_cast<double *>(pk + o_cand), L, lab_cur, reinterpret_cast<double *>(pk + o_after),
			                            reinterpret_cast<int64_t *>(pk + o_cnt2)));
			if (slots)
				PXB_CUDA(cudaMemcpyAsync(hout, pk, pack_bytes, cudaMemcpyDeviceToHost, ctx_->stream));
			else
				PXB_TRY(api_d2h(ctx_, hout, pk, pack_bytes));
			return PXB_OK;
		};
		// replayable when the sweep is the single-block greedy kernel (no smoothness term, N <= 16384: the alpha-expansion has
		// a host move loop, the multi-block greedy sweep is a cooperative launch)
		if (slots && !smooth && N_ <= 16384) {
			PXB_TRY(run_chain(chain_key(kChainPearl, {buffers_key(), (uint64_t)L, (uint64_t)init_prev, bits(s_.threshold), bits(s_.lambda),
			                           bits(label_cost), bits(d_w)}), enqueue));
			energy_dev = ctx_->outB.as<double>(); // the greedy sweep's energy sits in the pack (replays do not run the lambda)
		} else {
			PXB_TRY(enqueue());
		}
		have_labels = true;
		PXB_TRY(api_sync(ctx_));
		const unsigned char *h = hout;
		if (energy_dev) std::memcpy(&energy, h, sizeof(double));
		const double *before = reinterpret_cast<const double *>(h + o_before), *after = reinterpret_cast<const double *>(h + o_after);
		const double *fitted = reinterpret_cast<const double *>(h + o_fit);
		const int32_t *ok = reinterpret_cast<const int32_t *>(h + o_ok);
		counts.assign(reinterpret_cast<const int64_t *>(h + o_cnt), reinterpret_cast<const int64_t *>(h + o_cnt) + L);
		bool model_parameters_changed = false;
		model_rejected = false;
		size_t outliers = (size_t)N_;
		for (int64_t l = 0; l < L; ++l) outliers -= (size_t)counts[l];
		for (int64_t l = 0; l < L; ++l) {
			if ((size_t)counts[l] < non_minimal_sample_size()) continue; // :363-365
			if (ok[l] && after[l] < before[l]) { // :393-399
				models_[l].model.assign(fitted + l * ms_, fitted + (l + 1) * ms_);
				model_parameters_changed = true;
			}
		}
		if (s_.do_logging) {
			fprintf(stdout, "[pxb]   PEARL it %zu: L=%lld energy %.4f (prev %.4f) init=%d points per instance:", iteration_number,
			        (long long)L, energy, previous_energy, init_prev ? 1 : 0);
			for (int64_t l = 0; l < L; ++l) fprintf(stdout, " %lld", (long long)counts[l]);
			fprintf(stdout, " outliers %zu%s\n", outliers, model_parameters_changed ? " (refit accepted)" : "");
		}
		// ---- rejectInstances (back to front) ----
		for (int64_t l = L - 1; l >= 0; --l)
			if ((size_t)counts[l] < s_.min_inliers) {
				outliers += (size_t)counts[l];
				PXB_TRY(erase_pref_row((size_t)l, models_.size()));
				models_.erase(models_.begin() + l);
				model_rejected = true;
			}
		pearl_outliers_ = outliers;
		if (!model_rejected && !model_parameters_changed && std::fabs(energy - previous_energy) < 1e-5 &&
		    iteration_number > 1)
			convergence = true;
		previous_energy = energy;
	}
	// getLabeling (:218-249): raw gco labels of the last labeling
	std::vector<int32_t> labels((size_t)N_, 0);
	if (have_labels) {
		PXB_TRY(api_d2h(ctx_, labels.data(), lab_cur, sizeof(int32_t) * (size_t)N_));
		PXB_TRY(api_sync(ctx_));
	}
	labeling_.assign(labels.begin(), labels.end());
	return PXB_OK;
}

// px/include/progressive_x.h:495-513
size_t Driver::predicted_unseen_inliers(size_t iterations, size_t compound_inliers) const {
	const size_t unseen_point_number = (size_t)N_ - compound_inliers;
	const double one_over_iteration_number = 1.0 / (double)iterations;
	const double one_over_sample_size = 1.0 / (double)m_;
	const double inlier_ratio =
	    std::pow(1.0 - std::pow(1.0 - s_.confidence, one_over_iteration_number), one_over_sample_size);
	return static_cast<size_t>(std::round((double)unseen_point_number * inlier_ratio));
}

// ---- sharded runs: the request loop of ranks > 0 and the final broadcast ------------------------------------------------
int Driver::shard_worker() {
	bool has_compound = false;
	const size_t bytes = shard_msg_bytes();
	PXB_TRY(ctx_->shard_msg.reserve(bytes));
	for (;;) {
		PXB_TRY(shard_broadcast(ctx_, ctx_->shard_msg.ptr, bytes, 0));
		ShardMsg h;
		PXB_CUDA(cudaMemcpyAsync(&h, ctx_->shard_msg.ptr, sizeof(h), cudaMemcpyDeviceToHost, ctx_->stream));
		PXB_TRY(ctx_wait(ctx_));
		if (h.op == kShardBlock) {
			if (h.want < 0 || (size_t)h.want > shard_block_cap()) {
				set_error("sharded block request of %lld samples exceeds the agreed capacity", (long long)h.want);
				return PXB_ERR_STATE;
			}
			PXB_TRY(shard_compute_and_gather((size_t)h.want, h.T2, h.has_compound != 0 && has_compound));
			++shard_blocks_;
		} else if (h.op == kShardCompound) {
			PXB_TRY(ctx_->cpref.reserve(sizeof(double) * (size_t)N_));
			PXB_TRY(shard_broadcast(ctx_, ctx_->cpref.ptr, sizeof(double) * (size_t)N_, 0));
			has_compound = true;
		} else if (h.op == kShardDone) {
			if (h.result < 0) {
				set_error("the coordinator rank reported status %lld", (long long)h.result);
				return (int)h.result;
			}
			const size_t M = (size_t)h.result, pay = sizeof(double) * M * ms_ + sizeof(int64_t) * (size_t)N_;
			PXB_TRY(ctx_->staging.reserve(pay));
			PXB_TRY(shard_broadcast(ctx_, ctx_->staging.ptr, pay, 0));
			std::vector<unsigned char> host(pay);
			PXB_CUDA(cudaMemcpyAsync(host.data(), ctx_->staging.ptr, pay, cudaMemcpyDeviceToHost, ctx_->stream));
			PXB_TRY(ctx_wait(ctx_));
			models_.assign(M, Instance());
			for (size_t k = 0; k < M; ++k) {
				models_[k].model.resize(ms_);
				std::memcpy(models_[k].model.data(), host.data() + sizeof(double) * k * ms_, sizeof(double) * ms_);
			}
			labeling_.resize((size_t)N_);
			std::memcpy(labeling_.data(), host.data() + sizeof(double) * M * ms_, sizeof(int64_t) * (size_t)N_);
			return PXB_OK;
		} else {
			set_error("unknown shard request %d", h.op);
			return PXB_ERR_STATE;
		}
	}
}

int Driver::shard_finish(int result) {
	ShardMsg h{};
	h.op = kShardDone;
	h.result = result < 0 ? result : (int64_t)models_.size();
	PXB_TRY(shard_send(h, nullptr, 0));
	if (result < 0) return api_sync(ctx_);
	const size_t M = models_.size(), pay = sizeof(double) * M * ms_ + sizeof(int64_t) * (size_t)N_;
	std::vector<unsigned char> host(pay);
	for (size_t k = 0; k < M; ++k) std::memcpy(host.data() + sizeof(double) * k * ms_, models_[k].model.data(), sizeof(double) * ms_);
	std::memcpy(host.data() + sizeof(double) * M * ms_, labeling_.data(), sizeof(int64_t) * (size_t)N_);
	PXB_TRY(ctx_->staging.reserve(pay));
	PXB_CUDA(cudaMemcpyAsync(ctx_->staging.ptr, host.data(), pay, cudaMemcpyHostToDevice, ctx_->stream));
	PXB_TRY(shard_broadcast(ctx_, ctx_->staging.ptr, pay, 0));
	return api_sync(ctx_);
}

int Driver::run(int setup_status) {
	PXB_CUDA(cudaSetDevice(ctx_->device)); // the driver launches kernels directly as well: bind this host thread
	if (setup_status != PXB_OK) { // the coordinator failed before the loop: release the other ranks
		if (sharded() && !is_worker()) (void)shard_finish(setup_status);
		return setup_status;
	}
	if (!sharded()) return run_local();
	if (is_worker()) return shard_worker();
	const int rc = run_local();
	const int rc2 = shard_finish(rc);
	return rc != PXB_OK ? rc : rc2;
}

// px/include/progressive_x.h:251-489
int Driver::run_local() {
	labeling_.assign((size_t)N_, 0);
	models_.clear();
	defer_inliers_ = true;
	PXB_TRY(reserve_round_buffers());
	size_t number_of_ransac_iterations = 0, unaccepted = 0;
	// statistics.inliers_of_each_model.size(): one entry is appended whenever an instance is added while it is the only one
	// (:375-381); the reference passes this COUNT as the compound inlier number whenever one instance remains (:447-451)
	size_t inliers_of_each_model_size = 0;
	// ---- statistics (progressive_x.h:78-104): CUDA events on the stream where the reference reads its clock ----
	pxb_multi_model_statistics &st = ctx_->last_statistics;
	const bool timed = s_.allow_shard; // nested (DEGENSAC) runs do not touch the context's statistics
	const int64_t launches0 = ctx_->launches;
	if (timed) {
		st = pxb_multi_model_statistics{};
		if (!ctx_->timing_events_ready) {
			for (cudaEvent_t &e : ctx_->timing_events) PXB_CUDA(cudaEventCreate(&e));
			ctx_->timing_events_ready = true;
		}
		PXB_CUDA(cudaEventRecord(ctx_->timing_events[5 * PXB_MAX_ROUNDS], ctx_->stream));
	}
	struct RoundMarks {
		int first_event;
		pxb_iteration_statistics stat;
	};
	std::vector<RoundMarks> rounds;
	int next_event = 0;
	auto mark = [&]() -> int { // records the next event of the current round
		if (timed) cudaEventRecord(ctx_->timing_events[next_event], ctx_->stream);
		return next_event++;
	};
	for (size_t it = 0; it < PXB_MAX_ROUNDS; ++it) { // :272 hard cap
		std::vector<double> model;
		bool found = false;
		next_event = 5 * (int)rounds.size();
		const int e_start = mark();
		PXB_TRY(propose(s_.seed * 1000003ull + it, model, found));
		double tanimoto = 0.0;
		if (found) {
			mark(); // end of the proposal engine / start of the validation
			PXB_TRY(proposal_finish(model.data(), proposal_inliers_, tanimoto));
		}
		if (s_.do_logging)
			fprintf(stdout, "[pxb] proposal %zu: %s, %zu inliers, %zu iterations, %zu LO runs, %zu graph cuts, DEGENSAC %zu/%zu\n",
			        it + 1, found ? "found" : "none", proposal_inliers_.size(), iteration_number_, lo_number_, graph_cut_number_,
			        degensac_updates_, degensac_degenerate_);
		if (!found) continue; // :301-303
		number_of_ransac_iterations += iteration_number_;
		RoundMarks rm{};
		rm.first_event = e_start;
		rm.stat.ransac_iteration_number = iteration_number_;
		rm.stat.local_optimization_number = lo_number_;
		rm.stat.graph_cut_number = graph_cut_number_;
		rm.stat.proposal_inlier_number = proposal_inliers_.size();
		// isPutativeModelValid (:565-591): enough inliers, and `maximum_tanimoto_similarity < similarity` rejects (NaN
		// compares false: accepted)
		const bool valid = proposal_inliers_.size() >= std::max((size_t)m_, s_.min_inliers) && !(s_.max_tanimoto < tanimoto);
		if (!valid) { // :334-346 (the counter is never reset)
			++unaccepted;
			if (unaccepted == s_.max_proposals_without_change) break;
			continue;
		}
		mark(); // end of the validation / start of the optimisation
		Instance inst;
		inst.model = model;
		PXB_CUDA(cudaMemcpyAsync(pref_row(models_.size()), ctx_->pref2.ptr, sizeof(double) * (size_t)N_, cudaMemcpyDeviceToDevice, ctx_->stream));
		models_.push_back(std::move(inst));
		if (models_.size() == 1) { // :375-385
			std::fill(labeling_.begin(), labeling_.end(), 1);
			for (int64_t i : proposal_inliers_) labeling_[(size_t)i] = 0;
			++inliers_of_each_model_size;
		} else {
			PXB_TRY(pearl()); // :390-396
		}
		mark(); // end of the optimisation / start of the compound update
		// updateCompoundModel (:597-624): max over the *stored* preference vectors
		if (!models_.empty()) PXB_TRY(publish_compound());
		mark(); // end of this round's compound update (every round has its own five events: they are read after the run)
		rm.stat.number_of_instances = models_.size();
		if (rounds.size() < PXB_MAX_ROUNDS) rounds.push_back(rm);
		size_t unseen;
		if (models_.size() == 1) // evaluated AFTER the optimisation: also when PEARL pruned the set back to one instance
			unseen = predicted_unseen_inliers(number_of_ransac_iterations, inliers_of_each_model_size);
		else
			unseen = predicted_unseen_inliers(number_of_ransac_iterations, (size_t)N_ - pearl_outliers_);
		if (s_.do_logging)
			fprintf(stdout, "[pxb] round %zu: %zu instances, %zu RANSAC iterations, predicted unseen inliers %zu\n", it + 1,
			        models_.size(), number_of_ransac_iterations, unseen);
		if (unseen < s_.min_inliers) break;              // :468
		if (models_.size() >= s_.max_models) break;      // :472
	}
	if (timed) { // addIterationStatistics (:91-100) + processing_time (:483-488)
		cudaEvent_t *ev = ctx_->timing_events;
		PXB_CUDA(cudaEventRecord(ev[5 * PXB_MAX_ROUNDS + 1], ctx_->stream));
		PXB_TRY(ctx_wait(ctx_));
		for (RoundMarks &rm : rounds) {
			float ms[4] = {0, 0, 0, 0};
			for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&ms[k], ev[rm.first_event + k], ev[rm.first_event + k + 1]);
			rm.stat.time_of_proposal_engine = ms[0] * 1e-3;
			rm.stat.time_of_model_validation = ms[1] * 1e-3;
			rm.stat.time_of_optimization = ms[2] * 1e-3;
			rm.stat.time_of_compound_model_update = ms[3] * 1e-3;
			st.iteration_statistics[st.iteration_statistics_size++] = rm.stat;
			st.total_time_of_proposal_engine += rm.stat.time_of_proposal_engine;
			st.total_time_of_model_validation += rm.stat.time_of_model_validation;
			st.total_time_of_optimization += rm.stat.time_of_optimization;
			st.total_time_of_compound_model_calculation += rm.stat.time_of_compound_model_update;
		}
		float total_ms = 0;
		cudaEventElapsedTime(&total_ms, ev[5 * PXB_MAX_ROUNDS], ev[5 * PXB_MAX_ROUNDS + 1]);
		st.processing_time = total_ms * 1e-3;
		st.model_number = models_.size();
		st.inliers_of_each_model_size = inliers_of_each_model_size;
		st.kernel_launches = (size_t)(ctx_->launches - launches0);
	}
	return PXB_OK;
}

// the caller's model buffer must hold every instance: the labeling refers to all of them
int check_model_capacity(int64_t M, int64_t max_models_out) {
	if (M <= max_models_out) return PXB_OK;
	set_error("models_out holds %lld models but %lld instances were found", (long long)max_models_out, (long long)M);
	return PXB_ERR_ARGUMENT;
}

int run_two_view(pxb_ctx *ctx, int type, const double *corr, int64_t N, int64_t *labeling_out, double *models_out,
                 int64_t max_models_out, const double (&sizes)[4], double lambda, double threshold, double confidence, double radius,
                 double max_tanimoto, size_t max_iters, size_t min_points, int max_models, size_t sampler_id,
                 double scoring_exponent, bool set_exponent, int do_logging, uint64_t seed) {
	if (!ctx || !corr || !labeling_out || !models_out || N < 4) {
		set_error("bad argument");
		return PXB_ERR_ARGUMENT;
	}
	if (sampler_id > 3) { // progressivex_python.cpp:240-245
		fprintf(stderr, "Unknown sampler identifier: %zu. The accepted samplers are 0 (uniform sampling), 1 (PROSAC "
		                "sampling), 2 (P-NAPSAC sampling)\n", sampler_id);
		return 0;
	}
	PXB_TRY(pxb_upload_points(ctx, type, corr, N));
	Settings s;
	s.type = type;
	s.min_inliers = min_points;
	s.threshold = threshold;
	s.confidence = confidence;
	s.max_tanimoto = max_tanimoto;
	s.lambda = lambda;
	s.max_iters = max_iters;
	if (max_models > 0) s.max_models = (size_t)max_models;
	s.sampler_id = sampler_id;
	s.napsac = sampler_id == 3;
	s.progressive_napsac = sampler_id == 2 && sizes[0] > 0 && sizes[1] > 0 && sizes[2] > 0 && sizes[3] > 0;
	for (int d = 0; d < 4; ++d) s.sizes[d] = sizes[d];
	if (set_exponent) s.exponent = (int)scoring_exponent; // setExponent(const int) truncates (progressive_x.h:551)
	s.do_logging = do_logging != 0;
	s.rows_host = corr;
	if (const char *e = getenv("PXB_DEGENSAC")) s.use_degensac = atoi(e) != 0; // A/B switch (default on, as the reference)
	if (seed == 0) seed = (uint64_t)std::chrono::high_resolution_clock::now().time_since_epoch().count();
	s.seed = seed;
	Driver drv(ctx, s);
	int setup = PXB_OK;
	if (!drv.is_worker()) { // sampler state and the neighbourhood graph are only needed where the control flow runs
		if (s.progressive_napsac) drv.build_grid_layers(corr);
		if (lambda > 0.0 || sampler_id == 3) {
			Scoped t(drv.prof_, "build_graph");
			setup = drv.build_graph(radius, graph_degree());
		}
	}
	{
		Scoped t(drv.prof_, "run (inclusive)");
		PXB_TRY(drv.run(setup));
	}
	drv.prof_.print();
	const auto &inst = drv.instances();
	const int ms = model_size(type);
	const int64_t M = (int64_t)inst.size();
	PXB_TRY(check_model_capacity(M, max_models_out));
	for (int64_t k = 0; k < M; ++k)
		std::memcpy(models_out + k * ms, inst[k].model.data(), sizeof(double) * ms);
	std::memcpy(labeling_out, drv.labeling().data(), sizeof(int64_t) * (size_t)N);
	return (int)M;
}

// findVanishingPoints_ / findLines_ (px/src/progressivex_python.cpp:306-423, 425-535)
int run_points_family(pxb_ctx *ctx, int type, const double *rows, const double *weights, int64_t N, int64_t *labeling_out,
                      double *models_out, int64_t max_models_out, double lambda, double threshold, double confidence,
                      double radius, double max_tanimoto, size_t max_iters, size_t min_points, int max_models,
                      size_t sampler_id, size_t max_sampler_id, size_t napsac_id, double scoring_exponent, int do_logging,
                      uint64_t seed) {
	if (!ctx || !rows || !labeling_out || !models_out || N < 2) {
		set_error("bad argument");
		return PXB_ERR_ARGUMENT;
	}
	if (sampler_id > max_sampler_id) { // :353-367 / :463-481: unknown sampler -> message on stderr, 0 models
		fprintf(stderr, "Unknown sampler identifier: %zu. The accepted samplers are 0 (uniform sampling), 1 (PROSAC "
		                "sampling), 2 (P-NAPSAC sampling)\n", sampler_id);
		return 0;
	}
	PXB_TRY(pxb_upload_points(ctx, type, rows, N));
	Settings s;
	s.type = type;
	s.min_inliers = min_points;
	s.threshold = threshold;
	s.confidence = confidence;
	s.max_tanimoto = max_tanimoto;
	s.lambda = lambda;
	s.max_iters = max_iters;
	if (max_models > 0) s.max_models = (size_t)max_models;
	s.sampler_id = sampler_id;
	s.napsac = sampler_id == napsac_id;
	s.exponent = (int)scoring_exponent; // setScoringExponent -> setExponent(const int)
	s.do_logging = do_logging != 0;
	if (weights && type == PXB_MODEL_VANISHING_POINT) s.point_weights.assign(weights, weights + N); // :381
	if (seed == 0) seed = (uint64_t)std::chrono::high_resolution_clock::now().time_since_epoch().count();
	s.seed = seed;
	Driver drv(ctx, s);
	int setup = PXB_OK;
	if (!drv.is_worker() && (lambda > 0.0 || s.napsac)) setup = drv.build_graph(radius, graph_degree());
	PXB_TRY(drv.run(setup));
	const auto &inst = drv.instances();
	const int64_t M = (int64_t)inst.size();
	PXB_TRY(check_model_capacity(M, max_models_out));
	for (int64_t k = 0; k < M; ++k) std::memcpy(models_out + k * 3, inst[k].model.data(), sizeof(double) * 3);
	std::memcpy(labeling_out, drv.labeling().data(), sizeof(int64_t) * (size_t)N);
	return (int)M;
}

} // namespace
} // namespace pxb

using namespace pxb;

extern "C" {

int pxb_find_homographies(pxb_ctx *ctx, const double *correspondences, int64_t N, int64_t *labeling_out,
                          double *models_out, int64_t max_models_out, size_t w1, size_t h1, size_t w2, size_t h2,
                          double spatial_coherence_weight, double threshold, double confidence,
                          double neighborhood_ball_radius, double maximum_tanimoto_similarity, size_t max_iters,
                          size_t minimum_point_number, int maximum_model_number, size_t sampler_id,
                          double scoring_exponent, int do_logging, uint64_t seed) {
	const double sizes[4] = {(double)w1, (double)h1, (double)w2, (double)h2};
	return run_two_view(ctx, PXB_MODEL_HOMOGRAPHY, correspondences, N, labeling_out, models_out, max_models_out, sizes,
	                    spatial_coherence_weight, threshold, confidence, neighborhood_ball_radius,
	                    maximum_tanimoto_similarity, max_iters, minimum_point_number, maximum_model_number, sampler_id,
	                    scoring_exponent, true, do_logging, seed);
}

int pxb_find_two_view_motions(pxb_ctx *ctx, const double *correspondences, int64_t N, int64_t *labeling_out,
                              double *models_out, int64_t max_models_out, size_t w1, size_t h1, size_t w2, size_t h2,
                              double spatial_coherence_weight, double threshold, double confidence,
                              double neighborhood_ball_radius, double maximum_tanimoto_similarity, size_t max_iters,
                              size_t minimum_point_number, int maximum_model_number, size_t sampler_id,
                              double scoring_exponent, int do_logging, uint64_t seed) {
	// findTwoViewMotions_ never forwards scoring_exponent (progressivex_python.cpp:621-638): the exponent stays 2
	const double sizes[4] = {(double)w1, (double)h1, (double)w2, (double)h2};
	return run_two_view(ctx, PXB_MODEL_FUNDAMENTAL, correspondences, N, labeling_out, models_out, max_models_out, sizes,
	                    spatial_coherence_weight, threshold, confidence, neighborhood_ball_radius,
	                    maximum_tanimoto_similarity, max_iters, minimum_point_number, maximum_model_number, sampler_id,
	                    scoring_exponent, false, do_logging, seed);
}

// find6DPoses_ (progressivex_python.cpp:41-171): K^-1-normalised image points for estimation, threshold / f, the
// neighbourhood graph on the UN-normalised [u v X Y Z] rows (:104 vs :143), uniform samplers.
int pxb_find_6d_poses(pxb_ctx *ctx, const double *image_points, const double *world_points, const double *K,
                      int64_t N, int64_t *labeling_out, double *poses_out, int64_t max_models_out,
                      double spatial_coherence_weight, double threshold, double confidence,
                      double neighborhood_ball_radius, double maximum_tanimoto_similarity, size_t max_iters,
                      size_t minimum_point_number, int maximum_model_number, uint64_t seed) {
	if (!ctx || !image_points || !world_points || !K || !labeling_out || !poses_out || N < 3) {
		set_error("bad argument");
		return PXB_ERR_ARGUMENT;
	}
	// Eigen Matrix3d::inverse(): cofactors / determinant
	double Kinv[9];
	{
		const double *m = K;
		const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
		const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
		const double id = 1.0 / det;
		Kinv[0] = c00 * id;
		Kinv[1] = (m[2] * m[7] - m[1] * m[8]) * id;
		Kinv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
		Kinv[3] = c01 * id;
		Kinv[4] = (m[0] * m[8] - m[2] * m[6]) * id;
		Kinv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
		Kinv[6] = c02 * id;
		Kinv[7] = (m[1] * m[6] - m[0] * m[7]) * id;
		Kinv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
	}
	std::vector<double> raw((size_t)N * 5), nrm((size_t)N * 5);
	for (int64_t i = 0; i < N; ++i) {
		const double u = image_points[2 * i], v = image_points[2 * i + 1];
		raw[5 * i] = u;
		raw[5 * i + 1] = v;
		nrm[5 * i] = Kinv[0] * u + Kinv[1] * v + Kinv[2] * 1;
		nrm[5 * i + 1] = Kinv[3] * u + Kinv[4] * v + Kinv[5] * 1;
		for (int c = 0; c < 3; ++c) raw[5 * i + 2 + c] = nrm[5 * i + 2 + c] = world_points[3 * i + c];
	}
	const double f = 0.5 * (K[0] + K[4]);
	Settings s;
	s.type = PXB_MODEL_PNP;
	s.min_inliers = minimum_point_number;
	s.threshold = threshold / f;
	s.confidence = confidence;
	s.max_tanimoto = maximum_tanimoto_similarity;
	s.lambda = spatial_coherence_weight;
	s.max_iters = max_iters;
	if (maximum_model_number > 0) s.max_models = (size_t)maximum_model_number;
	s.sampler_id = 0;
	if (seed == 0) seed = (uint64_t)std::chrono::high_resolution_clock::now().time_since_epoch().count();
	s.seed = seed;
	// the neighbourhood graph is built on the raw rows [u v X Y Z] (progressivex_python.cpp:60-75), everything else runs on
	// the normalised ones: the raw rows only travel when a graph is needed
	const bool needs_graph = spatial_coherence_weight > 0.0;
	if (needs_graph) PXB_TRY(pxb_upload_points(ctx, PXB_MODEL_PNP, raw.data(), N));
	else PXB_TRY(pxb_upload_points(ctx, PXB_MODEL_PNP, nrm.data(), N));
	Driver drv(ctx, s);
	int setup = PXB_OK;
	if (needs_graph) {
		if (!drv.is_worker()) {
			Scoped t(drv.prof_, "build_graph");
			setup = drv.build_graph(neighborhood_ball_radius, graph_degree());
		}
		if (setup == PXB_OK) setup = pxb_upload_points(ctx, PXB_MODEL_PNP, nrm.data(), N);
	}
	{
		Scoped t(drv.prof_, "run (inclusive)");
		PXB_TRY(drv.run(setup));
	}
	drv.prof_.print();
	const auto &inst = drv.instances();
	const int64_t M = (int64_t)inst.size();
	PXB_TRY(check_model_capacity(M, max_models_out));
	for (int64_t k = 0; k < M; ++k) std::memcpy(poses_out + k * 12, inst[k].model.data(), sizeof(double) * 12);
	std::memcpy(labeling_out, drv.labeling().data(), sizeof(int64_t) * (size_t)N);
	return (int)M;
}

int pxb_find_vanishing_points(pxb_ctx *ctx, const double *lines, const double *weights, int64_t N, int64_t *labeling_out,
                              double *vanishing_points_out, int64_t max_models_out, size_t, size_t,
                              double spatial_coherence_weight, double threshold, double confidence,
                              double neighborhood_ball_radius, double maximum_tanimoto_similarity, size_t max_iters,
                              size_t minimum_point_number, int maximum_model_number, size_t sampler_id,
                              double scoring_exponent, int do_logging, uint64_t seed) {
	// samplers 0 (uniform) and 1 (PROSAC, falls back to uniform here) only; no NAPSAC for this entry
	return run_points_family(ctx, PXB_MODEL_VANISHING_POINT, lines, weights, N, labeling_out, vanishing_points_out,
	                         max_models_out, spatial_coherence_weight, threshold, confidence, neighborhood_ball_radius,
	                         maximum_tanimoto_similarity, max_iters, minimum_point_number, maximum_model_number, sampler_id, 1,
	                         SIZE_MAX, scoring_exponent, do_logging, seed);
}

int pxb_find_lines(pxb_ctx *ctx, const double *points, const double *weights, int64_t N, int64_t *labeling_out,
                   double *lines_out, int64_t max_models_out, size_t, size_t, double spatial_coherence_weight,
                   double threshold, double confidence, double neighborhood_ball_radius,
                   double maximum_tanimoto_similarity, size_t max_iters, size_t minimum_point_number,
                   int maximum_model_number, size_t sampler_id, double scoring_exponent, int do_logging, uint64_t seed) {
	// samplers 0, 1 and 2 = NapsacSampler (progressivex_python.cpp:476-478)
	return run_points_family(ctx, PXB_MODEL_LINE2D, points, weights, N, labeling_out, lines_out, max_models_out,
	                         spatial_coherence_weight, threshold, confidence, neighborhood_ball_radius,
	                         maximum_tanimoto_similarity, max_iters, minimum_point_number, maximum_model_number, sampler_id, 2,
	                         2, scoring_exponent, do_logging, seed);
}

int pxb_h_degenerate_sample(const double *rows, const int64_t *sample7, const double *F, double *H_out, int32_t *degenerate) {
	if (!rows || !sample7 || !F || !H_out || !degenerate) {
		set_error("null argument");
		return PXB_ERR_ARGUMENT;
	}
	*degenerate = h_degenerate_sample(rows, sample7, F, H_out) ? 1 : 0;
	return PXB_OK;
}

} // extern "C"
