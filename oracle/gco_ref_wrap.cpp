// gco_ref_wrap.cpp -- extern "C" wrapper around the REFERENCE's own gco-v3 / Boykov-Kolmogorov sources.
// TEST INFRASTRUCTURE ONLY (see pxo_oracle.h). The reference sources are compiled where they lie under
// /root/reference (oracle/Makefile, target _ref/libgco_ref.so); nothing is copied into this repository.
//
// Reference sources linked: gcr/GCoptimization.cpp, gcr/LinkedBlockList.cpp, gcr/graph.cpp, gcr/maxflow.cpp
// (gcr/ = /root/reference/graph-cut-ransac/src/pygcransac/include/). They have no third-party includes.
//
// The two entry points drive those sources exactly the way the reference does:
//   gco_ref_pearl_label    <- pearl::PEARL::labeling            px/include/PEARL.h:476-555
//   gco_ref_lo_labeling    <- gcransac::GCRANSAC::labeling      gcr/GCRANSAC.h:914-1022
#include <cstdint>
#include <set>
#include <utility>
#include <vector>

#include "GCoptimization.h"

namespace {
struct Info {
	const double *D;
	int L1;
	double lambda;
};
// px/include/PEARL.h:82-128 reads a cached value per (point,label); here the matrix is an input
double data_fn(int site, int label, void *p) {
	const Info *info = reinterpret_cast<const Info *>(p);
	return info->D[(size_t)site * info->L1 + label];
}
// px/include/PEARL.h:59-79
double smooth_fn(int, int, int l1, int l2, void *p) {
	const Info *info = reinterpret_cast<const Info *>(p);
	return l1 != l2 ? info->lambda : 0;
}
} // namespace

extern "C" {

// Returns the energy returned by expansion(it, 1000). nbr_off/nbr_idx: directed neighbour lists exactly as
// NeighborhoodGraph::getNeighbors(i) yields them (duplicates and one-sided entries preserved).
double gco_ref_pearl_label(int N, int L1, const double *D, double lambda, double label_cost, const int32_t *nbr_off,
                           const int32_t *nbr_idx, const int32_t *init_labels, int32_t *labels_out, int *cycles) {
	GCoptimizationGeneralGraph *gc = new GCoptimizationGeneralGraph(N, L1); // PEARL.h:507-508
	Info info{D, L1, lambda};
	double energy = 0;
	try {
		gc->setDataCost(&data_fn, &info);                     // :519
		if (lambda > 0.0) gc->setSmoothCost(&smooth_fn, &info); // :523-525
		if (label_cost > 0.0) gc->setLabelCost(label_cost);     // :528-529
		if (lambda > 0.0)                                       // :532-536
			for (int i = 0; i < N; ++i)
				for (int32_t e = nbr_off[i]; e < nbr_off[i + 1]; ++e)
					if (i != nbr_idx[e]) gc->setNeighbors(i, nbr_idx[e]);
		if (init_labels) // :541-547
			for (int i = 0; i < N; ++i) gc->setLabel(i, init_labels[i]);
		int it = 0;
		energy = gc->expansion(it, 1000); // :550
		if (cycles) *cycles = it;
		for (int i = 0; i < N; ++i) labels_out[i] = gc->whatLabel(i);
	} catch (GCException &e) {
		e.Report();
		delete gc;
		return -1.0;
	}
	delete gc;
	return energy;
}

// Energy of an arbitrary labelling under the same model (compute_energy()).
double gco_ref_energy(int N, int L1, const double *D, double lambda, double label_cost, const int32_t *nbr_off,
                      const int32_t *nbr_idx, const int32_t *labels) {
	GCoptimizationGeneralGraph gc(N, L1);
	Info info{D, L1, lambda};
	gc.setDataCost(&data_fn, &info);
	if (lambda > 0.0) gc.setSmoothCost(&smooth_fn, &info);
	if (label_cost > 0.0) gc.setLabelCost(label_cost);
	if (lambda > 0.0)
		for (int i = 0; i < N; ++i)
			for (int32_t e = nbr_off[i]; e < nbr_off[i + 1]; ++e)
				if (i != nbr_idx[e]) gc.setNeighbors(i, nbr_idx[e]);
	for (int i = 0; i < N; ++i) gc.setLabel(i, labels[i]);
	return gc.compute_energy();
}

// gcr/GCRANSAC.h:914-1022 with the unary terms (e0,e1) and the clamped distances d given (a13 oracle output).
// The reference's N x N used_edges matrix (:964) is replaced by a set with identical semantics.
// inlier_out[i] = 1 iff what_segment(i) == SINK (:1015-1018). Returns the flow/energy of minimize().
double gco_ref_lo_labeling(int N, const double *e0, const double *e1, const double *d, double lambda,
                           const int32_t *nbr_off, const int32_t *nbr_idx, uint8_t *inlier_out) {
	typedef Energy<double, double, double> E;
	int64_t edges = nbr_off[N];
	E *g = new E(N, (int)edges, NULL);
	for (int i = 0; i < N; ++i) g->add_node();
	for (int i = 0; i < N; ++i) g->add_term1(i, e0[i], e1[i]);
	if (lambda > 0) {
		std::set<std::pair<int, int>> used;
		const double e11 = 0;
		for (int i = 0; i < N; ++i) {
			const double energy1 = d[i];
			for (int32_t e = nbr_off[i]; e < nbr_off[i + 1]; ++e) {
				const int j = nbr_idx[e];
				if (j == i) continue;
				if (used.count({j, i}) || used.count({i, j})) continue;
				used.insert({j, i});
				used.insert({i, j});
				const double energy2 = d[j];
				const double energy_sum = energy1 + energy2;
				const double e00 = 0.5 * energy_sum;
				g->add_term2(i, j, e00 * lambda, lambda, lambda, e11 * lambda);
			}
		}
	}
	const double en = g->minimize();
	for (int i = 0; i < N; ++i) inlier_out[i] = g->what_segment(i) == Graph<double, double, double>::SINK ? 1 : 0;
	delete g;
	return en;
}

} // extern "C"
