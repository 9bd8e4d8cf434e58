/*
 * pxb200_scoring_adapter.h -- header-only binding of libpxb200.so to the reference's scoring seam.
 *
 * GpuScoringWithCompoundModel<Estimator> implements gcransac::ScoringFunction<Estimator>
 * (graph-cut-ransac/src/pygcransac/include/scoring_function.h:74-100) the way
 * MSACScoringFunctionWithCompoundModel does (src/pyprogressivex/include/scoring_function_with_compound_model.h:18-125):
 * it is injected as GCRANSAC's third template argument (src/pyprogressivex/include/progressive_x.h:114-121) and evaluates
 * getScore's N-point loop on the GPU through the C ABI (pxb_upload_points / pxb_score_compound / pxb_inliers).
 *
 * Include it AFTER the reference's scoring_function.h and progx_model.h (it only names types those headers define):
 * gcransac::ScoringFunction, gcransac::Score, gcransac::Model, progx::Model, cv::Mat, Eigen::VectorXd.
 * tests/test_integration_adapter.py type-checks it against the reference's own declaration of the interface.
 */
#ifndef PXB200_SCORING_ADAPTER_H
#define PXB200_SCORING_ADAPTER_H

#include <cmath>
#include <cstdint>
#include <vector>

#include "pxb200.h"

template <class Estimator>
class GpuScoringWithCompoundModel : public gcransac::ScoringFunction<Estimator> {
	pxb_ctx *ctx = nullptr;
	int model_type; /* PXB_MODEL_HOMOGRAPHY / _FUNDAMENTAL / _PNP / _VANISHING_POINT / _LINE2D */
	double T2 = 0;
	size_t N = 0;
	int exponent = 2; /* scoring_function_with_compound_model.h:20 (an int there as well) */
	const std::vector<progx::Model<Estimator>> *compound = nullptr;
	const Eigen::VectorXd *compound_pref = nullptr;
	mutable const double *uploaded = nullptr;

  public:
	explicit GpuScoringWithCompoundModel(int type, int device = 0) : model_type(type) { pxb_ctx_create(device, &ctx); }
	~GpuScoringWithCompoundModel() override { pxb_ctx_destroy(ctx); }
	GpuScoringWithCompoundModel(const GpuScoringWithCompoundModel &) = delete;
	GpuScoringWithCompoundModel &operator=(const GpuScoringWithCompoundModel &) = delete;

	/* scoring_function_with_compound_model.h:39-58 */
	void setExponent(const int e) { exponent = e; }
	void setCompoundModel(const std::vector<progx::Model<Estimator>> *models, const Eigen::VectorXd *preference) {
		compound = models;
		compound_pref = preference;
	}

	void initialize(const double squared_truncated_threshold, const size_t point_number) override {
		T2 = squared_truncated_threshold;
		N = point_number;
	}

	gcransac::Score getScore(const cv::Mat &points, gcransac::Model &model, const Estimator &, const double,
	                         std::vector<size_t> &inliers, const gcransac::Score &best = gcransac::Score(),
	                         const bool store_inliers = true,
	                         const std::vector<const std::vector<size_t> *> * = nullptr) const override {
		if (uploaded != points.template ptr<double>(0)) { /* cv::Mat(N, d, CV_64F) is the ABI's layout already */
			pxb_upload_points(ctx, model_type, points.template ptr<double>(0), points.rows);
			uploaded = points.template ptr<double>(0);
		}
		double m[12]; /* Eigen stores column-major: export row-major like progressivex_python.cpp:292-300 */
		const int rows = (int)model.descriptor.rows(), cols = (int)model.descriptor.cols();
		for (int r = 0; r < rows; ++r)
			for (int c = 0; c < cols; ++c) m[r * cols + c] = model.descriptor(r, c);
		const bool has_compound = compound != nullptr && !compound->empty() && compound_pref != nullptr;
		int64_t count = 0;
		double value = 0, shared = 0;
		pxb_score_compound(ctx, m, 1, T2, has_compound ? compound_pref->data() : nullptr, &count, &value, &shared);
		gcransac::Score s;
		if ((size_t)count + 1 < best.inlier_number) return s; /* :105-106 */
		s.inlier_number = (size_t)count;
		s.value = value - (has_compound ? std::pow(shared, exponent) : 0.0); /* :110-121 */
		if (store_inliers) {
			std::vector<int64_t> idx(N);
			int64_t n = 0;
			pxb_inliers(ctx, m, T2, idx.data(), &n);
			inliers.assign(idx.begin(), idx.begin() + n);
		}
		return s;
	}
};

#endif /* PXB200_SCORING_ADAPTER_H */
