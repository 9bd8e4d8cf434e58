// mini_cv.h -- the cv::Mat spellings used by the reference function bodies compiled by oracle/extract_ref.py:
// a non-owning row-major view with rows/cols/data, row(i), ptr<T>(r), at<T>(i). TEST INFRASTRUCTURE ONLY; written from
// scratch. MIN/MAX are spelled exactly as opencv2/core/cvdef.h spells them (their NaN behaviour matters).
#pragma once
#include <cstddef>

#ifndef MIN
#define MIN(a, b) ((a) > (b) ? (b) : (a))
#endif
#ifndef MAX
#define MAX(a, b) ((a) < (b) ? (b) : (a))
#endif
#define CV_64F 6

namespace cv {
class Mat {
  public:
	int rows = 0, cols = 0;
	unsigned char *data = nullptr;
	Mat() {}
	Mat(int r, int c, int /*type*/, void *p) : rows(r), cols(c), data(reinterpret_cast<unsigned char *>(p)) {}
	Mat row(int i) const { return Mat(1, cols, CV_64F, data + sizeof(double) * (size_t)i * cols); }
	template <class T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data) + (size_t)r * cols; }
	template <class T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data) + (size_t)r * cols; }
	template <class T> T &at(int i) const { return reinterpret_cast<T *>(data)[i]; }
	template <class T> T &at(int i, int j) const { return reinterpret_cast<T *>(data)[(size_t)i * cols + j]; }
};
} // namespace cv
