// mini_eigen.h -- the handful of Eigen spellings used by the reference function bodies that oracle/extract_ref.py
// compiles verbatim (TEST INFRASTRUCTURE ONLY; written from scratch, not Eigen code). Only element access, comma
// initialisation, Zero(), hasNaN(), resize(), row().cross() and determinant() exist. All arithmetic of the pinned
// functions is spelled out in scalar code in the reference sources themselves; the only arithmetic supplied HERE is
// determinant() (partial-pivot LU, as Eigen does for dynamic matrices), cross() and the fixed-size product operator*.
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>

namespace Eigen {
constexpr int Dynamic = -1;

template <class T, int R, int C> class Matrix;

template <class T, int R, int C> struct CommaInit {
	Matrix<T, R, C> &m;
	int k;
	CommaInit &operator,(const T &v) {
		m.data_[(size_t)k++] = v; // row-major fill order, like Eigen's comma initialiser
		return *this;
	}
};

template <class T, int R, int C> class Matrix {
  public:
	std::vector<T> data_; // row-major
	int rows_, cols_;
	Matrix() : rows_(R > 0 ? R : 0), cols_(C > 0 ? C : 0) { data_.assign((size_t)rows_ * cols_, T()); }
	Matrix(int r, int c) : rows_(r), cols_(c) { data_.assign((size_t)r * c, T()); }
	template <int R2, int C2> Matrix(const Matrix<T, R2, C2> &o) { assign_from(o); }
	template <int R2, int C2> Matrix &operator=(const Matrix<T, R2, C2> &o) {
		assign_from(o);
		return *this;
	}
	template <int R2, int C2> void assign_from(const Matrix<T, R2, C2> &o) {
		if (R > 0 && C > 0 && o.rows_ * o.cols_ == R * C) { // vectors may be assigned across orientations
			rows_ = R;
			cols_ = C;
		} else {
			rows_ = o.rows_;
			cols_ = o.cols_;
		}
		data_ = o.data_;
	}
	int rows() const { return rows_; }
	int cols() const { return cols_; }
	size_t size() const { return data_.size(); }
	void resize(int r, int c = 1) {
		rows_ = r;
		cols_ = c;
		data_.assign((size_t)r * c, T());
	}
	T &operator()(int i, int j) { return data_[(size_t)i * cols_ + j]; }
	const T &operator()(int i, int j) const { return data_[(size_t)i * cols_ + j]; }
	T &operator()(int i) { return data_[(size_t)i]; }
	const T &operator()(int i) const { return data_[(size_t)i]; }
	T &operator[](int i) { return data_[(size_t)i]; }
	const T &operator[](int i) const { return data_[(size_t)i]; }
	CommaInit<T, R, C> operator<<(const T &v) {
		data_[0] = v;
		return CommaInit<T, R, C>{*this, 1};
	}
	static Matrix Zero(int r, int c = 1) { return Matrix(r, c); }
	bool hasNaN() const {
		for (const T &v : data_)
			if (v != v) return true;
		return false;
	}
	Matrix<T, 1, Dynamic> row(int i) const {
		Matrix<T, 1, Dynamic> r(1, cols_);
		for (int j = 0; j < cols_; ++j) r.data_[(size_t)j] = (*this)(i, j);
		return r;
	}
	template <int R2, int C2> Matrix<T, 3, 1> cross(const Matrix<T, R2, C2> &b) const {
		Matrix<T, 3, 1> r;
		const std::vector<T> &a = data_;
		r[0] = a[1] * b[2] - a[2] * b[1];
		r[1] = a[2] * b[0] - a[0] * b[2];
		r[2] = a[0] * b[1] - a[1] * b[0];
		return r;
	}
	// PartialPivLU determinant (what Eigen's determinant() does for a dynamic-size matrix)
	T determinant() const {
		const int n = rows_;
		std::vector<T> a = data_;
		int sign = 1;
		for (int k = 0; k < n; ++k) {
			int piv = k;
			T best = std::fabs(a[(size_t)k * n + k]);
			for (int i = k + 1; i < n; ++i)
				if (std::fabs(a[(size_t)i * n + k]) > best) {
					best = std::fabs(a[(size_t)i * n + k]);
					piv = i;
				}
			if (best == T(0)) continue;
			if (piv != k) {
				for (int j = 0; j < n; ++j) std::swap(a[(size_t)k * n + j], a[(size_t)piv * n + j]);
				sign = -sign;
			}
			for (int i = k + 1; i < n; ++i) a[(size_t)i * n + k] = a[(size_t)i * n + k] / a[(size_t)k * n + k];
			for (int i = k + 1; i < n; ++i)
				for (int j = k + 1; j < n; ++j) a[(size_t)i * n + j] = a[(size_t)i * n + j] - a[(size_t)i * n + k] * a[(size_t)k * n + j];
		}
		T d = a[0];
		for (int k = 1; k < n; ++k) d = d * a[(size_t)k * n + k];
		return T(sign) * d;
	}
};

// Coefficient-wise product as Eigen evaluates a fixed-size 3x3 lazy product: every entry is the inner product of a row
// and a column accumulated left to right, (a0 b0 + a1 b1) + a2 b2. Used only by the DEGENSAC plane-and-parallax solver.
template <class T, int R, int C, int C2> Matrix<T, R, C2> operator*(const Matrix<T, R, C> &a, const Matrix<T, C, C2> &b) {
	Matrix<T, R, C2> r(a.rows_, b.cols_);
	for (int i = 0; i < a.rows_; ++i)
		for (int j = 0; j < b.cols_; ++j) {
			T acc = a(i, 0) * b.data_[(size_t)0 * b.cols_ + j];
			for (int k = 1; k < a.cols_; ++k) acc = acc + a(i, k) * b.data_[(size_t)k * b.cols_ + j];
			r.data_[(size_t)i * b.cols_ + j] = acc;
		}
	return r;
}

typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 3, 1> Vector3d;
} // namespace Eigen
