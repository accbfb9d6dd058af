// TooN stand-in (TEST INFRASTRUCTURE, not product code).
//
// The reference (domantasjurkus/slambench) depends on the external header-only
// library TooN, pinned in its Makefile:20-24 at commit
// 92241416d2a4874fd2334e08a5d417dfea6a1a3f.  TooN is neither vendored in
// /root/reference nor installed in this image and there is no network, so the
// reference C++ backend cannot be compiled as shipped.  This directory holds a
// from-scratch, minimal restatement of exactly the TooN surface the reference's
// cpp path touches, so that kfusion/src/cpp/kernels.cpp, kfusion/src/benchmark.cpp
// and kfusion/include/*.h compile UNMODIFIED (see oracle/Makefile).
//
// Surface provided (call sites in the reference):
//   Vector<N,P,Base>, slice<S,L>(), +=, norm           commons.h:380-404, cpp/kernels.cpp:487-489,765-771
//   Matrix<R,C,P,Layout>, Reference::RowMajor, ()/[]   cpp/kernels.cpp:487,764,782
//   wrapMatrix<R,C>(P*)                                commons.h:367-376,410
//   Identity, Zeros, makeVector                        commons.h:366,382,408; kernels.h:108
//   matrix * matrix                                    commons.h:375
//   gaussian_elimination(Matrix<N,N>, Matrix<N,R>)     commons.h:369
//   SE3<P> (se3.h), GR_SVD<R,C> (GR_SVD.h)             commons.h:402-411; cpp/kernels.cpp:766
//
// Semantics are restated from the published TooN algorithms (partial-pivot Gaussian
// elimination with a double `factor`, Rodrigues/SE3 exponential with TooN's
// small-angle branches, pseudo-inverse back-substitution with a condition cutoff).
// They cannot be diffed against the pinned TooN commit here: "parity unpinned" at
// this boundary.
#ifndef TOON_SHIM_TOON_H
#define TOON_SHIM_TOON_H

#include <cmath>
#include <cstddef>
#include <cassert>
#include <iostream>
#include <algorithm>
#include <type_traits>

namespace TooN {

struct IdentityTag {};
struct ZerosTag {};
static const IdentityTag Identity = IdentityTag();
static const ZerosTag Zeros = ZerosTag();

struct RowMajor {};
namespace Reference {
struct RowMajor {};
}
namespace Internal {
struct VOwn {};
struct VRef {};
template <class A, class B> struct Promote {
	typedef decltype(typename std::remove_const<A>::type() * typename std::remove_const<B>::type()) type;
};
}

template <int N, class P = double, class Base = Internal::VOwn> struct Vector;

// ----------------------------------------------------------------- owning vector
template <int N, class P> struct Vector<N, P, Internal::VOwn> {
	P d[N];
	Vector() {}
	Vector(const ZerosTag&) { for (int i = 0; i < N; ++i) d[i] = 0; }
	template <class P2, class B2> Vector(const Vector<N, P2, B2>& o) { for (int i = 0; i < N; ++i) d[i] = o[i]; }
	template <class P2, class B2> Vector& operator=(const Vector<N, P2, B2>& o) { for (int i = 0; i < N; ++i) d[i] = o[i]; return *this; }
	P& operator[](int i) { return d[i]; }
	const P& operator[](int i) const { return d[i]; }
	int size() const { return N; }
	template <int S, int L> Vector<L, P, Internal::VRef> slice() { return Vector<L, P, Internal::VRef>(d + S); }
	template <int S, int L> Vector<L, const P, Internal::VRef> slice() const { return Vector<L, const P, Internal::VRef>(d + S); }
	template <class P2, class B2> Vector& operator+=(const Vector<N, P2, B2>& o) { for (int i = 0; i < N; ++i) d[i] += o[i]; return *this; }
	template <class P2, class B2> Vector& operator-=(const Vector<N, P2, B2>& o) { for (int i = 0; i < N; ++i) d[i] -= o[i]; return *this; }
	Vector& operator*=(const P& s) { for (int i = 0; i < N; ++i) d[i] *= s; return *this; }
};

// -------------------------------------------------------------- reference vector
template <int N, class P> struct Vector<N, P, Internal::VRef> {
	P* d;
	explicit Vector(P* p) : d(p) {}
	Vector(const Vector& o) : d(o.d) {}
	// assignment copies ELEMENTS (a view never rebinds)
	const Vector& operator=(const Vector& o) const { for (int i = 0; i < N; ++i) d[i] = o[i]; return *this; }
	template <class P2, class B2> const Vector& operator=(const Vector<N, P2, B2>& o) const { for (int i = 0; i < N; ++i) d[i] = o[i]; return *this; }
	P& operator[](int i) const { return d[i]; }
	int size() const { return N; }
	template <int S, int L> Vector<L, P, Internal::VRef> slice() const { return Vector<L, P, Internal::VRef>(d + S); }
	template <class P2, class B2> const Vector& operator+=(const Vector<N, P2, B2>& o) const { for (int i = 0; i < N; ++i) d[i] += o[i]; return *this; }
	template <class P2, class B2> const Vector& operator-=(const Vector<N, P2, B2>& o) const { for (int i = 0; i < N; ++i) d[i] -= o[i]; return *this; }
	template <class S> const Vector& operator*=(const S& s) const { for (int i = 0; i < N; ++i) d[i] *= s; return *this; }
};

template <int N, class P1, class B1, class P2, class B2>
typename Internal::Promote<P1, P2>::type operator*(const Vector<N, P1, B1>& a, const Vector<N, P2, B2>& b) {
	typename Internal::Promote<P1, P2>::type s = 0;
	for (int i = 0; i < N; ++i) s += a[i] * b[i];
	return s;
}
template <int N, class P1, class B1, class P2, class B2>
Vector<N, typename Internal::Promote<P1, P2>::type> operator+(const Vector<N, P1, B1>& a, const Vector<N, P2, B2>& b) {
	Vector<N, typename Internal::Promote<P1, P2>::type> r;
	for (int i = 0; i < N; ++i) r[i] = a[i] + b[i];
	return r;
}
template <int N, class P1, class B1>
Vector<N, typename Internal::Promote<P1, double>::type> operator*(double s, const Vector<N, P1, B1>& a) {
	Vector<N, typename Internal::Promote<P1, double>::type> r;
	for (int i = 0; i < N; ++i) r[i] = s * a[i];
	return r;
}
template <int N, class P1, class B1>
Vector<N, typename Internal::Promote<P1, double>::type> operator*(const Vector<N, P1, B1>& a, double s) {
	Vector<N, typename Internal::Promote<P1, double>::type> r;
	for (int i = 0; i < N; ++i) r[i] = a[i] * s;
	return r;
}
// 3-vector cross product (TooN spells it operator^)
template <class P1, class B1, class P2, class B2>
Vector<3, typename Internal::Promote<P1, P2>::type> operator^(const Vector<3, P1, B1>& a, const Vector<3, P2, B2>& b) {
	Vector<3, typename Internal::Promote<P1, P2>::type> r;
	r[0] = a[1] * b[2] - a[2] * b[1];
	r[1] = a[2] * b[0] - a[0] * b[2];
	r[2] = a[0] * b[1] - a[1] * b[0];
	return r;
}
template <int N, class P, class B> typename std::remove_const<P>::type norm(const Vector<N, P, B>& v) {
	using std::sqrt;
	return sqrt(v * v);
}

inline Vector<3> makeVector(double a, double b, double c) {
	Vector<3> v; v[0] = a; v[1] = b; v[2] = c; return v;
}
inline Vector<6> makeVector(double a, double b, double c, double d, double e, double f) {
	Vector<6> v; v[0] = a; v[1] = b; v[2] = c; v[3] = d; v[4] = e; v[5] = f; return v;
}

// ------------------------------------------------------------------------ matrix
template <int R, int C = R, class P = double, class Layout = RowMajor> struct Matrix;

template <int R, int C, class P> struct Matrix<R, C, P, RowMajor> {
	P d[R * C];
	Matrix() {}
	Matrix(const ZerosTag&) { for (int i = 0; i < R * C; ++i) d[i] = 0; }
	Matrix(const IdentityTag&) { for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) d[r * C + c] = (r == c) ? 1 : 0; }
	template <class P2, class L2> Matrix(const Matrix<R, C, P2, L2>& o) { for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) d[r * C + c] = o(r, c); }
	template <class P2, class L2> Matrix& operator=(const Matrix<R, C, P2, L2>& o) { for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) d[r * C + c] = o(r, c); return *this; }
	Vector<C, P, Internal::VRef> operator[](int r) { return Vector<C, P, Internal::VRef>(d + r * C); }
	Vector<C, const P, Internal::VRef> operator[](int r) const { return Vector<C, const P, Internal::VRef>(d + r * C); }
	P& operator()(int r, int c) { return d[r * C + c]; }
	const P& operator()(int r, int c) const { return d[r * C + c]; }
	int num_rows() const { return R; }
	int num_cols() const { return C; }
};

template <int R, int C, class P> struct Matrix<R, C, P, Reference::RowMajor> {
	P* d;
	Matrix(P* p) : d(p) {}
	Matrix(const Matrix& o) : d(o.d) {}
	const Matrix& operator=(const Matrix& o) const { for (int i = 0; i < R * C; ++i) d[i] = o.d[i]; return *this; }
	template <class P2, class L2> const Matrix& operator=(const Matrix<R, C, P2, L2>& o) const { for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) d[r * C + c] = o(r, c); return *this; }
	Vector<C, P, Internal::VRef> operator[](int r) const { return Vector<C, P, Internal::VRef>(d + r * C); }
	P& operator()(int r, int c) const { return d[r * C + c]; }
	int num_rows() const { return R; }
	int num_cols() const { return C; }
};

template <int R, int C, class P> Matrix<R, C, P, Reference::RowMajor> wrapMatrix(P* p) {
	return Matrix<R, C, P, Reference::RowMajor>(p);
}

// matrix * matrix: plain triple loop, k innermost, accumulation in the promoted type
template <int R, int K, int C, class P1, class L1, class P2, class L2>
Matrix<R, C, typename Internal::Promote<P1, P2>::type> operator*(const Matrix<R, K, P1, L1>& a, const Matrix<K, C, P2, L2>& b) {
	typedef typename Internal::Promote<P1, P2>::type PR;
	Matrix<R, C, PR> out;
	for (int r = 0; r < R; ++r)
		for (int c = 0; c < C; ++c) {
			PR s = 0;
			for (int k = 0; k < K; ++k) s += a(r, k) * b(k, c);
			out(r, c) = s;
		}
	return out;
}
template <int R, int C, class P1, class L1, class P2, class B2>
Vector<R, typename Internal::Promote<P1, P2>::type> operator*(const Matrix<R, C, P1, L1>& a, const Vector<C, P2, B2>& v) {
	typedef typename Internal::Promote<P1, P2>::type PR;
	Vector<R, PR> out;
	for (int r = 0; r < R; ++r) {
		PR s = 0;
		for (int c = 0; c < C; ++c) s += a(r, c) * v[c];
		out[r] = s;
	}
	return out;
}

// Gaussian elimination with partial pivoting, matrix right-hand side.  No
// singularity check: a singular A (the all-zero raycastPose of the first frames,
// cpp/kernels.cpp:53,948) propagates inf/NaN exactly like plain IEEE arithmetic.
template <int N, int R, class P>
Matrix<N, R, P> gaussian_elimination(Matrix<N, N, P> A, Matrix<N, R, P> b) {
	for (int i = 0; i < N; ++i) {
		int argmax = i;
		P maxval = std::abs(A(i, i));
		for (int ii = i + 1; ii < N; ++ii) {
			double v = std::abs(A(ii, i));
			if (v > maxval) { maxval = v; argmax = ii; }
		}
		P pivot = A(argmax, i);
		P inv_pivot = static_cast<P>(1) / pivot;
		if (argmax != i) {
			for (int j = i; j < N; ++j) std::swap(A(i, j), A(argmax, j));
			for (int j = 0; j < R; ++j) std::swap(b(i, j), b(argmax, j));
		}
		for (int j = i + 1; j < N; ++j) A(i, j) *= inv_pivot;
		for (int j = 0; j < R; ++j) b(i, j) *= inv_pivot;
		for (int u = i + 1; u < N; ++u) {
			double factor = A(u, i);
			for (int j = i + 1; j < N; ++j) A(u, j) -= factor * A(i, j);
			for (int j = 0; j < R; ++j) b(u, j) -= factor * b(i, j);
		}
	}
	Matrix<N, R, P> x;
	for (int i = N - 1; i >= 0; --i) {
		for (int c = 0; c < R; ++c) x(i, c) = b(i, c);
		for (int j = i + 1; j < N; ++j)
			for (int c = 0; c < R; ++c) x(i, c) -= A(i, j) * x(j, c);
	}
	return x;
}

}  // namespace TooN
#endif
