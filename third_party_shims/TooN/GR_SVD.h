// TooN stand-in: GR_SVD<R,C> — see TooN.h in this directory for why this exists.
// TEST INFRASTRUCTURE, not product code.
//
// The reference solves the 6x6 ICP normal equations with
//   TooN::GR_SVD<6,6> svd(C); x = svd.backsub(b, 1e6);        (commons.h:402-403)
// i.e. x = V * diag(w_i * 1e6 > w_max ? 1/w_i : 0) * U^T * b  (Moore-Penrose
// pseudo-inverse with a condition-number cutoff).  TooN factorises with
// Golub-Reinsch; any correct SVD gives the same x up to fp64 round-off (~1e-12
// relative on these well-scaled 6x6 systems), so this stand-in uses a one-sided
// Jacobi (Hestenes) SVD in double.
#ifndef TOON_SHIM_GR_SVD_H
#define TOON_SHIM_GR_SVD_H

#include <TooN/TooN.h>

namespace TooN {

template <int Rows, int Cols = Rows, class P = double> class GR_SVD {
public:
	template <class P2, class L2> GR_SVD(const Matrix<Rows, Cols, P2, L2>& m) {
		for (int r = 0; r < Rows; ++r) for (int c = 0; c < Cols; ++c) U(r, c) = m(r, c);
		for (int r = 0; r < Cols; ++r) for (int c = 0; c < Cols; ++c) V(r, c) = (r == c) ? 1 : 0;
		for (int sweep = 0; sweep < 60; ++sweep) {
			bool rotated = false;
			for (int p = 0; p < Cols - 1; ++p)
				for (int q = p + 1; q < Cols; ++q) {
					P alpha = 0, beta = 0, gamma = 0;
					for (int r = 0; r < Rows; ++r) {
						alpha += U(r, p) * U(r, p);
						beta += U(r, q) * U(r, q);
						gamma += U(r, p) * U(r, q);
					}
					if (gamma == 0 || std::abs(gamma) <= 1e-17 * std::sqrt(alpha * beta)) continue;
					rotated = true;
					const P zeta = (beta - alpha) / (2 * gamma);
					const P t = (zeta >= 0 ? 1 : -1) / (std::abs(zeta) + std::sqrt(1 + zeta * zeta));
					const P c = 1 / std::sqrt(1 + t * t), s = c * t;
					for (int r = 0; r < Rows; ++r) {
						const P up = U(r, p), uq = U(r, q);
						U(r, p) = c * up - s * uq;
						U(r, q) = s * up + c * uq;
					}
					for (int r = 0; r < Cols; ++r) {
						const P vp = V(r, p), vq = V(r, q);
						V(r, p) = c * vp - s * vq;
						V(r, q) = s * vp + c * vq;
					}
				}
			if (!rotated) break;
		}
		for (int c = 0; c < Cols; ++c) {
			P n = 0;
			for (int r = 0; r < Rows; ++r) n += U(r, c) * U(r, c);
			n = std::sqrt(n);
			W[c] = n;
			if (n > 0) for (int r = 0; r < Rows; ++r) U(r, c) /= n;
		}
	}

	template <class P2, class B2> Vector<Cols, P> backsub(const Vector<Rows, P2, B2>& b, const P condition = 1e9) const {
		P wmax = 0;
		for (int c = 0; c < Cols; ++c) wmax = std::max(wmax, std::abs(W[c]));
		Vector<Cols, P> x;
		P y[Cols];
		for (int c = 0; c < Cols; ++c) {
			P utb = 0;
			for (int r = 0; r < Rows; ++r) utb += U(r, c) * b[r];
			const P inv = (W[c] * condition > wmax) ? static_cast<P>(1) / W[c] : 0;
			y[c] = inv * utb;
		}
		for (int r = 0; r < Cols; ++r) {
			P s = 0;
			for (int c = 0; c < Cols; ++c) s += V(r, c) * y[c];
			x[r] = s;
		}
		return x;
	}

	const Matrix<Rows, Cols, P>& get_U() const { return U; }
	const Vector<Cols, P>& get_diagonal() const { return W; }
	const Matrix<Cols, Cols, P>& get_V() const { return V; }

private:
	Matrix<Rows, Cols, P> U;
	Matrix<Cols, Cols, P> V;
	Vector<Cols, P> W;
};

}  // namespace TooN
#endif
