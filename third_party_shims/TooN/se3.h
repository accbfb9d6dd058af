// TooN stand-in: SE3<P> — see TooN.h in this directory for why this exists.
// TEST INFRASTRUCTURE, not product code.
//
// Restates the published SE3 exponential map used by TooN: the 6-vector is
// (translation[3], rotation[3]); rotation by Rodrigues' formula with Taylor
// branches for theta^2 < 1e-8 and < 1e-6.  Call sites in the reference:
// kernels.h:106-109 (initial pose), cpp/kernels.cpp:766-767 (pose update),
// commons.h:406-412 (toMatrix4 = SE3 * Identity4).
#ifndef TOON_SHIM_SE3_H
#define TOON_SHIM_SE3_H

#include <TooN/TooN.h>

namespace TooN {

template <class P = double> class SE3 {
public:
	SE3() {
		for (int r = 0; r < 3; ++r) { t[r] = 0; for (int c = 0; c < 3; ++c) R(r, c) = (r == c) ? 1 : 0; }
	}
	template <class P2, class B2> SE3(const Vector<6, P2, B2>& v) { *this = exp(v); }

	template <class P2, class B2> static SE3 exp(const Vector<6, P2, B2>& mu) {
		using std::sqrt; using std::sin; using std::cos;
		static const P one_6th = 1.0 / 6.0;
		static const P one_20th = 1.0 / 20.0;
		SE3 result;
		Vector<3, P> w, tr;
		for (int i = 0; i < 3; ++i) { tr[i] = mu[i]; w[i] = mu[i + 3]; }
		const P theta_sq = w * w;
		const P theta = sqrt(theta_sq);
		P A, B;
		const Vector<3, P> cross = w ^ tr;
		if (theta_sq < 1e-8) {
			A = 1.0 - one_6th * theta_sq;
			B = 0.5;
			for (int i = 0; i < 3; ++i) result.t[i] = tr[i] + 0.5 * cross[i];
		} else {
			P C;
			if (theta_sq < 1e-6) {
				C = one_6th * (1.0 - one_20th * theta_sq);
				A = 1.0 - theta_sq * C;
				B = 0.5 - 0.25 * one_6th * theta_sq;
			} else {
				const P inv_theta = 1.0 / theta;
				A = sin(theta) * inv_theta;
				B = (1 - cos(theta)) * (inv_theta * inv_theta);
				C = (1 - A) * (inv_theta * inv_theta);
			}
			const Vector<3, P> wcross = w ^ cross;
			for (int i = 0; i < 3; ++i) result.t[i] = tr[i] + B * cross[i] + C * wcross[i];
		}
		// Rodrigues
		{
			const P wx2 = w[0] * w[0], wy2 = w[1] * w[1], wz2 = w[2] * w[2];
			result.R(0, 0) = 1.0 - B * (wy2 + wz2);
			result.R(1, 1) = 1.0 - B * (wx2 + wz2);
			result.R(2, 2) = 1.0 - B * (wx2 + wy2);
		}
		{
			const P a = A * w[2], b = B * (w[0] * w[1]);
			result.R(0, 1) = b - a;
			result.R(1, 0) = b + a;
		}
		{
			const P a = A * w[1], b = B * (w[0] * w[2]);
			result.R(0, 2) = b + a;
			result.R(2, 0) = b - a;
		}
		{
			const P a = A * w[0], b = B * (w[1] * w[2]);
			result.R(1, 2) = b - a;
			result.R(2, 1) = b + a;
		}
		return result;
	}

	const Matrix<3, 3, P>& get_rotation() const { return R; }
	const Vector<3, P>& get_translation() const { return t; }

	Matrix<3, 3, P> R;
	Vector<3, P> t;
};

// SE3 * (4 x C matrix): rows 0..2 = R * M[0..2] + t (x) M[3]; row 3 = M[3].
template <class P, int C, class P2, class L2>
Matrix<4, C, typename Internal::Promote<P, P2>::type> operator*(const SE3<P>& s, const Matrix<4, C, P2, L2>& m) {
	typedef typename Internal::Promote<P, P2>::type PR;
	Matrix<4, C, PR> out;
	for (int c = 0; c < C; ++c) {
		for (int r = 0; r < 3; ++r) {
			PR acc = 0;
			for (int k = 0; k < 3; ++k) acc += s.R(r, k) * m(k, c);
			out(r, c) = acc + s.t[r] * m(3, c);
		}
		out(3, c) = m(3, c);
	}
	return out;
}

}  // namespace TooN
#endif
