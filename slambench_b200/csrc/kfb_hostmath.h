// Pose algebra of the product path: the host side of updatePoseKernel / checkPoseKernel
// (kfusion/src/cpp/kernels.cpp:759-792) and the 4x4 helpers of commons.h:343-412.
// In the reference these go through TooN (external, pinned 92241416...; not in the tree):
//   inverse()        -> TooN::gaussian_elimination (float, partial pivoting, double factor)
//   operator*        -> TooN fixed-size float product (k innermost)
//   solve()          -> TooN::GR_SVD<6,6>::backsub(b, 1e6) in double
//   SE3<>::exp       -> Rodrigues with TooN's small-angle branches
// Every function is __host__ __device__: the same code runs on the CPU (host-solve mode,
// integrate/raycast matrices) and inside the ICP kernel (device-solve mode).
#ifndef KFB_HOSTMATH_H
#define KFB_HOSTMATH_H

#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define KFB_HM __host__ __device__ inline
#else
#define KFB_HM static inline
#endif

// commons.h:343-350
KFB_HM void hm_camera_matrix(float* K, const float* k) {
	for (int i = 0; i < 16; ++i) K[i] = 0.f;
	K[0] = k[0]; K[2] = k[2]; K[5] = k[1]; K[6] = k[3]; K[10] = 1.f; K[15] = 1.f;
}
// commons.h:352-359
KFB_HM void hm_inverse_camera_matrix(float* K, const float* k) {
	for (int i = 0; i < 16; ++i) K[i] = 0.f;
	K[0] = 1.0f / k[0]; K[2] = -k[2] / k[0]; K[5] = 1.0f / k[1]; K[6] = -k[3] / k[1]; K[10] = 1.f; K[15] = 1.f;
}
// commons.h:373-378
KFB_HM void hm_matmul4(float* out, const float* a, const float* b) {
	float r[16];
	for (int i = 0; i < 4; ++i)
		for (int j = 0; j < 4; ++j) {
			float s = 0.f;
			for (int k = 0; k < 4; ++k) s = s + a[4 * i + k] * b[4 * k + j];
			r[4 * i + j] = s;
		}
	for (int i = 0; i < 16; ++i) out[i] = r[i];
}
// commons.h:365-371.  No singularity check on purpose: the all-zero raycastPose of frames
// 0-3 must yield NaN so that tracking rejects every pixel exactly like the reference does.
KFB_HM void hm_inverse4(float* out, const float* in) {
	float A[4][4], b[4][4], x[4][4];
	for (int i = 0; i < 4; ++i)
		for (int j = 0; j < 4; ++j) { A[i][j] = in[4 * i + j]; b[i][j] = (i == j) ? 1.f : 0.f; }
	for (int i = 0; i < 4; ++i) {
		int arg = i;
		float maxval = fabsf(A[i][i]);
		for (int ii = i + 1; ii < 4; ++ii) {
			const double v = fabsf(A[ii][i]);
			if (v > maxval) { maxval = (float) v; arg = ii; }
		}
		const float inv_pivot = 1.0f / A[arg][i];
		if (arg != i) {
			for (int j = i; j < 4; ++j) { const float t = A[i][j]; A[i][j] = A[arg][j]; A[arg][j] = t; }
			for (int j = 0; j < 4; ++j) { const float t = b[i][j]; b[i][j] = b[arg][j]; b[arg][j] = t; }
		}
		for (int j = i + 1; j < 4; ++j) A[i][j] *= inv_pivot;
		for (int j = 0; j < 4; ++j) b[i][j] *= inv_pivot;
		for (int u = i + 1; u < 4; ++u) {
			const double factor = A[u][i];
			for (int j = i + 1; j < 4; ++j) A[u][j] = (float) ((double) A[u][j] - factor * (double) A[i][j]);
			for (int j = 0; j < 4; ++j) b[u][j] = (float) ((double) b[u][j] - factor * (double) b[i][j]);
		}
	}
	for (int i = 3; i >= 0; --i) {
		for (int c = 0; c < 4; ++c) x[i][c] = b[i][c];
		for (int j = i + 1; j < 4; ++j)
			for (int c = 0; c < 4; ++c) x[i][c] = x[i][c] - A[i][j] * x[j][c];
	}
	for (int i = 0; i < 4; ++i)
		for (int j = 0; j < 4; ++j) out[4 * i + j] = x[i][j];
}

// commons.h:380-404: x = V diag(w_i*1e6 > w_max ? 1/w_i : 0) U^T b for the symmetric PSD
// 6x6 JtJ; cyclic Jacobi eigen-decomposition in double.
KFB_HM void hm_solve6(double* x6, const float* vals27) {
	double b[6], C[6][6], V[6][6];
	for (int i = 0; i < 6; ++i) b[i] = vals27[i];
	int idx = 6;
	for (int r = 0; r < 6; ++r)
		for (int c = r; c < 6; ++c) { C[r][c] = vals27[idx++]; C[c][r] = C[r][c]; }
	for (int r = 0; r < 6; ++r)
		for (int c = 0; c < 6; ++c) V[r][c] = (r == c) ? 1.0 : 0.0;
	for (int sweep = 0; sweep < 64; ++sweep) {
		double off = 0;
		for (int p = 0; p < 5; ++p)
			for (int q = p + 1; q < 6; ++q) off += C[p][q] * C[p][q];
		if (off == 0) break;
		for (int p = 0; p < 5; ++p)
			for (int q = p + 1; q < 6; ++q) {
				const double apq = C[p][q];
				if (apq == 0) continue;
				const double theta = (C[q][q] - C[p][p]) / (2 * apq);
				const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
				const double c = 1 / sqrt(t * t + 1), s = t * c;
				for (int k = 0; k < 6; ++k) { const double kp = C[k][p], kq = C[k][q]; C[k][p] = c * kp - s * kq; C[k][q] = s * kp + c * kq; }
				for (int k = 0; k < 6; ++k) { const double pk = C[p][k], qk = C[q][k]; C[p][k] = c * pk - s * qk; C[q][k] = s * pk + c * qk; }
				for (int k = 0; k < 6; ++k) { const double kp = V[k][p], kq = V[k][q]; V[k][p] = c * kp - s * kq; V[k][q] = s * kp + c * kq; }
			}
	}
	double wmax = 0;
	for (int i = 0; i < 6; ++i) { const double w = fabs(C[i][i]); if (w > wmax) wmax = w; }
	double y[6];
	for (int i = 0; i < 6; ++i) {
		const double lam = C[i][i];
		double vtb = 0;
		for (int k = 0; k < 6; ++k) vtb += V[k][i] * b[k];
		y[i] = (fabs(lam) * 1e6 > wmax) ? vtb / lam : 0.0;
	}
	for (int r = 0; r < 6; ++r) { double s = 0; for (int i = 0; i < 6; ++i) s += V[r][i] * y[i]; x6[r] = s; }
}

// Fast path of the same solve for the device-resident ICP loop.  For a symmetric positive-definite
// JtJ whose condition number is certified below the reference's pseudo-inverse cut-off
// (cond <= trace(C) * trace(C^-1) < 1e6  =>  every singular value satisfies sigma * 1e6 > sigma_max
// => GR_SVD::backsub keeps them all => x = C^-1 b), a Cholesky solve in double gives the same x to
// ~cond * 1e-16.  Returns 0 (caller falls back to hm_solve6) when the certificate fails: not
// positive definite (start-up frames: C == 0), NaN, or possibly ill-conditioned.
#define KFB_TRI(i, j) ((i) * ((i) + 1) / 2 + (j))   // packed lower triangle, j <= i
KFB_HM int hm_solve6_chol(double* x6, const float* vals27) {
	// packed lower triangles only (21 + 21 doubles) so that the whole solve stays in registers on the device
	double b[6], L[21], M[21];
#pragma unroll
	for (int i = 0; i < 6; ++i) b[i] = vals27[i];
	{
		// vals27[6..26] is the UPPER triangle row by row (commons.h:385-392): C[r][c], c >= r  ->  L slot (c, r)
		int idx = 6;
#pragma unroll
		for (int r = 0; r < 6; ++r)
#pragma unroll
			for (int c = r; c < 6; ++c) L[KFB_TRI(c, r)] = vals27[idx++];
	}
	double trC = 0;
#pragma unroll
	for (int i = 0; i < 6; ++i) trC += L[KFB_TRI(i, i)];
	// in-place Cholesky, column by column; M's diagonal takes the reciprocal pivots
#pragma unroll
	for (int j = 0; j < 6; ++j) {
		double s = L[KFB_TRI(j, j)];
#pragma unroll
		for (int k = 0; k < j; ++k) s -= L[KFB_TRI(j, k)] * L[KFB_TRI(j, k)];
		if (!(s > 0)) return 0;
#if defined(__CUDA_ARCH__)
		const double inv = rsqrt(s), d = s * inv;   // one special-function step on the serial path instead of sqrt + divide
#else
		const double d = sqrt(s), inv = 1.0 / d;
#endif
		L[KFB_TRI(j, j)] = d; M[KFB_TRI(j, j)] = inv;
#pragma unroll
		for (int i = j + 1; i < 6; ++i) {
			double t = L[KFB_TRI(i, j)];
#pragma unroll
			for (int k = 0; k < j; ++k) t -= L[KFB_TRI(i, k)] * L[KFB_TRI(j, k)];
			L[KFB_TRI(i, j)] = t * inv;
		}
	}
	// M = L^-1 (lower triangular); trace(C^-1) = ||M||_F^2
	double trInv = 0;
#pragma unroll
	for (int j = 0; j < 6; ++j) {
		trInv += M[KFB_TRI(j, j)] * M[KFB_TRI(j, j)];
#pragma unroll
		for (int i = j + 1; i < 6; ++i) {
			double t = 0;
#pragma unroll
			for (int k = j; k < i; ++k) t -= L[KFB_TRI(i, k)] * M[KFB_TRI(k, j)];
			M[KFB_TRI(i, j)] = t * M[KFB_TRI(i, i)];
			trInv += M[KFB_TRI(i, j)] * M[KFB_TRI(i, j)];
		}
	}
	if (!(trC * trInv < 0.99e6)) return 0;
	double y[6];
#pragma unroll
	for (int i = 0; i < 6; ++i) {
		double t = 0;
#pragma unroll
		for (int j = 0; j <= i; ++j) t += M[KFB_TRI(i, j)] * b[j];
		y[i] = t;
	}
#pragma unroll
	for (int j = 0; j < 6; ++j) {
		double t = 0;
#pragma unroll
		for (int i = j; i < 6; ++i) t += M[KFB_TRI(i, j)] * y[i];
		x6[j] = t;
	}
	return 1;
}

// TooN::SE3<double>::exp followed by toMatrix4 (commons.h:406-412)
KFB_HM void hm_se3_exp(float* out16, const double* mu) {
	const double w0 = mu[3], w1 = mu[4], w2 = mu[5], t0 = mu[0], t1 = mu[1], t2 = mu[2];
	const double theta_sq = w0 * w0 + w1 * w1 + w2 * w2;
	const double theta = sqrt(theta_sq);
	const double c0 = w1 * t2 - w2 * t1, c1 = w2 * t0 - w0 * t2, c2 = w0 * t1 - w1 * t0;
	double A, B, T0, T1, T2;
	if (theta_sq < 1e-8) {
		A = 1.0 - (1.0 / 6.0) * theta_sq;
		B = 0.5;
		T0 = t0 + 0.5 * c0; T1 = t1 + 0.5 * c1; T2 = t2 + 0.5 * c2;
	} else {
		double C;
		if (theta_sq < 1e-6) {
			C = (1.0 / 6.0) * (1.0 - (1.0 / 20.0) * theta_sq);
			A = 1.0 - theta_sq * C;
			B = 0.5 - 0.25 * (1.0 / 6.0) * theta_sq;
		} else {
			const double inv_theta = 1.0 / theta;
			A = sin(theta) * inv_theta;
			B = (1 - cos(theta)) * (inv_theta * inv_theta);
			C = (1 - A) * (inv_theta * inv_theta);
		}
		const double d0 = w1 * c2 - w2 * c1, d1 = w2 * c0 - w0 * c2, d2 = w0 * c1 - w1 * c0;
		T0 = t0 + B * c0 + C * d0; T1 = t1 + B * c1 + C * d1; T2 = t2 + B * c2 + C * d2;
	}
	const double wx2 = w0 * w0, wy2 = w1 * w1, wz2 = w2 * w2;
	double R[3][3];
	R[0][0] = 1.0 - B * (wy2 + wz2); R[1][1] = 1.0 - B * (wx2 + wz2); R[2][2] = 1.0 - B * (wx2 + wy2);
	double a = A * w2, b = B * (w0 * w1);
	R[0][1] = b - a; R[1][0] = b + a;
	a = A * w1; b = B * (w0 * w2);
	R[0][2] = b + a; R[2][0] = b - a;
	a = A * w0; b = B * (w1 * w2);
	R[1][2] = b - a; R[2][1] = b + a;
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c) out16[4 * r + c] = (float) R[r][c];
	out16[3] = (float) T0; out16[7] = (float) T1; out16[11] = (float) T2;
	out16[12] = 0.f; out16[13] = 0.f; out16[14] = 0.f; out16[15] = 1.f;
}

// cpp/kernels.cpp:759-775.  `red` = the 32 reduced sums; returns 1 when ||x|| < icp_threshold.
KFB_HM int hm_update_pose(float* pose, const float* red, float icp_threshold) {
	double x[6];
	hm_solve6(x, red + 1);
	float d[16];
	hm_se3_exp(d, x);
	hm_matmul4(pose, d, pose);
	double n = 0;
	for (int i = 0; i < 6; ++i) n += x[i] * x[i];
	return sqrt(n) < (double) icp_threshold;
}
// same, with the certified Cholesky fast path (device-resident ICP loop)
KFB_HM int hm_update_pose_fast(float* pose, const float* red, float icp_threshold) {
	double x[6];
	if (!hm_solve6_chol(x, red + 1)) hm_solve6(x, red + 1);
	float d[16];
	hm_se3_exp(d, x);
	hm_matmul4(pose, d, pose);
	double n = 0;
	for (int i = 0; i < 6; ++i) n += x[i] * x[i];
	return sqrt(n) < (double) icp_threshold;
}
// cpp/kernels.cpp:777-792.  NaN (0/0 when nothing tracked) compares false, so the inlier-ratio
// test decides — exactly the start-up behaviour of frames 0-3.
KFB_HM int hm_check_pose(float* pose, const float* old_pose, const float* red, unsigned w, unsigned h, float track_threshold) {
	if (((double) sqrtf(red[0] / red[28]) > 2e-2) || (red[28] / (float) (w * h) < track_threshold)) {
		for (int i = 0; i < 16; ++i) pose[i] = old_pose[i];
		return 0;
	}
	return 1;
}

#endif
