/* ORACLE — TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product path
 * (slambench_b200/csrc, libkfb200.so) never does and has no CPU fallback.
 *
 * Plain-C restatement of the reference's KinectFusion per-frame pipeline
 * (domantasjurkus/slambench, kfusion/src/cpp/kernels.cpp + kfusion/include/commons.h
 * + kfusion/thirdparty/cutil_math.h).  Every function cites the reference lines it
 * follows.  Arithmetic is un-fused IEEE fp32 in the reference's operation order
 * (compile with -ffp-contract=off; the reference is built for baseline x86-64, which
 * has no FMA).
 *
 * PINNING: tests/test_oracle_vs_ref.py compares every function here, bit for bit,
 * against oracle/_ref/libkfusion_ref.so (the unmodified reference sources compiled in
 * place, see oracle/Makefile) whenever that library is present, and
 * tests/test_oracle_golden.py compares it against the golden vectors in tests/golden/
 * that were generated from that same reference build (tests/golden/make_golden.py).
 * The reference ships NO tests or golden vectors of its own.
 * The one boundary that stays "parity unpinned" is the TooN dependency (external,
 * pinned at 92241416d2a4874fd2334e08a5d417dfea6a1a3f in the reference Makefile:22,
 * absent from /root/reference): the 4x4 float Gaussian-elimination inverse, the SE3
 * exponential and the 6x6 SVD back-substitution are restated here from TooN's
 * published algorithms and cross-checked only against our own TooN stand-in.
 *
 * Exported names (kfo_*) are identical to oracle/ref_harness.cpp so one Python
 * wrapper (tests/cpu_backend.py) drives either library.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y, z; } f3;
typedef struct { short x, y; } s2;
typedef struct { int result; float error; float J[6]; } TrackData; /* commons.h:325-329 */
typedef struct { unsigned sx, sy, sz; float dx, dy, dz; s2* data; } Vol; /* commons.h:149-153 */

#define KF_INVALID (-2.0f) /* commons.h:14 */
/* constant_parameters.h:15-23 */
static const float c_e_delta = 0.1f;
static const int c_radius = 2;
static const float c_dist_threshold = 0.1f;
static const float c_normal_threshold = 0.8f;
static const float c_track_threshold = 0.15f;
static const float c_maxweight = 100.0f;
static const float c_nearPlane = 0.4f;
static const float c_farPlane = 4.0f;
static const float c_delta = 4.0f;
static const f3 c_light = { 1.f, 1.f, -1.f };
static const f3 c_ambient = { 0.1f, 0.1f, 0.1f };

/* cutil_math.h:43-57 — host fminf/fmaxf/min/max are plain ternaries */
static inline float kminf(float a, float b) { return a < b ? a : b; }
static inline float kmaxf(float a, float b) { return a > b ? a : b; }
static inline int kmini(int a, int b) { return a < b ? a : b; }
static inline int kmaxi(int a, int b) { return a > b ? a : b; }
static inline float kclampf(float f, float a, float b) { return kmaxf(a, kminf(f, b)); } /* :972 */
/* cutil_math.h:978-980 clamp(uint,uint,uint): on the host `min`/`max` resolve to the
 * (int,int) overloads (:51-57), so a wrapped-around x-1 == 0xffffffff becomes -1 and
 * clamps to 0. */
static inline unsigned kclampu(unsigned f, unsigned a, unsigned b) { return (unsigned) kmaxi((int) a, kmini((int) f, (int) b)); }
static inline float sq(float r) { return r * r; } /* commons.h:82 */
static inline float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } /* cutil_math.h:1063-1071 */
static inline f3 mk3(float x, float y, float z) { f3 r = { x, y, z }; return r; }
static inline f3 sub3(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline f3 add3(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline f3 mul3s(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
static inline f3 mul3(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline f3 cross3(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); } /* :1244 */
static inline float length3(f3 v) { return sqrtf(dot3(v, v)); } /* :1117-1124 */
/* cutil_math.h:1147-1151 with host rsqrtf = 1.0f / sqrtf(x) (:59-61) */
static inline f3 normalize3(f3 v) { float inv = 1.0f / sqrtf(dot3(v, v)); return mul3s(v, inv); }

/* Matrix4 is 4 rows of float4 (commons.h:317-319): m[4*r + c] */
static inline f3 mat_mul_pt(const float* M, f3 v) { /* commons.h:331-336 */
	return mk3(dot3(mk3(M[0], M[1], M[2]), v) + M[3], dot3(mk3(M[4], M[5], M[6]), v) + M[7],
			dot3(mk3(M[8], M[9], M[10]), v) + M[11]);
}
static inline f3 mat_rotate(const float* M, f3 v) { /* commons.h:338-341 */
	return mk3(dot3(mk3(M[0], M[1], M[2]), v), dot3(mk3(M[4], M[5], M[6]), v), dot3(mk3(M[8], M[9], M[10]), v));
}

const char* kfo_impl_name(void) { return "oracle-port"; }

/* ------------------------------------------------------------------ host 4x4 math */
void kfo_camera_matrix(float* K, const float* k) { /* commons.h:343-350 */
	const float m[16] = { k[0], 0, k[2], 0, 0, k[1], k[3], 0, 0, 0, 1, 0, 0, 0, 0, 1 };
	memcpy(K, m, sizeof m);
}
void kfo_inverse_camera_matrix(float* K, const float* k) { /* commons.h:352-359 */
	const float m[16] = { 1.0f / k[0], 0, -k[2] / k[0], 0, 0, 1.0f / k[1], -k[3] / k[1], 0, 0, 0, 1, 0, 0, 0, 0, 1 };
	memcpy(K, m, sizeof m);
}
void kfo_matmul(float* out, const float* a, const float* b) { /* commons.h:373-378 (TooN float product, k innermost) */
	float r[16];
	for (int i = 0; i < 4; ++i)
		for (int j = 0; j < 4; ++j) {
			float s = 0;
			for (int k = 0; k < 4; ++k) s += a[4 * i + k] * b[4 * k + j];
			r[4 * i + j] = s;
		}
	memcpy(out, r, sizeof r);
}
/* commons.h:365-371: TooN::gaussian_elimination(A, Identity) in float with partial
 * pivoting, elimination factor held in a double, no singularity check (a zero matrix
 * yields NaN, which the start-up frames rely on — SURVEY §8a a18). */
void kfo_inverse(float* out, const float* in) {
	float A[4][4], b[4][4], x[4][4];
	for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { A[i][j] = in[4 * i + j]; b[i][j] = (i == j) ? 1.f : 0.f; }
	for (int i = 0; i < 4; ++i) {
		int argmax = i;
		float maxval = fabsf(A[i][i]);
		for (int ii = i + 1; ii < 4; ++ii) {
			double v = fabsf(A[ii][i]);
			if (v > maxval) { maxval = (float) v; argmax = ii; }
		}
		float pivot = A[argmax][i];
		float inv_pivot = 1.0f / pivot;
		if (argmax != i) {
			for (int j = i; j < 4; ++j) { float t = A[i][j]; A[i][j] = A[argmax][j]; A[argmax][j] = t; }
			for (int j = 0; j < 4; ++j) { float t = b[i][j]; b[i][j] = b[argmax][j]; b[argmax][j] = t; }
		}
		for (int j = i + 1; j < 4; ++j) A[i][j] *= inv_pivot;
		for (int j = 0; j < 4; ++j) b[i][j] *= inv_pivot;
		for (int u = i + 1; u < 4; ++u) {
			double factor = A[u][i];
			for (int j = i + 1; j < 4; ++j) A[u][j] = (float) (A[u][j] - factor * A[i][j]);
			for (int j = 0; j < 4; ++j) b[u][j] = (float) (b[u][j] - factor * b[i][j]);
		}
	}
	for (int i = 3; i >= 0; --i) {
		for (int c = 0; c < 4; ++c) x[i][c] = b[i][c];
		for (int j = i + 1; j < 4; ++j)
			for (int c = 0; c < 4; ++c) x[i][c] -= A[i][j] * x[j][c];
	}
	for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out[4 * i + j] = x[i][j];
}

/* commons.h:380-404: x = pinv(JtJ) * Jte in double; singular values with
 * w * 1e6 <= w_max are dropped (TooN GR_SVD::backsub(b, 1e6)).  JtJ is symmetric PSD,
 * so a cyclic Jacobi eigen-decomposition gives the SVD directly (w = |lambda|). */
void kfo_solve(double* x6, const float* vals27) {
	double b[6], C[6][6], V[6][6];
	for (int i = 0; i < 6; ++i) b[i] = vals27[i];
	/* makeJTJ: upper triangle row-major from vals[6..26], mirrored (commons.h:381-395) */
	int idx = 6;
	for (int r = 0; r < 6; ++r) for (int c = r; c < 6; ++c) { C[r][c] = vals27[idx++]; C[c][r] = C[r][c]; }
	for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) V[r][c] = (r == c);
	for (int sweep = 0; sweep < 64; ++sweep) {
		double off = 0;
		for (int p = 0; p < 5; ++p) for (int q = p + 1; q < 6; ++q) off += C[p][q] * C[p][q];
		if (off == 0) break;
		for (int p = 0; p < 5; ++p)
			for (int q = p + 1; q < 6; ++q) {
				if (C[p][q] == 0) continue;
				double theta = (C[q][q] - C[p][p]) / (2 * C[p][q]);
				double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
				double c = 1 / sqrt(t * t + 1), s = t * c;
				for (int k = 0; k < 6; ++k) { double kp = C[k][p], kq = C[k][q]; C[k][p] = c * kp - s * kq; C[k][q] = s * kp + c * kq; }
				for (int k = 0; k < 6; ++k) { double pk = C[p][k], qk = C[q][k]; C[p][k] = c * pk - s * qk; C[q][k] = s * pk + c * qk; }
				for (int k = 0; k < 6; ++k) { double kp = V[k][p], kq = V[k][q]; V[k][p] = c * kp - s * kq; V[k][q] = s * kp + c * kq; }
			}
	}
	double wmax = 0;
	for (int i = 0; i < 6; ++i) if (fabs(C[i][i]) > wmax) wmax = fabs(C[i][i]);
	double y[6];
	for (int i = 0; i < 6; ++i) {
		double lam = C[i][i], w = fabs(lam), vtb = 0;
		for (int k = 0; k < 6; ++k) vtb += V[k][i] * b[k];
		y[i] = (w * 1e6 > wmax) ? vtb / lam : 0.0;
	}
	for (int r = 0; r < 6; ++r) { double s = 0; for (int i = 0; i < 6; ++i) s += V[r][i] * y[i]; x6[r] = s; }
}

/* TooN::SE3<double>::exp then toMatrix4 (commons.h:406-412): x = (translation, rotation) */
void kfo_se3_exp(float* out16, const double* mu) {
	const double w[3] = { mu[3], mu[4], mu[5] }, tr[3] = { mu[0], mu[1], mu[2] };
	const double theta_sq = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
	const double theta = sqrt(theta_sq);
	const double cr[3] = { w[1] * tr[2] - w[2] * tr[1], w[2] * tr[0] - w[0] * tr[2], w[0] * tr[1] - w[1] * tr[0] };
	double A, B, t[3], R[3][3];
	if (theta_sq < 1e-8) {
		A = 1.0 - (1.0 / 6.0) * theta_sq;
		B = 0.5;
		for (int i = 0; i < 3; ++i) t[i] = tr[i] + 0.5 * cr[i];
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
		const double wc[3] = { w[1] * cr[2] - w[2] * cr[1], w[2] * cr[0] - w[0] * cr[2], w[0] * cr[1] - w[1] * cr[0] };
		for (int i = 0; i < 3; ++i) t[i] = tr[i] + B * cr[i] + C * wc[i];
	}
	{
		const double wx2 = w[0] * w[0], wy2 = w[1] * w[1], wz2 = w[2] * w[2];
		R[0][0] = 1.0 - B * (wy2 + wz2); R[1][1] = 1.0 - B * (wx2 + wz2); R[2][2] = 1.0 - B * (wx2 + wy2);
		double a = A * w[2], b = B * (w[0] * w[1]);
		R[0][1] = b - a; R[1][0] = b + a;
		a = A * w[1]; b = B * (w[0] * w[2]);
		R[0][2] = b + a; R[2][0] = b - a;
		a = A * w[0]; b = B * (w[1] * w[2]);
		R[1][2] = b - a; R[2][1] = b + a;
	}
	for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) out16[4 * r + c] = (float) R[r][c]; out16[4 * r + 3] = (float) t[r]; }
	out16[12] = 0; out16[13] = 0; out16[14] = 0; out16[15] = 1;
}

/* ----------------------------------------------------------------------- kernels */
void kfo_init_volume(short* data, const unsigned* size, const float* dim) { /* cpp/kernels.cpp:147-157, commons.h:176-180 */
	(void) dim;
	size_t n = (size_t) size[0] * size[1] * size[2];
	s2* d = (s2*) data;
	for (size_t i = 0; i < n; ++i) { d[i].x = (short) (1.0f * 32766.0f); d[i].y = (short) 0.0f; }
}

void kfo_gaussian(float* out5) { /* cpp/kernels.cpp:101-107: integer x, integer -(x*x) */
	for (unsigned i = 0; i < (unsigned) (c_radius * 2 + 1); i++) {
		int x = (int) i - 2;
		out5[i] = expf(-(x * x) / (2 * c_delta * c_delta));
	}
}

void kfo_mm2meters(float* out, unsigned ow, unsigned oh, const unsigned short* in, unsigned iw, unsigned ih) { /* :562-589 */
	if (iw < ow || ih < oh || iw % ow != 0 || ih % oh != 0 || iw / ow != ih / oh) {
		fprintf(stderr, "Invalid ratio.\n");
		exit(1);
	}
	int ratio = iw / ow;
#pragma omp parallel for
	for (unsigned y = 0; y < oh; y++)
		for (unsigned x = 0; x < ow; x++) out[x + ow * y] = in[x * ratio + iw * y * ratio] / 1000.0f;
}

void kfo_bilateral(float* out, const float* in, unsigned w, unsigned h, const float* gauss, float e_d, int r) { /* :159-198 */
	const float e_d_squared_2 = e_d * e_d * 2;
#pragma omp parallel for
	for (unsigned y = 0; y < h; y++)
		for (unsigned x = 0; x < w; x++) {
			unsigned pos = x + y * w;
			if (in[pos] == 0) { out[pos] = 0; continue; }
			float sum = 0.0f, t = 0.0f;
			const float center = in[pos];
			for (int i = -r; i <= r; ++i)
				for (int j = -r; j <= r; ++j) {
					unsigned cx = kclampu(x + i, 0u, w - 1), cy = kclampu(y + j, 0u, h - 1);
					const float curPix = in[cx + cy * w];
					if (curPix > 0) {
						const float mod = sq(curPix - center);
						const float factor = gauss[i + r] * gauss[j + r] * expf(-mod / e_d_squared_2);
						t += factor * curPix;
						sum += factor;
					}
				}
			out[pos] = t / sum;
		}
}

void kfo_halfsample(float* out, const float* in, unsigned iw, unsigned ih, float e_d, int r) { /* :591-626 */
	const unsigned ow = iw / 2, oh = ih / 2;
#pragma omp parallel for
	for (unsigned y = 0; y < oh; y++)
		for (unsigned x = 0; x < ow; x++) {
			const unsigned cx = 2 * x, cy = 2 * y;
			float sum = 0.0f, t = 0.0f;
			const float center = in[cx + cy * iw];
			for (int i = -r + 1; i <= r; ++i)
				for (int j = -r + 1; j <= r; ++j) {
					int px = kmaxi(0, kmini((int) cx + j, 2 * (int) ow - 1));
					int py = kmaxi(0, kmini((int) cy + i, 2 * (int) oh - 1));
					float current = in[px + py * iw];
					if (fabsf(current - center) < e_d) { sum += 1.0f; t += current; }
				}
			out[x + y * ow] = t / sum;
		}
}

void kfo_depth2vertex(float* vtx, const float* depth, unsigned w, unsigned h, const float* invK) { /* :200-218 */
	f3* v = (f3*) vtx;
#pragma omp parallel for
	for (unsigned y = 0; y < h; y++)
		for (unsigned x = 0; x < w; x++) {
			if (depth[x + y * w] > 0) v[x + y * w] = mul3s(mat_rotate(invK, mk3((float) x, (float) y, 1.f)), depth[x + y * w]);
			else v[x + y * w] = mk3(0, 0, 0);
		}
}

void kfo_vertex2normal(float* out_, const float* in_, unsigned w, unsigned h) { /* :220-249 */
	f3* out = (f3*) out_;
	const f3* in = (const f3*) in_;
#pragma omp parallel for
	for (unsigned y = 0; y < h; y++)
		for (unsigned x = 0; x < w; x++) {
			const unsigned lx = kmaxi((int) x - 1, 0), rx = kmini(x + 1, (int) w - 1);
			const unsigned uy = kmaxi((int) y - 1, 0), dy = kmini(y + 1, (int) h - 1);
			const f3 left = in[lx + w * y], right = in[rx + w * y], up = in[x + w * uy], down = in[x + w * dy];
			if (left.z == 0 || right.z == 0 || up.z == 0 || down.z == 0) { out[x + y * w].x = KF_INVALID; continue; } /* only .x */
			const f3 dxv = sub3(right, left), dyv = sub3(down, up);
			out[x + y * w] = normalize3(cross3(dyv, dxv));
		}
}

void kfo_track(void* trackdata, const float* inV_, const float* inN_, unsigned w, unsigned h, const float* refV_,
		const float* refN_, unsigned rw, unsigned rh, const float* Ttrack, const float* view, float dist_thr,
		float normal_thr) { /* :497-560 */
	TrackData* output = (TrackData*) trackdata;
	const f3 *inV = (const f3*) inV_, *inN = (const f3*) inN_, *refV = (const f3*) refV_, *refN = (const f3*) refN_;
#pragma omp parallel for
	for (unsigned py = 0; py < h; py++)
		for (unsigned px = 0; px < w; px++) {
			TrackData* row = &output[px + py * rw]; /* stride = reference (full-res) width */
			if (inN[px + py * w].x == KF_INVALID) { row->result = -1; continue; }
			const f3 projectedVertex = mat_mul_pt(Ttrack, inV[px + py * w]);
			const f3 projectedPos = mat_mul_pt(view, projectedVertex);
			const float ppx = projectedPos.x / projectedPos.z + 0.5f, ppy = projectedPos.y / projectedPos.z + 0.5f;
			if (ppx < 0 || ppx > rw - 1 || ppy < 0 || ppy > rh - 1) { row->result = -2; continue; }
			const unsigned rx = (unsigned) ppx, ry = (unsigned) ppy;
			const f3 referenceNormal = refN[rx + ry * rw];
			if (referenceNormal.x == KF_INVALID) { row->result = -3; continue; }
			const f3 diff = sub3(refV[rx + ry * rw], projectedVertex);
			const f3 projectedNormal = mat_rotate(Ttrack, inN[px + py * w]);
			if (length3(diff) > dist_thr) { row->result = -4; continue; }
			if (dot3(projectedNormal, referenceNormal) < normal_thr) { row->result = -5; continue; }
			row->result = 1;
			row->error = dot3(referenceNormal, diff);
			const f3 c = cross3(projectedVertex, referenceNormal);
			row->J[0] = referenceNormal.x; row->J[1] = referenceNormal.y; row->J[2] = referenceNormal.z;
			row->J[3] = c.x; row->J[4] = c.y; row->J[5] = c.z;
		}
}

/* :251-495 (non-OLDREDUCE path, serial order of the cpp build): 8 interleaved row
 * groups, each summed serially in float in (y, x) order, then rows 1..7 added into
 * row 0 in order (:487-489). */
void kfo_reduce(float* out, void* trackdata, unsigned jw, unsigned jh, unsigned w, unsigned h) {
	const TrackData* J = (const TrackData*) trackdata;
	(void) jh;
	for (int blockIndex = 0; blockIndex < 8; blockIndex++) {
		float s[32];
		for (int i = 0; i < 32; ++i) s[i] = 0.0f;
		for (unsigned y = blockIndex; y < h; y += 8)
			for (unsigned x = 0; x < w; x++) {
				const TrackData* row = &J[x + y * jw];
				if (row->result < 1) {
					s[29] += row->result == -4 ? 1 : 0;
					s[30] += row->result == -5 ? 1 : 0;
					s[31] += row->result > -4 ? 1 : 0;
					continue;
				}
				s[0] += row->error * row->error;
				for (int i = 0; i < 6; ++i) s[i + 1] += row->error * row->J[i];
				int k = 7;
				for (int a = 0; a < 6; ++a) for (int b = a; b < 6; ++b) s[k++] += row->J[a] * row->J[b];
				s[28] += 1;
			}
		for (int i = 0; i < 32; ++i) out[blockIndex * 32 + i] = s[i];
	}
	for (int j = 1; j < 8; ++j) for (int i = 0; i < 32; ++i) out[i] += out[j * 32 + i];
}

int kfo_update_pose(float* pose, const float* output, float icp_threshold) { /* :759-775 */
	double x[6];
	kfo_solve(x, output + 1);
	float d[16], np[16];
	kfo_se3_exp(d, x);
	kfo_matmul(np, d, pose);
	memcpy(pose, np, sizeof np);
	double n = 0;
	for (int i = 0; i < 6; ++i) n += x[i] * x[i];
	return sqrt(n) < icp_threshold;
}

int kfo_check_pose(float* pose, const float* old_pose, const float* output, unsigned w, unsigned h, float thr) { /* :777-792 */
	if ((sqrtf(output[0] / output[28]) > 2e-2) || (output[28] / (w * h) < thr)) {
		memcpy(pose, old_pose, sizeof(float) * 16);
		return 0;
	}
	return 1;
}

static inline Vol mkvol(short* data, const unsigned* size, const float* dim) {
	Vol v = { size[0], size[1], size[2], dim[0], dim[1], dim[2], (s2*) data };
	return v;
}
/* commons.h:172-174; index arithmetic is 32-bit unsigned exactly like the reference */
static inline float vs2(const Vol* v, unsigned x, unsigned y, unsigned z) { return v->data[x + y * v->sx + z * v->sx * v->sy].x; }

void kfo_integrate(short* data, const unsigned* size, const float* dim, const float* depth, unsigned w, unsigned h,
		const float* invTrack, const float* K, float mu, float maxweight) { /* :628-673 */
	Vol vol = mkvol(data, size, dim);
	const f3 delta = mat_rotate(invTrack, mk3(0, 0, vol.dz / vol.sz));
	const f3 cameraDelta = mat_rotate(K, delta);
#pragma omp parallel for
	for (unsigned y = 0; y < vol.sy; y++)
		for (unsigned x = 0; x < vol.sx; x++) {
			/* Volume::pos (commons.h:186-189) */
			f3 pos = mat_mul_pt(invTrack, mk3((x + 0.5f) * vol.dx / vol.sx, (y + 0.5f) * vol.dy / vol.sy, (0 + 0.5f) * vol.dz / vol.sz));
			f3 cameraX = mat_mul_pt(K, pos);
			for (unsigned z = 0; z < vol.sz; ++z, pos = add3(pos, delta), cameraX = add3(cameraX, cameraDelta)) {
				if (pos.z < 0.0001f) continue;
				const float pxf = cameraX.x / cameraX.z + 0.5f, pyf = cameraX.y / cameraX.z + 0.5f;
				if (pxf < 0 || pxf > w - 1 || pyf < 0 || pyf > h - 1) continue;
				const unsigned px = (unsigned) pxf, py = (unsigned) pyf;
				if (depth[px + py * w] == 0) continue;
				const float diff = (depth[px + py * w] - cameraX.z) * sqrtf(1 + sq(pos.x / pos.z) + sq(pos.y / pos.z));
				if (diff > -mu) {
					const float sdf = kminf(1.f, diff / mu);
					s2* vox = &vol.data[x + y * vol.sx + z * vol.sx * vol.sy];
					float dx_ = vox->x * 0.00003051944088f, dy_ = vox->y; /* commons.h:160-163 */
					dx_ = kclampf((dy_ * dx_ + sdf) / (dy_ + 1), -1.f, 1.f);
					dy_ = kminf(dy_ + 1, maxweight);
					vox->x = (short) (dx_ * 32766.0f); /* commons.h:182-185, C truncation */
					vox->y = (short) dy_;
				}
			}
		}
}

static float vol_interp(const Vol* v, f3 pos) { /* commons.h:191-213 */
	const f3 sp = mk3((pos.x * v->sx / v->dx) - 0.5f, (pos.y * v->sy / v->dy) - 0.5f, (pos.z * v->sz / v->dz) - 0.5f);
	const float fx = floorf(sp.x), fy = floorf(sp.y), fz = floorf(sp.z);
	const int bx = (int) fx, by = (int) fy, bz = (int) fz;
	const f3 f = mk3(sp.x - fx, sp.y - fy, sp.z - fz);
	const int lx = kmaxi(bx, 0), ly = kmaxi(by, 0), lz = kmaxi(bz, 0);
	const int ux = kmini(bx + 1, (int) v->sx - 1), uy = kmini(by + 1, (int) v->sy - 1), uz = kmini(bz + 1, (int) v->sz - 1);
	return (((vs2(v, lx, ly, lz) * (1 - f.x) + vs2(v, ux, ly, lz) * f.x) * (1 - f.y)
			+ (vs2(v, lx, uy, lz) * (1 - f.x) + vs2(v, ux, uy, lz) * f.x) * f.y) * (1 - f.z)
			+ ((vs2(v, lx, ly, uz) * (1 - f.x) + vs2(v, ux, ly, uz) * f.x) * (1 - f.y)
					+ (vs2(v, lx, uy, uz) * (1 - f.x) + vs2(v, ux, uy, uz) * f.x) * f.y) * f.z) * 0.00003051944088f;
}

static f3 vol_grad(const Vol* v, f3 pos) { /* commons.h:215-301 */
	const f3 sp = mk3((pos.x * v->sx / v->dx) - 0.5f, (pos.y * v->sy / v->dy) - 0.5f, (pos.z * v->sz / v->dz) - 0.5f);
	const float flx = floorf(sp.x), fly = floorf(sp.y), flz = floorf(sp.z);
	const int bx = (int) flx, by = (int) fly, bz = (int) flz;
	const f3 f = mk3(sp.x - flx, sp.y - fly, sp.z - flz);
	const int mx = (int) v->sx - 1, my = (int) v->sy - 1, mz = (int) v->sz - 1;
	const int llx = kmaxi(bx - 1, 0), lly = kmaxi(by - 1, 0), llz = kmaxi(bz - 1, 0);     /* lower_lower */
	const int lx = kmaxi(bx, 0), ly = kmaxi(by, 0), lz = kmaxi(bz, 0);                    /* lower_upper == lower */
	const int ux = kmini(bx + 1, mx), uy = kmini(by + 1, my), uz = kmini(bz + 1, mz);     /* upper_lower == upper */
	const int uux = kmini(bx + 2, mx), uuy = kmini(by + 2, my), uuz = kmini(bz + 2, mz);  /* upper_upper */
	f3 g;
#define V(a, b, c) vs2(v, a, b, c)
	g.x = (((V(ux, ly, lz) - V(llx, ly, lz)) * (1 - f.x) + (V(uux, ly, lz) - V(lx, ly, lz)) * f.x) * (1 - f.y)
			+ ((V(ux, uy, lz) - V(llx, uy, lz)) * (1 - f.x) + (V(uux, uy, lz) - V(lx, uy, lz)) * f.x) * f.y) * (1 - f.z)
			+ (((V(ux, ly, uz) - V(llx, ly, uz)) * (1 - f.x) + (V(uux, ly, uz) - V(lx, ly, uz)) * f.x) * (1 - f.y)
					+ ((V(ux, uy, uz) - V(llx, uy, uz)) * (1 - f.x) + (V(uux, uy, uz) - V(lx, uy, uz)) * f.x) * f.y) * f.z;
	g.y = (((V(lx, uy, lz) - V(lx, lly, lz)) * (1 - f.x) + (V(ux, uy, lz) - V(ux, lly, lz)) * f.x) * (1 - f.y)
			+ ((V(lx, uuy, lz) - V(lx, ly, lz)) * (1 - f.x) + (V(ux, uuy, lz) - V(ux, ly, lz)) * f.x) * f.y) * (1 - f.z)
			+ (((V(lx, uy, uz) - V(lx, lly, uz)) * (1 - f.x) + (V(ux, uy, uz) - V(ux, lly, uz)) * f.x) * (1 - f.y)
					+ ((V(lx, uuy, uz) - V(lx, ly, uz)) * (1 - f.x) + (V(ux, uuy, uz) - V(ux, ly, uz)) * f.x) * f.y) * f.z;
	g.z = (((V(lx, ly, uz) - V(lx, ly, llz)) * (1 - f.x) + (V(ux, ly, uz) - V(ux, ly, llz)) * f.x) * (1 - f.y)
			+ ((V(lx, uy, uz) - V(lx, uy, llz)) * (1 - f.x) + (V(ux, uy, uz) - V(ux, uy, llz)) * f.x) * f.y) * (1 - f.z)
			+ (((V(lx, ly, uuz) - V(lx, ly, lz)) * (1 - f.x) + (V(ux, ly, uuz) - V(ux, ly, lz)) * f.x) * (1 - f.y)
					+ ((V(lx, uy, uuz) - V(lx, uy, lz)) * (1 - f.x) + (V(ux, uy, uuz) - V(ux, uy, lz)) * f.x) * f.y) * f.z;
#undef V
	return mul3s(mul3(g, mk3(v->dx / v->sx, v->dy / v->sy, v->dz / v->sz)), (0.5f * 0.00003051944088f));
}

/* :674-725; returns hit.xyz and *tw = hit.w */
static f3 raycast_one(const Vol* v, unsigned px, unsigned py, const float* view, float nearP, float farP, float step,
		float largestep, float* tw) {
	const f3 origin = mk3(view[3], view[7], view[11]);
	const f3 direction = mat_rotate(view, mk3((float) px, (float) py, 1.f));
	const f3 invR = mk3(1.0f / direction.x, 1.0f / direction.y, 1.0f / direction.z);
	const f3 tbot = mul3(mul3s(invR, -1.f), origin); /* -1 * invR * origin */
	const f3 ttop = mul3(invR, sub3(mk3(v->dx, v->dy, v->dz), origin));
	const f3 tmin = mk3(kminf(ttop.x, tbot.x), kminf(ttop.y, tbot.y), kminf(ttop.z, tbot.z));
	const f3 tmax = mk3(kmaxf(ttop.x, tbot.x), kmaxf(ttop.y, tbot.y), kmaxf(ttop.z, tbot.z));
	const float largest_tmin = kmaxf(kmaxf(tmin.x, tmin.y), kmaxf(tmin.x, tmin.z));   /* x used twice, as in the reference */
	const float smallest_tmax = kminf(kminf(tmax.x, tmax.y), kminf(tmax.x, tmax.z));
	const float tnear = kmaxf(largest_tmin, nearP);
	const float tfar = kminf(smallest_tmax, farP);
	if (tnear < tfar) {
		float t = tnear;
		float stepsize = largestep;
		float f_t = vol_interp(v, add3(origin, mul3s(direction, t)));
		float f_tt = 0;
		if (f_t > 0) {
			for (; t < tfar; t += stepsize) {
				f_tt = vol_interp(v, add3(origin, mul3s(direction, t)));
				if (f_tt < 0) break;
				if (f_tt < 0.8f) stepsize = step;
				f_t = f_tt;
			}
			if (f_tt < 0) {
				t = t + stepsize * f_tt / (f_t - f_tt);
				*tw = t;
				return add3(origin, mul3s(direction, t));
			}
		}
	}
	*tw = 0;
	return mk3(0, 0, 0);
}

void kfo_raycast(float* vtx, float* nrm, unsigned w, unsigned h, short* data, const unsigned* size, const float* dim,
		const float* view, float nearP, float farP, float step, float largestep) { /* :726-757 */
	Vol vol = mkvol(data, size, dim);
	f3 *vertex = (f3*) vtx, *normal = (f3*) nrm;
#pragma omp parallel for schedule(dynamic, 4)
	for (unsigned y = 0; y < h; y++)
		for (unsigned x = 0; x < w; x++) {
			float hw;
			const f3 hit = raycast_one(&vol, x, y, view, nearP, farP, step, largestep, &hw);
			if (hw > 0.0) {
				vertex[x + y * w] = hit;
				f3 surfNorm = vol_grad(&vol, hit);
				if (length3(surfNorm) == 0) normal[x + y * w].x = KF_INVALID; /* only .x */
				else normal[x + y * w] = normalize3(surfNorm);
			} else {
				vertex[x + y * w] = mk3(0, 0, 0);
				normal[x + y * w] = mk3(KF_INVALID, 0, 0);
			}
		}
}

/* ------------------------------------------------------------------ render kernels */
static void gs2rgb(double h, unsigned char* rgba) { /* commons.h:86-147 */
	double v = 0.75, r = 0, g = 0, b = 0;
	double m = 0.25, sv = 0.6667;
	h *= 6.0;
	int sextant = (int) h;
	double fract = h - sextant, vsf = v * sv * fract, mid1 = m + vsf, mid2 = v - vsf;
	switch (sextant) {
	case 0: r = v; g = mid1; b = m; break;
	case 1: r = mid2; g = v; b = m; break;
	case 2: r = m; g = v; b = mid1; break;
	case 3: r = m; g = mid2; b = v; break;
	case 4: r = mid1; g = m; b = v; break;
	case 5: r = v; g = m; b = mid2; break;
	default: r = 0; g = 0; b = 0; break;
	}
	rgba[0] = (unsigned char) (r * 255); rgba[1] = (unsigned char) (g * 255); rgba[2] = (unsigned char) (b * 255); rgba[3] = 0;
}
void kfo_render_depth(unsigned char* out, float* depth, unsigned w, unsigned h, float nearP, float farP) { /* :814-842 */
	float rangeScale = 1 / (farP - nearP);
#pragma omp parallel for
	for (unsigned y = 0; y < h; y++)
		for (unsigned x = 0; x < w; x++) {
			unsigned pos = y * w + x;
			unsigned char* o = out + 4 * (size_t) pos;
			if (depth[pos] < nearP) { o[0] = 255; o[1] = 255; o[2] = 255; o[3] = 0; }
			else if (depth[pos] > farP) { o[0] = 0; o[1] = 0; o[2] = 0; o[3] = 0; }
			else { const float d = (depth[pos] - nearP) * rangeScale; gs2rgb(d, o); }
		}
}
void kfo_render_track(unsigned char* out, const void* trackdata, unsigned w, unsigned h) { /* :844-878 */
	const TrackData* data = (const TrackData*) trackdata;
	static const unsigned char col[7][4] = { { 128, 128, 128, 0 }, { 0, 0, 0, 0 }, { 255, 0, 0, 0 }, { 0, 255, 0, 0 },
			{ 0, 0, 255, 0 }, { 255, 255, 0, 0 }, { 255, 128, 128, 0 } };
	for (unsigned pos = 0; pos < w * h; ++pos) {
		int r = data[pos].result, c;
		if (r == 1) c = 0; else if (r <= -1 && r >= -5) c = -r; else c = 6;
		memcpy(out + 4 * (size_t) pos, col[c], 4);
	}
}
void kfo_render_volume(unsigned char* out, unsigned w, unsigned h, short* data, const unsigned* size, const float* dim,
		const float* view, float nearP, float farP, float step, float largestep) { /* :880-913 */
	Vol vol = mkvol(data, size, dim);
#pragma omp parallel for schedule(dynamic, 4)
	for (unsigned y = 0; y < h; y++)
		for (unsigned x = 0; x < w; x++) {
			unsigned char* o = out + 4 * (size_t) (x + y * w);
			float hw;
			const f3 test = raycast_one(&vol, x, y, view, nearP, farP, step, largestep, &hw);
			o[0] = o[1] = o[2] = o[3] = 0;
			if (hw > 0) {
				const f3 surfNorm = vol_grad(&vol, test);
				if (length3(surfNorm) > 0) {
					const f3 diff = normalize3(sub3(c_light, test));
					const float dir = kmaxf(dot3(normalize3(surfNorm), diff), 0.f);
					const f3 col = mul3s(mk3(kclampf(dir + c_ambient.x, 0.f, 1.f), kclampf(dir + c_ambient.y, 0.f, 1.f),
							kclampf(dir + c_ambient.z, 0.f, 1.f)), 255);
					o[0] = (unsigned char) col.x; o[1] = (unsigned char) col.y; o[2] = (unsigned char) col.z;
				}
			}
		}
}

/* ------------------------------------------------ whole pipeline (class Kfusion) */
typedef struct {
	unsigned cw, ch;
	unsigned vres[3];
	float vdim[3];
	int n_levels, iters[16];
	float step;
	float pose[16], oldPose[16], raycastPose[16];
	float gaussian[5];
	s2* volume;
	f3 *vertex, *normal;
	TrackData* trackingResult;
	float reduction[8 * 32];
	float* scaledDepth[16];
	f3* inputVertex[16];
	f3* inputNormal[16];
	float* floatDepth;
} KF;
static KF* g = NULL;

int kfo_kf_create(unsigned cw, unsigned ch, const unsigned* vres, const float* vdim, const float* init_pos,
		const int* pyramid, int n_levels) { /* kernels.h:99-119 + cpp/kernels.cpp:67-112 */
	if (g) return 1;
	g = (KF*) calloc(1, sizeof(KF));
	g->cw = cw; g->ch = ch;
	memcpy(g->vres, vres, sizeof g->vres);
	memcpy(g->vdim, vdim, sizeof g->vdim);
	g->n_levels = n_levels;
	for (int i = 0; i < n_levels; ++i) g->iters[i] = pyramid[i];
	/* pose = SE3::exp((initPose, 0, 0, 0)) = identity rotation + translation */
	const double mu[6] = { init_pos[0], init_pos[1], init_pos[2], 0, 0, 0 };
	kfo_se3_exp(g->pose, mu);
	float mind = kminf(kminf(vdim[0], vdim[1]), vdim[2]);
	unsigned maxr = vres[0] > vres[1] ? vres[0] : vres[1];
	if (vres[2] > maxr) maxr = vres[2];
	g->step = mind / maxr; /* kernels.h:116 */
	size_t npx = (size_t) cw * ch;
	for (int i = 0; i < n_levels; ++i) { /* cpp/kernels.cpp:79-89 */
		size_t n = npx / (size_t) (int) pow(2, i);
		g->scaledDepth[i] = (float*) calloc(n, sizeof(float));
		g->inputVertex[i] = (f3*) calloc(n, sizeof(f3));
		g->inputNormal[i] = (f3*) calloc(n, sizeof(f3));
	}
	g->floatDepth = (float*) calloc(npx, sizeof(float));
	g->vertex = (f3*) calloc(npx, sizeof(f3));
	g->normal = (f3*) calloc(npx, sizeof(f3));
	g->trackingResult = (TrackData*) calloc(npx, sizeof(TrackData));
	kfo_gaussian(g->gaussian);
	g->volume = (s2*) malloc((size_t) vres[0] * vres[1] * vres[2] * sizeof(s2));
	kfo_init_volume((short*) g->volume, g->vres, g->vdim);
	return 0;
}
void kfo_kf_destroy(void) {
	if (!g) return;
	for (int i = 0; i < g->n_levels; ++i) { free(g->scaledDepth[i]); free(g->inputVertex[i]); free(g->inputNormal[i]); }
	free(g->floatDepth); free(g->vertex); free(g->normal); free(g->trackingResult); free(g->volume);
	free(g);
	g = NULL;
}
void kfo_kf_reset(void) { kfo_init_volume((short*) g->volume, g->vres, g->vdim); }

int kfo_kf_preprocess(const unsigned short* depth, unsigned iw, unsigned ih) { /* :915-922 */
	kfo_mm2meters(g->floatDepth, g->cw, g->ch, depth, iw, ih);
	kfo_bilateral(g->scaledDepth[0], g->floatDepth, g->cw, g->ch, g->gaussian, c_e_delta, c_radius);
	return 1;
}

int kfo_kf_track(const float* k, float icp_threshold, unsigned tracking_rate, unsigned frame) { /* :924-971 */
	if (frame % tracking_rate != 0) return 0;
	for (int i = 1; i < g->n_levels; ++i)
		kfo_halfsample(g->scaledDepth[i], g->scaledDepth[i - 1], g->cw / (int) pow(2, i - 1), g->ch / (int) pow(2, i - 1),
				c_e_delta * 3, 1);
	unsigned lw = g->cw, lh = g->ch;
	for (int i = 0; i < g->n_levels; ++i) {
		const float s = (float) (1 << i);
		const float ks[4] = { k[0] / s, k[1] / s, k[2] / s, k[3] / s }; /* float4 / float (cutil_math.h) */
		float invK[16];
		kfo_inverse_camera_matrix(invK, ks);
		kfo_depth2vertex((float*) g->inputVertex[i], g->scaledDepth[i], lw, lh, invK);
		kfo_vertex2normal((float*) g->inputNormal[i], (float*) g->inputVertex[i], lw, lh);
		lw /= 2; lh /= 2;
	}
	memcpy(g->oldPose, g->pose, sizeof g->pose);
	float K[16], invRP[16], projectReference[16];
	kfo_camera_matrix(K, k);
	kfo_inverse(invRP, g->raycastPose);
	kfo_matmul(projectReference, K, invRP);
	for (int level = g->n_levels - 1; level >= 0; --level) {
		const unsigned w = g->cw / (int) pow(2, level), h = g->ch / (int) pow(2, level);
		for (int i = 0; i < g->iters[level]; ++i) {
			kfo_track(g->trackingResult, (float*) g->inputVertex[level], (float*) g->inputNormal[level], w, h,
					(float*) g->vertex, (float*) g->normal, g->cw, g->ch, g->pose, projectReference, c_dist_threshold,
					c_normal_threshold);
			kfo_reduce(g->reduction, g->trackingResult, g->cw, g->ch, w, h);
			if (kfo_update_pose(g->pose, g->reduction, icp_threshold)) break;
		}
	}
	return kfo_check_pose(g->pose, g->oldPose, g->reduction, g->cw, g->ch, c_track_threshold);
}

int kfo_kf_integrate(const float* k, unsigned integration_rate, float mu, unsigned frame) { /* :988-1004 */
	int doIntegrate = kfo_check_pose(g->pose, g->oldPose, g->reduction, g->cw, g->ch, c_track_threshold);
	if ((doIntegrate && ((frame % integration_rate) == 0)) || (frame <= 3)) {
		float inv[16], K[16];
		kfo_inverse(inv, g->pose);
		kfo_camera_matrix(K, k);
		kfo_integrate((short*) g->volume, g->vres, g->vdim, g->floatDepth, g->cw, g->ch, inv, K, mu, c_maxweight);
		return 1;
	}
	return 0;
}

int kfo_kf_raycast(const float* k, float mu, unsigned frame) { /* :973-986 */
	if (frame > 2) {
		memcpy(g->raycastPose, g->pose, sizeof g->pose);
		float invK[16], view[16];
		kfo_inverse_camera_matrix(invK, k);
		kfo_matmul(view, g->raycastPose, invK);
		kfo_raycast((float*) g->vertex, (float*) g->normal, g->cw, g->ch, (short*) g->volume, g->vres, g->vdim, view,
				c_nearPlane, c_farPlane, g->step, 0.75f * mu);
	}
	return 0;
}
void kfo_kf_get_pose(float* out16) { memcpy(out16, g->pose, sizeof g->pose); }
void kfo_kf_render_depth(unsigned char* out, unsigned w, unsigned h) { kfo_render_depth(out, g->floatDepth, w, h, c_nearPlane, c_farPlane); }
void kfo_kf_render_track(unsigned char* out, unsigned w, unsigned h) { kfo_render_track(out, g->trackingResult, w, h); }
void kfo_kf_render_volume(unsigned char* out, unsigned w, unsigned h, int frame, int rate, const float* k, float largestep) { /* :1032-1038 */
	if (frame % rate == 0) {
		float invK[16], view[16];
		kfo_inverse_camera_matrix(invK, k);
		kfo_matmul(view, g->pose, invK); /* viewPose defaults to &pose (kernels.h:117) */
		kfo_render_volume(out, w, h, (short*) g->volume, g->vres, g->vdim, view, c_nearPlane, c_farPlane * 2.0f, g->step, largestep);
	}
}
void kfo_kf_dump_volume(const char* path) { /* :1006-1030: tsdf shorts only, x fastest */
	FILE* f = fopen(path, "wb");
	if (!f) { printf("Error opening file: %s\n", path); exit(1); }
	size_t n = (size_t) g->vres[0] * g->vres[1] * g->vres[2];
	for (size_t i = 0; i < n; ++i) fwrite(&g->volume[i].x, sizeof(short), 1, f);
	fclose(f);
}
void* kfo_kf_buffer(int which, int level) {
	switch (which) {
	case 0: return g->volume;
	case 1: return g->vertex;
	case 2: return g->normal;
	case 3: return g->floatDepth;
	case 4: return g->scaledDepth[level];
	case 5: return g->inputVertex[level];
	case 6: return g->inputNormal[level];
	case 7: return g->reduction;
	case 8: return g->trackingResult;
	case 9: return g->raycastPose;
	case 10: return g->oldPose;
	case 11: return g->gaussian;
	}
	return NULL;
}
