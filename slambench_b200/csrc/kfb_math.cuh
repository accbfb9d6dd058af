// Small-vector / 4x4 math with the EXACT operation order of the reference's host helpers
// (kfusion/thirdparty/cutil_math.h, kfusion/include/commons.h:317-378).  Everything in this
// library is compiled with --fmad=false -prec-div=true -prec-sqrt=true so that a*b+c is two
// IEEE roundings, `/` and sqrtf are correctly rounded, and results are bit-identical to the
// reference's baseline-x86-64 (no FMA) arithmetic.
#ifndef KFB_MATH_CUH
#define KFB_MATH_CUH

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define KFB_INVALID (-2.0f)  // commons.h:14

#define KFB_HDI __host__ __device__ __forceinline__

struct Mat4 {
	float m[16];  // row-major: m[4*r + c]  (commons.h:317-319: float4 data[4])
};

// cutil_math.h:43-57 — on the host fminf/fmaxf/min/max are plain ternaries (NaN-propagation
// differs from CUDA's fminf/fmaxf, so spell them out)
KFB_HDI float kminf(float a, float b) { return a < b ? a : b; }
KFB_HDI float kmaxf(float a, float b) { return a > b ? a : b; }
KFB_HDI int kmini(int a, int b) { return a < b ? a : b; }
KFB_HDI int kmaxi(int a, int b) { return a > b ? a : b; }
KFB_HDI float kclampf(float f, float a, float b) { return kmaxf(a, kminf(f, b)); }  // :972
KFB_HDI float ksq(float r) { return r * r; }                                          // commons.h:82

KFB_HDI float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
KFB_HDI float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
KFB_HDI float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
KFB_HDI float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
KFB_HDI float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
KFB_HDI float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
KFB_HDI float kdot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // :1063-1071
KFB_HDI float3 kcross(float3 a, float3 b) {                                            // :1244-1247
	return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
KFB_HDI float klength(float3 v) { return sqrtf(kdot(v, v)); }                         // :1117-1124
// :1147-1151 with the host rsqrtf(x) = 1.0f / sqrtf(x) (:59-61)
KFB_HDI float3 knormalize(float3 v) {
	const float inv = 1.0f / sqrtf(kdot(v, v));
	return v * inv;
}

// commons.h:331-336  Matrix4 * float3 (point transform)
KFB_HDI float3 mat_point(const Mat4& M, float3 v) {
	return f3(kdot(f3(M.m[0], M.m[1], M.m[2]), v) + M.m[3], kdot(f3(M.m[4], M.m[5], M.m[6]), v) + M.m[7],
			kdot(f3(M.m[8], M.m[9], M.m[10]), v) + M.m[11]);
}
// commons.h:338-341  rotate
KFB_HDI float3 mat_rotate(const Mat4& M, float3 v) {
	return f3(kdot(f3(M.m[0], M.m[1], M.m[2]), v), kdot(f3(M.m[4], M.m[5], M.m[6]), v),
			kdot(f3(M.m[8], M.m[9], M.m[10]), v));
}

// packed float3 arrays (12-byte stride, the reference's layout: vector_types.h:200)
KFB_HDI float3 ld3(const float* p, size_t i) { return f3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
KFB_HDI void st3(float* p, size_t i, float3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }

#endif
