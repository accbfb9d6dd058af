// Hand-written sm_100a kernels of the KinectFusion per-frame pipeline.
// Each kernel cites the reference function whose results it reproduces
// (kfusion/src/cpp/kernels.cpp, kfusion/include/commons.h).  None of these stages is a dense
// contraction, so there is no tensor-core work: they are HBM/L2-bound streaming, stencil,
// gather and reduction kernels.  Compiled with --fmad=false -prec-div=true -prec-sqrt=true:
// parity with the reference's un-fused IEEE fp32 arithmetic is bit-exact by construction for
// every stage except the (order-dependent) ICP sums.
#ifndef KFB_KERNELS_CUH
#define KFB_KERNELS_CUH

#include "kfb_math.cuh"
#include "kfb_expf.h"
#include "kfb_hostmath.h"

// ------------------------------------------------------------------------------------------
// initVolumeKernel (cpp/kernels.cpp:147-157): every voxel <- short2(32766, 0).  128-bit stores.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_init_volume(uint4* __restrict__ vol4, size_t n4, short2* __restrict__ vol, size_t n) {
	const uint32_t v = (uint32_t) (uint16_t) 32766;  // x = 32766, y = 0
	const size_t stride = (size_t) gridDim.x * blockDim.x;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) vol4[i] = make_uint4(v, v, v, v);
	// tail (n not a multiple of 4 voxels)
	for (size_t i = n4 * 4 + (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) vol[i] = make_short2(32766, 0);
}

// ------------------------------------------------------------------------------------------
// mm2metersKernel + bilateralFilterKernel fused (cpp/kernels.cpp:562-589, 159-198).
// One 32x8 output tile per CTA; the (32+4)x(8+4) halo tile of metres is staged in shared
// memory straight from the uint16 sensor frame, so the raw float depth is never re-read.
// Writes BOTH the raw depth (integrate consumes it, :995) and the filtered depth (:918).
// Border taps clamp to the edge exactly like the host `clamp(uint,..)` does.
// ------------------------------------------------------------------------------------------
struct Gauss5 { float g[5]; };

#define PP_BX 32
#define PP_BY 8
#define PP_R 2
__global__ void __launch_bounds__(PP_BX* PP_BY) k_mm2m_bilateral(const uint16_t* __restrict__ in, uint32_t iw, int ratio,
		float* __restrict__ raw, float* __restrict__ filt, uint32_t w, uint32_t h, Gauss5 gs, float e_d,
		unsigned int* __restrict__ dmax_bits, unsigned int* __restrict__ dmax_next) {
	__shared__ float tile[PP_BY + 2 * PP_R][PP_BX + 2 * PP_R + 1];
	if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0) *dmax_next = 0u;  // re-arm the other slot
	const int x0 = blockIdx.x * PP_BX, y0 = blockIdx.y * PP_BY;
	const int tid = threadIdx.y * PP_BX + threadIdx.x;
	for (int i = tid; i < (PP_BY + 2 * PP_R) * (PP_BX + 2 * PP_R); i += PP_BX * PP_BY) {
		const int ty = i / (PP_BX + 2 * PP_R), tx = i - ty * (PP_BX + 2 * PP_R);
		const int gx = kmaxi(0, kmini(x0 + tx - PP_R, (int) w - 1));
		const int gy = kmaxi(0, kmini(y0 + ty - PP_R, (int) h - 1));
		tile[ty][tx] = (float) in[(size_t) gx * ratio + (size_t) iw * gy * ratio] / 1000.0f;
	}
	__syncthreads();
	const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
	const bool inside = x < (int) w && y < (int) h;
	const float center = tile[threadIdx.y + PP_R][threadIdx.x + PP_R];
	// max of the raw depth image (non-negative floats order like their bit patterns): integrate's far cull
	{
		const unsigned int m = __reduce_max_sync(0xffffffffu, inside ? __float_as_uint(center) : 0u);
		if (threadIdx.x == 0 && m) atomicMax(dmax_bits, m);
	}
	if (!inside) return;
	const size_t pos = (size_t) x + (size_t) y * w;
	raw[pos] = center;
	if (center == 0) { filt[pos] = 0; return; }
	const float e_d_squared_2 = e_d * e_d * 2;
	float sum = 0.0f, t = 0.0f;
#pragma unroll
	for (int i = -PP_R; i <= PP_R; ++i) {      // i walks x, j walks y — the reference's order
#pragma unroll
		for (int j = -PP_R; j <= PP_R; ++j) {
			const float curPix = tile[threadIdx.y + PP_R + j][threadIdx.x + PP_R + i];
			if (curPix > 0) {
				const float mod = ksq(curPix - center);
				// expf(-0) == 1 exactly: equal depths (flat, fronto-parallel surfaces; the centre tap) skip the fp64 routine
				const float factor = gs.g[i + PP_R] * gs.g[j + PP_R] * (mod == 0.f ? 1.0f : kfb_expf_nonpos(-mod / e_d_squared_2));
				t += factor * curPix;
				sum += factor;
			}
		}
	}
	filt[pos] = t / sum;
}

// ------------------------------------------------------------------------------------------
// Pyramid + vertex/normal maps in ONE launch (cpp/kernels.cpp:931-945):
//   halfSampleRobustImageKernel (:591-626) for levels 1..L-1,
//   depth2vertexKernel (:200-218) and vertex2normalKernel (:220-249) for levels 0..L-1.
// One thread per output pixel of every level.  Coarse depths are RECOMPUTED from level 0
// (4 or 16 L1-resident loads) instead of being produced by a chain of dependent launches;
// the recomputation performs the identical fp32 operations, so the result is bit-identical.
// ------------------------------------------------------------------------------------------
struct PyrParams {
	const float* d0;          // ScaledDepth[0]
	float* depth[3];          // ScaledDepth[l] (l>=1 written here)
	float* vertex[3];
	float* normal[3];
	uint32_t w[3], h[3];
	uint32_t first[4];        // first linear thread id of each level (prefix sums), first[L] = total
	Mat4 invK[3];
	int levels;
	float e_d;                // e_delta * 3
};

// robust 2x2 mean of `in` (size iw x ..) at output pixel (x,y): r = 1 => offsets {0,1}
template <class F> __device__ __forceinline__ float half_sample(F in, int x, int y, float e_d) {
	const int cx = 2 * x, cy = 2 * y;
	float sum = 0.0f, t = 0.0f;
	const float center = in(cx, cy);
#pragma unroll
	for (int i = 0; i <= 1; ++i)
#pragma unroll
		for (int j = 0; j <= 1; ++j) {
			const float current = in(cx + j, cy + i);
			if (fabsf(current - center) < e_d) { sum += 1.0f; t += current; }
		}
	return t / sum;
}

__device__ __forceinline__ float pyr_depth(const PyrParams& p, int level, int x, int y) {
	const float* d0 = p.d0;
	const uint32_t w0 = p.w[0];
	auto l0 = [&](int xx, int yy) { return __ldg(d0 + (size_t) xx + (size_t) yy * w0); };
	if (level == 0) return l0(x, y);
	auto l1 = [&](int xx, int yy) { return half_sample(l0, xx, yy, p.e_d); };
	if (level == 1) return l1(x, y);
	return half_sample(l1, x, y, p.e_d);
}

__device__ __forceinline__ float3 pyr_vertex(const PyrParams& p, int level, int x, int y) {
	const float d = pyr_depth(p, level, x, y);
	if (d > 0) return d * mat_rotate(p.invK[level], f3((float) x, (float) y, 1.f));
	return f3(0, 0, 0);
}

__global__ void __launch_bounds__(256) k_pyramid(PyrParams p) {
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	if (tid >= p.first[p.levels]) return;
	int level = 0;
	if (p.levels > 1 && tid >= p.first[1]) level = 1;
	if (p.levels > 2 && tid >= p.first[2]) level = 2;
	const uint32_t local = tid - p.first[level];
	const int w = p.w[level], h = p.h[level];
	const int x = local % w, y = local / w;
	if (level > 0) p.depth[level][local] = pyr_depth(p, level, x, y);
	const float3 v = pyr_vertex(p, level, x, y);
	st3(p.vertex[level], local, v);
	const float3 left = pyr_vertex(p, level, kmaxi(x - 1, 0), y);
	const float3 right = pyr_vertex(p, level, kmini(x + 1, w - 1), y);
	const float3 up = pyr_vertex(p, level, x, kmaxi(y - 1, 0));
	const float3 down = pyr_vertex(p, level, x, kmini(y + 1, h - 1));
	if (left.z == 0 || right.z == 0 || up.z == 0 || down.z == 0) {
		p.normal[level][3 * (size_t) local] = KFB_INVALID;  // only .x, like the reference (:240)
		return;
	}
	const float3 dxv = right - left, dyv = down - up;
	st3(p.normal[level], local, knormalize(kcross(dyv, dxv)));
}

// Levels beyond the third (the reference accepts any `-y` list, default_parameters.h:394-396): one launch per level,
// each from the STORED depth of the level above — same halfSample / depth2vertex / vertex2normal arithmetic.
__global__ void __launch_bounds__(256) k_pyramid_level(const float* __restrict__ dprev, uint32_t pw, float* __restrict__ depth,
		float* __restrict__ vertex, float* __restrict__ normal, uint32_t w, uint32_t h, Mat4 invK, float e_d) {
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	if (tid >= w * h) return;
	const int x = tid % w, y = tid / w;
	auto lp = [&](int xx, int yy) { return __ldg(dprev + (size_t) xx + (size_t) yy * pw); };
	auto dep = [&](int xx, int yy) { return half_sample(lp, xx, yy, e_d); };
	auto vtx = [&](int xx, int yy) {
		const float d = dep(xx, yy);
		return d > 0 ? d * mat_rotate(invK, f3((float) xx, (float) yy, 1.f)) : f3(0, 0, 0);
	};
	depth[tid] = dep(x, y);
	st3(vertex, tid, vtx(x, y));
	const float3 left = vtx(kmaxi(x - 1, 0), y), right = vtx(kmini(x + 1, (int) w - 1), y);
	const float3 up = vtx(x, kmaxi(y - 1, 0)), down = vtx(x, kmini(y + 1, (int) h - 1));
	if (left.z == 0 || right.z == 0 || up.z == 0 || down.z == 0) {
		normal[3 * (size_t) tid] = KFB_INVALID;
		return;
	}
	st3(normal, tid, knormalize(kcross(down - up, right - left)));
}

// ------------------------------------------------------------------------------------------
// trackKernel + reduceKernel FUSED (cpp/kernels.cpp:497-560, 251-495).
// The reference writes a 32-byte TrackData per pixel to memory and re-reads it in a second
// kernel; here each thread folds its pixels straight into 27 register sums + 4 counters,
// warps combine with shuffles, CTAs through shared memory, and the last CTA to finish sums
// the per-CTA partials in a FIXED order (deterministic, no float atomics) into the 32-float
// result: [0]=sum e^2, [1..6]=J^T e, [7..27]=upper-tri J^T J, [28]=#inliers, [29]=#(-4),
// [30]=#(-5), [31]=#(-1,-2,-3).  Cross-thread accumulation is fp64, so the result does not
// depend on the grid shape beyond ~1e-16 and multi-GPU partial sums compose.
// ------------------------------------------------------------------------------------------
#define TR_THREADS 256
#define TR_MAX_BLOCKS 1184  // 148 SMs x 8

struct TrackParams {
	const float* inV; const float* inN;     // packed float3[w*h] of this level
	const float* refV; const float* refN;   // packed float3[rw*rh] (raycast maps, world frame)
	uint32_t w, h, rw, rh;
	uint32_t row0, row1;                    // this context's share of input rows [row0,row1) (multi-GPU: a band)
	Mat4 Ttrack, view;                      // pose, projectReference (used when pose_dev == nullptr)
	const float* pose_dev;                  // optional: pose / view live in device memory (device-side ICP loop)
	const float* view_dev;
	float dist_threshold, normal_threshold;
	double* partials;                       // [gridDim.x][32]
	unsigned int* counter;                  // last-block ticket
	float* out32;                           // device result
	float* out32_host; volatile uint32_t* seq_host; uint32_t seq;  // optional mapped-host mirror + sequence flag
	int8_t* status;                         // optional per-pixel result plane (stride rw), for renderTrack
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}


// per-thread accumulators of the 32 reduction outputs (reduceKernel, cpp/kernels.cpp:251-495)
struct TrackAcc {
	float s[28];
	int c28, c29, c30, c31;
	__device__ __forceinline__ void clear() {
#pragma unroll
		for (int i = 0; i < 28; ++i) s[i] = 0.f;
		c28 = c29 = c30 = c31 = 0;
	}
};

// trackKernel for ONE input pixel (cpp/kernels.cpp:497-560) folded straight into the sums
__device__ __forceinline__ void track_pixel(TrackAcc& a, const float* __restrict__ inV, const float* __restrict__ inN,
		const float* __restrict__ refV, const float* __restrict__ refN, uint32_t w, uint32_t rw, uint32_t rh, uint32_t px, uint32_t py,
		const Mat4& T, const Mat4& V, float dist_threshold, float normal_threshold, int8_t* status) {
	const size_t idx = (size_t) px + (size_t) py * w;
	int result;
	float err = 0.f;
	float3 Ja = f3(0, 0, 0), Jb = f3(0, 0, 0);
	const float3 n = ld3(inN, idx);
	if (n.x == KFB_INVALID) {
		result = -1;
	} else {
		const float3 pv = mat_point(T, ld3(inV, idx));
		const float3 pp = mat_point(V, pv);
		const float pixx = pp.x / pp.z + 0.5f, pixy = pp.y / pp.z + 0.5f;
		if (pixx < 0 || pixx > (float) (rw - 1) || pixy < 0 || pixy > (float) (rh - 1)) {
			result = -2;
		} else {
			// (uint)NaN is 0 on x86-64 (cvttss2si, low 32 bits) and here (cvt.rzi) — start-up frames
			const uint32_t rx = (uint32_t) pixx, ry = (uint32_t) pixy;
			const size_t ridx = (size_t) rx + (size_t) ry * rw;
			const float3 rn = ld3(refN, ridx);
			if (rn.x == KFB_INVALID) {
				result = -3;
			} else {
				const float3 diff = ld3(refV, ridx) - pv;
				const float3 pn = mat_rotate(T, n);
				if (klength(diff) > dist_threshold) result = -4;
				else if (kdot(pn, rn) < normal_threshold) result = -5;
				else {
					result = 1;
					err = kdot(rn, diff);
					Ja = rn;
					Jb = kcross(pv, rn);
				}
			}
		}
	}
	if (status) status[(size_t) px + (size_t) py * rw] = (int8_t) result;
	if (result < 1) {
		a.c29 += (result == -4);
		a.c30 += (result == -5);
		a.c31 += (result > -4);
	} else {
		const float J[6] = { Ja.x, Ja.y, Ja.z, Jb.x, Jb.y, Jb.z };
		a.s[0] += err * err;
#pragma unroll
		for (int k = 0; k < 6; ++k) a.s[1 + k] += err * J[k];
		int q = 7;
#pragma unroll
		for (int i = 0; i < 6; ++i)
#pragma unroll
			for (int j = i; j < 6; ++j) a.s[q++] += J[i] * J[j];
		a.c28 += 1;
	}
}

// the same for U pixels at once, staged so that the U x (inN, inV) loads and then the U x (refN, refV) gathers are
// all in flight together (the per-pixel chain input -> projection -> gather is latency-bound otherwise).
// Invalid / out-of-image pixels gather from index 0 (in bounds) and discard the values: results are identical.
template <int U>
__device__ __forceinline__ void track_pixels(TrackAcc& a, const float* __restrict__ inV, const float* __restrict__ inN,
		const float* __restrict__ refV, const float* __restrict__ refN, uint32_t w, uint32_t rw, uint32_t rh, const uint32_t (&pix)[U],
		const bool (&on)[U], const Mat4& T, const Mat4& V, float dist_threshold, float normal_threshold, int8_t* status) {
	float3 n[U], vin[U], pv[U], rn[U], rv[U];
	int res[U];
	size_t ridx[U];
#pragma unroll
	for (int u = 0; u < U; ++u) {
		const size_t idx = on[u] ? (size_t) pix[u] : 0;
		n[u] = ld3(inN, idx);
		vin[u] = ld3(inV, idx);
	}
#pragma unroll
	for (int u = 0; u < U; ++u) {
		res[u] = 0;
		ridx[u] = 0;
		pv[u] = mat_point(T, vin[u]);
		if (n[u].x == KFB_INVALID) res[u] = -1;
		else {
			const float3 pp = mat_point(V, pv[u]);
			const float pixx = pp.x / pp.z + 0.5f, pixy = pp.y / pp.z + 0.5f;
			if (pixx < 0 || pixx > (float) (rw - 1) || pixy < 0 || pixy > (float) (rh - 1)) res[u] = -2;
			else ridx[u] = (size_t) (uint32_t) pixx + (size_t) (uint32_t) pixy * rw;   // (uint)NaN == 0, like x86-64
		}
	}
#pragma unroll
	for (int u = 0; u < U; ++u) {
		rn[u] = ld3(refN, ridx[u]);
		rv[u] = ld3(refV, ridx[u]);
	}
#pragma unroll
	for (int u = 0; u < U; ++u) {
		if (!on[u]) continue;
		int result = res[u];
		float err = 0.f;
		float3 Ja = f3(0, 0, 0), Jb = f3(0, 0, 0);
		if (result == 0) {
			if (rn[u].x == KFB_INVALID) result = -3;
			else {
				const float3 diff = rv[u] - pv[u];
				const float3 pn = mat_rotate(T, n[u]);
				if (klength(diff) > dist_threshold) result = -4;
				else if (kdot(pn, rn[u]) < normal_threshold) result = -5;
				else {
					result = 1;
					err = kdot(rn[u], diff);
					Ja = rn[u];
					Jb = kcross(pv[u], rn[u]);
				}
			}
		}
		if (status) status[(size_t) (pix[u] % w) + (size_t) (pix[u] / w) * rw] = (int8_t) result;
		if (result < 1) {
			a.c29 += (result == -4);
			a.c30 += (result == -5);
			a.c31 += (result > -4);
		} else {
			const float J[6] = { Ja.x, Ja.y, Ja.z, Jb.x, Jb.y, Jb.z };
			a.s[0] += err * err;
#pragma unroll
			for (int k = 0; k < 6; ++k) a.s[1 + k] += err * J[k];
			int q = 7;
#pragma unroll
			for (int i = 0; i < 6; ++i)
#pragma unroll
				for (int j = i; j < 6; ++j) a.s[q++] += J[i] * J[j];
			a.c28 += 1;
		}
	}
}

__global__ void __launch_bounds__(TR_THREADS) k_track_reduce(TrackParams p) {
	__shared__ double sm[TR_THREADS / 32][32];
	__shared__ bool is_last;
	Mat4 T, V;
	if (p.pose_dev) {
#pragma unroll
		for (int i = 0; i < 16; ++i) { T.m[i] = p.pose_dev[i]; V.m[i] = p.view_dev[i]; }
	} else { T = p.Ttrack; V = p.view; }

	TrackAcc acc;
	acc.clear();
	const uint32_t npx = (p.row1 - p.row0) * p.w;
	for (uint32_t i = blockIdx.x * TR_THREADS + threadIdx.x; i < npx; i += gridDim.x * TR_THREADS) {
		const uint32_t py = p.row0 + i / p.w, px = i % p.w;
		track_pixel(acc, p.inV, p.inN, p.refV, p.refN, p.w, p.rw, p.rh, px, py, T, V, p.dist_threshold, p.normal_threshold, p.status);
	}
	float* s = acc.s;
	const int c28 = acc.c28, c29 = acc.c29, c30 = acc.c30, c31 = acc.c31;

	// warp -> CTA -> grid, all in fp64 and in a fixed order
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
	for (int i = 0; i < 28; ++i) {
		const double v = warp_sum((double) s[i]);
		if (lane == 0) sm[wid][i] = v;
	}
	{
		const double v28 = warp_sum((double) c28), v29 = warp_sum((double) c29), v30 = warp_sum((double) c30), v31 = warp_sum((double) c31);
		if (lane == 0) { sm[wid][28] = v28; sm[wid][29] = v29; sm[wid][30] = v30; sm[wid][31] = v31; }
	}
	__syncthreads();
	if (wid == 0) {
		double v = 0;
#pragma unroll
		for (int w = 0; w < TR_THREADS / 32; ++w) v += sm[w][lane];
		p.partials[(size_t) blockIdx.x * 32 + lane] = v;
		__threadfence();
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		const unsigned int ticket = atomicAdd(p.counter, 1u);
		is_last = (ticket == gridDim.x - 1);
	}
	__syncthreads();
	if (!is_last) return;
	__threadfence();
	// last CTA: warp w sums partial rows w, w+8, ... ; then the 8 warp sums are added in order
	{
		double v = 0;
		for (uint32_t b = wid; b < gridDim.x; b += TR_THREADS / 32) v += __ldcg(p.partials + (size_t) b * 32 + lane);
		sm[wid][lane] = v;
	}
	__syncthreads();
	if (wid == 0) {
		double v = 0;
#pragma unroll
		for (int w = 0; w < TR_THREADS / 32; ++w) v += sm[w][lane];
		const float r = (float) v;
		p.out32[lane] = r;
		if (p.out32_host) p.out32_host[lane] = r;
		if (lane == 0) *p.counter = 0;  // re-arm for the next launch
		__syncwarp();
		if (p.seq_host) {
			__threadfence_system();
			if (lane == 0) *p.seq_host = p.seq;
		}
	}
}

// ------------------------------------------------------------------------------------------
// The WHOLE ICP loop of Kfusion::tracking (cpp/kernels.cpp:950-967) as ONE persistent
// cooperative kernel: for level = L-1..0, for i < iterations[level]: fused track+reduce over this
// level's pixels -> per-CTA fp64 partials -> grid barrier; the LAST CTA to arrive sums the partials
// in a fixed order, solves the 6x6 system (updatePoseKernel, :759-775: certified Cholesky fast
// path, Jacobi pseudo-inverse fallback), composes the new pose, decides the per-level `break`,
// mirrors pose + sums to mapped host memory and releases the barrier.  No launch, no host round
// trip and no skipped work between iterations; the host waits once per frame.
// Launched with cudaLaunchCooperativeKernel (all CTAs co-resident: they wait on one another).
// ------------------------------------------------------------------------------------------
// Frame state produced ON THE DEVICE by the ICP kernel's tail (checkPoseKernel cpp/kernels.cpp:777-792, inverse(pose)
// commons.h:365-371, raycastPose * invK :979) and consumed by the integrate / raycast kernels of the same frame, so that
// Kfusion::computeFrame can enqueue the whole frame before the host has seen the pose.
struct DevFrame {
	float pose[16];        // pose after checkPose (oldPose again when tracking failed)
	float invTrack[16];    // inverse(pose): integrate
	float view[16];        // pose * getInverseCameraMatrix(k): raycast
	float ca[9], tz[3];    // integrate's classification geometry: K.rot * invTrack.rot; third row of invTrack.rot
	int tracked, do_integrate;
};
struct IcpTail {
	DevFrame* out;         // nullptr: no tail (the host evaluates checkPose / the matrices itself)
	Mat4 K, invK;          // getCameraMatrix(k), getInverseCameraMatrix(k)
	float track_threshold;
	int force_integrate;   // frame <= 3                      (cpp/kernels.cpp:994)
	int rate_ok;           // frame % integration_rate == 0
};

#define ICP_MAX_LEVELS 8
struct IcpParams {
	const float* inV[ICP_MAX_LEVELS]; const float* inN[ICP_MAX_LEVELS];
	uint32_t w[ICP_MAX_LEVELS], h[ICP_MAX_LEVELS];
	int iterations[ICP_MAX_LEVELS];
	int levels;
	const float* refV; const float* refN;
	uint32_t rw, rh;
	Mat4 pose0, view;                       // pose at entry (:947), projectReference (:948)
	float dist_threshold, normal_threshold, icp_threshold;
	double* partials;                       // [gridDim.x][32]
	unsigned int* bar;                      // [0] arrival counter, [1] generation, [2] converged flag of the last solve, [3] iterations
	float* out32;                           // [32] device copy of the sums
	unsigned long long* prof;               // optional phase timers (KFB_ICP_PROFILE): see k_icp
	float* pose_dev;                        // [16] current pose (global; rewritten by the last CTA each iteration)
	float* out_host;                        // mapped: [0..31] sums, [32] seq, [33] error, [48..63] pose, [64] iterations
	uint32_t seq;
	int8_t* status;
	IcpTail tail;
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
	unsigned int v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
	asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int* p) {
	unsigned int v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

// hm_solve6_chol (kfb_hostmath.h) spread over lanes 0..5 of a warp: lane i owns row i of the Cholesky factor and column i of
// its inverse, so the ~250 dependent fp64 operations of the one-thread version become chains of ~40.  Every entry is
// computed with the same operations in the same order as there (bit-identical x); only the certificate's trace(C^-1) is
// summed column-wise.  All 32 lanes must call; returns 0 (uniformly) when the certificate fails.
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
#define KFB_UP(r, c) (6 + 6 * (r) - (r) * ((r) - 1) / 2 + ((c) - (r)))   // vals27 index of C[r][c], c >= r (commons.h:385-392)
__device__ __forceinline__ int icp_solve6_chol_warp(double* x6, const float* __restrict__ vals27, int lane) {
	const int i = lane < 6 ? lane : 5;    // lanes >= 6 mirror lane 5 (their results are never read)
	double Lr[6];
#pragma unroll
	for (int k = 0; k < 6; ++k) Lr[k] = (k <= i) ? (double) vals27[KFB_UP(k, i)] : 0.0;
	double diag = 0;
#pragma unroll
	for (int k = 0; k < 6; ++k) if (k == i) diag = Lr[k];
	double trC = 0;
#pragma unroll
	for (int r = 0; r < 6; ++r) trC += shfl_d(diag, r);
	double invd[6];
	bool ok = true;
#pragma unroll
	for (int j = 0; j < 6; ++j) {
		double s = Lr[j];
#pragma unroll
		for (int k = 0; k < j; ++k) s -= Lr[k] * shfl_d(Lr[k], j);
		const double sj = shfl_d(s, j);
		if (!(sj > 0)) ok = false;
		const double inv = rsqrt(sj), d = sj * inv;
		invd[j] = inv;
		Lr[j] = (i == j) ? d : s * inv;   // rows i < j: never read again
	}
	if (!ok) return 0;
	// all of L to every lane, then lane j inverts column j:  M(j,j) = 1 / L(j,j);  M(i,j) = -(sum_{k=j}^{i-1} L(i,k) M(k,j)) / L(i,i)
	double Lf[6][6];
#pragma unroll
	for (int r = 1; r < 6; ++r)
#pragma unroll
		for (int k = 0; k < r; ++k) Lf[r][k] = shfl_d(Lr[k], r);
	double Mc[6];
	double sq = 0;
#pragma unroll
	for (int r = 0; r < 6; ++r) {
		double t = 0;
#pragma unroll
		for (int k = 0; k < r; ++k) if (k >= i) t -= Lf[r][k] * Mc[k];
		Mc[r] = (r == i) ? invd[r] : ((r > i) ? t * invd[r] : 0.0);
		sq += Mc[r] * Mc[r];
	}
	double trInv = 0;
#pragma unroll
	for (int r = 0; r < 6; ++r) trInv += shfl_d(sq, r);
	if (!(trC * trInv < 0.99e6)) return 0;
	const double bi = (double) vals27[i];
	double y[6];
#pragma unroll
	for (int r = 0; r < 6; ++r) {
		const double term = Mc[r] * bi;   // M(r, i) b_i, zero for i > r
		double t = 0;
#pragma unroll
		for (int j = 0; j <= r; ++j) t += shfl_d(term, j);
		y[r] = t;
	}
	double xi = 0;
#pragma unroll
	for (int r = 0; r < 6; ++r) if (r >= i) xi += Mc[r] * y[r];
#pragma unroll
	for (int j = 0; j < 6; ++j) x6[j] = shfl_d(xi, j);
	return 1;
}

// warp 0 of the last CTA: updatePoseKernel (:759-775) on the device copy of the pose
__device__ __noinline__ int icp_solve(float* pose_dev, const float* pose_cur, const float* red32, float icp_threshold, int lane) {
	double x[6];
	const int fast = icp_solve6_chol_warp(x, red32 + 1, lane);
	int conv = 0;
	if (lane == 0) {
		if (!fast) hm_solve6(x, red32 + 1);   // not positive definite / possibly ill-conditioned: the pseudo-inverse
		float pose[16], d[16];
#pragma unroll
		for (int i = 0; i < 16; ++i) pose[i] = pose_cur[i];   // this CTA's shared-memory copy of the current pose: no L2 round trip
		hm_se3_exp(d, x);
		hm_matmul4(pose, d, pose);
#pragma unroll
		for (int i = 0; i < 16; ++i) pose_dev[i] = pose[i];
		double n = 0;
		for (int i = 0; i < 6; ++i) n += x[i] * x[i];
		conv = sqrt(n) < (double) icp_threshold;
	}
	return __shfl_sync(0xffffffffu, conv, 0);
}

__device__ __forceinline__ unsigned long long gtime_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}

// one thread, after the last iteration: what Kfusion::tracking / integration / raycasting compute on the host between the
// kernels (cpp/kernels.cpp:968, 991-996, 978-981), with the same __host__ __device__ routines the host path uses
__device__ __noinline__ void icp_tail(const IcpParams& p) {
	float pose[16], old[16], red[32];
#pragma unroll
	for (int i = 0; i < 16; ++i) { pose[i] = __ldcg(p.pose_dev + i); old[i] = p.pose0.m[i]; }
#pragma unroll
	for (int i = 0; i < 32; ++i) red[i] = __ldcg(p.out32 + i);
	const int tracked = hm_check_pose(pose, old, red, p.rw, p.rh, p.tail.track_threshold);
	DevFrame* f = p.tail.out;
	float inv[16], view[16];
	hm_inverse4(inv, pose);
	hm_matmul4(view, pose, p.tail.invK.m);
#pragma unroll
	for (int i = 0; i < 16; ++i) { f->pose[i] = pose[i]; f->invTrack[i] = inv[i]; f->view[i] = view[i]; p.pose_dev[i] = pose[i]; }
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c)
			f->ca[3 * r + c] = p.tail.K.m[4 * r + 0] * inv[c] + p.tail.K.m[4 * r + 1] * inv[4 + c] + p.tail.K.m[4 * r + 2] * inv[8 + c];
	f->tz[0] = inv[8]; f->tz[1] = inv[9]; f->tz[2] = inv[10];
	f->tracked = tracked;
	f->do_integrate = (p.tail.force_integrate || (tracked && p.tail.rate_ok)) ? 1 : 0;
	volatile unsigned int* h = reinterpret_cast<volatile unsigned int*>(p.out_host);
	h[65] = (unsigned int) tracked;
	h[66] = (unsigned int) f->do_integrate;
}

#ifndef ICP_THREADS
#define ICP_THREADS 512   // one CTA per SM: half the partial rows for the last CTA to sum (measured 12.0 -> 11.3 us / iteration)
#endif
#ifndef ICP_PX
#define ICP_PX 2   // measured: 4 pixels per round is slower (5.3 vs 4.8 us of per-iteration compute: registers, longer reduction tail)
#endif
#define ICP_NW (ICP_THREADS / 32)          // warps per CTA
#define ICP_VPW (32 / ICP_NW)              // reduction outputs summed by each warp
#define ICP_SMEM_BYTES (32 * ICP_THREADS * sizeof(float))
__global__ void __launch_bounds__(ICP_THREADS, 512 / ICP_THREADS) k_icp(const __grid_constant__ IcpParams p) {
	__shared__ double sm[ICP_NW][32];
	extern __shared__ float xs_raw[];       // [32][ICP_THREADS]: per-thread sums, transposed (dynamic: 64 KB at 512 threads)
	float (*xs)[ICP_THREADS] = reinterpret_cast<float (*)[ICP_THREADS]>(xs_raw);
	__shared__ float red32[32];
	__shared__ Mat4 Tsh;
	__shared__ int flag_sh;   // 1: this CTA arrived last; 2: barrier time-out
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const unsigned int gen0 = ld_acquire_u32(p.bar + 1);   // nobody advances it before every CTA has arrived once
	unsigned int step = 0;
	const Mat4 V = p.view;

	for (int level = p.levels - 1; level >= 0; --level) {
		const uint32_t w = p.w[level], npx = p.w[level] * p.h[level];
		for (int it = 0; it < p.iterations[level]; ++it) {
			unsigned long long tp0 = 0, tp1 = 0, tp2 = 0, tp3 = 0;
			if (p.prof && threadIdx.x == 0) tp0 = gtime_ns();
			if (threadIdx.x < 16) Tsh.m[threadIdx.x] = (step == 0) ? p.pose0.m[threadIdx.x] : __ldcg(p.pose_dev + threadIdx.x);
			__syncthreads();
			const Mat4 T = Tsh;
			TrackAcc acc;
			acc.clear();
			// ICP_PX pixels per thread and round, all their loads in flight together
			for (uint32_t i = blockIdx.x * ICP_THREADS + threadIdx.x; i < npx; i += ICP_PX * gridDim.x * ICP_THREADS) {
				uint32_t pix[ICP_PX];
				bool on[ICP_PX];
#pragma unroll
				for (int u = 0; u < ICP_PX; ++u) { pix[u] = i + u * gridDim.x * ICP_THREADS; on[u] = pix[u] < npx; }
				track_pixels<ICP_PX>(acc, p.inV[level], p.inN[level], p.refV, p.refN, w, p.rw, p.rh, pix, on, T, V, p.dist_threshold,
						p.normal_threshold, p.status);
			}
			// thread -> CTA through shared memory, fp64, fixed order: value i of thread t sits at xs[i][t]; warp w
			// sums values ICP_VPW*w .. (lane l adds threads l, l+32, .. in order, then a 5-step butterfly)
			__syncthreads();
#pragma unroll
			for (int i = 0; i < 28; ++i) xs[i][threadIdx.x] = acc.s[i];
			xs[28][threadIdx.x] = (float) acc.c28; xs[29][threadIdx.x] = (float) acc.c29;
			xs[30][threadIdx.x] = (float) acc.c30; xs[31][threadIdx.x] = (float) acc.c31;
			__syncthreads();
#pragma unroll
			for (int q = 0; q < ICP_VPW; ++q) {
				const int i = wid * ICP_VPW + q;
				double v = 0;
#pragma unroll
				for (int k = 0; k < ICP_NW; ++k) v += (double) xs[i][lane + 32 * k];
				v = warp_sum(v);
				if (lane == 0) __stcg(p.partials + (size_t) blockIdx.x * 32 + i, v);
			}
			__threadfence();
			__syncthreads();
			++step;
			if (p.prof && threadIdx.x == 0) tp1 = gtime_ns();
			if (threadIdx.x == 0) flag_sh = (atomicAdd(p.bar, 1u) == gridDim.x - 1) ? 1 : 0;
			__syncthreads();
			if (flag_sh) {
				// last CTA: every partial row is visible (each writer fenced before its ticket).  Warp w sums rows
				// w, w + ICP_NW, ... with eight independent accumulators (loads overlap), always in the same order.
				__threadfence();
				double va[8];
#pragma unroll
				for (int j = 0; j < 8; ++j) va[j] = 0;
				uint32_t b = wid;
				const double* part = p.partials + lane;
				for (; b + 7 * ICP_NW < gridDim.x; b += 8 * ICP_NW) {
					double a[8];
#pragma unroll
					for (int j = 0; j < 8; ++j) a[j] = __ldcg(part + (size_t) (b + ICP_NW * j) * 32);
#pragma unroll
					for (int j = 0; j < 8; ++j) va[j] += a[j];
				}
				{
					double a[8];
#pragma unroll
					for (int j = 0; j < 8; ++j) a[j] = (b + ICP_NW * j < gridDim.x) ? __ldcg(part + (size_t) (b + ICP_NW * j) * 32) : 0.0;
#pragma unroll
					for (int j = 0; j < 8; ++j) va[j] += a[j];
				}
				sm[wid][lane] = ((va[0] + va[1]) + (va[2] + va[3])) + ((va[4] + va[5]) + (va[6] + va[7]));
				__syncthreads();
				if (wid == 0) {
					double t = 0;
#pragma unroll
					for (int k = 0; k < ICP_NW; ++k) t += sm[k][lane];
					const float r = (float) t;
					red32[lane] = r;
					p.out32[lane] = r;            // device copy; mirrored to the host once, at the end
					__syncwarp();
					if (lane == 0 && p.prof) tp2 = gtime_ns();
					const int c = icp_solve(p.pose_dev, Tsh.m, red32, p.icp_threshold, lane);   // the whole warp: lane-parallel Cholesky
					if (lane == 0) {
						if (p.prof) {   // last CTA: [0] own compute, [1] final reduce, [2] solve (ns, summed over iterations)
							tp3 = gtime_ns();
							atomicAdd(p.prof + 0, tp1 - tp0); atomicAdd(p.prof + 1, tp2 - tp1); atomicAdd(p.prof + 2, tp3 - tp2);
							atomicAdd(p.prof + 4, 1ull);
						}
						p.bar[2] = (unsigned int) c;
						p.bar[3] = step;                    // iterations executed so far
						p.bar[0] = 0;                       // re-arm the arrival counter
						st_release_u32(p.bar + 1, gen0 + step);
					}
				}
			} else if (threadIdx.x == 0) {
				unsigned int spins = 0;
				while (ld_relaxed_u32(p.bar + 1) != gen0 + step) {
					__nanosleep(64);
					if (++spins > (1u << 22)) { flag_sh = 2; break; }   // ~1 s: a lost CTA must not hang the GPU
				}
				__threadfence();   // acquire: the last CTA's pose / flags are visible below
			}
			__syncthreads();
			if (flag_sh == 2) {   // barrier time-out (cannot happen under a cooperative launch): report and leave
				if (threadIdx.x == 0) { reinterpret_cast<volatile unsigned int*>(p.out_host)[33] = 1u; __threadfence_system(); }
				return;
			}
			const unsigned int converged = __ldcg(p.bar + 2);
			if (p.prof && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.prof + 3, gtime_ns() - tp0);   // [3] whole iteration seen by CTA 0
			__syncthreads();
			if (converged) break;   // updatePoseKernel returned true: next level (:963-964)
		}
	}
	// results to the host, once per frame: sums of the last iteration, final pose, iteration count, then the flag
	if (blockIdx.x == 0 && wid == 0) {
		p.out_host[lane] = __ldcg(p.out32 + lane);
		if (p.tail.out) {
			// the frame's remaining host algebra, here: checkPoseKernel, inverse(pose), raycastPose * invK
			if (lane == 0) icp_tail(p);
			__syncwarp();
			__threadfence();
		}
		if (lane < 16) p.out_host[48 + lane] = __ldcg(p.pose_dev + lane);
		if (lane == 0) reinterpret_cast<volatile unsigned int*>(p.out_host)[64] = __ldcg(p.bar + 3);
		__syncwarp();
		__threadfence_system();
		if (lane == 0) *reinterpret_cast<volatile unsigned int*>(p.out_host + 32) = p.seq;
	}
}

// ------------------------------------------------------------------------------------------
// Brick flags: one byte per 8^3 voxels, set (never cleared) as soon as any voxel of the brick — or of the
// one-voxel halo on its upper sides, so that the 2x2x2 taps of a trilinear sample whose base voxel lies in
// the brick are all covered — holds a tsdf below BRICK_T.  A clear flag therefore PROVES that every tap of
// such a sample is >= BRICK_T, i.e. the interpolated value is >= 0.82 > 0.8: the raycaster's decisions for
// that sample (`f < 0` and `f < 0.8`, cpp/kernels.cpp:711-714) are known without reading a voxel.
// Maintained by integrate (and rebuilt when a volume is written from the host); 256 KB at 512^3.
// ------------------------------------------------------------------------------------------
#define BRICK_T 27000
#define BRICK_SHIFT 3
#define KFB_MAX_SLABS 8
struct BrickMap {
	unsigned char* flag;     // [bnz][bny][bnx]; nullptr = not maintained
	uint32_t bnx, bny, bnz;
	// second level, same rule one octave up: one byte per 8^3 BRICKS (64^3 voxels), set whenever one of its bricks is.  It
	// lives right behind the brick flags in the same allocation (so a peer's copy is found the same way): the raycaster
	// leaps whole runs of COARSE steps through a clear super-brick (raycast_one).
	size_t n_bricks;         // bnx * bny * bnz: offset of the super-brick flags
	uint32_t snx, sny;
	// z-slab mode over peer memory: every rank keeps a map of the WHOLE volume, and a rank that flags a brick of its slab
	// stores the byte into all peers' maps as well (NVLink P2P; "set to 1" is idempotent, so there is nothing to merge)
	unsigned char* peer[KFB_MAX_SLABS - 1];
	int n_peer;
};
__device__ __forceinline__ void brick_set(const BrickMap& b, uint32_t bx, uint32_t by, uint32_t bz) {
	const size_t idx = ((size_t) bz * b.bny + by) * b.bnx + bx;
	const size_t sidx = b.n_bricks + ((size_t) (bz >> 3) * b.sny + (by >> 3)) * b.snx + (bx >> 3);
	// Flags are never cleared between resets, so after the first frames almost every call finds its flag already set: look
	// before storing (a local L1/L2 hit) — on a z-slab group each store would otherwise be 2 x (1 + peers) single-byte writes,
	// most of them over NVLink (measured on 8 GPUs: the surface slabs of 1024^3 577 -> 157 us, 1024^3 860 -> 1329 fps, 2048^3
	// 298 -> 478 fps).  Whoever set the local flag
	// (this rank, or a peer whose halo reaches here) stored to every map in the same call, and the barrier that ends
	// integrate orders those stores before any raycast.
	if (__ldcg(b.flag + idx) != 0 && __ldcg(b.flag + sidx) != 0) return;
	b.flag[idx] = 1; b.flag[sidx] = 1;
	for (int i = 0; i < b.n_peer; ++i) { b.peer[i][idx] = 1; b.peer[i][sidx] = 1; }
}
__device__ __noinline__ void brick_mark(const BrickMap b, uint32_t x, uint32_t y, uint32_t z) {
	const uint32_t bx1 = x >> BRICK_SHIFT, by1 = y >> BRICK_SHIFT, bz1 = z >> BRICK_SHIFT;
	const uint32_t bx0 = (((x & 7u) == 0u) && x) ? bx1 - 1 : bx1, by0 = (((y & 7u) == 0u) && y) ? by1 - 1 : by1,
			bz0 = (((z & 7u) == 0u) && z) ? bz1 - 1 : bz1;
	for (uint32_t bz = bz0; bz <= bz1; ++bz)
		for (uint32_t by = by0; by <= by1; ++by)
			for (uint32_t bx = bx0; bx <= bx1; ++bx) {
				if (b.flag[((size_t) bz * b.bny + by) * b.bnx + bx] == 0) brick_set(b, bx, by, bz);
			}
}
// rebuild from a volume (after kfb_write_buffer / for tests)
__global__ void __launch_bounds__(256) k_brick_rebuild(BrickMap b, const short2* __restrict__ vol, uint32_t sx, uint32_t sy, uint32_t sz, uint32_t z_begin) {
	const size_t n = (size_t) sx * sy * sz, stride = (size_t) gridDim.x * blockDim.x;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		if (vol[i].x < BRICK_T) {
			const uint32_t x = (uint32_t) (i % sx), y = (uint32_t) ((i / sx) % sy), z = z_begin + (uint32_t) (i / ((size_t) sx * sy));
			brick_mark(b, x, y, z);
		}
	}
}

// ------------------------------------------------------------------------------------------
// integrateKernel (cpp/kernels.cpp:628-673): TSDF running average over the voxels that
// project into the depth image.  The reference walks each (x,y) column from z = 0 and
// advances `pos` / `cameraX` by REPEATED fp32 addition; reproducing those exact values is
// what keeps every voxel bit-identical (recomputing pos0 + z*delta is off by several LSB,
// SURVEY §7).  Only ~8 % of the voxels are updated by a frame, and memory is touched for
// those alone, so the cost that matters is DECIDING, for the other 92 %, that nothing
// happens.  Three layers keep that cheap without changing a single result:
//   1. per column, a conservative z-interval [za, zb) outside which the voxel provably fails
//      the reference's `pos.z < 1e-4` / pixel-bounds tests (closed form on the un-rounded
//      line, widened by a bound on the accumulated rounding error); a warp (32 x-adjacent
//      columns -> 128 contiguous bytes per z-slice) takes the union of its lanes' intervals;
//   2. the additions up to the interval start are REPLAYED in registers (6 independent FADD
//      chains, no memory traffic, no tests) so the first visited voxel sees the reference's
//      exact accumulated values;
//   3. inside the interval the pixel is first located with an approximate division; only when
//      the quotient is within 1e-3 of an integer (a pixel edge or the image border) is the
//      IEEE division executed.  `e = depth - cameraX.z` then decides exactly, by monotonicity
//      of correctly-rounded * and / and lambda = sqrt(1 + ..) >= 1:  e > mu  =>  sdf == 1;
//      e < -mu  =>  no update;  otherwise the reference's full sqrt/division expression runs.
// The kernel is launched on the slab [z_begin, z_end) this context owns; `vol` points at the
// slab's first voxel.  N_upd (voxels actually updated) is counted exactly.
// ------------------------------------------------------------------------------------------
#ifndef INT_U
#define INT_U 8   // slices per batch (must divide 8: a batch never straddles a brick layer)
#endif
struct IntegrateParams {
	short2* vol;
	uint32_t sx, sy, sz;       // full volume resolution
	float dx, dy, dz;          // volume dimensions (metres)
	uint32_t z_begin, z_end;   // slab owned by this context
	uint32_t zchunk;           // z-steps per blockIdx.z
	const float* depth; uint32_t dw, dh;
	Mat4 invTrack, K;
	const DevFrame* dev;       // optional: invTrack (and the integrate gate) come from the ICP kernel's tail instead
	float mu, maxweight;
	const float* dmax;         // optional: max of the depth image (device scalar); nullptr = unknown
	int cull;                  // 0 = visit every voxel (debug / A-B), 1 = interval + fast tests
	unsigned long long* n_upd;
	BrickMap brick;            // flags for the raycaster (flag == nullptr: not maintained)
	uint2* queue;              // work list: warp-columns with a non-empty visited interval
	unsigned int* piece_ctr;   // per work-list entry: next unclaimed piece
	unsigned int* queue_count; // [0] items appended by the plan pass, [1] next item (this frame's slot)
	unsigned int* queue_head;
	unsigned int* queue_next;  // the other slot's two counters, zeroed by the plan pass for the next frame
};

__device__ __forceinline__ float rcp_approx(float x) {   // MUFU.RCP, <= 1 ulp, no denormal handling
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}
// packed fp32 pair add (Blackwell FADD2: one issue slot for two IEEE additions)
struct F2 { unsigned long long v; };
__device__ __forceinline__ F2 f2_make(float lo, float hi) {
	F2 r;
	r.v = ((unsigned long long) __float_as_uint(hi) << 32) | (unsigned long long) __float_as_uint(lo);
	return r;
}
__device__ __forceinline__ float f2_lo(F2 a) { return __uint_as_float((unsigned int) a.v); }
__device__ __forceinline__ float f2_hi(F2 a) { return __uint_as_float((unsigned int) (a.v >> 32)); }
__device__ __forceinline__ F2 f2_add(F2 a, F2 b) {
	F2 r;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
	return r;
}

// tighten [tlo, thi] with the constraint a + b*t >= -s.  Approximate division: the caller widens the
// result by a whole voxel.  NaNs never tighten; b == 0 decides on a alone.
__device__ __forceinline__ void clip_line(float a, float b, float s, float& tlo, float& thi) {
	const float r = (-s - a) * rcp_approx(b);
	if (b > 0.f) { if (r > tlo) tlo = r; }
	else if (b < 0.f) { if (r < thi) thi = r; }
	else if (a < -s) { tlo = 1.f; thi = 0.f; }
}

// the reference's full per-voxel expression (cpp/kernels.cpp:647-661): returns sdf, or -4 for "no update"
__device__ __noinline__ float integrate_exact(float Px, float Py, float Pz, float Cx, float Cy, float Cz, const float* __restrict__ depth,
		uint32_t dw, float dwm1, float dhm1, float mu) {
	const float pxf = Cx / Cz + 0.5f, pyf = Cy / Cz + 0.5f;
	if (pxf < 0 || pxf > dwm1 || pyf < 0 || pyf > dhm1) return -4.f;
	const uint32_t px = (uint32_t) pxf, py = (uint32_t) pyf;
	const float d = __ldg(depth + (px + py * dw));
	if (d == 0) return -4.f;
	const float diff = (d - Cz) * sqrtf(1 + ksq(Px / Pz) + ksq(Py / Pz));
	if (diff > -mu) return kminf(1.f, diff / mu);
	return -4.f;
}
// same, for a voxel whose pixel is already known exactly: only the sdf part (:655-661); e = depth[px] - cameraX.z
__device__ __noinline__ float integrate_exact_sdf(float Px, float Py, float Pz, float e, float mu) {
	const float diff = e * sqrtf(1 + ksq(Px / Pz) + ksq(Py / Pz));
	if (diff > -mu) return kminf(1.f, diff / mu);
	return -4.f;
}

// per-column state at z = 0 (cpp/kernels.cpp:638-644) and the conservative visited interval of one column
struct IntColumn { float3 pos0, cam0, delta, cameraDelta; };
__device__ __forceinline__ IntColumn int_column(const IntegrateParams& p, const Mat4& invTrack, uint32_t x, uint32_t y) {
	IntColumn c;
	c.delta = mat_rotate(invTrack, f3(0, 0, p.dz / (float) p.sz));
	c.cameraDelta = mat_rotate(p.K, c.delta);
	// Volume::pos (commons.h:186-189) at z = 0
	c.pos0 = mat_point(invTrack, f3(((float) x + 0.5f) * p.dx / (float) p.sx, ((float) y + 0.5f) * p.dy / (float) p.sy,
			(0 + 0.5f) * p.dz / (float) p.sz));
	c.cam0 = mat_point(p.K, c.pos0);
	return c;
}
__device__ __forceinline__ IntColumn int_column(const IntegrateParams& p, uint32_t x, uint32_t y) { return int_column(p, p.invTrack, x, y); }

// Pass 1: one warp per 32 x-adjacent columns: conservative interval, appended to the work list.
// entry = { xtile | y << 16, za | zb << 16 } plus a zeroed piece counter.
__global__ void __launch_bounds__(256) k_integrate_plan(IntegrateParams p) {
	const uint32_t x = blockIdx.x * 32 + threadIdx.x;
	const uint32_t y = blockIdx.y * blockDim.y + threadIdx.y;   // warp-uniform
	if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0) { p.queue_next[0] = 0u; p.queue_next[1] = 0u; }  // re-arm the other slot
	const bool valid = x < p.sx && y < p.sy;
	int za = (int) p.z_begin, zb = (int) p.z_end;
	if (!valid) { za = 0x7fffffff; zb = 0; }
	else if (p.cull) {
		const IntColumn c = int_column(p, x, y);
		const float3 pos0 = c.pos0, cam0 = c.cam0, delta = c.delta, cameraDelta = c.cameraDelta;
		const float dwm1 = (float) (p.dw - 1), dhm1 = (float) (p.dh - 1);
		// conservative interval on the un-rounded line; slack = bound on the accumulated rounding error
		const float n = (float) p.sz;
		const float Mx = kmaxf(fabsf(cam0.x), fabsf(cam0.x + n * cameraDelta.x));
		const float My = kmaxf(fabsf(cam0.y), fabsf(cam0.y + n * cameraDelta.y));
		const float Mz = kmaxf(fabsf(cam0.z), fabsf(cam0.z + n * cameraDelta.z));
		const float Mp = kmaxf(fabsf(pos0.z), fabsf(pos0.z + n * delta.z));
		const float eps = (n + 64.f) * 2.3841858e-7f;   // (N + 64) * 2^-22: > 2x the worst-case drift of N additions
		const float wq = dwm1 - 0.5f, hq = dhm1 - 0.5f;
		float tlo = -1e9f, thi = 1e9f;
		clip_line(pos0.z - 0.0001f, delta.z, eps * Mp, tlo, thi);                                            // pos.z >= 1e-4
		clip_line(cam0.x + 0.5f * cam0.z, cameraDelta.x + 0.5f * cameraDelta.z, eps * (Mx + Mz), tlo, thi);   // px >= 0
		clip_line(wq * cam0.z - cam0.x, wq * cameraDelta.z - cameraDelta.x, eps * (Mx + (wq + 2.f) * Mz), tlo, thi);  // px <= w-1
		clip_line(cam0.y + 0.5f * cam0.z, cameraDelta.y + 0.5f * cameraDelta.z, eps * (My + Mz), tlo, thi);   // py >= 0
		clip_line(hq * cam0.z - cam0.y, hq * cameraDelta.z - cameraDelta.y, eps * (My + (hq + 2.f) * Mz), tlo, thi);  // py <= h-1
		if (p.dmax) {
			// no update unless depth - cameraX.z > -mu  (lambda >= 1)  =>  cameraX.z < max(depth) + mu
			const float far = *p.dmax + p.mu;
			if (far == far) clip_line(far - cam0.z, -cameraDelta.z, eps * Mz + 1e-6f * fabsf(far), tlo, thi);
		}
		if (tlo > thi) { za = 0x7fffffff; zb = 0; }
		else {
			const float lo = floorf(tlo) - 2.f, hi = ceilf(thi) + 3.f;   // approximate quotient + rounding: a voxel each side
			if (lo > (float) za) za = (lo < 2e9f) ? (int) lo : 0x7fffffff;
			if (hi < (float) zb) zb = (hi > -2e9f) ? (int) hi : 0;
		}
	}
	za = __reduce_min_sync(0xffffffffu, za);
	zb = __reduce_max_sync(0xffffffffu, zb);
	if (za >= zb) return;
	za = max((int) p.z_begin, za & ~(INT_U - 1));   // batches on multiples of INT_U (brick layers); still conservative
	if (threadIdx.x == 0) {
		const uint32_t e = atomicAdd(p.queue_count, 1u);
		p.queue[e] = make_uint2(blockIdx.x | (y << 16), (uint32_t) za | ((uint32_t) zb << 16));
		p.piece_ctr[e] = 0u;
	}
}

// Pass 2: persistent warps.  A warp takes a warp-column from the list and sweeps it piece by piece (`zchunk`
// slices, claimed one at a time from the column's counter), so the reference's additions are replayed only once
// per column; warps that run out of fresh columns lap around the list and claim pieces of columns still in
// progress (they replay the additions up to their piece), which removes the tail.  Every piece is processed
// exactly once.
#define INT_LAPS 4
#ifndef KFB_INT_MINBLOCKS
#define KFB_INT_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(256, KFB_INT_MINBLOCKS) k_integrate_run(IntegrateParams p) {
	const uint32_t lane = threadIdx.x & 31;
	const float dwm1 = (float) (p.dw - 1), dhm1 = (float) (p.dh - 1);
	// fast tests need: mu > 0, image small enough that an approximate quotient beyond +-2048 is conclusively outside
	const bool fast = p.cull && p.mu > 0.f && p.dw <= 2040 && p.dh <= 2040;
	// |approximate - exact| <= |q| 2^-22 + ulp: scale the "too close to an integer" band with the image size
	const float tol = (float) max(p.dw, p.dh) * 4.0e-7f + 1.0e-5f;
	const float mu = p.mu;
	const float* __restrict__ depth = p.depth;
	const uint32_t dw = p.dw;
	const size_t plane = (size_t) p.sx * p.sy;
	const unsigned int count = *p.queue_count;
	const unsigned int visits = count * INT_LAPS;
	unsigned int updated = 0;

	for (;;) {
		unsigned int it = 0;
		if (lane == 0) it = atomicAdd(p.queue_head, 1u);
		it = __shfl_sync(0xffffffffu, it, 0);
		if (it >= visits) break;
		const unsigned int e = it % count;
		const uint2 item = __ldcg(p.queue + e);
		const uint32_t x = (item.x & 0xffffu) * 32 + lane, y = item.x >> 16;
		const int col_za = (int) (item.y & 0xffffu), col_zb = (int) (item.y >> 16);
		const unsigned int npieces = ((unsigned int) (col_zb - col_za) + p.zchunk - 1) / p.zchunk;
		const bool valid = x < p.sx;
		F2 A, B, C, dA, dB, dC;
		int zstate = -1;   // slice the running values correspond to; -1 = not computed yet
	  for (;;) {
		unsigned int pc = 0;
		if (lane == 0) pc = atomicAdd(p.piece_ctr + e, 1u);
		pc = __shfl_sync(0xffffffffu, pc, 0);
		if (pc >= npieces) break;
		const int za = col_za + (int) (pc * p.zchunk), zb = min(col_zb, za + (int) p.zchunk);
		if (zstate < 0 || zstate > za) {
			const IntColumn c = int_column(p, x, y);
			// running values, packed in pairs: (pos.x, pos.y) (pos.z, cam.x) (cam.y, cam.z)
			A = f2_make(c.pos0.x, c.pos0.y); B = f2_make(c.pos0.z, c.cam0.x); C = f2_make(c.cam0.y, c.cam0.z);
			dA = f2_make(c.delta.x, c.delta.y); dB = f2_make(c.delta.z, c.cameraDelta.x); dC = f2_make(c.cameraDelta.y, c.cameraDelta.z);
			zstate = 0;
		}
		// replay the reference's additions up to the first slice of the piece (nothing to do when the warp
		// continues from the previous piece of the same column)
#pragma unroll 8
		for (int z = zstate; z < za; ++z) { A = f2_add(A, dA); B = f2_add(B, dB); C = f2_add(C, dC); }
		zstate = za + (int) (((unsigned int) (zb - za) + INT_U - 1) / INT_U * INT_U);   // the batches below advance in units of INT_U
		short2* col = p.vol + (size_t) x + (size_t) y * p.sx + (size_t) ((uint32_t) za - p.z_begin) * plane;
		// INT_U consecutive slices per batch: all decisions first, straight-line (depth gathers hit L1/L2), then
		// all voxel loads back to back (INT_U independent 128-byte requests in flight per warp: the
		// read-modify-write is latency-bound otherwise), then the updates and stores.
		for (int z = za; z < zb; z += INT_U, col += INT_U * plane) {
			float sdf[INT_U];
			short2 v[INT_U];
			unsigned int low = 0;   // bit u: this lane stored a tsdf below BRICK_T in slice z + u
#pragma unroll
			for (int u = 0; u < INT_U; ++u) {
				const float Px = f2_lo(A), Py = f2_hi(A), Pz = f2_lo(B), Cx = f2_hi(B), Cy = f2_lo(C), Cz = f2_hi(C);
				A = f2_add(A, dA); B = f2_add(B, dB); C = f2_add(C, dC);
				const bool act = valid && (z + u < zb) && !(Pz < 0.0001f);
				float s = -4.f;   // "no update" (a real sdf is > -1)
				if (fast) {
					const float r = rcp_approx(Cz);
					const float pxf = Cx * r + 0.5f, pyf = Cy * r + 0.5f;
					// the truncated pixel and the bounds tests (against the integers 0, w-1, h-1) can only differ from
					// the exact ones when the value is within `tol` of an integer.  NaN/inf fail both comparisons.
					const bool sure = (fabsf(pxf - rintf(pxf)) >= tol) && (fabsf(pyf - rintf(pyf)) >= tol);
					const bool inb = !(pxf < 0 || pxf > dwm1 || pyf < 0 || pyf > dhm1);
					const uint32_t idx = inb ? ((uint32_t) pxf + (uint32_t) pyf * dw) : 0u;
					const float d = __ldg(depth + idx);
					const float e = d - Cz;
					// sure & inside: e > mu => sdf == 1 exactly; e < -mu or d == 0 => no update (header comment);
					// sure & outside: no update; |e| <= mu: the reference's sqrt/division expression; not sure: all of it
					if (act && sure && inb && e > mu && d != 0) s = 1.f;
					if (act && sure && inb && d != 0 && !(e > mu) && !(e < -mu)) s = integrate_exact_sdf(Px, Py, Pz, e, mu);
					if (act && !sure) s = integrate_exact(Px, Py, Pz, Cx, Cy, Cz, depth, dw, dwm1, dhm1, mu);
				} else if (act) s = integrate_exact(Px, Py, Pz, Cx, Cy, Cz, depth, dw, dwm1, dhm1, mu);
				sdf[u] = s;
			}
#pragma unroll
			for (int u = 0; u < INT_U; ++u)
				if (sdf[u] > -2.f) v[u] = __ldcs(col + u * plane);
#pragma unroll
			for (int u = 0; u < INT_U; ++u) low |= (sdf[u] > -2.f && sdf[u] < 0.9f) ? (1u << u) : 0u;
			// Brick flags for the raycaster, while the voxel loads are in flight.  A stored tsdf can only fall below
			// BRICK_T (0.824) in an update whose sdf is below 0.9: with sdf >= 0.9 the running average
			// (w t + sdf) / (w + 1), w + 1 <= 101, either stays above 0.89 or rises by more than the 1 LSB the truncation
			// can take away.  So flagging every slice that holds an update with sdf < 0.9 keeps the invariant
			// "flag clear => every voxel of the brick (and halo) >= BRICK_T" without looking at the stored values.
			// The warp's 32 voxels of a slice span 4 bricks in x (lanes 8k..8k+7 -> brick k; lane 8k also lies in the
			// halo of brick k-1), one or two in y and in z: lanes 0..4 each flag one x-brick.
			// Batches start at multiples of 8 when the flags are maintained (the plan pass aligns za), so a batch is
			// one brick layer bz = z / 8 plus, through its first slice, the halo of layer bz - 1.
			if (p.brick.flag) {
				const unsigned int m_any = __ballot_sync(0xffffffffu, low != 0u), m_0 = ((z & 7) == 0) ? __ballot_sync(0xffffffffu, low & 1u) : 0u;
				if (m_any && lane < 5) {
					const int j = (int) lane - 1;   // x-brick relative to the warp's first brick: -1 (halo of the previous tile) .. 3
					unsigned int any = 0, first = 0;
					if (j >= 0) { any = (m_any >> (8 * j)) & 0xffu; first = (m_0 >> (8 * j)) & 0xffu; }
					if (j < 3) { any |= (m_any >> (8 * (j + 1))) & 1u; first |= (m_0 >> (8 * (j + 1))) & 1u; }   // next brick's first voxel: halo
					const uint32_t xb = (item.x & 0xffffu) * 4, bx = xb + (uint32_t) j;
					if (any && (j >= 0 || xb > 0) && bx < p.brick.bnx) {
						const uint32_t by1 = y >> BRICK_SHIFT, by0 = (((y & 7u) == 0u) && y) ? by1 - 1 : by1;
						const uint32_t bz1 = (uint32_t) z >> BRICK_SHIFT, bz0 = (first && bz1) ? bz1 - 1 : bz1;
						for (uint32_t bz = bz0; bz <= bz1; ++bz)
							for (uint32_t by = by0; by <= by1; ++by) brick_set(p.brick, bx, by, bz);
					}
				}
			}
#pragma unroll
			for (int u = 0; u < INT_U; ++u)
				if (sdf[u] > -2.f) {
					float tsdf = (float) v[u].x * 0.00003051944088f, wgt = (float) v[u].y;   // commons.h:160-163
					tsdf = kclampf((wgt * tsdf + sdf[u]) / (wgt + 1), -1.f, 1.f);
					wgt = kminf(wgt + 1, p.maxweight);
					__stcs(col + u * plane, make_short2((short) (int) (tsdf * 32766.0f), (short) (int) wgt)); // commons.h:182-185 (truncation)
					++updated;
				}
		}
	  }
	}
	// exact N_upd: one atomic per warp
	updated = __reduce_add_sync(0xffffffffu, updated);
	if (lane == 0 && updated) atomicAdd(p.n_upd, (unsigned long long) updated);
}

// ------------------------------------------------------------------------------------------
// raycastKernel (cpp/kernels.cpp:674-757) with Volume::interp / Volume::grad
// (commons.h:191-301).  One thread per pixel; `t` advances by repeated fp32 addition and the
// coarse->fine step switch is history dependent, so rays are marched exactly as the reference
// does.  The volume may be split into z-slabs owned by different GPUs: slab s holds
// z in [slab_z[s], slab_z[s+1]) at slab_ptr[s] (peer memory over NVLink when s is remote).
// ------------------------------------------------------------------------------------------
struct VolView {
	const short2* slab_ptr[KFB_MAX_SLABS];
	uint32_t slab_z[KFB_MAX_SLABS + 1];
	int n_slabs;
	uint32_t sx, sy, sz;
	float dx, dy, dz;
	// 1/dim (correctly rounded) and "the 3-instruction division by this constant is exact" (verified exhaustively
	// over all 2^23 significands on the host at kfb_create, see kfb_fastdiv_ok)
	float rdx, rdy, rdz;
	int fastdiv;
	const unsigned char* brick;   // brick flags (see BrickMap) or nullptr
	uint32_t bnx, bny;
	const unsigned char* super;   // super-brick flags (64^3 voxels each) or nullptr: coarse steps are then taken one by one
	uint32_t snx, sny;
	int no_leap;                  // A/B switch (KFB_RAY_NO_LEAP=1): take every fine step through clear bricks one by one
};

// fl(a / d) for a per-launch constant d with rd = fl(1/d): q = a*rd, r = fma(-d, q, a) (exact), q + r*rd rounds
// correctly (Markstein).  Only used when the host verified it for this d; out-of-range a takes the IEEE path.
__device__ __forceinline__ float div_const(float a, float d, float rd, int ok) {
	if (ok && fabsf(a) < 1e30f && fabsf(a) > 1e-30f) {
		const float q = a * rd;
		return __fmaf_rn(__fmaf_rn(-d, q, a), rd, q);
	}
	return a / d;
}

__device__ __forceinline__ float vol_vs2(const VolView& v, int x, int y, int z) {  // commons.h:172-174
	int s = 0;
	if (v.n_slabs > 1) {
#pragma unroll
		for (int i = 1; i < KFB_MAX_SLABS; ++i) s += (i < v.n_slabs && (uint32_t) z >= v.slab_z[i]);
	}
	const short2* base = v.slab_ptr[s];
	const size_t idx = (size_t) x + (size_t) y * v.sx + (size_t) ((uint32_t) z - v.slab_z[s]) * v.sx * v.sy;
	return (float) __ldg(reinterpret_cast<const short*>(base + idx));
}

__device__ __forceinline__ const short2* vol_plane(const VolView& v, int z) {   // first voxel of slice z (maybe in a peer's slab)
	int s = 0;
	if (v.n_slabs > 1) {
#pragma unroll
		for (int i = 1; i < KFB_MAX_SLABS; ++i) s += (i < v.n_slabs && (uint32_t) z >= v.slab_z[i]);
	}
	return v.slab_ptr[s] + (size_t) ((uint32_t) z - v.slab_z[s]) * v.sx * v.sy;
}

__device__ __forceinline__ float div_cell(float a, float d, float rd, int ok) {
	if (ok) {
		const float q = a * rd;
		return __fmaf_rn(__fmaf_rn(-d, q, a), rd, q);
	}
	return a / d;
}
struct VolCell { int bx, by, bz; float fx, fy, fz; };
__device__ __forceinline__ VolCell vol_cell(const VolView& v, float3 pos) {   // commons.h:192-197: scaled position -> base voxel + fraction
	// div_cell: the range guard of div_const is not needed here — a zero / denormal numerator gives 0 - 0.5 = -0.5 on
	// both paths, and a non-finite position is outside anything the reference can sample
	const float spx = div_cell(pos.x * (float) v.sx, v.dx, v.rdx, v.fastdiv) - 0.5f;
	const float spy = div_cell(pos.y * (float) v.sy, v.dy, v.rdy, v.fastdiv) - 0.5f;
	const float spz = div_cell(pos.z * (float) v.sz, v.dz, v.rdz, v.fastdiv) - 0.5f;
	const float flx = floorf(spx), fly = floorf(spy), flz = floorf(spz);
	VolCell c;
	// base is in [-1, N-1] for every position the reference samples (inside the volume box); clamping the base itself
	// changes nothing there and keeps the taps in bounds for any other position (NaN pose, renderVolume's 2x far plane)
	c.bx = kmini(kmaxi((int) flx, -1), (int) v.sx - 1); c.by = kmini(kmaxi((int) fly, -1), (int) v.sy - 1);
	c.bz = kmini(kmaxi((int) flz, -1), (int) v.sz - 1);
	c.fx = spx - flx; c.fy = spy - fly; c.fz = spz - flz;
	return c;
}
// true: all 8 taps of this cell are proven >= BRICK_T (the sample is >= 0.82) without touching the volume
__device__ __forceinline__ bool vol_cell_free(const VolView& v, const VolCell& c) {
	if (!v.brick) return false;
	const uint32_t bx = (uint32_t) kmaxi(c.bx, 0) >> BRICK_SHIFT, by = (uint32_t) kmaxi(c.by, 0) >> BRICK_SHIFT, bz = (uint32_t) kmaxi(c.bz, 0) >> BRICK_SHIFT;
	return __ldg(v.brick + ((bz * v.bny + by) * v.bnx + bx)) == 0;   // < 2^32 bricks
}
__device__ __forceinline__ float vol_interp_cell(const VolView& v, const VolCell& c) {  // commons.h:198-212
	const int bx = c.bx, by = c.by, bz = c.bz;
	const float fx = c.fx, fy = c.fy, fz = c.fz;
	const int lx = kmaxi(bx, 0), ly = kmaxi(by, 0), lz = kmaxi(bz, 0);
	const int ux = kmini(bx + 1, (int) v.sx - 1), uy = kmini(by + 1, (int) v.sy - 1), uz = kmini(bz + 1, (int) v.sz - 1);
	// two slice bases (z may straddle slabs), two row offsets, two column offsets: 8 taps from 3 adds each
	const short2* pl = vol_plane(v, lz);
	const short2* pu = (v.n_slabs > 1) ? vol_plane(v, uz) : pl + (size_t) (uz - lz) * v.sx * v.sy;
	const uint32_t rl = (uint32_t) ly * v.sx, ru = (uint32_t) uy * v.sx;
#define TAP(P, R, X) ((float) __ldg(reinterpret_cast<const short*>((P) + ((R) + (uint32_t) (X)))))
	const float v000 = TAP(pl, rl, lx), v100 = TAP(pl, rl, ux), v010 = TAP(pl, ru, lx), v110 = TAP(pl, ru, ux);
	const float v001 = TAP(pu, rl, lx), v101 = TAP(pu, rl, ux), v011 = TAP(pu, ru, lx), v111 = TAP(pu, ru, ux);
#undef TAP
	const float gx = 1 - fx, gy = 1 - fy, gz = 1 - fz;
	return (((v000 * gx + v100 * fx) * gy + (v010 * gx + v110 * fx) * fy) * gz
			+ ((v001 * gx + v101 * fx) * gy + (v011 * gx + v111 * fx) * fy) * fz) * 0.00003051944088f;
}
__device__ __forceinline__ float vol_interp(const VolView& v, float3 pos) {  // commons.h:191-213
	return vol_interp_cell(v, vol_cell(v, pos));
}

__device__ __forceinline__ float3 vol_grad(const VolView& v, float3 pos) {  // commons.h:215-301
	const float3 sp = f3(div_const(pos.x * (float) v.sx, v.dx, v.rdx, v.fastdiv) - 0.5f, div_const(pos.y * (float) v.sy, v.dy, v.rdy, v.fastdiv) - 0.5f,
			div_const(pos.z * (float) v.sz, v.dz, v.rdz, v.fastdiv) - 0.5f);
	const float flx = floorf(sp.x), fly = floorf(sp.y), flz = floorf(sp.z);
	const int bx = (int) flx, by = (int) fly, bz = (int) flz;
	const float3 f = f3(sp.x - flx, sp.y - fly, sp.z - flz);
	const int mx = (int) v.sx - 1, my = (int) v.sy - 1, mz = (int) v.sz - 1;
	const int llx = kmaxi(bx - 1, 0), lly = kmaxi(by - 1, 0), llz = kmaxi(bz - 1, 0);     // lower_lower
	const int lx = kmaxi(bx, 0), ly = kmaxi(by, 0), lz = kmaxi(bz, 0);                    // lower_upper == lower
	const int ux = kmini(bx + 1, mx), uy = kmini(by + 1, my), uz = kmini(bz + 1, mz);     // upper_lower == upper
	const int uux = kmini(bx + 2, mx), uuy = kmini(by + 2, my), uuz = kmini(bz + 2, mz);  // upper_upper
	float3 g;
	// the 32 taps come from 4 slices x 4 rows x 4 columns: slice bases (each maybe in a peer's slab) and row offsets once,
	// then every tap is two adds away (vol_vs2 recomputes a 64-bit index per tap)
	const short2* Pll = vol_plane(v, llz);
	const short2* Pl = vol_plane(v, lz);
	const short2* Pu = vol_plane(v, uz);
	const short2* Puu = vol_plane(v, uuz);
	const uint32_t Rll = (uint32_t) lly * v.sx, Rl = (uint32_t) ly * v.sx, Ru = (uint32_t) uy * v.sx, Ruu = (uint32_t) uuy * v.sx;
#define VS(X, R, P) ((float) __ldg(reinterpret_cast<const short*>((P) + ((R) + (uint32_t) (X)))))
	g.x = (((VS(ux, Rl, Pl) - VS(llx, Rl, Pl)) * (1 - f.x) + (VS(uux, Rl, Pl) - VS(lx, Rl, Pl)) * f.x) * (1 - f.y)
			+ ((VS(ux, Ru, Pl) - VS(llx, Ru, Pl)) * (1 - f.x) + (VS(uux, Ru, Pl) - VS(lx, Ru, Pl)) * f.x) * f.y) * (1 - f.z)
			+ (((VS(ux, Rl, Pu) - VS(llx, Rl, Pu)) * (1 - f.x) + (VS(uux, Rl, Pu) - VS(lx, Rl, Pu)) * f.x) * (1 - f.y)
					+ ((VS(ux, Ru, Pu) - VS(llx, Ru, Pu)) * (1 - f.x) + (VS(uux, Ru, Pu) - VS(lx, Ru, Pu)) * f.x) * f.y) * f.z;
	g.y = (((VS(lx, Ru, Pl) - VS(lx, Rll, Pl)) * (1 - f.x) + (VS(ux, Ru, Pl) - VS(ux, Rll, Pl)) * f.x) * (1 - f.y)
			+ ((VS(lx, Ruu, Pl) - VS(lx, Rl, Pl)) * (1 - f.x) + (VS(ux, Ruu, Pl) - VS(ux, Rl, Pl)) * f.x) * f.y) * (1 - f.z)
			+ (((VS(lx, Ru, Pu) - VS(lx, Rll, Pu)) * (1 - f.x) + (VS(ux, Ru, Pu) - VS(ux, Rll, Pu)) * f.x) * (1 - f.y)
					+ ((VS(lx, Ruu, Pu) - VS(lx, Rl, Pu)) * (1 - f.x) + (VS(ux, Ruu, Pu) - VS(ux, Rl, Pu)) * f.x) * f.y) * f.z;
	g.z = (((VS(lx, Rl, Pu) - VS(lx, Rl, Pll)) * (1 - f.x) + (VS(ux, Rl, Pu) - VS(ux, Rl, Pll)) * f.x) * (1 - f.y)
			+ ((VS(lx, Ru, Pu) - VS(lx, Ru, Pll)) * (1 - f.x) + (VS(ux, Ru, Pu) - VS(ux, Ru, Pll)) * f.x) * f.y) * (1 - f.z)
			+ (((VS(lx, Rl, Puu) - VS(lx, Rl, Pl)) * (1 - f.x) + (VS(ux, Rl, Puu) - VS(ux, Rl, Pl)) * f.x) * (1 - f.y)
					+ ((VS(lx, Ru, Puu) - VS(lx, Ru, Pl)) * (1 - f.x) + (VS(ux, Ru, Puu) - VS(ux, Ru, Pl)) * f.x) * f.y) * f.z;
#undef VS
	return g * f3(v.dx / (float) v.sx, v.dy / (float) v.sy, v.dz / (float) v.sz) * (0.5f * 0.00003051944088f);
}

#ifndef RAY_STATS
#define RAY_STATS 0   // development: KFB_BUF_RAYTILECOST holds, per tile, the max over its rays of 1 loop iterations,
#endif                // 2 samples read from the volume, 3 leaps, 4 iterations in fine mode, 5 leapt steps — instead of cycles
#define RAY_STAT(K) do { if (RAY_STATS == (K)) ++*stat; } while (0)
// cpp/kernels.cpp:674-725; returns hit.xyz, *tw = hit.w
__device__ __forceinline__ float3 raycast_one(const VolView& v, uint32_t px, uint32_t py, const Mat4& view, float nearPlane,
		float farPlane, float step, float largestep, float* tw, unsigned int* stat = nullptr) {
	const float3 origin = f3(view.m[3], view.m[7], view.m[11]);
	const float3 direction = mat_rotate(view, f3((float) px, (float) py, 1.f));
	const float3 invR = f3(1.0f / direction.x, 1.0f / direction.y, 1.0f / direction.z);
	const float3 tbot = (-1.f * invR) * origin;
	const float3 ttop = invR * (f3(v.dx, v.dy, v.dz) - origin);
	const float3 tmin = f3(kminf(ttop.x, tbot.x), kminf(ttop.y, tbot.y), kminf(ttop.z, tbot.z));
	const float3 tmax = f3(kmaxf(ttop.x, tbot.x), kmaxf(ttop.y, tbot.y), kmaxf(ttop.z, tbot.z));
	const float largest_tmin = kmaxf(kmaxf(tmin.x, tmin.y), kmaxf(tmin.x, tmin.z));   // x twice, as in the reference (:693)
	const float smallest_tmax = kminf(kminf(tmax.x, tmax.y), kminf(tmax.x, tmax.z));
	const float tnear = kmaxf(largest_tmin, nearPlane);
	const float tfar = kminf(smallest_tmax, farPlane);
	if (tnear < tfar) {
		float t = tnear;
		float stepsize = largestep;
		float f_t = vol_interp(v, origin + direction * t);
		float f_tt = 0;
		if (f_t > 0) {
			// Samples whose cell lies in a brick with a clear flag are >= 0.82: neither `f_tt < 0` nor `f_tt < 0.8`
			// can fire, so the step is taken without reading the volume.  The value itself is only ever needed as
			// f_t of the zero crossing: it is then evaluated at the remembered t (same expression, same value).
			bool lazy = false;
			float t_lazy = t;
			const bool leap = !v.no_leap;
			for (; t < tfar; t += stepsize) {
				const VolCell c = vol_cell(v, origin + direction * t);
				RAY_STAT(1);
				if (stepsize == step) RAY_STAT(4);
				if (vol_cell_free(v, c)) {
					f_tt = 1.f; lazy = true; t_lazy = t;
					// Fine steps (one voxel each) inside a clear brick: the next samples whose base voxel provably stays in
					// THIS brick are clear as well, so only the reference's `t += stepsize` is replayed for them (the exact
					// sequence of t values is kept; nothing else depends on those samples).  A ray that has grazed a surface
					// walks the rest of its way in fine steps: these are the longest chains of the kernel.
					// The same one level up for COARSE steps: a clear SUPER-brick (64^3 voxels) holds only clear bricks, and a ray
					// crosses it in several coarse steps — most of the march through free space in front of the surface.
					if (leap && c.bx >= 0 && c.by >= 0 && c.bz >= 0 && c.bx < (int) v.sx - 1 && c.by < (int) v.sy - 1 && c.bz < (int) v.sz - 1) {
						const bool fine = stepsize == step;
						bool region_clear = fine;   // the brick itself is known clear
						if (!fine && v.super) region_clear = __ldg(v.super + ((((uint32_t) c.bz >> 6) * v.sny + ((uint32_t) c.by >> 6)) * v.snx + ((uint32_t) c.bx >> 6))) == 0;
						if (region_clear) {
							const int msk = fine ? 7 : 63;
							const float hi = fine ? 7.95f : 63.95f;   // the region's far face, less a margin that dwarfs every rounding involved
							// voxels per unit t along each axis, and the room (in t) to the region's faces (positions are exact to ~1e-3 voxel)
							const float3 ds = f3(direction.x * ((float) v.sx * v.rdx), direction.y * ((float) v.sy * v.rdy), direction.z * ((float) v.sz * v.rdz));
							const float px_ = (float) (c.bx & msk) + c.fx, py_ = (float) (c.by & msk) + c.fy, pz_ = (float) (c.bz & msk) + c.fz;   // position inside the region
							const float rx = (ds.x > 0.f ? hi - px_ : px_ - 0.05f) * rcp_approx(fabsf(ds.x) + 1e-20f);
							const float ry = (ds.y > 0.f ? hi - py_ : py_ - 0.05f) * rcp_approx(fabsf(ds.y) + 1e-20f);
							const float rz = (ds.z > 0.f ? hi - pz_ : pz_ - 0.05f) * rcp_approx(fabsf(ds.z) + 1e-20f);
							const float room = kminf(kminf(rx, ry), rz);
							int n = (room > 0.f) ? (int) kminf(room * rcp_approx(stepsize), 64.f) - 1 : 0;
							if (n > 0) RAY_STAT(3);
							while (n > 0) {
								const float tn = t + stepsize;
								if (!(tn < tfar)) break;   // the loop's own increment and test end the march
								t = tn; t_lazy = tn; --n;
								RAY_STAT(5);
							}
						}
					}
					continue;
				}
				f_tt = vol_interp_cell(v, c);
				RAY_STAT(2);
				if (f_tt < 0) break;
				if (f_tt < 0.8f) stepsize = step;
				f_t = f_tt;
				lazy = false;
			}
			if (f_tt < 0) {
				if (lazy) f_t = vol_interp(v, origin + direction * t_lazy);
				t = t + stepsize * f_tt / (f_t - f_tt);
				*tw = t;
				return origin + direction * t;
			}
		}
	}
	*tw = 0;
	return f3(0, 0, 0);
}

// Slow tiles first.  The kernel ends when its slowest tile does, and a tile's cost hardly changes from one frame to the
// next (the camera moves a little): every launch records the cycles each tile kept its warp, and lists the tiles above
// 1.5 x the previous launch's mean for the NEXT launch, which hands those out before the rest in index order.  Which warp
// computes which tile changes nothing in the maps.  Three rotating slots: read by this launch / appended to for the next /
// zeroed for the one after.  (Handing a slow tile out as four 4x2-pixel quarters to four warps was measured slower: a
// quarter's chain is as long as the tile's — profiles/r2_summary.md.)
struct RaySched {
	unsigned int* cost;                         // [tiles] cycles of the last launch (also KFB_BUF_RAYTILECOST)
	unsigned int* stamp;                        // [tiles] launch number for which the tile is on the slow list
	const unsigned int* slow_cur; unsigned int* slow_next;
	const unsigned int* n_cur; unsigned int* n_next; unsigned int* n_zero;
	const unsigned int* sum_cur; unsigned int* sum_next; unsigned int* sum_zero;   // sum of cost >> 6
	unsigned int launch;                        // this launch's number (from 1)
	int enabled;                                // 0 (KFB_RAY_NO_SCHED=1): index order, costs still recorded
};
struct RaycastParams {
	VolView vol;
	float* vertex; float* normal;   // packed float3[w*h]
	uint32_t w, h;
	uint32_t row0, row1;            // rows handled by this context (multi-GPU: a band of pixels)
	Mat4 view;
	const float* view_dev;          // optional: the view matrix comes from the ICP kernel's tail (DevFrame::view)
	float nearPlane, farPlane, step, largestep;
	unsigned int* tile_next;        // dynamic tile counter of this launch; tile_reset is zeroed for the next one
	unsigned int* tile_reset;
	RaySched sched;
};

#define RC_BX 16
#define RC_BY 16
// CTA = 32x8 pixels as 8 warps of 8x4 pixels: neighbouring rays walk neighbouring voxels (L1 reuse of the taps)
#define RCK_BX 32
#define RCK_BY 4
// Persistent warps pull 8x4-pixel tiles from a counter: rays differ a lot in length (near objects vs the far wall),
// and with a static grid the last wave leaves most SMs idle.
#ifndef KFB_RAY_MINBLOCKS
#define KFB_RAY_MINBLOCKS 9   // 56 registers: the kernel needs its warps (measured in round 1: monotonically faster up to 9 CTAs / SM)
#endif
__global__ void __launch_bounds__(RCK_BX* RCK_BY, KFB_RAY_MINBLOCKS) k_raycast(RaycastParams p) {
	const uint32_t lane = threadIdx.x & 31;
	if (blockIdx.x == 0 && threadIdx.x == 0) *p.tile_reset = 0u;
	__shared__ Mat4 view;   // by value from the host, or from the ICP kernel's tail (shared memory: no registers held for it)
	if (threadIdx.x < 16) view.m[threadIdx.x] = p.view_dev ? __ldg(p.view_dev + threadIdx.x) : p.view.m[threadIdx.x];
	__syncthreads();
	const uint32_t tiles_x = (p.w + 7) / 8, tiles_y = (p.row1 - p.row0 + 3) / 4, tiles = tiles_x * tiles_y;
	if (blockIdx.x == 0 && threadIdx.x == 0) { *p.sched.n_zero = 0u; *p.sched.sum_zero = 0u; }
	const uint32_t n_slow = __ldcg(p.sched.n_cur);
	const uint32_t mean = __ldcg(p.sched.sum_cur) / (tiles ? tiles : 1u);
	const uint32_t slow_thr = (mean && p.sched.enabled) ? mean + (mean >> 1) : 0xffffffffu;
	for (;;) {
		uint32_t t = 0;
		if (lane == 0) t = atomicAdd(p.tile_next, 1u);
		t = __shfl_sync(0xffffffffu, t, 0);
		if (t < n_slow) {
			t = __ldcg(p.sched.slow_cur + t);
			if (t >= tiles) continue;                                  // the band shrank since the list was made
		} else {
			t -= n_slow;
			if (t >= tiles) break;
			if (__ldcg(p.sched.stamp + t) == p.sched.launch) continue;   // handed out from the slow list
		}
		const uint32_t x = (t % tiles_x) * 8 + (lane & 7);
		const uint32_t y = p.row0 + (t / tiles_x) * 4 + (lane >> 3);
		const long long c0 = clock64();
		unsigned int stat = 0;
		if (x < p.w && y < p.row1) {
			const size_t idx = (size_t) x + (size_t) y * p.w;
			float hw;
			const float3 hit = raycast_one(p.vol, x, y, view, p.nearPlane, p.farPlane, p.step, p.largestep, &hw, &stat);
			if (hw > 0.0f) {
				st3(p.vertex, idx, hit);
				const float3 surfNorm = vol_grad(p.vol, hit);
				if (klength(surfNorm) == 0) p.normal[3 * idx] = KFB_INVALID;  // only .x (:745)
				else st3(p.normal, idx, knormalize(surfNorm));
			} else {
				st3(p.vertex, idx, f3(0, 0, 0));
				st3(p.normal, idx, f3(KFB_INVALID, 0, 0));
			}
		}
		__syncwarp();
		if (RAY_STATS) stat = __reduce_max_sync(0xffffffffu, stat);
		if (lane == 0) {   // how long this tile kept its warp: the next launch's schedule (and KFB_BUF_RAYTILECOST)
			const unsigned int cost = (unsigned int) (clock64() - c0), cs = cost >> 6;
			p.sched.cost[t] = RAY_STATS ? stat : cost;
			atomicAdd(p.sched.sum_next, cs);
			if (cs > slow_thr) {
				p.sched.slow_next[atomicAdd(p.sched.n_next, 1u)] = t;   // <= tiles entries: a tile is computed once per launch
				p.sched.stamp[t] = p.sched.launch + 1u;
			}
		}
	}
}

// ------------------------------------------------------------------------------------------
// render kernels (cpp/kernels.cpp:814-913) — visualisation only, outside "computation"
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uchar4 gs2rgb_dev(double h) {  // commons.h:86-147
	double r = 0, g = 0, b = 0;
	const double v = 0.75, m = 0.25, sv = 0.6667;
	h *= 6.0;
	const int sextant = (int) h;
	const double fract = h - sextant, vsf = v * sv * fract, mid1 = m + vsf, mid2 = v - vsf;
	switch (sextant) {
	case 0: r = v; g = mid1; b = m; break;
	case 1: r = mid2; g = v; b = m; break;
	case 2: r = m; g = v; b = mid1; break;
	case 3: r = m; g = mid2; b = v; break;
	case 4: r = mid1; g = m; b = v; break;
	case 5: r = v; g = m; b = mid2; break;
	default: r = 0; g = 0; b = 0; break;
	}
	return make_uchar4((unsigned char) (int) (r * 255), (unsigned char) (int) (g * 255), (unsigned char) (int) (b * 255), 0);
}

__global__ void __launch_bounds__(256) k_render_depth(uchar4* out, const float* depth, uint32_t n, float nearPlane, float farPlane) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float rangeScale = 1 / (farPlane - nearPlane);
	const float d = depth[i];
	if (d < nearPlane) out[i] = make_uchar4(255, 255, 255, 0);
	else if (d > farPlane) out[i] = make_uchar4(0, 0, 0, 0);
	else out[i] = gs2rgb_dev((d - nearPlane) * rangeScale);
}

__global__ void __launch_bounds__(256) k_render_track(uchar4* out, const int8_t* status, uint32_t n) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uchar4 c;
	switch (status[i]) {
	case 1: c = make_uchar4(128, 128, 128, 0); break;
	case -1: c = make_uchar4(0, 0, 0, 0); break;
	case -2: c = make_uchar4(255, 0, 0, 0); break;
	case -3: c = make_uchar4(0, 255, 0, 0); break;
	case -4: c = make_uchar4(0, 0, 255, 0); break;
	case -5: c = make_uchar4(255, 255, 0, 0); break;
	default: c = make_uchar4(255, 128, 128, 0); break;
	}
	out[i] = c;
}

struct RenderVolumeParams {
	VolView vol;
	uchar4* out;
	uint32_t w, h;
	Mat4 view;
	float nearPlane, farPlane, step, largestep;
	float3 light, ambient;
};
__global__ void __launch_bounds__(RC_BX* RC_BY) k_render_volume(RenderVolumeParams p) {
	const uint32_t x = blockIdx.x * RC_BX + threadIdx.x, y = blockIdx.y * RC_BY + threadIdx.y;
	if (x >= p.w || y >= p.h) return;
	float hw;
	const float3 test = raycast_one(p.vol, x, y, p.view, p.nearPlane, p.farPlane, p.step, p.largestep, &hw);
	uchar4 o = make_uchar4(0, 0, 0, 0);
	if (hw > 0) {
		const float3 surfNorm = vol_grad(p.vol, test);
		if (klength(surfNorm) > 0) {
			const float3 diff = knormalize(p.light - test);
			const float dir = kmaxf(kdot(knormalize(surfNorm), diff), 0.f);
			const float3 col = f3(kclampf(dir + p.ambient.x, 0.f, 1.f), kclampf(dir + p.ambient.y, 0.f, 1.f),
					kclampf(dir + p.ambient.z, 0.f, 1.f)) * 255.f;
			o = make_uchar4((unsigned char) (int) col.x, (unsigned char) (int) col.y, (unsigned char) (int) col.z, 0);
		}
	}
	p.out[(size_t) x + (size_t) y * p.w] = o;
}

// ------------------------------------------------------------------------------------------
// z-slab group: this rank's band of the raycast maps (rows [row0, row1): contiguous in memory) goes to every peer's maps —
// the all-gather, as coalesced 16-byte stores over NVLink right behind k_raycast.  (Storing each pixel to the peers from
// inside k_raycast — 1.6 M scattered 4-byte remote writes per rank and frame — measured within 3 % on 8 GPUs, 1329 vs
// 1287 - 1309 fps at 1024^3: the stage is bound by the rays' voxel reads from peer slabs, not by the exchange.  This form
// keeps k_raycast free of peer pointers.)  The band as it stands in this rank's maps is copied, including the components an x-only normal write left
// from the previous frame: the owner of a pixel never changes, so every rank's maps stay bit-identical to the owner's.
// ------------------------------------------------------------------------------------------
struct BandPushParams {
	const uint4* vertex; const uint4* normal;   // this rank's band (first byte of row0)
	uint4* peer_vertex[KFB_MAX_SLABS - 1]; uint4* peer_normal[KFB_MAX_SLABS - 1];
	int n_peer;
	uint32_t n16;                               // 16-byte words per map band
};
__global__ void __launch_bounds__(256) k_band_push(BandPushParams p) {
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n16; i += stride) {
		const uint4 v = __ldcg(p.vertex + i), n = __ldcg(p.normal + i);
		for (int r = 0; r < p.n_peer; ++r) { p.peer_vertex[r][i] = v; p.peer_normal[r][i] = n; }
	}
}

// ------------------------------------------------------------------------------------------
// Stream-ordered barrier between the ranks of a z-slab group, over peer memory (no NCCL): barrier number `n` is complete
// for this rank when every peer has stored n into this rank's arrival slots.  One thread per rank.  Everything this rank
// wrote before (its slab, the flags and map bands it stored into peers) is fenced system-wide before its arrival is
// published; a peer that never arrives is reported after ~4 s instead of hanging the GPU.
// ------------------------------------------------------------------------------------------
struct PeerSync { unsigned int arrive[KFB_MAX_SLABS]; };
struct PeerSyncTable { PeerSync* p[KFB_MAX_SLABS]; };
__global__ void k_peer_barrier(PeerSyncTable all, int rank, int world, unsigned int n, volatile unsigned int* err_host) {
	const int r = threadIdx.x;
	if (r >= world) return;
	__threadfence_system();
	if (r != rank) {
		asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&all.p[r]->arrive[rank]), "r"(n) : "memory");
		const unsigned long long t0 = gtime_ns();
		unsigned int v;
		for (;;) {
			asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(&all.p[rank]->arrive[r]) : "memory");
			if ((int) (v - n) >= 0) break;
			__nanosleep(100);
			if (gtime_ns() - t0 > 4000000000ull) { if (err_host) *err_host = 0x100u + (unsigned int) r; break; }
		}
	}
	__threadfence_system();
}

// dumpVolume helper (cpp/kernels.cpp:1022-1026): gather the tsdf shorts
__global__ void __launch_bounds__(256) k_extract_tsdf(short* out, const short2* vol, size_t n) {
	const size_t stride = (size_t) gridDim.x * blockDim.x;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = vol[i].x;
}

#endif
