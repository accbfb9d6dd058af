// integrateKernel (cpp/kernels.cpp:628-673), second generation: decisions at BRICK granularity.
//
// Round 1's kernels (k_integrate_plan / k_integrate_run in kfb_kernels.cuh, kept behind KFB_FLAG_INTEGRATE_V1 for A/B
// runs) spent 115 of their 158 us deciding, voxel by voxel, what the reference's per-voxel tests would say; only ~45 %
// of the voxels they visited were updated at all, and 85 % of those with sdf == 1 (free space in front of the surface).
// Here the decision is taken once per 8^3 brick, before any voxel is looked at:
//
//   k_depth_mip        (min, max) pyramid of the raw depth image, texels of 8, 16, 32, ... pixels, one launch.
//   k_integrate_plan2  one warp per brick COLUMN (bx, by); each lane classifies bricks (bx, by, bz = lane, lane + 32, ..)
//                         SKIP   no voxel of the brick is updated: behind the camera plane; outside one of the four image
//                                half-spaces (linear in the voxel position, so exact over the box); or max depth over the
//                                brick's pixel footprint + mu < min camera z of the brick
//                         FREE   EVERY voxel is updated with sdf == 1 exactly: footprint strictly inside the image and
//                                min depth over it - max camera z > mu; then e = depth - z > mu for each voxel, and by
//                                monotonicity of the correctly rounded * and / the reference's min(1, e*lambda/mu) is 1
//                         MIXED  anything else: decided voxel by voxel with the exact tests of round 1 (_IN: the footprint
//                                is inside the image and in front of the camera, so those two tests are known)
//                      with margins that cover the accumulated rounding of the reference's repeated additions (the bound
//                      round 1's interval plan uses).  The warp then turns the column into work items: runs of MIXED
//                      bricks (at most 8 layers, per 8x4-voxel half of the column) and runs of FREE bricks.
//   k_integrate_run2   persistent warps, two queues.  MIXED items first (long, compute-bound): the warp's 32 lanes are
//                      8 (x) x 4 (y) columns; the reference's additions are replayed from z = 0 to the run (3 FADD2 per
//                      slice) and the run's voxels decided one by one — approximate-reciprocal pixel with an exactness
//                      guard, e vs mu fast decisions, the full sqrt/division expression only for |e| <= mu.  Then FREE
//                      items (short, memory-bound, they fill the tail): a pure 128-bit streaming update with no per-column
//                      state at all; a voxel that still holds tsdf 32766 (t == 1.0f exactly) stays 32766 for any weight,
//                      so its update is w <- min(w + 1, maxweight) in integer arithmetic; any other value runs the
//                      reference's expression with sdf = 1.
// Every voxel the reference updates gets the reference's exact bits; which path produced them is invisible
// (tests: cull vs no-cull voxel-for-voxel over random poses, oracle parity at 64^3 .. 2048^3).
#ifndef KFB_INTEGRATE2_CUH
#define KFB_INTEGRATE2_CUH

#include "kfb_kernels.cuh"

#ifndef INT_V
#define INT_V 4   // slices per step of the per-voxel path (divides 8: a step never straddles a brick layer)
#endif
static_assert(INT_V == 4, "a step of the per-voxel path is one 8 x 4 x 4 cell");

// ------------------------------------------------------------------------------------------ depth (min, max) mip
#define MIP_MAX_LEVELS 9
#define MIP_LOCAL_LEVELS 4   // levels 0..3 (texels of 8..64 pixels) are built by the CTA that owns the 64x64-pixel region
struct DepthMip {
	float2* lvl[MIP_MAX_LEVELS];          // level l: texels of (8 << l)^2 pixels, row-major [h[l]][w[l]], .x = min, .y = max
	uint32_t w[MIP_MAX_LEVELS], h[MIP_MAX_LEVELS];
	int n;
};

// One CTA per 64x64 pixels: level 0 (8x8 texels) from the image, levels 1..3 from shared memory; the last CTA to finish
// builds the remaining (tiny) levels from level 3.  Depths are >= 0; 0 (= invalid) pulls the minimum to 0.
__global__ void __launch_bounds__(256) k_depth_mip(const float* __restrict__ depth, uint32_t w, uint32_t h, DepthMip mip, unsigned int* ticket) {
	__shared__ float2 s0[8][8], s1[4][4], s2[2][2];
	__shared__ bool is_last;
	const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const uint32_t px0 = blockIdx.x * 64, py0 = blockIdx.y * 64;
	// warp `wid` owns texel row `wid` (8 pixel rows x 64 columns): lane -> 2 adjacent pixels per row
	float mn = 3.0e38f, mx = 0.f;
#pragma unroll
	for (int r = 0; r < 8; ++r) {
		const uint32_t y = py0 + wid * 8 + r, x = px0 + lane * 2;
		if (y < h && x < w) {
			const float a = __ldg(depth + (size_t) y * w + x);
			mn = fminf(mn, a); mx = fmaxf(mx, a);
			if (x + 1 < w) { const float b = __ldg(depth + (size_t) y * w + x + 1); mn = fminf(mn, b); mx = fmaxf(mx, b); }
		}
	}
#pragma unroll
	for (int o = 1; o < 4; o <<= 1) {   // 4 lanes (8 pixels) per texel
		mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
		mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
	}
	if ((lane & 3) == 0) s0[wid][lane >> 2] = make_float2(mn, mx);
	__syncthreads();
	if (tid < 64) {
		const uint32_t tx = blockIdx.x * 8 + (tid & 7), ty = blockIdx.y * 8 + (tid >> 3);
		if (tx < mip.w[0] && ty < mip.h[0]) mip.lvl[0][(size_t) ty * mip.w[0] + tx] = s0[tid >> 3][tid & 7];
	}
	if (tid < 16) {
		const uint32_t x = tid & 3, y = tid >> 2;
		const float2 a = s0[2 * y][2 * x], b = s0[2 * y][2 * x + 1], c = s0[2 * y + 1][2 * x], d = s0[2 * y + 1][2 * x + 1];
		const float2 r = make_float2(fminf(fminf(a.x, b.x), fminf(c.x, d.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y)));
		s1[y][x] = r;
		const uint32_t tx = blockIdx.x * 4 + x, ty = blockIdx.y * 4 + y;
		if (mip.n > 1 && tx < mip.w[1] && ty < mip.h[1]) mip.lvl[1][(size_t) ty * mip.w[1] + tx] = r;
	}
	__syncthreads();
	if (tid < 4) {
		const uint32_t x = tid & 1, y = tid >> 1;
		const float2 a = s1[2 * y][2 * x], b = s1[2 * y][2 * x + 1], c = s1[2 * y + 1][2 * x], d = s1[2 * y + 1][2 * x + 1];
		const float2 r = make_float2(fminf(fminf(a.x, b.x), fminf(c.x, d.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y)));
		s2[y][x] = r;
		const uint32_t tx = blockIdx.x * 2 + x, ty = blockIdx.y * 2 + y;
		if (mip.n > 2 && tx < mip.w[2] && ty < mip.h[2]) mip.lvl[2][(size_t) ty * mip.w[2] + tx] = r;
	}
	__syncthreads();
	if (tid == 0 && mip.n > 3) {
		const float2 a = s2[0][0], b = s2[0][1], c = s2[1][0], d = s2[1][1];
		__stcg(mip.lvl[3] + (size_t) blockIdx.y * mip.w[3] + blockIdx.x,
				make_float2(fminf(fminf(a.x, b.x), fminf(c.x, d.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y))));
	}
	if (mip.n <= MIP_LOCAL_LEVELS) return;
	__threadfence();
	__syncthreads();
	if (tid == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1);
	__syncthreads();
	if (!is_last) return;
	__threadfence();
	for (int l = MIP_LOCAL_LEVELS; l < mip.n; ++l) {
		const uint32_t wl = mip.w[l], hl = mip.h[l], wp = mip.w[l - 1], hp = mip.h[l - 1];
		for (uint32_t i = tid; i < wl * hl; i += 256) {
			const uint32_t tx = i % wl, ty = i / wl;
			float a = 3.0e38f, b = 0.f;
#pragma unroll
			for (int dy = 0; dy < 2; ++dy)
#pragma unroll
				for (int dx = 0; dx < 2; ++dx) {
					const uint32_t qx = 2 * tx + dx, qy = 2 * ty + dy;
					if (qx < wp && qy < hp) {
						const float2 t = __ldcg(mip.lvl[l - 1] + (size_t) qy * wp + qx);
						a = fminf(a, t.x); b = fmaxf(b, t.y);
					}
				}
			__stcg(mip.lvl[l] + i, make_float2(a, b));
		}
		__threadfence();
		__syncthreads();
	}
	if (tid == 0) *ticket = 0u;   // re-arm
}

// (min, max) of the depth over the pixel rectangle [x0, x1] x [y0, y1] (inside the image): the finest level whose texels
// cover the rectangle with at most 4 x 4 of them (conservative: the texels may reach beyond it).  The 16 loads are
// independent (clamped coordinates repeat texels for smaller rectangles); the top level is a single texel, so the search
// always ends with a span <= 3.
__device__ __forceinline__ float2 mip_range(const DepthMip& mip, int x0, int y0, int x1, int y1) {
	int l = 0, s = 3;
	while (l < mip.n - 1 && (((x1 >> s) - (x0 >> s)) > 3 || ((y1 >> s) - (y0 >> s)) > 3)) { ++l; ++s; }
	const float2* t = mip.lvl[l];
	const int wl = (int) mip.w[l], tx0 = x0 >> s, tx1 = x1 >> s, ty0 = y0 >> s, ty1 = y1 >> s;
	float2 v[16];
#pragma unroll
	for (int j = 0; j < 4; ++j)
#pragma unroll
		for (int i = 0; i < 4; ++i) v[4 * j + i] = __ldg(t + (size_t) min(ty0 + j, ty1) * wl + min(tx0 + i, tx1));
	float a = v[0].x, b = v[0].y;
#pragma unroll
	for (int i = 1; i < 16; ++i) { a = fminf(a, v[i].x); b = fmaxf(b, v[i].y); }
	return make_float2(a, b);
}

// ------------------------------------------------------------------------------------------ classification + plan
#define CLS_SKIP 0
#define CLS_FREE 1
#define CLS_MIXED_IN 2     // per voxel; every voxel projects inside the image with pos.z >= 1e-4
#define CLS_MIXED_EDGE 3   // per voxel, all tests

struct Integrate2Params {
	IntegrateParams b;        // volume, slab, depth, matrices, mu (the work-list members of the v1 kernels are unused)
	DepthMip mip;             // (min, max) pyramid of b.depth
	unsigned char* cls;       // [bnz_slab][bny][bnx] brick classes of this launch (diagnostics / tests)
	uint32_t bnx, bny;        // bricks per row / column of the volume
	int maxw_i;               // floor(maxweight) when the integer weight update is exact (1 <= maxweight <= 32767), else -1
	int vec_ok;               // sx % 8 == 0: 128-bit voxel accesses are aligned and every x-brick is complete
	int std_k;                // K's third row is (0, 0, 1, 0): cameraX.z == pos.z bit for bit
	// classification geometry (approximate values: the margins cover them); ca / tz follow invTrack: from the host, or from
	// the ICP kernel's tail (b.dev)
	float vsz[3];             // voxel size (m)
	float ca[9];              // K.rot * invTrack.rot, row-major: camera-space step per metre along the volume's x, y, z
	float tz[3];              // third row of invTrack.rot: pos.z per metre along x, y, z
	uint2* q_mixed; uint2* q_free; uint4* q_replay;   // work items (see k_integrate_plan2)
	unsigned long long* ckpt; // [item][3][32]: the running values (3 packed pairs per column) at the first slice of MIXED item `item`
	unsigned int ckpt_cap;    // items that have a checkpoint slot; the others replay the additions from z = 0 themselves
	unsigned int* ready;      // [pass][by][bx][half]: == seq once this launch's checkpoints of that column half are written
	unsigned int seq;         // launch number (never 0)
	unsigned int* ctr;        // this launch's [0] #mixed items, [1] #free items, [2] next item to claim, [3] #replay jobs
	unsigned int* ctr_next;   // the other slot's four counters, zeroed by the plan pass for the next launch
};

// invTrack and what is derived from it, in shared memory: by value from the host (staged API) or from the ICP kernel's tail
struct IntGeom { Mat4 invTrack; float ca[9], tz[3]; };
// returns false when the device-side gate says "no integrate this frame" (cpp/kernels.cpp:994)
__device__ __forceinline__ bool int_geom_load(IntGeom& g, const Integrate2Params& q, uint32_t tid) {
	const DevFrame* d = q.b.dev;
	const int gate = d ? __ldcg(&d->do_integrate) : 1;   // in flight together with the loads below
	if (tid < 16) g.invTrack.m[tid] = d ? __ldcg(d->invTrack + tid) : q.b.invTrack.m[tid];
	else if (tid < 25) g.ca[tid - 16] = d ? __ldcg(d->ca + (tid - 16)) : q.ca[tid - 16];
	else if (tid < 28) g.tz[tid - 25] = d ? __ldcg(d->tz + (tid - 25)) : q.tz[tid - 25];
	__syncthreads();
	return gate != 0;
}

// Class of the box of voxels [x0, x1] x [y0, y1] x [z0, z1] (inclusive), from linear bounds over the box of their centres
// and its 8 projected corners (see the file header).  FREE is only claimed where the caller may use it (`allow_free`).
// `quick`: only the linear tests (camera plane, image half-spaces, farthest depth) — SKIP or "cannot tell" (MIXED_EDGE).
__device__ __noinline__ int classify_box(const Integrate2Params& q, const IntGeom& g, uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1,
		uint32_t z0, uint32_t z1, float dmax_all, bool allow_free, bool quick = false) {
	const IntegrateParams& p = q.b;
	const float vx = q.vsz[0], vy = q.vsz[1], vz = q.vsz[2];
	// box of the voxel CENTRES (Volume::pos, commons.h:186-189): centre and half extents in metres
	const float3 ctr = f3(((float) (x0 + x1) * 0.5f + 0.5f) * vx, ((float) (y0 + y1) * 0.5f + 0.5f) * vy, ((float) (z0 + z1) * 0.5f + 0.5f) * vz);
	const float hx = (float) (x1 - x0) * 0.5f * vx, hy = (float) (y1 - y0) * 0.5f * vy, hz = (float) (z1 - z0) * 0.5f * vz;
	const float3 pc = mat_point(g.invTrack, ctr);
	const float3 cc = mat_point(p.K, pc);
	// camera-space offsets of the three box axes: columns of (K.rot * T.rot) scaled by the half extents
	const float3 a0 = f3(g.ca[0], g.ca[3], g.ca[6]) * hx, a1 = f3(g.ca[1], g.ca[4], g.ca[7]) * hy, a2 = f3(g.ca[2], g.ca[5], g.ca[8]) * hz;
	const float hpz = fabsf(g.tz[0]) * hx + fabsf(g.tz[1]) * hy + fabsf(g.tz[2]) * hz;   // half extent of pos.z
	const float hcx = fabsf(a0.x) + fabsf(a1.x) + fabsf(a2.x), hcy = fabsf(a0.y) + fabsf(a1.y) + fabsf(a2.y), hcz = fabsf(a0.z) + fabsf(a1.z) + fabsf(a2.z);
	// bound on the drift of the reference's accumulated values from the exact line: (N + 64) * 2^-22 times the largest
	// magnitude along the column (k_integrate_plan uses the same bound); the column spans dz in z
	const float eps = ((float) p.sz + 64.f) * 2.3841858e-7f;
	const float3 kz = f3(g.ca[2], g.ca[5], g.ca[8]) * p.dz;   // change of cam over the whole column
	const float cxm = fabsf(cc.x) + hcx, cym = fabsf(cc.y) + hcy, czm = fabsf(cc.z) + hcz;   // magnitudes inside the box
	const float e_pz = eps * (fabsf(pc.z) + hpz + fabsf(g.tz[2]) * p.dz);
	const float e_cx = eps * (cxm + fabsf(kz.x)), e_cy = eps * (cym + fabsf(kz.y)), e_cz = eps * (czm + fabsf(kz.z));
	const float pz_min = pc.z - hpz - e_pz, pz_max = pc.z + hpz + e_pz;
	const float cz_min = cc.z - hcz - e_cz, cz_max = cc.z + hcz + e_cz;
	if (!(pz_max == pz_max && cz_min == cz_min && cxm == cxm && cym == cym)) return CLS_MIXED_EDGE;   // NaN matrices: no shortcut
	if (pz_max < 0.0000999f) return CLS_SKIP;   // every voxel fails `pos.z < 0.0001f`
	const float dwm1 = (float) (p.dw - 1), dhm1 = (float) (p.dh - 1);
	// image half-spaces: for a voxel with cameraX.z > 0,  pixel.x < 0  <=>  Cx + 0.5 Cz < 0,  pixel.x > w-1  <=>  (w-1.5) Cz - Cx < 0
	// (same in y).  Both sides are linear in the voxel position: their maximum over the box is centre + |half extents|.
	// Voxels that survive the pos.z test have Cz == pos.z >= 1e-4 > 0 when K is a camera matrix; otherwise need cz_min > 0.
	if (q.std_k || cz_min > 0.f) {
		const float wq = dwm1 - 0.5f, hq = dhm1 - 0.5f;
		const float gl = cc.x + 0.5f * cc.z + (fabsf(a0.x + 0.5f * a0.z) + fabsf(a1.x + 0.5f * a1.z) + fabsf(a2.x + 0.5f * a2.z));
		const float gr = wq * cc.z - cc.x + (fabsf(wq * a0.z - a0.x) + fabsf(wq * a1.z - a1.x) + fabsf(wq * a2.z - a2.x));
		const float gt = cc.y + 0.5f * cc.z + (fabsf(a0.y + 0.5f * a0.z) + fabsf(a1.y + 0.5f * a1.z) + fabsf(a2.y + 0.5f * a2.z));
		const float gb = hq * cc.z - cc.y + (fabsf(hq * a0.z - a0.y) + fabsf(hq * a1.z - a1.y) + fabsf(hq * a2.z - a2.y));
		// margins: the drifts, the division's rounding (relative 2^-24 of a quotient up to w), and this evaluation's own
		const float ml = e_cx + e_cz + 1e-5f * (cxm + czm), mr = e_cx + (wq + 2.f) * e_cz + 2e-6f * (cxm + (wq + 2.f) * czm);
		const float mt = e_cy + e_cz + 1e-5f * (cym + czm), mb = e_cy + (hq + 2.f) * e_cz + 2e-6f * (cym + (hq + 2.f) * czm);
		if (gl < -ml || gr < -mr || gt < -mt || gb < -mb) return CLS_SKIP;
	}
	// beyond the farthest depth of the whole image: e < -mu (or depth == 0) for every pixel
	if (dmax_all + p.mu + (p.mu * 1e-5f + 1e-6f * (dmax_all + cz_max)) < cz_min) return CLS_SKIP;
	if (quick || !(cz_min >= 0.05f)) return CLS_MIXED_EDGE;                          // too close to the camera plane for a footprint
	float umin = 3.0e38f, umax = -3.0e38f, vmin = 3.0e38f, vmax = -3.0e38f;
#pragma unroll
	for (int c = 0; c < 8; ++c) {
		const float sx_ = (c & 1) ? 1.f : -1.f, sy_ = (c & 2) ? 1.f : -1.f, sz_ = (c & 4) ? 1.f : -1.f;
		const float X = cc.x + sx_ * a0.x + sy_ * a1.x + sz_ * a2.x;
		const float Y = cc.y + sx_ * a0.y + sy_ * a1.y + sz_ * a2.y;
		const float Z = cc.z + sx_ * a0.z + sy_ * a1.z + sz_ * a2.z;
		const float r = rcp_approx(Z);   // Z >= 0.05
		const float u = X * r, v = Y * r;
		umin = fminf(umin, u); umax = fmaxf(umax, u); vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
	}
	const float ua = fmaxf(fabsf(umin), fabsf(umax)), va = fmaxf(fabsf(vmin), fabsf(vmax));
	const float rz = rcp_approx(cz_min) * 1.000001f;
	// |computed pixel - ideal pixel|: perturbation of the quotient by the drifts, the divisions' rounding, and ours
	const float mu_ = (e_cx + (ua + 1.f) * e_cz) * rz + 0.02f + ua * 1e-5f;
	const float mv_ = (e_cy + (va + 1.f) * e_cz) * rz + 0.02f + va * 1e-5f;
	const float pxlo = umin + 0.5f - mu_, pxhi = umax + 0.5f + mu_, pylo = vmin + 0.5f - mv_, pyhi = vmax + 0.5f + mv_;
	if (pxhi < 0.f || pxlo > dwm1 || pyhi < 0.f || pylo > dhm1) return CLS_SKIP;       // every voxel fails the pixel bounds test
	if (!(fabsf(pxlo) < 1e6f && fabsf(pxhi) < 1e6f && fabsf(pylo) < 1e6f && fabsf(pyhi) < 1e6f)) return CLS_MIXED_EDGE;
	const int X0 = max(0, (int) floorf(pxlo)), X1 = min((int) p.dw - 1, (int) floorf(pxhi));
	const int Y0 = max(0, (int) floorf(pylo)), Y1 = min((int) p.dh - 1, (int) floorf(pyhi));
	if (X0 > X1 || Y0 > Y1) return CLS_SKIP;
	const float2 dr = mip_range(q.mip, X0, Y0, X1, Y1);
	// e = fl(depth - z): the subtraction's own rounding is below 2^-23 of the larger operand
	const float slack = p.mu * 1e-5f + 1e-6f * (dr.y + cz_max);
	if (dr.y + p.mu + slack < cz_min) return CLS_SKIP;                                   // e < -mu everywhere (or depth == 0)
	const bool inside = pxlo >= 0.f && pxhi <= dwm1 && pylo >= 0.f && pyhi <= dhm1 && pz_min >= 0.000101f;
	if (inside && allow_free && dr.x - cz_max > p.mu + slack) return CLS_FREE;           // e > mu everywhere, depth > 0
	return inside ? CLS_MIXED_IN : CLS_MIXED_EDGE;
}

// Work items.
//   MIXED  { bx | by << 12 | half << 24,  z_first | n_slices << 16 | cells << 24 }: at most INT_MIXED_CAP (<= 2) layers of
//          per-voxel bricks in one 8x4-voxel half of a brick column.  `cells`: 2 bits per step of 4 slices — the class of
//          that 8 x 4 x 4 cell of voxels, classified like a brick but on a quarter of its volume: inside a brick that is
//          MIXED as a whole, about 40 % of the cells are still all-SKIP or all-FREE.
//   FREE   { bx_first | by << 12 | log2(len) << 24,  bz }: 2^k (<= 16) FREE bricks in a row ALONG X (k_integrate_free_runs).
//   REPLAY { bx | by << 12 | half << 24, first MIXED item, #items, 0 }: one per column half that has MIXED items (they
//          are contiguous and in z order): the reference's additions (cpp/kernels.cpp:646-647) replayed ONCE from z = 0,
//          the running values stored as a CHECKPOINT at the first slice of every item.
// Items are small on purpose: the run kernel's time is the per-warp chain of its longest item, and checkpoints are what
// makes a small MIXED item possible.
#ifndef INT_MIXED_CAP
#define INT_MIXED_CAP 2
#endif
static_assert(INT_MIXED_CAP == 1 || INT_MIXED_CAP == 2, "the cell classes of an item take 8 bits: at most 2 layers");
#define PLAN_GROUPS 8   // 32-layer groups per pass (8 x 32 x 8 = 2048 slices)

// starts of the items inside one 32-layer group: every `cap`-th layer of each run of set bits, counted from the run's start
__device__ __forceinline__ unsigned int item_starts(unsigned int m, uint32_t lane, unsigned int cap) {
	const unsigned int below = ~m & ((1u << lane) - 1u);                                  // clear layers below this lane
	const unsigned int run0 = below ? 32u - (unsigned int) __clz((int) below) : 0u;      // first layer of this lane's run
	return __ballot_sync(0xffffffffu, ((m >> lane) & 1u) && ((lane - run0) % cap == 0u));
}
__device__ __forceinline__ unsigned int item_len(unsigned int m, uint32_t lane, unsigned int cap) {   // layers of the item starting at `lane`
	unsigned int len = (unsigned int) __ffs((int) ~(m >> lane)) - 1u;   // consecutive set layers from here (none clear: wraps, capped below)
	if (len > cap) len = cap;
	if (len > 32u - lane) len = 32u - lane;
	return len;
}

#ifndef KFB_PLAN_MINBLOCKS
#define KFB_PLAN_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(256, KFB_PLAN_MINBLOCKS) k_integrate_plan2(const __grid_constant__ Integrate2Params q) {
	__shared__ unsigned int s_mm[8][PLAN_GROUPS], s_ms[8][PLAN_GROUPS];
	__shared__ IntGeom geom;
	const IntegrateParams& p = q.b;
	const uint32_t lane = threadIdx.x;
	const uint32_t bx = blockIdx.x, by = blockIdx.y * blockDim.y + threadIdx.y;   // warp-uniform
	if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 4 && threadIdx.y == 0) q.ctr_next[threadIdx.x] = 0u;   // re-arm the other slot
	if (!int_geom_load(geom, q, threadIdx.y * 32 + threadIdx.x)) return;   // gate closed: no items, the run pass finds none
	if (by >= q.bny) return;
	const uint32_t bz0 = p.z_begin >> 3, bz1 = (p.z_end + 7) >> 3;
	const bool fast = p.cull && p.mu > 0.f && p.dw <= 2040 && p.dh <= 2040;
	const float dmax_all = __ldg(q.mip.lvl[q.mip.n - 1]).y;   // the top level is one texel: max over the image
	const uint32_t halves = (by * 8 + 4 < p.sy) ? 2 : 1;
	const unsigned int lt = (1u << lane) - 1u;
	const uint32_t x0 = bx * 8, x1 = min(x0 + 7, p.sx - 1), y0 = by * 8, y1 = min(y0 + 7, p.sy - 1);
	// the whole column first (most columns lie outside the view frustum: one test instead of one per brick)
	if (fast && classify_box(q, geom, x0, x1, y0, y1, p.z_begin, p.z_end - 1, dmax_all, false, true) == CLS_SKIP) {
		for (uint32_t bz = bz0 + lane; bz < bz1; bz += 32) q.cls[((size_t) (bz - bz0) * q.bny + by) * q.bnx + bx] = CLS_SKIP;
		return;
	}
	for (uint32_t pass0 = bz0; pass0 < bz1; pass0 += 32 * PLAN_GROUPS) {
		unsigned int* mm = s_mm[threadIdx.y];   // per group: MIXED layers / first layers of the MIXED items (rolled loops: small code)
		unsigned int* ms = s_ms[threadIdx.y];
		unsigned int n_half = 0;
		// pass 1: classes, and the number of items of the column
#pragma unroll 1
		for (int g = 0; g < PLAN_GROUPS; ++g) {
			if (lane == 0) { mm[g] = 0u; ms[g] = 0u; }
			if (pass0 + 32 * g >= bz1) continue;   // warp-uniform
			const uint32_t bz = pass0 + 32 * g + lane;
			int c = CLS_SKIP;
			if (bz < bz1) {
				c = fast ? classify_box(q, geom, x0, x1, y0, y1, bz * 8, min(bz * 8 + 7, p.sz - 1), dmax_all, q.vec_ok != 0) : CLS_MIXED_EDGE;
				q.cls[((size_t) (bz - bz0) * q.bny + by) * q.bnx + bx] = (unsigned char) c;
			}
			const unsigned int m_mixed = __ballot_sync(0xffffffffu, c >= CLS_MIXED_IN), m_starts = item_starts(m_mixed, lane, INT_MIXED_CAP);
			if (lane == 0) { mm[g] = m_mixed; ms[g] = m_starts; }
			n_half += __popc(m_starts);
		}
		if (n_half == 0) continue;
		// One round trip to the queue counters per column (they are hot: 4096 warps): the column's MIXED items take one
		// contiguous range, [half 0's items in z order][half 1's items in z order], plus one REPLAY job per half.  (The FREE
		// items are cut along x from the class array by k_integrate_free_runs.)
		unsigned int base_m = 0, base_r = 0;
		if (lane == 0) { base_m = atomicAdd(q.ctr + 0, n_half * halves); base_r = atomicAdd(q.ctr + 3, halves); }
		base_m = __shfl_sync(0xffffffffu, base_m, 0);
		base_r = __shfl_sync(0xffffffffu, base_r, 0);
		__syncwarp();
		if (lane < halves) q.q_replay[base_r + lane] = make_uint4(bx | (by << 12) | (lane << 24), base_m + lane * n_half, n_half, 0u);
		unsigned int at_m = base_m;
		__syncwarp();
		// pass 2: the 8 x 4 x 4 cells of the MIXED bricks, and the items
#pragma unroll 1
		for (int g = 0; g < PLAN_GROUPS; ++g) {
			const unsigned int m_mixed = mm[g], m_starts = ms[g];
			if (m_mixed == 0u) continue;           // warp-uniform
			const uint32_t bz = pass0 + 32 * g + lane;
			// One cell per lane and round: task t = 4 * (rank of the MIXED brick in this group) + (2 h + s), so a column's
			// cells are classified side by side instead of four after another by the lane that owns the brick.
			unsigned int cells = 0;                // bits 2 (2 h + s): class of cell (half h, slices 4 s .. 4 s + 3) of this lane's brick
			const bool mine = (m_mixed >> lane) & 1u;
			const unsigned int n_tasks = 4u * (unsigned int) __popc(m_mixed), my_task0 = 4u * (unsigned int) __popc(m_mixed & lt);
			for (unsigned int r = 0; r * 32u < n_tasks; ++r) {
				const unsigned int t = r * 32u + lane;
				int c = CLS_MIXED_EDGE;            // !fast: every cell per voxel, all tests
				if (t < n_tasks && fast) {
					const uint32_t tbz = pass0 + 32 * g + __fns(m_mixed, 0, (int) (t >> 2) + 1), hs = t & 3u;
					const uint32_t cy0 = y0 + 4 * (hs >> 1), cz0 = tbz * 8 + 4 * (hs & 1);
					c = CLS_SKIP;
					if (cy0 < p.sy && cz0 < p.z_end) c = classify_box(q, geom, x0, x1, cy0, min(cy0 + 3, p.sy - 1), cz0, min(cz0 + 3, p.sz - 1), dmax_all, true);
				}
#pragma unroll
				for (unsigned int hs = 0; hs < 4; ++hs) {
					const unsigned int src = my_task0 + hs;
					const int v = __shfl_sync(0xffffffffu, c, (int) (src & 31u));
					if (mine && (src >> 5) == r) cells |= (unsigned int) v << (2 * hs);
				}
			}
			const unsigned int next_cells = __shfl_down_sync(0xffffffffu, cells, 1);
			if ((m_starts >> lane) & 1u) {
				const unsigned int len = item_len(m_mixed, lane, INT_MIXED_CAP);
				const uint32_t za = bz * 8, zb = min(p.z_end, (bz + len) * 8);
				const unsigned int at = at_m + __popc(m_starts & lt);
				for (uint32_t h = 0; h < halves; ++h) {
					unsigned int cc = (cells >> (4 * h)) & 0xfu;
					if (len == 2) cc |= ((next_cells >> (4 * h)) & 0xfu) << 4;
					q.q_mixed[at + h * n_half] = make_uint2(bx | (by << 12) | (h << 24), za | ((zb - za) << 16) | (cc << 24));
				}
			}
			at_m += __popc(m_starts);
		}
		__syncwarp();
	}
}

// FREE items are cut ALONG X: a row of `len` consecutive FREE bricks is len * 32 contiguous bytes per voxel row, and DRAM
// wants long contiguous bursts — single bricks (32 bytes per row, rows 2-8 KB apart) streamed at only ~25 % of the HBM
// rate.  One warp per (by, bz) row of bricks, lanes over bx; runs are decomposed into aligned power-of-two pieces of at
// most 16 bricks (index arithmetic by shifts in the run kernel; 32 KB per item keeps the tail short — cutting the rows inside
// the run kernel instead, one whole brick row per claim, was measured: 81 -> 132 us, the row jobs are too long).
//   FREE  { bx_first | by << 12 | log2(len) << 24,  bz }
__global__ void __launch_bounds__(256) k_integrate_free_runs(const __grid_constant__ Integrate2Params q) {
	const IntegrateParams& p = q.b;
	if (p.dev && __ldcg(&p.dev->do_integrate) == 0) return;
	const uint32_t lane = threadIdx.x, by = blockIdx.x, bz_rel = blockIdx.y * blockDim.y + threadIdx.y;   // warp-uniform
	const uint32_t bz0 = p.z_begin >> 3, bz1 = (p.z_end + 7) >> 3;
	if (bz0 + bz_rel >= bz1) return;
	const unsigned char* row = q.cls + ((size_t) bz_rel * q.bny + by) * q.bnx;
	for (uint32_t base = 0; base < q.bnx; base += 32) {
		const uint32_t bx = base + lane;
		const unsigned int m = __ballot_sync(0xffffffffu, bx < q.bnx && __ldg(row + bx) == CLS_FREE);
		if (m == 0u) continue;
		// lane l starts a piece of 2^k bricks when l is a multiple of 2^k, bits l .. l + 2^k - 1 are set, and l is not inside
		// a larger aligned piece: the largest k that fits at the aligned position wins
		int k_here = -1;   // log2 of the piece starting at this lane
		bool covered = false;
#pragma unroll
		for (int k = 4; k >= 0; --k) {
			const unsigned int len = 1u << k, al = lane & ~(len - 1u);
			const unsigned int need = ((len == 32u) ? 0xffffffffu : ((1u << len) - 1u)) << al;
			const bool full = (m & need) == need;            // the aligned block of 2^k bricks that contains this lane is all FREE
			if (full && !covered) { covered = true; if (al == lane) k_here = k; }
		}
		const unsigned int starts = __ballot_sync(0xffffffffu, k_here >= 0);
		unsigned int slot = 0;
		if (lane == 0) slot = atomicAdd(q.ctr + 1, (unsigned int) __popc(starts));
		slot = __shfl_sync(0xffffffffu, slot, 0);
		if (k_here >= 0) q.q_free[slot + __popc(starts & ((1u << lane) - 1u))] = make_uint2(bx | (by << 12) | ((uint32_t) k_here << 24), bz0 + bz_rel);
	}
}

// the running average with sdf (cpp/kernels.cpp:662-669; commons.h:160-163, 182-185).  `rcp` = shared table of RN(1/d),
// d = 0..127: for d in 1..127 and ANY normal fp32 numerator, q = n*y, q + fma(-d, q, n)*y (y = RN(1/d)) is the correctly
// rounded n / d — verified exhaustively over all 2^23 significands per divisor (tools/div_small_exhaustive.c).
__device__ __forceinline__ short2 tsdf_update(short2 v, float sdf, float maxweight, const float* rcp) {
	float tsdf = (float) v.x * 0.00003051944088f, wgt = (float) v.y;
	const float num = wgt * tsdf + sdf, den = wgt + 1;
	float quo;
	if ((unsigned int) v.y < 127u) {
		const float y = rcp[v.y + 1], q0 = num * y;
		quo = __fmaf_rn(__fmaf_rn(-den, q0, num), y, q0);
	} else quo = num / den;
	tsdf = kclampf(quo, -1.f, 1.f);
	wgt = kminf(wgt + 1, maxweight);
	return make_short2((short) (int) (tsdf * 32766.0f), (short) (int) wgt);   // truncation
}
// the same for sdf == 1 on a packed voxel (x = low half, weight = high half).  tsdf 32766 reads back as exactly 1.0f, so
// (w * 1 + 1) / (w + 1) == 1 and the stored value stays 32766 for every weight >= 0: only the weight moves.
__device__ __noinline__ uint32_t tsdf_update_free_slow(uint32_t v, float maxweight, const float* rcp) {
	const short2 r = tsdf_update(make_short2((short) (v & 0xffffu), (short) (v >> 16)), 1.f, maxweight, rcp);
	return ((uint32_t) (uint16_t) r.y << 16) | (uint32_t) (uint16_t) r.x;
}
__device__ __forceinline__ uint32_t tsdf_update_free(uint32_t v, float maxweight, int maxw_i, const float* rcp) {
	if (maxw_i > 0 && (v & 0x8000ffffu) == 32766u) return ((uint32_t) min((int) (v >> 16) + 1, maxw_i) << 16) | 32766u;
	return tsdf_update_free_slow(v, maxweight, rcp);
}

// x <- fl(x + d) applied n times, bit for bit, in O(binades crossed): inside one binade and sign every correctly rounded
// addition of the same d moves x by the same whole number of ulps (ties included from the second step in the binade on), so
// after two real steps there the rest is one integer multiply-add on the bit pattern; the step that leaves the binade is taken
// for real (and a sum that falls just below the binade's floor rounds on the finer grid: stay strictly above it).  Checked
// against the plain loop on ~6 M random, KinectFusion-like and tie-prone cases: tools/jump_ahead_check.c.  This is what makes
// the replay of the reference's additions independent of how far up the column a slab starts.
__device__ __forceinline__ float jump_ahead(float x, float d, int n) {
	int k = n;
	while (k > 0) {
		const float a = __fadd_rn(x, d);
		if (--k == 0) { x = a; break; }
		float b = __fadd_rn(a, d);
		--k;
		const uint32_t ux = __float_as_uint(x), ua = __float_as_uint(a), ub = __float_as_uint(b);
		const uint32_t ea = ua >> 23;   // sign + exponent
		if (k > 0 && (ux >> 23) == ea && (ub >> 23) == ea && (ea & 0xffu) != 0u && (ea & 0xffu) != 0xffu) {
			const int S = (int) (ub - ua);   // ulps per step (magnitude space: same sign, same exponent)
			if (S == 0) return b;            // stagnation: every further step returns b
			const uint32_t M = ub & 0x7fffffffu, lo = (ea & 0xffu) << 23, hi = lo + 0x7fffffu;
			// steps that stay inside the binade (strictly above its floor when moving down): a float quotient rounded DOWN by a
			// safe factor instead of an integer division — an underestimate only costs another round of the loop
			const uint32_t room = S > 0 ? hi - M : ((M == lo) ? 0u : M - lo - 1u);
			uint32_t m = (uint32_t) (__fdividef((float) room, (float) abs(S)) * 0.9999f);
			if (m > (uint32_t) k) m = (uint32_t) k;
			b = __uint_as_float(ub + (uint32_t) ((int) m * S));
			k -= (int) m;
		}
		x = b;
	}
	return x;
}
// the six running values of a column, `n` slices further.  MEASURED (profiles/r2_summary.md): the exact jump is SLOWER than the
// plain additions at every distance this code meets — 512^3 (470 slices on average): +25 us per frame; 1024^3 / 2048^3 far
// slabs on 8 GPUs (920 / 1870 slices): 292 -> 505 us and 1714 -> 2707 us for the slab that holds the surface — each value
// takes ~11 data-dependent rounds of ~60 instructions and the lanes of a warp diverge, while the replay is 3 packed FADD2
// per slice with no divergence.  It stays in the tree, switched off, as the record of the experiment (-DINT_JUMP_MIN=640).
#ifndef INT_JUMP_MIN
#define INT_JUMP_MIN 0x7fffffff
#endif
__device__ __forceinline__ void replay_advance(F2& A, F2& B, F2& C, F2 dA, F2 dB, F2 dC, int n, bool cz_is_pz) {
	if (n < INT_JUMP_MIN) {
#pragma unroll 8
		for (int i = 0; i < n; ++i) { A = f2_add(A, dA); B = f2_add(B, dB); C = f2_add(C, dC); }
		return;
	}
	const float px = jump_ahead(f2_lo(A), f2_lo(dA), n), py = jump_ahead(f2_hi(A), f2_hi(dA), n);
	const float pz = jump_ahead(f2_lo(B), f2_lo(dB), n), cx = jump_ahead(f2_hi(B), f2_hi(dB), n);
	const float cy = jump_ahead(f2_lo(C), f2_lo(dC), n);
	// K's third row (0, 0, 1, 0): cameraX.z and pos.z start equal and receive the same increments — one value
	const float cz = cz_is_pz ? pz : jump_ahead(f2_hi(C), f2_hi(dC), n);
	A = f2_make(px, py); B = f2_make(pz, cx); C = f2_make(cy, cz);
}

// flag of the REPLAY job that covers slice `za` of column half (bx, by, half): one job per plan pass of 32 * PLAN_GROUPS layers
__device__ __forceinline__ size_t ready_index(const Integrate2Params& q, uint32_t bx, uint32_t by, uint32_t half, uint32_t za) {
	const uint32_t pass = ((za >> 3) - (q.b.z_begin >> 3)) / (32u * PLAN_GROUPS);
	return (((size_t) pass * q.bny + by) * q.bnx + bx) * 2 + half;
}

// REPLAY job: the lanes are the 8 x 4 voxel columns of one column half; one sweep from z = 0 past every item start
__device__ __forceinline__ void integrate_replay_job(const Integrate2Params& q, const IntGeom& geom, uint4 job, uint32_t lane) {
	const IntegrateParams& p = q.b;
	const uint32_t bx = job.x & 0xfffu, by = (job.x >> 12) & 0xfffu, half = (job.x >> 24) & 1u;
	const uint32_t x = bx * 8 + (lane & 7), y = by * 8 + half * 4 + (lane >> 3);
	const IntColumn c = int_column(p, geom.invTrack, x, y);
	// running values, packed in pairs: (pos.x, pos.y) (pos.z, cam.x) (cam.y, cam.z)
	F2 A = f2_make(c.pos0.x, c.pos0.y), B = f2_make(c.pos0.z, c.cam0.x), C = f2_make(c.cam0.y, c.cam0.z);
	const F2 dA = f2_make(c.delta.x, c.delta.y), dB = f2_make(c.delta.z, c.cameraDelta.x), dC = f2_make(c.cameraDelta.y, c.cameraDelta.z);
	int z = 0;
	// cameraX.z == pos.z bit for bit when K's third row is (0, 0, 1, 0) AND the two start equal (they do: cam0.z = 0*x + 0*y + 1*z + 0)
	const bool cz_is_pz = q.std_k && f2_hi(C) == f2_lo(B) && f2_hi(dC) == f2_lo(dB);
	for (unsigned int k = 0; k < job.z; ++k) {
		const unsigned int item = job.y + k;
		const int za = (int) (__ldcg(&q.q_mixed[item].y) & 0xffffu);
		replay_advance(A, B, C, dA, dB, dC, za - z, cz_is_pz);
		z = za;
		if (item < q.ckpt_cap) {
			unsigned long long* o = q.ckpt + (size_t) item * 96 + lane;
			__stcg(o, A.v); __stcg(o + 32, B.v); __stcg(o + 64, C.v);
		}
	}
	__threadfence();
	__syncwarp();
	if (lane == 0) st_release_u32(q.ready + ready_index(q, bx, by, half, __ldcg(&q.q_mixed[job.y].y) & 0xffffu), q.seq);
}

__global__ void __launch_bounds__(256, KFB_INT_MINBLOCKS) k_integrate_run2(Integrate2Params q) {
	__shared__ float rcp[128];
	__shared__ IntGeom geom;
	const IntegrateParams& p = q.b;
	if (threadIdx.x < 128) rcp[threadIdx.x] = 1.0f / (float) threadIdx.x;   // [0] = inf, never used
	if (!int_geom_load(geom, q, threadIdx.x)) return;
	const uint32_t lane = threadIdx.x & 31;
	const float dwm1 = (float) (p.dw - 1), dhm1 = (float) (p.dh - 1);
	const bool fast = p.cull && p.mu > 0.f && p.dw <= 2040 && p.dh <= 2040;
	const float tol = (float) max(p.dw, p.dh) * 4.0e-7f + 1.0e-5f;
	const float mu = p.mu;
	const float* __restrict__ depth = p.depth;
	const uint32_t dw = p.dw;
	const size_t plane = (size_t) p.sx * p.sy;
	const unsigned int n_mixed = __ldcg(q.ctr + 0), n_free = __ldcg(q.ctr + 1), n_replay = __ldcg(q.ctr + 3);
	const unsigned int n_items = n_replay + n_mixed + n_free;
	// claim order: all REPLAY jobs (long serial chains: started first), then MIXED and FREE items interleaved k : 1 — the FREE
	// items are pure memory streaming, the MIXED ones latency / issue bound: together on an SM they overlap — then what is
	// left of either kind.  (A MIXED item claimed while its REPLAY job still runs waits on the job's flag.)
#ifndef INT_INTERLEAVE
#define INT_INTERLEAVE 1
#endif
	const unsigned int k_mix = n_free ? max(1u, n_mixed / n_free) : 1u;
	const unsigned int n_groups = min(n_free, n_mixed / k_mix);
	unsigned int updated = 0;

	for (;;) {
		unsigned int it = 0;
		if (lane == 0) it = atomicAdd(q.ctr + 2, 1u);   // (claiming one item ahead was measured: 83 -> 92 us, the reserve item delays the tail)
		it = __shfl_sync(0xffffffffu, it, 0);
		if (it >= n_items) break;
		if (it < n_replay) {
			integrate_replay_job(q, geom, __ldcg(q.q_replay + it), lane);
			continue;
		}
		it -= n_replay;
		bool is_free;
		unsigned int idx;
#if INT_INTERLEAVE
		if (it < n_groups * (k_mix + 1)) {
			const unsigned int grp = it / (k_mix + 1), pos = it - grp * (k_mix + 1);
			is_free = pos == k_mix;
			idx = is_free ? grp : grp * k_mix + pos;
		} else {
			const unsigned int rest = it - n_groups * (k_mix + 1), mixed_left = n_mixed - n_groups * k_mix;
			is_free = rest >= mixed_left;
			idx = is_free ? n_groups + (rest - mixed_left) : n_groups * k_mix + rest;
		}
#else
		is_free = it < n_free;
		idx = is_free ? it : it - n_free;
#endif
		if (is_free) {
			// ---------------- FREE item: 2^k bricks along x in brick row (by, bz): 64 voxel rows of 2^k * 32 contiguous bytes
			const uint2 item = __ldcg(q.q_free + idx);
			const uint32_t bx0 = item.x & 0xfffu, by = (item.x >> 12) & 0xfffu, lg = item.x >> 24, bz = item.y;
			const uint32_t lg_row = lg + 1;                         // log2 of the uint4 (4 voxels) per voxel row
			const uint32_t total = 64u << lg_row;                   // uint4 of the item
			uint4* base = reinterpret_cast<uint4*>(p.vol + (size_t) bx0 * 8 + (size_t) (by * 8) * p.sx + (size_t) (bz * 8 - p.z_begin) * plane);
			const uint32_t sx4 = p.sx >> 2;                         // uint4 per volume row
			const size_t plane4 = plane >> 2;
#pragma unroll 1
			for (uint32_t b0 = 0; b0 < total; b0 += 128) {          // 4 x 32 uint4 in flight per warp
				uint4 v[4];
				uint4* ptr[4];
				bool ok[4];
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					const uint32_t e = b0 + 32 * k + lane;
					const uint32_t r = e >> lg_row, cidx = e & ((1u << lg_row) - 1u);   // voxel row (y + 8 z) and position in it
					const uint32_t yy = by * 8 + (r & 7u), zz = bz * 8 + (r >> 3);
					ok[k] = e < total && yy < p.sy && zz < p.z_end;
					ptr[k] = base + (size_t) (r >> 3) * plane4 + (size_t) (r & 7u) * sx4 + cidx;
					if (ok[k]) v[k] = __ldcs(ptr[k]);
				}
#pragma unroll
				for (int k = 0; k < 4; ++k)
					if (ok[k]) {
						v[k].x = tsdf_update_free(v[k].x, p.maxweight, q.maxw_i, rcp); v[k].y = tsdf_update_free(v[k].y, p.maxweight, q.maxw_i, rcp);
						v[k].z = tsdf_update_free(v[k].z, p.maxweight, q.maxw_i, rcp); v[k].w = tsdf_update_free(v[k].w, p.maxweight, q.maxw_i, rcp);
						__stcs(ptr[k], v[k]);
						updated += 4;
					}
			}
			continue;
		}
		// ---------------- MIXED item: 8 x 4 columns, slices [za, zb), one class per step of 4 slices
		const uint2 item = __ldcg(q.q_mixed + idx);
		const unsigned int cells = item.y >> 24;
		if (cells == 0u) continue;   // every cell of this half is SKIP
		const uint32_t bx = item.x & 0xfffu, by = (item.x >> 12) & 0xfffu, half = (item.x >> 24) & 1u;
		const uint32_t x = bx * 8 + (lane & 7), y = by * 8 + half * 4 + (lane >> 3);
		const int za = (int) (item.y & 0xffffu), zb = za + (int) ((item.y >> 16) & 0xffu);
		const bool valid = x < p.sx && y < p.sy;
		// The tile's voxels are pulled into L2 ahead of their use (one 32-byte sector = 8 voxels in x per lane and
		// instruction: lane -> row lane >> 3, slice lane & 7), so the read-modify-write below does not wait for DRAM.
		// Sectors of voxels that turn out not to be updated are fetched in vain.
		{
			const short2* pf = p.vol + (size_t) (bx * 8) + (size_t) (by * 8 + half * 4 + (lane >> 3)) * p.sx;
			const bool pf_row = by * 8 + half * 4 + (lane >> 3) < p.sy;
#pragma unroll
			for (int l = 0; l < INT_MIXED_CAP; ++l) {
				const int zz = za + 8 * l + (int) (lane & 7);
				if (pf_row && zz < zb && ((cells >> (4 * l + 2 * ((lane & 7) >> 2))) & 3u))
					asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + (size_t) ((uint32_t) zz - p.z_begin) * plane));
			}
		}
		// running values, packed in pairs: (pos.x, pos.y) (pos.z, cam.x) (cam.y, cam.z), at slice za: from the REPLAY job's
		// checkpoint, or (more items than checkpoint slots / the job is not done after a long wait) by replaying the
		// reference's additions from z = 0 here
		F2 A, B, C, dA, dB, dC;
		{
			const float3 delta = mat_rotate(geom.invTrack, f3(0, 0, p.dz / (float) p.sz)), cameraDelta = mat_rotate(p.K, delta);
			dA = f2_make(delta.x, delta.y); dB = f2_make(delta.z, cameraDelta.x); dC = f2_make(cameraDelta.y, cameraDelta.z);
		}
		bool have = false;
		if (idx < q.ckpt_cap) {
			// The job was claimed before this item (queue order) by a warp that is running and never waits: the flag WILL
			// flip.  The poll is bounded anyway — a warp that gives up just does the replay itself.
			const unsigned int* flag = q.ready + ready_index(q, bx, by, half, (uint32_t) za);
			for (int spin = 0; spin < 4096; ++spin) {
				if (ld_acquire_u32(flag) == q.seq) { have = true; break; }
				__nanosleep(200);
			}
			have = __all_sync(0xffffffffu, have);
		}
		if (have) {
			const unsigned long long* o = q.ckpt + (size_t) idx * 96 + lane;
			A.v = __ldcg(o); B.v = __ldcg(o + 32); C.v = __ldcg(o + 64);
		} else {
			const IntColumn c = int_column(p, geom.invTrack, x, y);
			A = f2_make(c.pos0.x, c.pos0.y); B = f2_make(c.pos0.z, c.cam0.x); C = f2_make(c.cam0.y, c.cam0.z);
			replay_advance(A, B, C, dA, dB, dC, za, false);
		}
		short2* col = p.vol + (size_t) x + (size_t) y * p.sx + (size_t) ((uint32_t) za - p.z_begin) * plane;
		// INT_V consecutive slices per step: the decisions first, straight-line (depth gathers hit L1/L2), then the voxel
		// loads back to back, then the updates and stores.
#pragma unroll 1
		for (int z = za; z < zb; z += INT_V, col += INT_V * plane) {
			const unsigned int cls = (cells >> (2 * ((z - za) >> 2))) & 3u;   // warp-uniform
			if (cls == CLS_SKIP) {
#pragma unroll
				for (int u = 0; u < INT_V; ++u) { A = f2_add(A, dA); B = f2_add(B, dB); C = f2_add(C, dC); }
				continue;
			}
			const bool edge = cls == CLS_MIXED_EDGE;
			float sdf[INT_V];
			short2 v[INT_V];
			unsigned int low = 0;   // bit u: this lane stored a tsdf below BRICK_T in slice z + u
			if (cls == CLS_FREE) {
#pragma unroll
				for (int u = 0; u < INT_V; ++u) {
					A = f2_add(A, dA); B = f2_add(B, dB); C = f2_add(C, dC);
					sdf[u] = (valid && (z + u < zb)) ? 1.f : -4.f;
				}
			} else {
#pragma unroll
				for (int u = 0; u < INT_V; ++u) {
					const float Px = f2_lo(A), Py = f2_hi(A), Pz = f2_lo(B), Cx = f2_hi(B), Cy = f2_lo(C), Cz = f2_hi(C);
					A = f2_add(A, dA); B = f2_add(B, dB); C = f2_add(C, dC);
					bool act = valid && (z + u < zb);
					if (edge) act = act && !(Pz < 0.0001f);
					float s = -4.f;   // "no update" (a real sdf is > -1)
					if (fast) {
						const float r = rcp_approx(Cz);
						const float pxf = Cx * r + 0.5f, pyf = Cy * r + 0.5f;
						// the truncated pixel and the bounds tests (against the integers 0, w-1, h-1) can only differ from
						// the exact ones when the value is within `tol` of an integer.  NaN/inf fail both comparisons.
						const bool sure = (fabsf(pxf - rintf(pxf)) >= tol) && (fabsf(pyf - rintf(pyf)) >= tol);
						bool inb = true;
						if (edge) inb = !(pxf < 0 || pxf > dwm1 || pyf < 0 || pyf > dhm1);
						const uint32_t pix = inb ? ((uint32_t) pxf + (uint32_t) pyf * dw) : 0u;
						const float d = __ldg(depth + pix);
						const float e_ = d - Cz;
						// sure & inside: e > mu => sdf == 1 exactly; e < -mu or d == 0 => no update;
						// sure & outside: no update; |e| <= mu: the reference's sqrt/division expression; not sure: all of it
						if (act && sure && inb && e_ > mu && d != 0) s = 1.f;
						if (act && sure && inb && d != 0 && !(e_ > mu) && !(e_ < -mu)) s = integrate_exact_sdf(Px, Py, Pz, e_, mu);
						if (act && !sure) s = integrate_exact(Px, Py, Pz, Cx, Cy, Cz, depth, dw, dwm1, dhm1, mu);
					} else if (act) s = integrate_exact(Px, Py, Pz, Cx, Cy, Cz, depth, dw, dwm1, dhm1, mu);
					sdf[u] = s;
				}
			}
#pragma unroll
			for (int u = 0; u < INT_V; ++u)
				if (sdf[u] > -2.f) v[u] = __ldcs(col + u * plane);
#pragma unroll
			for (int u = 0; u < INT_V; ++u) low |= (sdf[u] > -2.f && sdf[u] < 0.9f) ? (1u << u) : 0u;
			// Brick flags for the raycaster (see BrickMap), while the voxel loads are in flight: an update with sdf < 0.9
			// flags the voxel's brick and, for a voxel on a lower face (x, y or z a multiple of 8), the neighbouring
			// brick(s) whose trilinear taps reach it.  The warp's tile is one brick: lanes 0..7 each own one of the 8
			// (dx, dy, dz) neighbours.
			if (p.brick.flag) {
				const unsigned int m_any = __ballot_sync(0xffffffffu, low != 0u);
				if (m_any) {
					const unsigned int m_z0 = ((z & 7) == 0) ? __ballot_sync(0xffffffffu, low & 1u) : 0u;   // items start on brick layers
					if (lane < 8) {
						const uint32_t ddx = lane & 1, ddy = (lane >> 1) & 1, ddz = lane >> 2;
						unsigned int m = ddz ? m_z0 : m_any;
						if (ddx) m &= 0x01010101u;                    // lanes with x % 8 == 0
						if (ddy) m &= (half == 0) ? 0x000000ffu : 0u;  // lanes with y % 8 == 0
						const uint32_t bz = (uint32_t) z >> BRICK_SHIFT;
						if (m && bx >= ddx && by >= ddy && bz >= ddz)
							brick_set(p.brick, bx - ddx, by - ddy, bz - ddz);
					}
				}
			}
#pragma unroll
			for (int u = 0; u < INT_V; ++u)
				if (sdf[u] > -2.f) {
					__stcs(col + u * plane, tsdf_update(v[u], sdf[u], p.maxweight, rcp));
					++updated;
				}
		}
	}
	updated = __reduce_add_sync(0xffffffffu, updated);
	if (lane == 0 && updated) atomicAdd(p.n_upd, (unsigned long long) updated);
}

#endif
