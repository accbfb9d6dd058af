// EXPERIMENT (north_star item 5, "TMA-staged bricks"): raycastKernel with the bricks along the
// fine-march segment staged into shared memory by the bulk-async copy engine (cp.async.bulk + mbarrier: UBLKCP in SASS)
// instead of being gathered tap by tap with __ldg (also across z-slabs: a row is copied from its owner).  Selected with KFB_RAY_BULK=1; results are bit-identical to k_raycast
// (same taps, same arithmetic — tests/test_gpu_kernels.py runs both), the numbers are in profiles/r2_summary.md.
//
// A warp marches its 8x4 pixel tile in lockstep.  Whenever some lane's sample lies in a FLAGGED brick (the only samples
// that read voxels, ~15 % of all), the warp stages that brick — voxels [8B, 8B + 8] per axis, the 2x2x2 taps of every
// sample whose base voxel is in the brick — as 81 rows of 48 bytes, three bulk copies per lane, completion counted in
// bytes on a per-warp mbarrier; lanes whose cell is in the staged brick interpolate from shared memory, the others (a
// neighbouring brick) take the global path of k_raycast.  A tile's rays sit within ~4 voxels of each other at 512^3, so
// one staged brick serves the whole warp for up to 8 fine steps.
#ifndef KFB_RAYCAST_BULK_CUH
#define KFB_RAYCAST_BULK_CUH

#include "kfb_kernels.cuh"

#define RB_ROW 12                      // voxels per staged row (48 bytes: bulk copies move multiples of 16 bytes)
#define RB_WORDS (9 * 9 * RB_ROW)      // short2 words per staged brick

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

// trilinear sample of cell c from the staged brick whose first voxel is (ox, oy, oz) — vol_interp_cell's arithmetic
__device__ __forceinline__ float interp_staged(const VolView& v, const VolCell& c, const uint32_t* __restrict__ brick, int ox, int oy, int oz) {
	const int lx = kmaxi(c.bx, 0) - ox, ly = kmaxi(c.by, 0) - oy, lz = kmaxi(c.bz, 0) - oz;
	const int ux = kmini(c.bx + 1, (int) v.sx - 1) - ox, uy = kmini(c.by + 1, (int) v.sy - 1) - oy, uz = kmini(c.bz + 1, (int) v.sz - 1) - oz;
#define STAP(Z, Y, X) ((float) (short) (brick[((Z) * 9 + (Y)) * RB_ROW + (X)] & 0xffffu))
	const float v000 = STAP(lz, ly, lx), v100 = STAP(lz, ly, ux), v010 = STAP(lz, uy, lx), v110 = STAP(lz, uy, ux);
	const float v001 = STAP(uz, ly, lx), v101 = STAP(uz, ly, ux), v011 = STAP(uz, uy, lx), v111 = STAP(uz, uy, ux);
#undef STAP
	const float fx = c.fx, fy = c.fy, fz = c.fz, gx = 1 - fx, gy = 1 - fy, gz = 1 - fz;
	return (((v000 * gx + v100 * fx) * gy + (v010 * gx + v110 * fx) * fy) * gz
			+ ((v001 * gx + v101 * fx) * gy + (v011 * gx + v111 * fx) * fy) * fz) * 0.00003051944088f;
}

__global__ void __launch_bounds__(RCK_BX* RCK_BY, 8) k_raycast_bulk(RaycastParams p, unsigned int* bulk_stats) {
	__shared__ __align__(128) uint32_t s_brick[RCK_BY][RB_WORDS];
	__shared__ __align__(8) unsigned long long s_bar[RCK_BY];
	__shared__ Mat4 view;
	const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	if (blockIdx.x == 0 && threadIdx.x == 0) *p.tile_reset = 0u;
	if (threadIdx.x < 16) view.m[threadIdx.x] = p.view_dev ? __ldg(p.view_dev + threadIdx.x) : p.view.m[threadIdx.x];
	const uint32_t bar = smem_u32(&s_bar[wid]);
	if (lane == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	const VolView& v = p.vol;
	uint32_t* brick = s_brick[wid];
	unsigned int phase = 0, staged = 0, served = 0;
	const float3 origin = f3(view.m[3], view.m[7], view.m[11]);
	const uint32_t tiles_x = (p.w + 7) / 8, tiles_y = (p.row1 - p.row0 + 3) / 4, tiles = tiles_x * tiles_y;
	for (;;) {
		uint32_t tl = 0;
		if (lane == 0) tl = atomicAdd(p.tile_next, 1u);
		tl = __shfl_sync(0xffffffffu, tl, 0);
		if (tl >= tiles) break;
		const uint32_t x = (tl % tiles_x) * 8 + (lane & 7), y = p.row0 + (tl / tiles_x) * 4 + (lane >> 3);
		const bool in_img = x < p.w && y < p.row1;
		// ---- ray set-up: cpp/kernels.cpp:674-707 (as raycast_one)
		const float3 direction = mat_rotate(view, f3((float) x, (float) y, 1.f));
		float t = 0, tfar = 0, stepsize = p.largestep, f_t = 0, f_tt = 0, t_lazy = 0;
		bool active = false, lazy = false, hit = false;
		if (in_img) {
			const float3 invR = f3(1.0f / direction.x, 1.0f / direction.y, 1.0f / direction.z);
			const float3 tbot = (-1.f * invR) * origin, ttop = invR * (f3(v.dx, v.dy, v.dz) - origin);
			const float3 tmin = f3(kminf(ttop.x, tbot.x), kminf(ttop.y, tbot.y), kminf(ttop.z, tbot.z));
			const float3 tmax = f3(kmaxf(ttop.x, tbot.x), kmaxf(ttop.y, tbot.y), kmaxf(ttop.z, tbot.z));
			const float tnear = kmaxf(kmaxf(kmaxf(tmin.x, tmin.y), kmaxf(tmin.x, tmin.z)), p.nearPlane);
			tfar = kminf(kminf(kminf(tmax.x, tmax.y), kminf(tmax.x, tmax.z)), p.farPlane);
			if (tnear < tfar) {
				t = tnear;
				f_t = vol_interp(v, origin + direction * t);
				active = f_t > 0 && t < tfar;
				t_lazy = t;
			}
		}
		int cbx = -1, cby = -1, cbz = -1;   // brick staged in shared memory (none)
		// ---- the march, in lockstep: one sample per lane and iteration
		while (__any_sync(0xffffffffu, active)) {
			VolCell c;
			bool need = false;
			if (active) {
				c = vol_cell(v, origin + direction * t);
				if (vol_cell_free(v, c)) { f_tt = 1.f; lazy = true; t_lazy = t; }
				else need = true;
			}
			const unsigned int m_need = __ballot_sync(0xffffffffu, need);
			if (m_need) {
				const int leader = __ffs((int) m_need) - 1;
				const int lbx = __shfl_sync(0xffffffffu, need ? (kmaxi(c.bx, 0) >> 3) : 0, leader), lby = __shfl_sync(0xffffffffu, need ? (kmaxi(c.by, 0) >> 3) : 0, leader),
						lbz = __shfl_sync(0xffffffffu, need ? (kmaxi(c.bz, 0) >> 3) : 0, leader);
				if (lbx != cbx || lby != cby || lbz != cbz) {
					// stage brick (lbx, lby, lbz): rows y, z in [8B, min(8B + 8, N - 1)], 12 voxels from x = 8 lbx (fewer at the edge)
					const int ox = lbx * 8, oy = lby * 8, oz = lbz * 8;
					const int ny = kmini(9, (int) v.sy - oy), nz = kmini(9, (int) v.sz - oz);
					const uint32_t row_bytes = (uint32_t) kmini(RB_ROW, (int) v.sx - ox) * 4u;
					__syncwarp();   // every lane is done reading the previous brick
					if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes * (uint32_t) (ny * nz)) : "memory");
					__syncwarp();
					for (int r = (int) lane; r < ny * nz; r += 32) {
						const int rz = r / ny, ry = r - rz * ny;
						const short2* src = vol_plane(v, oz + rz) + (size_t) ox + (size_t) (oy + ry) * v.sx;   // the slice's owner: maybe a peer (NVLink)
						asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
								::"r"(smem_u32(brick + (rz * 9 + ry) * RB_ROW)), "l"(src), "r"(row_bytes), "r"(bar) : "memory");
					}
					// wait for the bytes (bounded: a wait that never completes must not hang the GPU — the lanes then fall back)
					bool ok = false;
					for (int spin = 0; spin < (1 << 20) && !ok; ++spin) {
						unsigned int done;
						asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
						ok = done != 0;
					}
					phase ^= 1u;
					if (__all_sync(0xffffffffu, ok)) { cbx = lbx; cby = lby; cbz = lbz; ++staged; }
					else { cbx = cby = cbz = -1; if (lane == 0 && bulk_stats) atomicAdd(bulk_stats + 2, 1u); }
				}
			}
			if (need) {
				const bool in_cache = (kmaxi(c.bx, 0) >> 3) == cbx && (kmaxi(c.by, 0) >> 3) == cby && (kmaxi(c.bz, 0) >> 3) == cbz;
				if (in_cache) { f_tt = interp_staged(v, c, brick, cbx * 8, cby * 8, cbz * 8); ++served; }
				else f_tt = vol_interp_cell(v, c);
				if (f_tt < 0) { hit = true; active = false; }       // cpp/kernels.cpp:711
				else { if (f_tt < 0.8f) stepsize = p.step; f_t = f_tt; lazy = false; }
			}
			if (active) { t += stepsize; active = t < tfar; }
		}
		// ---- results (cpp/kernels.cpp:718-757)
		if (in_img) {
			const size_t idx = (size_t) x + (size_t) y * p.w;
			if (hit) {
				if (lazy) f_t = vol_interp(v, origin + direction * t_lazy);
				t = t + stepsize * f_tt / (f_t - f_tt);
				const float3 vtx = origin + direction * t;
				if (t > 0.0f) {
					st3(p.vertex, idx, vtx);
					const float3 surfNorm = vol_grad(v, vtx);
					if (klength(surfNorm) == 0) p.normal[3 * idx] = KFB_INVALID;
					else st3(p.normal, idx, knormalize(surfNorm));
				} else { st3(p.vertex, idx, f3(0, 0, 0)); st3(p.normal, idx, f3(KFB_INVALID, 0, 0)); }
			} else { st3(p.vertex, idx, f3(0, 0, 0)); st3(p.normal, idx, f3(KFB_INVALID, 0, 0)); }
		}
	}
	if (bulk_stats) {
		staged = __reduce_add_sync(0xffffffffu, lane == 0 ? staged : 0u);
		served = __reduce_add_sync(0xffffffffu, served);
		if (lane == 0) { atomicAdd(bulk_stats + 0, staged); atomicAdd(bulk_stats + 1, served); }
	}
}

#endif
