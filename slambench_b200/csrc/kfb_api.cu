// libkfb200.so — context, stage drivers and the C ABI declared in include/kfb200.h.
// Host-side control flow follows the reference's stage drivers
// (kfusion/src/cpp/kernels.cpp:915-1055); all per-pixel / per-voxel work is in
// kfb_kernels.cuh.  There is NO CPU fallback: every entry point that computes needs a
// CUDA device and fails with an error code otherwise.
#include "../../include/kfb200.h"
#include "kfb_kernels.cuh"
#include "kfb_integrate2.cuh"
#include "kfb_raycast_bulk.cuh"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <unistd.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <vector>

// constant_parameters.h:15-23
static const float c_e_delta = 0.1f;
static const int c_radius = 2;
static const float c_dist_threshold = 0.1f;
static const float c_normal_threshold = 0.8f;
static const float c_track_threshold = 0.15f;
static const float c_maxweight = 100.0f;
static const float c_nearPlane = 0.4f;
static const float c_farPlane = 4.0f;
static const float c_delta = 4.0f;

static thread_local char g_err[512] = "";
static int set_err(int code, const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
	return code;
}
#define KFB_E_CUDA 2
#define KFB_E_ARG 3
#define KFB_E_STATE 4
#define CK(call)                                                                                             \
	do {                                                                                                     \
		cudaError_t e_ = (call);                                                                             \
		if (e_ != cudaSuccess) return set_err(KFB_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
	} while (0)

#define NUPD_SLOTS 4096

struct EvPair { cudaEvent_t a, b; };
struct StageTimer {
	std::vector<EvPair> pending, pool;
	double total_ms = 0;
	float last_ms = 0;
	uint64_t count = 0;
};

struct kfb_ctx {
	kfb_config cfg;
	int device;
	cudaStream_t stream;
	// Second stream: when the LAST things enqueued on `stream` are an integrate and / or a raycast, the next frame's H2D
	// copy, preprocessing and pyramid do not depend on them (they rewrite buffers whose readers all precede that integrate;
	// the raw depth and its maximum, which integrate reads, are double- / triple-buffered) and are enqueued here, behind an
	// event recorded just before the integrate; `stream` then waits for them.  They fill the tails of the two persistent
	// kernels (whose warps leave the SMs one by one) instead of starting after them.  Any other enqueue closes the window.
	cudaStream_t side;
	cudaEvent_t ev_window, ev_side_done;
	bool overlap_enabled, overlap_ok, side_pending;
	bool window_late;           // the overlap window opens before the raycast, not before the integrate (KFB_WINDOW_EARLY=1: round 1's choice)
	uint32_t cw, ch;
	int levels;
	uint32_t lw[KFB_MAX_LEVELS], lh[KFB_MAX_LEVELS];
	float step;
	// host pose state (the reference's `pose` member and oldPose / raycastPose globals)
	float pose[16], oldPose[16], raycastPose[16];
	float reduction[32];
	float gaussian[5];
	// device buffers
	short2* d_vol;
	uint32_t z0, z1;            // slab
	size_t slab_voxels;
	float *d_vertex, *d_normal; // raycast maps
	float* d_floatDepth;        // the CURRENT raw depth (= d_fd[fd_cur]); preprocessing writes the other buffer and flips
	float* d_fd[2]; int fd_cur;
	float* d_scaled[KFB_MAX_LEVELS];
	float* d_inV[KFB_MAX_LEVELS];
	float* d_inN[KFB_MAX_LEVELS];
	uint16_t* d_input; size_t input_bytes;
	uint16_t* h_stage[2]; size_t stage_bytes[2]; cudaEvent_t ev_stage[2]; int stage_cur;   // pinned staging for pageable callers
	int8_t* d_status;
	double* d_partials;
	unsigned int* d_counter;
	float* d_out32;
	float* h_out32;             // mapped pinned: [0..31] result, [32] seq flag, [48..63] pose, [64] iterations
	unsigned int* d_bar;        // k_icp grid barrier: [0] arrivals, [1] generation, [2] converged
	float* d_pose;              // k_icp: current pose [16]
	DevFrame* d_frame;          // frame state written by k_icp's tail (kfb_compute_frame's whole-frame enqueue)
	bool no_async;              // KFB_NO_ASYNC=1: kfb_compute_frame always takes the staged path (A/B)
	cudaEvent_t ev_h2d; bool ev_h2d_pending;   // the last asynchronous copy out of a caller's buffer
	int icp_grid;               // co-resident CTAs for the cooperative launch
	unsigned long long* d_icp_prof;   // phase timers, only with KFB_ICP_PROFILE=1
	float* h_out32_dev;
	uint32_t seq;
	unsigned long long* d_nupd; // NUPD_SLOTS per-integrate counters
	unsigned long long nupd_folded;   // counts of recycled slots
	unsigned int* d_dmax;       // three rotating slots: bit pattern of max(floatDepth), written by preprocess
	uint32_t int_zchunk;        // integrate piece length override (KFB_INT_ZCHUNK, tuning)
	uint2* d_queue; size_t queue_cap;   // integrate work list
	// integrate v2: (min, max) depth pyramids (one per raw-depth buffer), brick classes of the current launch
	float2* d_mip[2]; DepthMip mip[2]; bool mip_valid[2];
	unsigned int* d_mip_ticket;
	unsigned char* d_cls; size_t cls_bytes;
	uint2 *d_qmixed, *d_qfree; uint4* d_qreplay;   // integrate v2 work items (worst-case sized at create)
	unsigned int* d_q2ctr;               // 2 slots x {#mixed, #free, next, #replay}
	unsigned int* d_ready;               // [by][bx][half] checkpoint-ready flags (== launch number)
	unsigned long long* d_ckpt; unsigned int ckpt_cap;   // running-value checkpoints of the MIXED items (768 B each)
	BrickMap brick;             // brick flags for the raycaster (whole-volume contexts only)
	bool brick_off;
	unsigned int* d_queue_ctr;  // 2 slots x {count, head}
	int int_grid, int_grid2;    // persistent CTAs of k_integrate_run / k_integrate_run2
	int ray_grid;               // persistent CTAs of k_raycast
	unsigned int* d_tile_ctr;   // two alternating tile counters
	unsigned int* d_tile_cost;  // raycast schedule (RaySched): cost[tiles], stamp[tiles], slow[3][tiles], n[3], sum[3]
	uint32_t ray_tiles_cap;
	int ray_sched;              // slow tiles first (KFB_RAY_NO_SCHED=1: index order)
	uint64_t ray_launches;
	bool ray_bulk; unsigned int* d_bulk_stats; int ray_bulk_grid;   // KFB_RAY_BULK=1: the bulk-async (TMA engine) staging experiment
	uint64_t int_launches;
	int dmax_slot;              // slot holding the max of the CURRENT floatDepth, -1 = unknown
	uint64_t preprocess_count;
	uint64_t integrate_count;
	uchar4* d_render; size_t render_bytes;
	// multi-GPU
	int rank, world;
	VolView view_all;           // slab table for raycast
	void* peer_ptrs[KFB_MAX_SLABS];
	// z-slab mode over peer memory (kfb_ipc_import): peers' maps, flag maps and barrier slots, opened through CUDA IPC
	bool peer_mode;
	void* peer_open[KFB_MAX_SLABS][4];      // vertex, normal, bricks, sync of rank r (to close)
	float* peer_vertex[KFB_MAX_SLABS]; float* peer_normal[KFB_MAX_SLABS]; unsigned char* peer_bricks[KFB_MAX_SLABS];
	PeerSync* d_sync; PeerSyncTable sync_all;
	unsigned int barrier_count;
	uint32_t band0, band1;      // pixel rows handled by this context (multi-GPU); whole image by default
	// host buffers page-locked on the caller's explicit request (kfb_register_host_buffer)
	const void* reg_ptr[8]; int n_reg;
	// stats
	kfb_stats st;
	uint32_t timing;            // bitmask: 1 preprocess, 2 track, 4 integrate, 8 raycast
	StageTimer t_pre, t_track, t_int, t_ray;
};

// ------------------------------------------------------------------------------------------
static void timer_begin(kfb_ctx* c, StageTimer& t, uint32_t bit) {
	if (!(c->timing & bit)) return;
	EvPair p;
	if (!t.pool.empty()) { p = t.pool.back(); t.pool.pop_back(); }
	else { cudaEventCreate(&p.a); cudaEventCreate(&p.b); }
	cudaEventRecord(p.a, c->stream);
	t.pending.push_back(p);
}
static void timer_end(kfb_ctx* c, StageTimer& t, uint32_t bit) {
	if (!(c->timing & bit)) return;
	cudaEventRecord(t.pending.back().b, c->stream);
}
static void timer_resolve(kfb_ctx* c, StageTimer& t) {
	for (auto& p : t.pending) {
		cudaEventSynchronize(p.b);
		float ms = 0;
		cudaEventElapsedTime(&ms, p.a, p.b);
		t.total_ms += ms; t.last_ms = ms; t.count++;
		t.pool.push_back(p);
	}
	t.pending.clear();
}
static void timer_free(StageTimer& t) {
	for (auto& p : t.pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
	for (auto& p : t.pool) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
	t.pending.clear(); t.pool.clear();
}

// Is q = a*rd, q + fma(-d,q,a)*rd == a/d for EVERY float a (all 2^23 significands, two binades)?  Checked once per
// divisor on the host (fmaf is exact); the raycaster then divides by the volume dimensions in 3 instructions.
static int kfb_fastdiv_ok(float d) {
	static std::mutex mu;            // contexts may be created from several threads
	std::lock_guard<std::mutex> lock(mu);
	static float cache_d[8]; static int cache_r[8]; static int n_cache = 0;
	for (int i = 0; i < n_cache; ++i) if (cache_d[i] == d) return cache_r[i];
	int ok = (d == d) && d > 1e-20f && d < 1e20f;
	if (ok) {
		const float rd = 1.0f / d;
		for (uint32_t e = 127; e <= 128 && ok; ++e)
			for (uint32_t m = 0; m < (1u << 23); ++m) {
				const uint32_t u = (e << 23) | m;
				float a;
				memcpy(&a, &u, 4);
				const float q = a * rd;
				const float q1 = fmaf(fmaf(-d, q, a), rd, q);
				if (q1 != a / d) { ok = 0; break; }
			}
	}
	if (n_cache < 8) { cache_d[n_cache] = d; cache_r[n_cache] = ok; ++n_cache; }
	return ok;
}

static inline Mat4 toMat(const float* m) { Mat4 r; memcpy(r.m, m, sizeof r.m); return r; }
// every enqueue on the main stream ends the window in which the next frame's preprocessing may overlap (see kfb_ctx::side)
#define LAUNCHED(c) ((c)->st.kernel_launches++, (c)->overlap_ok = false, (c)->side_pending = false)

// brick flags + super-brick flags (one allocation)
static size_t brick_bytes(const kfb_ctx* c) {
	return c->brick.n_bricks + (size_t) c->brick.snx * c->brick.sny * ((c->brick.bnz + 7) / 8);
}
static int launch_init_volume(kfb_ctx* c) {
	const size_t n = c->slab_voxels, n4 = n / 4;
	k_init_volume<<<148 * 8, 256, 0, c->stream>>>((uint4*) c->d_vol, n4, c->d_vol, n);
	LAUNCHED(c);
	if (c->brick.flag) CK(cudaMemsetAsync(c->brick.flag, 0, brick_bytes(c), c->stream));   // 32766 everywhere: bricks and super-bricks clear
	CK(cudaGetLastError());
	return 0;
}

extern "C" {

int kfb_abi_version(void) { return KFB_ABI_VERSION; }
const char* kfb_last_error(void) { return g_err; }

void kfb_inverse4(float out[16], const float in[16]) { hm_inverse4(out, in); }
void kfb_matmul4(float out[16], const float a[16], const float b[16]) { hm_matmul4(out, a, b); }
void kfb_camera_matrix(float out[16], const float k[4]) { hm_camera_matrix(out, k); }
void kfb_inverse_camera_matrix(float out[16], const float k[4]) { hm_inverse_camera_matrix(out, k); }
int kfb_k_update_pose(float pose[16], const float red[32], float icp_threshold, int* converged) {
	const int r = hm_update_pose(pose, red, icp_threshold);
	if (converged) *converged = r;
	return 0;
}
int kfb_k_check_pose(float pose[16], const float old_pose[16], const float red[32], uint32_t w, uint32_t h, float thr, int* ok) {
	const int r = hm_check_pose(pose, old_pose, red, w, h, thr);
	if (ok) *ok = r;
	return 0;
}

// everything kfb_create allocates; on any failure the caller (kfb_create) destroys the partially built context
static int create_impl(const kfb_config* cfg, kfb_ctx* c) {
	CK(cudaSetDevice(cfg->device));
	memset(&c->st, 0, sizeof c->st);
	c->cfg = *cfg;
	{ const char* e = getenv("KFB_FLAGS"); if (e) c->cfg.flags |= (uint32_t) strtoul(e, nullptr, 0); }   // experiments: OR extra flags in
	c->device = cfg->device;
	c->cw = cfg->compute_w; c->ch = cfg->compute_h;
	c->levels = cfg->n_levels;
	for (int l = 0; l < c->levels; ++l) { c->lw[l] = c->cw >> l; c->lh[l] = c->ch >> l; }
	memcpy(c->pose, cfg->init_pose, sizeof c->pose);
	memset(c->oldPose, 0, sizeof c->oldPose);        // zero-initialised globals (cpp/kernels.cpp:52-53)
	memset(c->raycastPose, 0, sizeof c->raycastPose);
	memset(c->reduction, 0, sizeof c->reduction);    // calloc'd (cpp/kernels.cpp:73)
	// step = min(volumeDimensions) / max(volumeResolution)   kernels.h:116
	{
		const float mind = kminf(kminf(cfg->volume_dim[0], cfg->volume_dim[1]), cfg->volume_dim[2]);
		uint32_t maxr = cfg->volume_res[0] > cfg->volume_res[1] ? cfg->volume_res[0] : cfg->volume_res[1];
		if (cfg->volume_res[2] > maxr) maxr = cfg->volume_res[2];
		c->step = mind / (float) maxr;
	}
	// gaussian (cpp/kernels.cpp:101-107): integer x, integer -(x*x)
	for (unsigned i = 0; i < (unsigned) (c_radius * 2 + 1); ++i) {
		const int x = (int) i - 2;
		c->gaussian[i] = expf(-(x * x) / (2 * c_delta * c_delta));
	}
	c->z0 = cfg->slab_z0; c->z1 = cfg->slab_z1;
	if (c->z0 == 0 && c->z1 == 0) c->z1 = cfg->volume_res[2];
	if (c->z1 > cfg->volume_res[2] || c->z0 >= c->z1) return set_err(KFB_E_ARG, "bad z-slab [%u,%u)", cfg->slab_z0, cfg->slab_z1);
	// integrate decides per 8^3 brick and the raycaster's brick flags are per 8 slices: a slab starts on a brick layer
	if (c->z0 % 8 != 0) return set_err(KFB_E_ARG, "z-slab start %u is not a multiple of 8 (brick layers)", c->z0);
	c->slab_voxels = (size_t) cfg->volume_res[0] * cfg->volume_res[1] * (c->z1 - c->z0);
	c->timing = 0;
	c->rank = 0; c->world = 1; c->band0 = 0; c->band1 = cfg->compute_h;

	CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	CK(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
	CK(cudaEventCreateWithFlags(&c->ev_window, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&c->ev_side_done, cudaEventDisableTiming));
	{ const char* e = getenv("KFB_NO_OVERLAP"); c->overlap_enabled = !(e && atoi(e) > 0); }
	// measured in round 2 (profiles/r2_summary.md): with the brick-classified integrate the early window gives the same fps
	// (2600 vs 2620) but lets the next frame's preprocessing compete with k_integrate_run2 (133 vs 121 us)
	{ const char* e = getenv("KFB_WINDOW_EARLY"); c->window_late = !(e && atoi(e) > 0); }
	c->overlap_ok = false; c->side_pending = false;
	const size_t P = (size_t) c->cw * c->ch;
	CK(cudaMalloc(&c->d_vol, c->slab_voxels * sizeof(short2)));
	CK(cudaMalloc(&c->d_vertex, P * 3 * sizeof(float)));
	CK(cudaMalloc(&c->d_normal, P * 3 * sizeof(float)));
	CK(cudaMalloc(&c->d_fd[0], P * sizeof(float)));
	CK(cudaMalloc(&c->d_fd[1], P * sizeof(float)));
	CK(cudaMemsetAsync(c->d_fd[1], 0, P * sizeof(float), c->stream));
	c->fd_cur = 0; c->d_floatDepth = c->d_fd[0];
	// calloc'd in the reference (cpp/kernels.cpp:75-98): first-frame reads of `vertex`/`normal` see zeros
	CK(cudaMemsetAsync(c->d_vertex, 0, P * 3 * sizeof(float), c->stream));
	CK(cudaMemsetAsync(c->d_normal, 0, P * 3 * sizeof(float), c->stream));
	CK(cudaMemsetAsync(c->d_floatDepth, 0, P * sizeof(float), c->stream));
	for (int l = 0; l < c->levels; ++l) {
		const size_t n = (size_t) c->lw[l] * c->lh[l];
		CK(cudaMalloc(&c->d_scaled[l], n * sizeof(float)));
		CK(cudaMalloc(&c->d_inV[l], n * 3 * sizeof(float)));
		CK(cudaMalloc(&c->d_inN[l], n * 3 * sizeof(float)));
		CK(cudaMemsetAsync(c->d_scaled[l], 0, n * sizeof(float), c->stream));
		CK(cudaMemsetAsync(c->d_inV[l], 0, n * 3 * sizeof(float), c->stream));
		CK(cudaMemsetAsync(c->d_inN[l], 0, n * 3 * sizeof(float), c->stream));
	}
	CK(cudaMalloc(&c->d_status, P));
	CK(cudaMemsetAsync(c->d_status, 0, P, c->stream));
	CK(cudaMalloc(&c->d_partials, (size_t) TR_MAX_BLOCKS * 32 * sizeof(double)));
	CK(cudaMalloc(&c->d_counter, sizeof(unsigned int)));
	CK(cudaMemsetAsync(c->d_counter, 0, sizeof(unsigned int), c->stream));
	CK(cudaMalloc(&c->d_out32, 32 * sizeof(float)));
	CK(cudaMemsetAsync(c->d_out32, 0, 32 * sizeof(float), c->stream));
	CK(cudaHostAlloc(&c->h_out32, 128 * sizeof(float), cudaHostAllocMapped));
	memset(c->h_out32, 0, 128 * sizeof(float));
	CK(cudaMalloc(&c->d_bar, 4 * sizeof(unsigned int)));
	CK(cudaMemsetAsync(c->d_bar, 0, 4 * sizeof(unsigned int), c->stream));
	CK(cudaMalloc(&c->d_pose, 16 * sizeof(float)));
	CK(cudaMalloc(&c->d_sync, sizeof(PeerSync)));
	CK(cudaMemsetAsync(c->d_sync, 0, sizeof(PeerSync), c->stream));
	CK(cudaMalloc(&c->d_frame, sizeof(DevFrame)));
	CK(cudaMemsetAsync(c->d_frame, 0, sizeof(DevFrame), c->stream));
	{ const char* e = getenv("KFB_NO_ASYNC"); c->no_async = e && atoi(e) > 0; }
	CK(cudaEventCreateWithFlags(&c->ev_h2d, cudaEventDisableTiming));
	for (int i = 0; i < 2; ++i) { CK(cudaEventCreateWithFlags(&c->ev_stage[i], cudaEventDisableTiming)); CK(cudaEventRecord(c->ev_stage[i], c->stream)); }
	c->ev_h2d_pending = false;
	if (getenv("KFB_ICP_PROFILE")) { CK(cudaMalloc(&c->d_icp_prof, 8 * sizeof(unsigned long long))); CK(cudaMemsetAsync(c->d_icp_prof, 0, 8 * sizeof(unsigned long long), c->stream)); }
	{
		int coop = 0, per_sm = 0, sms = 0;
		CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device));
		CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
		CK(cudaFuncSetAttribute(k_icp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ICP_SMEM_BYTES));
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_icp, ICP_THREADS, ICP_SMEM_BYTES));
		if (!coop || per_sm < 1) return set_err(KFB_E_CUDA, "device cannot co-schedule the ICP kernel (cooperative launch %d, CTAs/SM %d)", coop, per_sm);
		if (per_sm > 512 / ICP_THREADS) per_sm = 512 / ICP_THREADS;
		c->icp_grid = sms * per_sm;
		if (c->icp_grid > TR_MAX_BLOCKS) c->icp_grid = TR_MAX_BLOCKS;
	}
	CK(cudaHostGetDevicePointer(&c->h_out32_dev, c->h_out32, 0));
	CK(cudaMalloc(&c->d_nupd, NUPD_SLOTS * sizeof(unsigned long long)));
	CK(cudaMemsetAsync(c->d_nupd, 0, NUPD_SLOTS * sizeof(unsigned long long), c->stream));
	CK(cudaMalloc(&c->d_dmax, 3 * sizeof(unsigned int)));
	CK(cudaMemsetAsync(c->d_dmax, 0, 3 * sizeof(unsigned int), c->stream));
	c->dmax_slot = -1; c->preprocess_count = 0;
	{ const char* e = getenv("KFB_INT_ZCHUNK"); c->int_zchunk = e ? (uint32_t) atoi(e) : 0; }
	c->ray_tiles_cap = ((c->cw + 7) / 8) * ((c->ch + 3) / 4);
	c->ray_sched = getenv("KFB_RAY_NO_SCHED") ? 0 : 1;
	CK(cudaMalloc(&c->d_tile_cost, ((size_t) c->ray_tiles_cap * 5 + 8) * sizeof(unsigned int)));
	CK(cudaMemsetAsync(c->d_tile_cost, 0, ((size_t) c->ray_tiles_cap * 5 + 8) * sizeof(unsigned int), c->stream));
	CK(cudaMalloc(&c->d_tile_ctr, 2 * sizeof(unsigned int)));
	CK(cudaMemsetAsync(c->d_tile_ctr, 0, 2 * sizeof(unsigned int), c->stream));
	{
		int per_sm = 0, sms = 0;
		CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_raycast, RCK_BX * RCK_BY, 0));
		if (per_sm < 1) per_sm = 1;
		const char* e = getenv("KFB_RAY_CTAS_PER_SM");   // tuning: fewer persistent warps = more tiles per warp
		if (e && atoi(e) > 0 && atoi(e) < per_sm) per_sm = atoi(e);
		c->ray_grid = sms * per_sm;
	}
	{
		const char* e = getenv("KFB_RAY_BULK");
		c->ray_bulk = e && atoi(e) > 0 && cfg->volume_res[0] % 8 == 0;
		if (c->ray_bulk) {
			int per_sm = 0, sms = 0;
			CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
			CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_raycast_bulk, RCK_BX * RCK_BY, 0));
			c->ray_bulk_grid = sms * (per_sm < 1 ? 1 : per_sm);
			CK(cudaMalloc(&c->d_bulk_stats, 4 * sizeof(unsigned int)));
			CK(cudaMemsetAsync(c->d_bulk_stats, 0, 4 * sizeof(unsigned int), c->stream));
		}
	}
	CK(cudaMalloc(&c->d_queue_ctr, 4 * sizeof(unsigned int)));
	CK(cudaMemsetAsync(c->d_queue_ctr, 0, 4 * sizeof(unsigned int), c->stream));
	{
		int per_sm = 0, per_sm2 = 0, sms = 0;
		CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_integrate_run, 256, 0));
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, k_integrate_run2, 256, 0));
		if (per_sm < 1) per_sm = 1;
		if (per_sm2 < 1) per_sm2 = 1;
		const char* e = getenv("KFB_INT_CTAS_PER_SM");
		if (e && atoi(e) > 0 && atoi(e) < per_sm) per_sm = atoi(e);
		if (e && atoi(e) > 0 && atoi(e) < per_sm2) per_sm2 = atoi(e);
		c->int_grid = sms * per_sm;
		c->int_grid2 = sms * per_sm2;
	}
	// integrate v2: two (min, max) depth pyramids (they follow the two raw-depth buffers), brick classes of the slab
	{
		size_t texels = 0;
		DepthMip m;
		memset(&m, 0, sizeof m);
		uint32_t w = (c->cw + 7) / 8, h = (c->ch + 7) / 8;
		for (m.n = 0; m.n < MIP_MAX_LEVELS;) {
			m.w[m.n] = w; m.h[m.n] = h; texels += (size_t) w * h; ++m.n;
			if (w == 1 && h == 1) break;
			w = (w + 1) / 2; h = (h + 1) / 2;
		}
		for (int b = 0; b < 2; ++b) {
			CK(cudaMalloc(&c->d_mip[b], texels * sizeof(float2)));
			c->mip[b] = m;
			size_t off = 0;
			for (int l = 0; l < m.n; ++l) { c->mip[b].lvl[l] = c->d_mip[b] + off; off += (size_t) m.w[l] * m.h[l]; }
			c->mip_valid[b] = false;
		}
		CK(cudaMalloc(&c->d_mip_ticket, sizeof(unsigned int)));
		CK(cudaMemsetAsync(c->d_mip_ticket, 0, sizeof(unsigned int), c->stream));
		const uint32_t bnx = (cfg->volume_res[0] + 7) / 8, bny = (cfg->volume_res[1] + 7) / 8, bnz = (c->z1 - c->z0 + 7) / 8;
		if (bnx > 4096 || bny > 4096 || cfg->volume_res[2] > 65535) return set_err(KFB_E_ARG, "volume too large for the integrate work list (x, y <= 32768, z <= 65535)");
		c->cls_bytes = (size_t) bnx * bny * bnz;
		CK(cudaMalloc(&c->d_cls, c->cls_bytes));
		// worst cases: every brick per-voxel (runs of <= 8 layers, two halves per column); FREE and other bricks alternating
		const size_t worst_mixed = (size_t) bnx * bny * 2 * ((bnz + INT_MIXED_CAP - 1) / INT_MIXED_CAP + 1);
		CK(cudaMalloc(&c->d_qmixed, worst_mixed * sizeof(uint2)));
		// checkpoints for a bounded number of items (a surface is 2-D: a frame has far fewer per-voxel bricks than the
		// worst case; items beyond the capacity replay the additions themselves)
		{
			size_t cap = worst_mixed / 8;
			if (cap < 32768) cap = 32768;
			if (cap > 524288) cap = 524288;
			if (cap > worst_mixed) cap = worst_mixed;
			c->ckpt_cap = (unsigned int) cap;
			CK(cudaMalloc(&c->d_ckpt, cap * 96 * sizeof(unsigned long long)));
		}
		CK(cudaMalloc(&c->d_qfree, (size_t) bnx * bny * bnz * sizeof(uint2)));   // worst case: every FREE brick an item of its own
		CK(cudaMalloc(&c->d_qreplay, (size_t) bnx * bny * 2 * ((bnz + 255) / 256) * sizeof(uint4)));
		CK(cudaMalloc(&c->d_ready, (size_t) bnx * bny * 2 * ((bnz + 255) / 256) * sizeof(unsigned int)));
		CK(cudaMemsetAsync(c->d_ready, 0, (size_t) bnx * bny * 2 * ((bnz + 255) / 256) * sizeof(unsigned int), c->stream));
		CK(cudaMalloc(&c->d_q2ctr, 8 * sizeof(unsigned int)));
		CK(cudaMemsetAsync(c->d_q2ctr, 0, 8 * sizeof(unsigned int), c->stream));
	}
	// single-slab view by default
	memset(&c->view_all, 0, sizeof c->view_all);
	c->view_all.n_slabs = 1;
	c->view_all.slab_ptr[0] = c->d_vol;
	c->view_all.slab_z[0] = c->z0; c->view_all.slab_z[1] = c->z1;
	c->view_all.sx = cfg->volume_res[0]; c->view_all.sy = cfg->volume_res[1]; c->view_all.sz = cfg->volume_res[2];
	c->view_all.dx = cfg->volume_dim[0]; c->view_all.dy = cfg->volume_dim[1]; c->view_all.dz = cfg->volume_dim[2];
	c->view_all.rdx = 1.0f / cfg->volume_dim[0]; c->view_all.rdy = 1.0f / cfg->volume_dim[1]; c->view_all.rdz = 1.0f / cfg->volume_dim[2];
	{ const char* e = getenv("KFB_RAY_NO_LEAP"); c->view_all.no_leap = (e && atoi(e) > 0) ? 1 : 0; }
	c->view_all.fastdiv = (kfb_fastdiv_ok(cfg->volume_dim[0]) && kfb_fastdiv_ok(cfg->volume_dim[1]) && kfb_fastdiv_ok(cfg->volume_dim[2])) ? 1 : 0;
	const bool whole = (c->z0 == 0 && c->z1 == cfg->volume_res[2]);
	if ((whole || (c->cfg.flags & KFB_FLAG_BRICKS_MERGED)) && !(c->cfg.flags & KFB_FLAG_RAYCAST_NO_SKIP)) {
		c->brick.bnx = (cfg->volume_res[0] + 7) / 8; c->brick.bny = (cfg->volume_res[1] + 7) / 8; c->brick.bnz = (cfg->volume_res[2] + 7) / 8;
		c->brick.n_bricks = (size_t) c->brick.bnx * c->brick.bny * c->brick.bnz;
		c->brick.snx = (c->brick.bnx + 7) / 8; c->brick.sny = (c->brick.bny + 7) / 8;
		CK(cudaMalloc(&c->brick.flag, brick_bytes(c)));
		c->view_all.brick = c->brick.flag; c->view_all.bnx = c->brick.bnx; c->view_all.bny = c->brick.bny;
		c->view_all.super = c->brick.flag + c->brick.n_bricks; c->view_all.snx = c->brick.snx; c->view_all.sny = c->brick.sny;
	}
	int rc = launch_init_volume(c);
	if (rc) return rc;
	CK(cudaStreamSynchronize(c->stream));
	return 0;
}

int kfb_create(const kfb_config* cfg, kfb_ctx** out) {
	if (!cfg || !out) return set_err(KFB_E_ARG, "null argument");
	*out = nullptr;
	if (cfg->n_levels < 1 || cfg->n_levels > KFB_MAX_LEVELS)
		return set_err(KFB_E_ARG, "pyramid levels must be 1..%d (got %d)", KFB_MAX_LEVELS, cfg->n_levels);
	if (cfg->compute_w == 0 || cfg->compute_h == 0 || cfg->volume_res[0] == 0 || cfg->volume_res[1] == 0 || cfg->volume_res[2] == 0)
		return set_err(KFB_E_ARG, "empty image or volume");
	if ((cfg->compute_w >> (cfg->n_levels - 1)) == 0 || (cfg->compute_h >> (cfg->n_levels - 1)) == 0)
		return set_err(KFB_E_ARG, "image too small for %d pyramid levels", cfg->n_levels);
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
		return set_err(KFB_E_CUDA, "no CUDA device: libkfb200 has no CPU fallback");
	if (cfg->device < 0 || cfg->device >= ndev) return set_err(KFB_E_ARG, "device %d out of range (%d devices)", cfg->device, ndev);
	kfb_ctx* c = new kfb_ctx();   // value-initialised: every pointer / handle below starts out null
	const int rc = create_impl(cfg, c);
	if (rc) {
		// one cleanup path for every failure: kfb_destroy tolerates a partially built context (it must not clobber the message)
		char keep[sizeof g_err];
		memcpy(keep, g_err, sizeof keep);
		kfb_destroy(c);
		cudaGetLastError();
		memcpy(g_err, keep, sizeof keep);
		return rc;
	}
	*out = c;
	return 0;
}

int kfb_destroy(kfb_ctx* c) {
	if (!c) return 0;
	cudaSetDevice(c->device);
	if (c->stream) cudaStreamSynchronize(c->stream);
	if (c->side) cudaStreamSynchronize(c->side);
	for (int i = 0; i < c->n_reg; ++i) cudaHostUnregister((void*) c->reg_ptr[i]);
	for (int i = 0; i < 2; ++i) { if (c->h_stage[i]) cudaFreeHost(c->h_stage[i]); if (c->ev_stage[i]) cudaEventDestroy(c->ev_stage[i]); }
	for (int i = 0; i < KFB_MAX_SLABS; ++i) if (c->peer_ptrs[i]) cudaIpcCloseMemHandle(c->peer_ptrs[i]);
	for (int i = 0; i < KFB_MAX_SLABS; ++i)
		for (int j = 0; j < 4; ++j) if (c->peer_open[i][j]) cudaIpcCloseMemHandle(c->peer_open[i][j]);
	cudaFree(c->d_sync);
	cudaFree(c->brick.flag);   // cudaFree(nullptr) is a no-op
	cudaFree(c->d_vol); cudaFree(c->d_vertex); cudaFree(c->d_normal); cudaFree(c->d_fd[0]); cudaFree(c->d_fd[1]);
	for (int l = 0; l < KFB_MAX_LEVELS; ++l) { cudaFree(c->d_scaled[l]); cudaFree(c->d_inV[l]); cudaFree(c->d_inN[l]); }
	cudaFree(c->d_status); cudaFree(c->d_partials); cudaFree(c->d_counter); cudaFree(c->d_out32); cudaFree(c->d_nupd); cudaFree(c->d_dmax);
	cudaFree(c->d_queue_ctr); cudaFree(c->d_tile_ctr); cudaFree(c->d_tile_cost); cudaFree(c->d_queue);
	cudaFree(c->d_mip[0]); cudaFree(c->d_mip[1]); cudaFree(c->d_mip_ticket); cudaFree(c->d_cls);
	cudaFree(c->d_qmixed); cudaFree(c->d_qfree); cudaFree(c->d_q2ctr); cudaFree(c->d_ckpt); cudaFree(c->d_qreplay); cudaFree(c->d_ready);
	if (c->d_icp_prof) {
		unsigned long long h[8];
		if (cudaMemcpy(h, c->d_icp_prof, sizeof h, cudaMemcpyDeviceToHost) == cudaSuccess && h[4])
			fprintf(stderr, "kfb icp profile: %llu iterations; per iteration: last-CTA compute %.2f us, final reduce %.2f us, solve %.2f us, whole iteration (CTA 0) %.2f us\n",
					h[4], h[0] * 1e-3 / h[4], h[1] * 1e-3 / h[4], h[2] * 1e-3 / h[4], h[3] * 1e-3 / h[4]);
		cudaFree(c->d_icp_prof);
	}
	if (c->d_bulk_stats) {
		unsigned int h[4];
		if (cudaMemcpy(h, c->d_bulk_stats, sizeof h, cudaMemcpyDeviceToHost) == cudaSuccess && c->ray_launches)
			fprintf(stderr, "kfb raycast bulk staging: %llu launches; per launch: %.0f bricks staged, %.0f samples served from shared memory, %u waits timed out in total\n",
					(unsigned long long) c->ray_launches, (double) h[0] / c->ray_launches, (double) h[1] / c->ray_launches, h[2]);
		cudaFree(c->d_bulk_stats);
	}
	if (c->h_out32) cudaFreeHost(c->h_out32);
	cudaFree(c->d_bar); cudaFree(c->d_pose); cudaFree(c->d_frame);
	if (c->ev_h2d) cudaEventDestroy(c->ev_h2d);
	cudaFree(c->d_input);
	cudaFree(c->d_render);
	timer_free(c->t_pre); timer_free(c->t_track); timer_free(c->t_int); timer_free(c->t_ray);
	if (c->ev_window) cudaEventDestroy(c->ev_window);
	if (c->ev_side_done) cudaEventDestroy(c->ev_side_done);
	if (c->side) cudaStreamDestroy(c->side);
	if (c->stream) cudaStreamDestroy(c->stream);
	cudaGetLastError();
	delete c;
	return 0;
}

int kfb_reset(kfb_ctx* c) {
	if (!c) return set_err(KFB_E_ARG, "null ctx");
	CK(cudaSetDevice(c->device));
	return launch_init_volume(c);
}

int kfb_sync(kfb_ctx* c) {
	if (!c) return set_err(KFB_E_ARG, "null ctx");
	CK(cudaStreamSynchronize(c->stream));
	const unsigned int e = *(volatile unsigned int*) (c->h_out32 + 34);
	if (e) { *(volatile unsigned int*) (c->h_out32 + 34) = 0; return set_err(KFB_E_STATE, "z-slab barrier timed out waiting for rank %u", e - 0x100u); }
	return 0;
}
int kfb_stream(kfb_ctx* c, void** s) { *s = (void*) c->stream; return 0; }
int kfb_get_pose(kfb_ctx* c, float pose[16]) { memcpy(pose, c->pose, sizeof c->pose); return 0; }
int kfb_set_pose(kfb_ctx* c, const float pose[16]) { memcpy(c->pose, pose, sizeof c->pose); return 0; }

// ------------------------------------------------------------------------- preprocessing
static int check_ratio(kfb_ctx* c, uint32_t iw, uint32_t ih, int* ratio) {
	// mm2metersKernel's input validation (cpp/kernels.cpp:565-577); the reference prints
	// "Invalid ratio." and exit(1)s — the C ABI reports it, the Kfusion glue exits.
	if (iw < c->cw || ih < c->ch || iw % c->cw != 0 || ih % c->ch != 0 || iw / c->cw != ih / c->ch)
		return set_err(KFB_E_ARG, "Invalid ratio.");
	*ratio = iw / c->cw;
	return 0;
}
// the stream the next frame's preprocessing goes to: the side stream (made to wait for everything before the raycast
// that is still running) inside the overlap window, else the main stream.  Stage timers keep everything serial.
static cudaStream_t preprocess_stream(kfb_ctx* c, bool* on_side) {
	*on_side = c->overlap_ok && !(c->timing & 3u);
	if (*on_side && cudaStreamWaitEvent(c->side, c->ev_window, 0) != cudaSuccess) { cudaGetLastError(); *on_side = false; }
	return *on_side ? c->side : c->stream;
}
// main stream continues only after the side stream's work; `side_pending`: the pyramid may follow on the side stream
static int join_side(kfb_ctx* c) {
	CK(cudaEventRecord(c->ev_side_done, c->side));
	CK(cudaStreamWaitEvent(c->stream, c->ev_side_done, 0));
	return 0;
}
static int launch_preprocess(kfb_ctx* c, const uint16_t* d_in, uint32_t iw, int ratio, cudaStream_t stream, bool on_side) {
	Gauss5 g;
	memcpy(g.g, c->gaussian, sizeof g.g);
	dim3 grid((c->cw + PP_BX - 1) / PP_BX, (c->ch + PP_BY - 1) / PP_BY), block(PP_BX, PP_BY);
	// the raw depth and its maximum go to buffers the integrate that may still be running does not read: the other
	// floatDepth buffer, and slot n % 3 of max(depth) (slot (n + 1) % 3 is zeroed for the next frame)
	const int slot = (int) (c->preprocess_count++ % 3);
	const int nb = c->fd_cur ^ 1;
	k_mm2m_bilateral<<<grid, block, 0, stream>>>(d_in, iw, ratio, c->d_fd[nb], c->d_scaled[0], c->cw, c->ch, g, c_e_delta,
			c->d_dmax + slot, c->d_dmax + (slot + 1) % 3);
	c->fd_cur = nb; c->d_floatDepth = c->d_fd[nb];
	c->dmax_slot = slot;
	LAUNCHED(c);
	// (min, max) pyramid of the new raw depth for integrate's brick classification: same stream, same overlap window
	k_depth_mip<<<dim3((c->cw + 63) / 64, (c->ch + 63) / 64), 256, 0, stream>>>(c->d_fd[nb], c->cw, c->ch, c->mip[nb], c->d_mip_ticket);
	c->mip_valid[nb] = true;
	c->st.kernel_launches++;
	CK(cudaGetLastError());
	if (on_side) {
		int rc = join_side(c);
		if (rc) return rc;
		c->side_pending = true;
	}
	return 0;
}
static int ensure_input(kfb_ctx* c, size_t bytes) {
	if (c->input_bytes >= bytes) return 0;
	if (c->d_input) CK(cudaFree(c->d_input));
	CK(cudaMalloc(&c->d_input, bytes));
	c->input_bytes = bytes;
	return 0;
}

int kfb_preprocess(kfb_ctx* c, const uint16_t* depth, uint32_t iw, uint32_t ih) {
	if (!c || !depth) return set_err(KFB_E_ARG, "null argument");
	CK(cudaSetDevice(c->device));
	int ratio, rc;
	if ((rc = check_ratio(c, iw, ih, &ratio))) return rc;
	const size_t bytes = (size_t) iw * ih * sizeof(uint16_t);
	if ((rc = ensure_input(c, bytes))) return rc;
	timer_begin(c, c->t_pre, 1u);
	// Is the caller's buffer page-locked RIGHT NOW (cudaHostAlloc, torch's pinned pool, cudaHostRegister, or
	// kfb_register_host_buffer)?  Asked on every call: a cached answer goes stale when the caller frees a buffer and a later
	// allocation reuses the address.  Pageable memory goes through one of two internal pinned staging buffers.
	bool pinned = false;
	{
		cudaPointerAttributes at;
		if (cudaPointerGetAttributes(&at, depth) == cudaSuccess && at.type == cudaMemoryTypeHost) pinned = true;
		else cudaGetLastError();
	}
	bool on_side = false;
	cudaStream_t ps = preprocess_stream(c, &on_side);
	if (pinned) {
		CK(cudaMemcpyAsync(c->d_input, depth, bytes, cudaMemcpyHostToDevice, ps));
		CK(cudaEventRecord(c->ev_h2d, ps));
		c->ev_h2d_pending = true;
	} else {
		const int sb = c->stage_cur ^= 1;
		if (c->stage_bytes[sb] < bytes) {
			if (c->h_stage[sb]) { CK(cudaEventSynchronize(c->ev_stage[sb])); CK(cudaFreeHost(c->h_stage[sb])); c->h_stage[sb] = nullptr; }
			CK(cudaHostAlloc(&c->h_stage[sb], bytes, cudaHostAllocDefault));
			c->stage_bytes[sb] = bytes;
		} else CK(cudaEventSynchronize(c->ev_stage[sb]));   // the copy that last used this staging buffer (two frames ago) has left it
		memcpy(c->h_stage[sb], depth, bytes);                 // the caller's buffer is free again when this call returns
		CK(cudaMemcpyAsync(c->d_input, c->h_stage[sb], bytes, cudaMemcpyHostToDevice, ps));
		CK(cudaEventRecord(c->ev_stage[sb], ps));
	}
	c->st.h2d_bytes += bytes;
	rc = launch_preprocess(c, c->d_input, iw, ratio, ps, on_side);
	timer_end(c, c->t_pre, 1u);
	return rc;
}

int kfb_register_host_buffer(kfb_ctx* c, const void* ptr, size_t bytes) {
	if (!c || !ptr || !bytes) return set_err(KFB_E_ARG, "null argument");
	CK(cudaSetDevice(c->device));
	for (int i = 0; i < c->n_reg; ++i) if (c->reg_ptr[i] == ptr) return 0;
	if (c->n_reg >= 8) return set_err(KFB_E_STATE, "too many registered host buffers (8)");
	CK(cudaHostRegister((void*) ptr, bytes, cudaHostRegisterDefault));
	c->reg_ptr[c->n_reg++] = ptr;
	return 0;
}
int kfb_unregister_host_buffer(kfb_ctx* c, const void* ptr) {
	if (!c || !ptr) return set_err(KFB_E_ARG, "null argument");
	CK(cudaSetDevice(c->device));
	for (int i = 0; i < c->n_reg; ++i)
		if (c->reg_ptr[i] == ptr) {
			CK(cudaStreamSynchronize(c->stream));
			CK(cudaStreamSynchronize(c->side));
			CK(cudaHostUnregister((void*) ptr));
			c->reg_ptr[i] = c->reg_ptr[--c->n_reg];
			return 0;
		}
	return set_err(KFB_E_ARG, "buffer was not registered with this context");
}

int kfb_preprocess_device(kfb_ctx* c, const uint16_t* d_depth, uint32_t iw, uint32_t ih) {
	if (!c || !d_depth) return set_err(KFB_E_ARG, "null argument");
	CK(cudaSetDevice(c->device));
	int ratio, rc;
	if ((rc = check_ratio(c, iw, ih, &ratio))) return rc;
	timer_begin(c, c->t_pre, 1u);
	bool on_side = false;
	cudaStream_t ps = preprocess_stream(c, &on_side);
	rc = launch_preprocess(c, d_depth, iw, ratio, ps, on_side);
	timer_end(c, c->t_pre, 1u);
	return rc;
}

// ------------------------------------------------------------------------------ tracking
static int launch_pyramid(kfb_ctx* c, const float k[4]) {
	PyrParams p;
	memset(&p, 0, sizeof p);
	p.d0 = c->d_scaled[0];
	const int fused = c->levels < 3 ? c->levels : 3;   // k_pyramid recomputes levels 1, 2 from level 0; deeper levels chain
	p.levels = fused;
	p.e_d = c_e_delta * 3;
	uint32_t first = 0;
	for (int l = 0; l < fused; ++l) {
		p.depth[l] = c->d_scaled[l]; p.vertex[l] = c->d_inV[l]; p.normal[l] = c->d_inN[l];
		p.w[l] = c->lw[l]; p.h[l] = c->lh[l];
		p.first[l] = first;
		first += c->lw[l] * c->lh[l];
		// getInverseCameraMatrix(k / float(1 << i))   cpp/kernels.cpp:940
		const float s = (float) (1 << l);
		const float ks[4] = { k[0] / s, k[1] / s, k[2] / s, k[3] / s };
		hm_inverse_camera_matrix(p.invK[l].m, ks);
	}
	p.first[fused] = first;
	// right behind a preprocessing that went to the side stream: follow it there (same window, same readers)
	const bool on_side = c->side_pending && !(c->timing & 3u);
	cudaStream_t st = on_side ? c->side : c->stream;
	k_pyramid<<<(first + 255) / 256, 256, 0, st>>>(p);
	LAUNCHED(c);
	for (int l = 3; l < c->levels; ++l) {
		const float s = (float) (1 << l);
		const float ks[4] = { k[0] / s, k[1] / s, k[2] / s, k[3] / s };
		Mat4 invK;
		hm_inverse_camera_matrix(invK.m, ks);
		const uint32_t n = c->lw[l] * c->lh[l];
		k_pyramid_level<<<(n + 255) / 256, 256, 0, st>>>(c->d_scaled[l - 1], c->lw[l - 1], c->d_scaled[l], c->d_inV[l], c->d_inN[l], c->lw[l], c->lh[l], invK, c_e_delta * 3);
		c->st.kernel_launches++;
	}
	CK(cudaGetLastError());
	if (on_side) return join_side(c);
	return 0;
}

static uint32_t track_blocks(uint32_t npx) {
	uint32_t b = (npx + TR_THREADS - 1) / TR_THREADS;
	if (b > 148 * 4) b = 148 * 4;
	if (b < 1) b = 1;
	return b;
}

// one fused track+reduce launch; result lands in c->h_out32 (mapped) when wait == true
static int wait_seq(kfb_ctx* c, uint32_t seq) {
	volatile uint32_t* flag = (volatile uint32_t*) (c->h_out32 + 32);
	// spin on the mapped flag; fall back to a stream query so a faulted kernel cannot hang us
	uint64_t spins = 0;
	while (*flag != seq) {
		if ((++spins & 0xfffff) == 0) {
			cudaError_t q = cudaStreamQuery(c->stream);
			if (q != cudaErrorNotReady && q != cudaSuccess) return set_err(KFB_E_CUDA, "track kernel failed: %s", cudaGetErrorString(q));
			if (q == cudaSuccess && *flag != seq) return set_err(KFB_E_CUDA, "track kernel finished without publishing its result");
		}
	}
	__sync_synchronize();
	return 0;
}

static int launch_track(kfb_ctx* c, int level, const float* T, const float* V, float dist, float nthr, bool wait) {
	TrackParams p;
	memset(&p, 0, sizeof p);
	p.inV = c->d_inV[level]; p.inN = c->d_inN[level];
	p.refV = c->d_vertex; p.refN = c->d_normal;
	p.w = c->lw[level]; p.h = c->lh[level]; p.rw = c->cw; p.rh = c->ch;
	p.row0 = c->band0 >> level; p.row1 = c->band1 >> level;   // whole level unless kfb_set_pixel_rows narrowed it
	if (c->band1 == c->ch) p.row1 = p.h;
	p.Ttrack = toMat(T); p.view = toMat(V);
	p.pose_dev = nullptr; p.view_dev = nullptr;
	p.dist_threshold = dist; p.normal_threshold = nthr;
	p.partials = c->d_partials; p.counter = c->d_counter; p.out32 = c->d_out32;
	p.out32_host = c->h_out32_dev;
	p.seq_host = (volatile uint32_t*) (c->h_out32_dev + 32);
	p.seq = ++c->seq;
	p.status = (c->cfg.flags & KFB_FLAG_TRACK_STATUS) ? c->d_status : nullptr;
	const uint32_t blocks = track_blocks(p.w * (p.row1 - p.row0));
	k_track_reduce<<<blocks, TR_THREADS, 0, c->stream>>>(p);
	LAUNCHED(c);
	CK(cudaGetLastError());
	if (wait) {
		int rc = wait_seq(c, p.seq);
		if (rc) return rc;
		memcpy(c->reduction, c->h_out32, 32 * sizeof(float));
		c->st.d2h_bytes += 32 * sizeof(float);
	}
	return 0;
}

int kfb_k_pyramid(kfb_ctx* c, const float k[4]) {
	CK(cudaSetDevice(c->device));
	return launch_pyramid(c, k);
}

int kfb_k_track_reduce(kfb_ctx* c, int level, const float T[16], const float V[16], float dist, float nthr, float out32[32]) {
	if (level < 0 || level >= c->levels) return set_err(KFB_E_ARG, "level %d out of range", level);
	CK(cudaSetDevice(c->device));
	int rc = launch_track(c, level, T, V, dist, nthr, true);
	if (rc) return rc;
	if (out32) memcpy(out32, c->reduction, 32 * sizeof(float));
	return 0;
}

// the whole ICP schedule of one frame as ONE persistent cooperative kernel (k_icp); `tail` (optional) makes its last CTA
// also evaluate checkPose and the matrices integrate / raycast need, into c->d_frame.  Returns the sequence number to wait for.
static int launch_icp(kfb_ctx* c, float icp_threshold, const float* projectReference, const IcpTail* tail, uint32_t* seq_out) {
	IcpParams p;
	memset(&p, 0, sizeof p);
	for (int l = 0; l < c->levels; ++l) {
		p.inV[l] = c->d_inV[l]; p.inN[l] = c->d_inN[l]; p.w[l] = c->lw[l]; p.h[l] = c->lh[l];
		p.iterations[l] = c->cfg.iterations[l];
	}
	p.levels = c->levels;
	p.refV = c->d_vertex; p.refN = c->d_normal; p.rw = c->cw; p.rh = c->ch;
	p.pose0 = toMat(c->pose); p.view = toMat(projectReference);
	p.dist_threshold = c_dist_threshold; p.normal_threshold = c_normal_threshold; p.icp_threshold = icp_threshold;
	p.partials = c->d_partials; p.bar = c->d_bar; p.pose_dev = c->d_pose; p.out32 = c->d_out32; p.prof = c->d_icp_prof;
	p.out_host = c->h_out32_dev;
	p.seq = ++c->seq;
	p.status = (c->cfg.flags & KFB_FLAG_TRACK_STATUS) ? c->d_status : nullptr;
	if (tail) p.tail = *tail;
	c->h_out32[33] = 0.f;
	void* args[] = { &p };
	CK(cudaLaunchCooperativeKernel((const void*) k_icp, dim3(c->icp_grid), dim3(ICP_THREADS), args, ICP_SMEM_BYTES, c->stream));
	LAUNCHED(c);
	*seq_out = p.seq;
	return 0;
}
// wait for that kernel's results (mapped host memory) and take them over: sums, pose, iteration count
static int collect_icp(kfb_ctx* c, uint32_t seq, uint64_t* iters) {
	int rc = wait_seq(c, seq);
	if (rc) return rc;
	if (*(volatile uint32_t*) (c->h_out32 + 33) != 0) {
		cudaStreamSynchronize(c->stream);
		cudaMemsetAsync(c->d_bar, 0, 4 * sizeof(unsigned int), c->stream);
		return set_err(KFB_E_CUDA, "ICP kernel: grid barrier timed out");
	}
	memcpy(c->reduction, c->h_out32, 32 * sizeof(float));
	memcpy(c->pose, c->h_out32 + 48, 16 * sizeof(float));
	*iters = *(volatile uint32_t*) (c->h_out32 + 64);
	c->st.d2h_bytes += (32 + 16 + 1) * sizeof(float);
	return 0;
}
static int icp_total_iterations(const kfb_ctx* c) {
	int total = 0;
	for (int level = 0; level < c->levels; ++level) total += c->cfg.iterations[level] > 0 ? c->cfg.iterations[level] : 0;
	return total;
}

int kfb_track(kfb_ctx* c, const float k[4], float icp_threshold, uint32_t tracking_rate, uint32_t frame, int* tracked) {
	if (!c || !k) return set_err(KFB_E_ARG, "null argument");
	CK(cudaSetDevice(c->device));
	if (tracked) *tracked = 0;
	if (tracking_rate == 0) return set_err(KFB_E_ARG, "tracking_rate must be > 0");
	if (frame % tracking_rate != 0) {                               // cpp/kernels.cpp:927
		// preprocessing() is synchronous with respect to the caller's buffer in the reference: the asynchronous copy of this
		// frame must have left it before the documented "reusable after the next kfb_track" holds
		if (c->ev_h2d_pending) { CK(cudaEventSynchronize(c->ev_h2d)); c->ev_h2d_pending = false; }
		return 0;
	}
	timer_begin(c, c->t_track, 2u);
	int rc = launch_pyramid(c, k);                                  // :931-945
	if (rc) return rc;
	memcpy(c->oldPose, c->pose, sizeof c->pose);                    // :947
	float K[16], invRP[16], projectReference[16];
	hm_camera_matrix(K, k);
	hm_inverse4(invRP, c->raycastPose);
	hm_matmul4(projectReference, K, invRP);                         // :948
	uint64_t iters = 0;
	if (c->cfg.flags & KFB_FLAG_ICP_HOST_SOLVE) {
		for (int level = c->levels - 1; level >= 0; --level) {         // :950-967
			for (int i = 0; i < c->cfg.iterations[level]; ++i) {
				rc = launch_track(c, level, c->pose, projectReference, c_dist_threshold, c_normal_threshold, true);
				if (rc) return rc;
				++iters;
				if (hm_update_pose(c->pose, c->reduction, icp_threshold)) break;
			}
		}
	} else if (icp_total_iterations(c) > 0) {
		// device-resident loop: ONE persistent cooperative kernel runs the whole schedule (solve, pose
		// update and the per-level `break` in its last-arriving CTA); ONE host wait per frame
		uint32_t seq;
		if ((rc = launch_icp(c, icp_threshold, projectReference, nullptr, &seq))) return rc;
		if ((rc = collect_icp(c, seq, &iters))) return rc;
	}
	c->ev_h2d_pending = false;   // the stream has passed the copy
	timer_end(c, c->t_track, 2u);
	c->st.icp_iterations_last = iters;
	c->st.icp_iterations_total += iters;
	const int ok = hm_check_pose(c->pose, c->oldPose, c->reduction, c->cw, c->ch, c_track_threshold);  // :968
	if (tracked) *tracked = ok;
	return 0;
}

// --------------------------------------------------------------------------- integration
static int launch_integrate(kfb_ctx* c, const float* invTrack, const float* K, float mu, float maxweight, const DevFrame* dev = nullptr) {
	// everything the next frame's preprocessing rewrites has been read by now, except the raw depth and its maximum,
	// which are double- / triple-buffered: its window (kfb_ctx::side) opens here
	if (c->overlap_enabled && !c->window_late) CK(cudaEventRecord(c->ev_window, c->stream));
	IntegrateParams p;
	p.vol = c->d_vol;
	p.sx = c->cfg.volume_res[0]; p.sy = c->cfg.volume_res[1]; p.sz = c->cfg.volume_res[2];
	p.dx = c->cfg.volume_dim[0]; p.dy = c->cfg.volume_dim[1]; p.dz = c->cfg.volume_dim[2];
	p.z_begin = c->z0; p.z_end = c->z1;
	p.depth = c->d_floatDepth; p.dw = c->cw; p.dh = c->ch;
	p.invTrack = toMat(invTrack); p.K = toMat(K);
	p.dev = dev;               // non-null: invTrack and the integrate gate come from the ICP kernel's tail
	p.mu = mu; p.maxweight = maxweight;
	p.cull = (c->cfg.flags & KFB_FLAG_INTEGRATE_NO_CULL) ? 0 : 1;
	p.brick = c->brick;
	p.brick.n_peer = 0;
	if (c->peer_mode && c->brick.flag)
		for (int r = 0; r < c->world; ++r) if (r != c->rank && c->peer_bricks[r]) p.brick.peer[p.brick.n_peer++] = c->peer_bricks[r];
	if (maxweight > 200.f && c->brick.flag) {
		// the flagging rule in k_integrate_run assumes w + 1 <= 201; beyond that stop using (and maintaining) the flags
		c->view_all.brick = nullptr; c->view_all.super = nullptr; p.brick.flag = nullptr; c->brick_off = true;
	}
	if (c->brick_off) p.brick.flag = nullptr;
	p.dmax = (c->dmax_slot >= 0) ? reinterpret_cast<const float*>(c->d_dmax + c->dmax_slot) : nullptr;
	const uint32_t slot = (uint32_t) (c->integrate_count % NUPD_SLOTS);
	if (c->integrate_count >= NUPD_SLOTS) {
		// a slot is recycled: its count moves into the host-side running total first (rare: once per 4096 integrates)
		unsigned long long old = 0;
		CK(cudaStreamSynchronize(c->stream));
		CK(cudaMemcpy(&old, c->d_nupd + slot, sizeof old, cudaMemcpyDeviceToHost));
		c->nupd_folded += old;
		CK(cudaMemsetAsync(c->d_nupd + slot, 0, sizeof(unsigned long long), c->stream));
	}
	p.n_upd = c->d_nupd + slot;
	// pass 1 cuts every warp-column's visited interval into pieces of `zchunk` slices; pass 2 is persistent
	// short pieces keep every warp's serial chain short (measured: 256^3 best at 16-32, 512^3 at 24-32)
	{
		const uint32_t nz = c->z1 - c->z0;
		uint32_t zc = (nz / 8) & ~7u;   // a multiple of INT_U so a warp can continue into the next piece
		const uint32_t zc_max = nz <= 512 ? 32u : 64u;   // measured: 512 slices 159 us at 24-32 vs 166 us at 64 (flat below)
		zc = zc < 16 ? 16 : (zc > zc_max ? zc_max : zc);
		p.zchunk = c->int_zchunk ? c->int_zchunk : zc;
	}
	const int qslot = (int) (c->int_launches++ & 1);   // never reset: the slots alternate strictly
	p.queue = c->d_queue;
	p.queue_count = c->d_queue_ctr + 2 * qslot; p.queue_head = p.queue_count + 1;
	p.queue_next = c->d_queue_ctr + 2 * (qslot ^ 1);
	{
		size_t need = (size_t) ((p.sx + 31) / 32) * p.sy;   // v1: one entry per warp-column; v2: two per brick column
		const size_t need2 = (size_t) ((p.sx + 7) / 8) * ((p.sy + 7) / 8) * 2;
		if (need2 > need) need = need2;
		if (need > c->queue_cap) {
			if (c->d_queue) { CK(cudaStreamSynchronize(c->stream)); CK(cudaFree(c->d_queue)); }
			CK(cudaMalloc(&c->d_queue, need * (sizeof(uint2) + sizeof(unsigned int))));
			c->queue_cap = need;
			p.queue = c->d_queue;
		}
		p.piece_ctr = reinterpret_cast<unsigned int*>(c->d_queue + c->queue_cap);
		if (p.zchunk % INT_U != 0 || p.zchunk == 0) return set_err(KFB_E_ARG, "integrate piece length must be a positive multiple of %d", INT_U);
	}
	if (c->cfg.flags & KFB_FLAG_INTEGRATE_V1) {
		dim3 block(32, 8), grid((p.sx + 31) / 32, (p.sy + 7) / 8);
		k_integrate_plan<<<grid, block, 0, c->stream>>>(p);
		LAUNCHED(c);
		k_integrate_run<<<c->int_grid, 256, 0, c->stream>>>(p);
		LAUNCHED(c);
	} else {
		Integrate2Params q;
		q.b = p;
		if (!c->mip_valid[c->fd_cur]) {   // raw depth written from the host (teacher-forced tests): build its pyramid now
			k_depth_mip<<<dim3((c->cw + 63) / 64, (c->ch + 63) / 64), 256, 0, c->stream>>>(c->d_floatDepth, c->cw, c->ch, c->mip[c->fd_cur], c->d_mip_ticket);
			LAUNCHED(c);
			c->mip_valid[c->fd_cur] = true;
		}
		q.mip = c->mip[c->fd_cur];
		q.cls = c->d_cls;
		q.bnx = (p.sx + 7) / 8; q.bny = (p.sy + 7) / 8;
		q.maxw_i = (maxweight >= 1.f && maxweight <= 32767.f) ? (int) maxweight : -1;
		q.vec_ok = (p.sx % 8 == 0) ? 1 : 0;
		q.std_k = (K[8] == 0.f && K[9] == 0.f && K[10] == 1.f && K[11] == 0.f) ? 1 : 0;
		q.vsz[0] = p.dx / (float) p.sx; q.vsz[1] = p.dy / (float) p.sy; q.vsz[2] = p.dz / (float) p.sz;
		for (int r = 0; r < 3; ++r)
			for (int cc = 0; cc < 3; ++cc)
				q.ca[3 * r + cc] = K[4 * r + 0] * invTrack[cc] + K[4 * r + 1] * invTrack[4 + cc] + K[4 * r + 2] * invTrack[8 + cc];
		q.tz[0] = invTrack[8]; q.tz[1] = invTrack[9]; q.tz[2] = invTrack[10];
		q.q_mixed = c->d_qmixed; q.q_free = c->d_qfree;
		q.q_replay = c->d_qreplay;
		q.ckpt = c->d_ckpt; q.ckpt_cap = c->ckpt_cap;
		q.ready = c->d_ready; q.seq = (unsigned int) (c->int_launches & 0xffffffffu);   // already incremented: >= 1
		q.ctr = c->d_q2ctr + 4 * qslot; q.ctr_next = c->d_q2ctr + 4 * (qslot ^ 1);
		dim3 block(32, 8), grid(q.bnx, (q.bny + 7) / 8);
		k_integrate_plan2<<<grid, block, 0, c->stream>>>(q);
		LAUNCHED(c);
		k_integrate_free_runs<<<dim3(q.bny, (((p.z_end + 7) >> 3) - (p.z_begin >> 3) + 7) / 8), dim3(32, 8), 0, c->stream>>>(q);
		LAUNCHED(c);
		k_integrate_run2<<<c->int_grid2, 256, 0, c->stream>>>(q);
		LAUNCHED(c);
	}
	CK(cudaGetLastError());
	c->integrate_count++;
	if (!dev) c->st.frames_integrated++;   // device-gated launches are counted when the host learns the gate
	c->overlap_ok = c->overlap_enabled && !c->window_late;   // until anything but the raycast is enqueued
	return 0;
}

int kfb_k_integrate(kfb_ctx* c, const float invTrack[16], const float K[16], float mu, float maxweight) {
	CK(cudaSetDevice(c->device));
	timer_begin(c, c->t_int, 4u);
	int rc = launch_integrate(c, invTrack, K, mu, maxweight);
	timer_end(c, c->t_int, 4u);
	return rc;
}

int kfb_integrate(kfb_ctx* c, const float k[4], uint32_t integration_rate, float mu, uint32_t frame, int* integrated) {
	if (!c || !k) return set_err(KFB_E_ARG, "null argument");
	if (integration_rate == 0) return set_err(KFB_E_ARG, "integration_rate must be > 0");
	CK(cudaSetDevice(c->device));
	int doIntegrate = hm_check_pose(c->pose, c->oldPose, c->reduction, c->cw, c->ch, c_track_threshold);   // cpp/kernels.cpp:991
	if ((doIntegrate && ((frame % integration_rate) == 0)) || (frame <= 3)) {                               // :994
		float inv[16], K[16];
		hm_inverse4(inv, c->pose);
		hm_camera_matrix(K, k);
		timer_begin(c, c->t_int, 4u);
		int rc = launch_integrate(c, inv, K, mu, c_maxweight);                                             // :995-996
		timer_end(c, c->t_int, 4u);
		if (rc) return rc;
		doIntegrate = 1;
	} else doIntegrate = 0;
	if (integrated) *integrated = doIntegrate;
	// z-slab group: every slab (and every flag stored into a peer) is complete before anybody raycasts through it.  All
	// ranks hold the same pose and flags, so all of them come through here, integrated or not.
	if (c->peer_mode) return kfb_peer_barrier(c);
	return 0;
}

// ---------------------------------------------------------------------------- raycasting
static int launch_raycast(kfb_ctx* c, const float* view, float nearP, float farP, float step, float largestep, const float* view_dev = nullptr) {
	RaycastParams p;
	p.view_dev = view_dev;
	p.vol = c->view_all;
	{
		const uint32_t cap = c->ray_tiles_cap;
		const uint64_t L64 = c->ray_launches + 1;
		const unsigned int L = (unsigned int) L64, cur = (unsigned int) (L64 % 3), nxt = (unsigned int) ((L64 + 1) % 3), zero = (unsigned int) ((L64 + 2) % 3);
		unsigned int* base = c->d_tile_cost;
		unsigned int* ctr = base + (size_t) cap * 5;   // n[3], sum[3]
		p.sched.cost = base; p.sched.stamp = base + cap;
		p.sched.slow_cur = base + (size_t) cap * (2 + cur); p.sched.slow_next = base + (size_t) cap * (2 + nxt);
		p.sched.n_cur = ctr + cur; p.sched.n_next = ctr + nxt; p.sched.n_zero = ctr + zero;
		p.sched.sum_cur = ctr + 3 + cur; p.sched.sum_next = ctr + 3 + nxt; p.sched.sum_zero = ctr + 3 + zero;
		p.sched.launch = L; p.sched.enabled = c->ray_sched;
	}
	p.vertex = c->d_vertex; p.normal = c->d_normal;
	p.w = c->cw; p.h = c->ch;
	p.row0 = c->band0; p.row1 = c->band1;
	p.view = toMat(view);
	p.nearPlane = nearP; p.farPlane = farP; p.step = step; p.largestep = largestep;
	const int slot = (int) (c->ray_launches++ & 1);
	p.tile_next = c->d_tile_ctr + slot; p.tile_reset = c->d_tile_ctr + (slot ^ 1);
	// window for the next frame's preprocessing: already open when this raycast directly follows an integrate
	if (c->overlap_enabled && !c->overlap_ok) CK(cudaEventRecord(c->ev_window, c->stream));
	if (c->ray_bulk) k_raycast_bulk<<<c->ray_bulk_grid, RCK_BX * RCK_BY, 0, c->stream>>>(p, c->d_bulk_stats);
	else k_raycast<<<c->ray_grid, RCK_BX * RCK_BY, 0, c->stream>>>(p);
	LAUNCHED(c);
	if (c->peer_mode && c->world > 1) {
		// the band this rank just wrote goes to every peer's maps (bands are whole rows: one contiguous block per map)
		const size_t off = (size_t) c->band0 * c->cw * 3 * sizeof(float), bytes = (size_t) (c->band1 - c->band0) * c->cw * 3 * sizeof(float);
		if (off % 16 || bytes % 16) return set_err(KFB_E_STATE, "pixel-row band is not 16-byte aligned (%u columns)", c->cw);
		BandPushParams bp;
		bp.vertex = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(c->d_vertex) + off);
		bp.normal = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(c->d_normal) + off);
		bp.n_peer = 0;
		for (int r = 0; r < c->world; ++r) if (r != c->rank) {
			bp.peer_vertex[bp.n_peer] = reinterpret_cast<uint4*>(reinterpret_cast<char*>(c->peer_vertex[r]) + off);
			bp.peer_normal[bp.n_peer] = reinterpret_cast<uint4*>(reinterpret_cast<char*>(c->peer_normal[r]) + off);
			++bp.n_peer;
		}
		bp.n16 = (uint32_t) (bytes / 16);
		if (bp.n16) { k_band_push<<<148, 256, 0, c->stream>>>(bp); LAUNCHED(c); }
	}
	CK(cudaGetLastError());
	c->overlap_ok = c->overlap_enabled;   // until anything else is enqueued
	return 0;
}

int kfb_k_raycast(kfb_ctx* c, const float view[16], float nearP, float farP, float step, float largestep) {
	CK(cudaSetDevice(c->device));
	timer_begin(c, c->t_ray, 8u);
	int rc = launch_raycast(c, view, nearP, farP, step, largestep);
	timer_end(c, c->t_ray, 8u);
	return rc;
}

int kfb_raycast(kfb_ctx* c, const float k[4], float mu, uint32_t frame) {
	if (!c || !k) return set_err(KFB_E_ARG, "null argument");
	CK(cudaSetDevice(c->device));
	if (frame > 2) {                                               // cpp/kernels.cpp:977
		memcpy(c->raycastPose, c->pose, sizeof c->pose);           // :978
		float invK[16], view[16];
		hm_inverse_camera_matrix(invK, k);
		hm_matmul4(view, c->raycastPose, invK);
		timer_begin(c, c->t_ray, 8u);
		int rc = launch_raycast(c, view, c_nearPlane, c_farPlane, c->step, 0.75f * mu);   // :979-981
		// z-slab group: all bands have landed in this rank's maps before the next ICP reads them, and nobody integrates the
		// next frame into a slab a peer's rays are still reading
		if (!rc && c->peer_mode) rc = kfb_peer_barrier(c);
		timer_end(c, c->t_ray, 8u);
		return rc;
	}
	return 0;
}

// track -> integrate -> raycast of Kfusion::computeFrame, after the preprocessing of either entry point
static int compute_frame_rest(kfb_ctx* c, const float k[4], uint32_t integration_rate, uint32_t tracking_rate, float icp_threshold, float mu,
		uint32_t frame, int* tracked, int* integrated) {
	int rc;
	// Kfusion::computeFrame (cpp/kernels.cpp:1048-1055) is the one entry point that knows the whole frame up front, so
	// the whole frame is ENQUEUED before the host has seen the pose: the ICP kernel's last CTA evaluates checkPose,
	// inverse(pose) and raycastPose * invK on the device (same routines as the host path: bit-identical) and the integrate
	// / raycast kernels read them from device memory; the host only waits for the ICP result (tracked, integrated, pose)
	// while integrate and raycast are already running.  Frames without tracking, the host-solve A/B mode, stage timers and
	// z-slab contexts (collectives of the caller between the stages) take the staged path.
	const bool whole = (frame % tracking_rate == 0) && !(c->cfg.flags & KFB_FLAG_ICP_HOST_SOLVE) && icp_total_iterations(c) > 0
			&& c->world == 1 && (c->timing & 3u) == 0 && !c->no_async;   // integrate / raycast timers are in-stream events: no sync
	if (!whole) {
		if ((rc = kfb_track(c, k, icp_threshold, tracking_rate, frame, tracked))) return rc;
		if ((rc = kfb_integrate(c, k, integration_rate, mu, frame, integrated))) return rc;
		return kfb_raycast(c, k, mu, frame);
	}
	CK(cudaSetDevice(c->device));
	if ((rc = launch_pyramid(c, k))) return rc;                     // :931-945
	memcpy(c->oldPose, c->pose, sizeof c->pose);                    // :947
	float K[16], invK[16], invRP[16], projectReference[16];
	hm_camera_matrix(K, k);
	hm_inverse_camera_matrix(invK, k);
	hm_inverse4(invRP, c->raycastPose);
	hm_matmul4(projectReference, K, invRP);                         // :948
	IcpTail tail;
	tail.out = c->d_frame;
	tail.K = toMat(K); tail.invK = toMat(invK);
	tail.track_threshold = c_track_threshold;
	tail.force_integrate = frame <= 3;                              // :994
	tail.rate_ok = (frame % integration_rate) == 0;
	uint32_t seq;
	if ((rc = launch_icp(c, icp_threshold, projectReference, &tail, &seq))) return rc;
	// integrate: launched unconditionally, gated on the device (DevFrame::do_integrate)
	timer_begin(c, c->t_int, 4u);
	rc = launch_integrate(c, c->pose /* unused */, K, mu, c_maxweight, c->d_frame);
	timer_end(c, c->t_int, 4u);
	if (rc) return rc;
	if (frame > 2) {                                                // :977-981
		timer_begin(c, c->t_ray, 8u);
		rc = launch_raycast(c, c->pose /* unused */, c_nearPlane, c_farPlane, c->step, 0.75f * mu, c->d_frame->view);
		timer_end(c, c->t_ray, 8u);
		if (rc) return rc;
	}
	uint64_t iters = 0;
	if ((rc = collect_icp(c, seq, &iters))) return rc;              // pose is already the one checkPose left (oldPose if lost)
	c->ev_h2d_pending = false;
	c->st.icp_iterations_last = iters;
	c->st.icp_iterations_total += iters;
	const int ok = (int) *(volatile uint32_t*) (c->h_out32 + 65), did = (int) *(volatile uint32_t*) (c->h_out32 + 66);
	c->st.d2h_bytes += 2 * sizeof(uint32_t);
	if (did) c->st.frames_integrated++;
	if (frame > 2) memcpy(c->raycastPose, c->pose, sizeof c->pose); // :978
	if (tracked) *tracked = ok;
	if (integrated) *integrated = did;
	return 0;
}

int kfb_compute_frame(kfb_ctx* c, const uint16_t* depth, uint32_t iw, uint32_t ih, const float k[4], uint32_t integration_rate,
		uint32_t tracking_rate, float icp_threshold, float mu, uint32_t frame, int* tracked, int* integrated) {
	int rc;
	if (!c || !k) return set_err(KFB_E_ARG, "null argument");
	if (tracking_rate == 0) return set_err(KFB_E_ARG, "tracking_rate must be > 0");
	if (integration_rate == 0) return set_err(KFB_E_ARG, "integration_rate must be > 0");
	if ((rc = kfb_preprocess(c, depth, iw, ih))) return rc;
	return compute_frame_rest(c, k, integration_rate, tracking_rate, icp_threshold, mu, frame, tracked, integrated);
}
int kfb_compute_frame_device(kfb_ctx* c, const uint16_t* d_depth, uint32_t iw, uint32_t ih, const float k[4], uint32_t integration_rate,
		uint32_t tracking_rate, float icp_threshold, float mu, uint32_t frame, int* tracked, int* integrated) {
	int rc;
	if (!c || !k) return set_err(KFB_E_ARG, "null argument");
	if (tracking_rate == 0) return set_err(KFB_E_ARG, "tracking_rate must be > 0");
	if (integration_rate == 0) return set_err(KFB_E_ARG, "integration_rate must be > 0");
	if ((rc = kfb_preprocess_device(c, d_depth, iw, ih))) return rc;
	return compute_frame_rest(c, k, integration_rate, tracking_rate, icp_threshold, mu, frame, tracked, integrated);
}

// ------------------------------------------------------------------------------- renders
static int ensure_render(kfb_ctx* c, size_t bytes) {
	if (c->render_bytes >= bytes) return 0;
	if (c->d_render) CK(cudaFree(c->d_render));
	CK(cudaMalloc(&c->d_render, bytes));
	c->render_bytes = bytes;
	return 0;
}
static int finish_render(kfb_ctx* c, uint8_t* out, size_t bytes) {
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(out, c->d_render, bytes, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	c->st.d2h_bytes += bytes;
	return 0;
}
int kfb_render_depth(kfb_ctx* c, uint8_t* out, uint32_t w, uint32_t h) {
	if (w != c->cw || h != c->ch) return set_err(KFB_E_ARG, "render size must equal the computation size");
	CK(cudaSetDevice(c->device));
	const uint32_t n = w * h;
	int rc = ensure_render(c, (size_t) n * 4);
	if (rc) return rc;
	k_render_depth<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_render, c->d_floatDepth, n, c_nearPlane, c_farPlane);
	LAUNCHED(c);
	return finish_render(c, out, (size_t) n * 4);
}
int kfb_render_track(kfb_ctx* c, uint8_t* out, uint32_t w, uint32_t h) {
	if (w != c->cw || h != c->ch) return set_err(KFB_E_ARG, "render size must equal the computation size");
	CK(cudaSetDevice(c->device));
	// the status plane is only maintained on request; the first call switches it on for later frames
	c->cfg.flags |= KFB_FLAG_TRACK_STATUS;
	const uint32_t n = w * h;
	int rc = ensure_render(c, (size_t) n * 4);
	if (rc) return rc;
	k_render_track<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_render, c->d_status, n);
	LAUNCHED(c);
	return finish_render(c, out, (size_t) n * 4);
}
int kfb_render_volume(kfb_ctx* c, uint8_t* out, uint32_t w, uint32_t h, int frame, int rate, const float k[4], float largestep,
		const float view_pose[16]) {
	if (rate <= 0) return set_err(KFB_E_ARG, "rendering rate must be > 0");
	if (frame % rate != 0) return 0;                               // cpp/kernels.cpp:1034
	CK(cudaSetDevice(c->device));
	const uint32_t n = w * h;
	int rc = ensure_render(c, (size_t) n * 4);
	if (rc) return rc;
	RenderVolumeParams p;
	p.vol = c->view_all;
	p.out = c->d_render; p.w = w; p.h = h;
	float invK[16], view[16];
	hm_inverse_camera_matrix(invK, k);
	hm_matmul4(view, view_pose ? view_pose : c->pose, invK);       // :1036
	p.view = toMat(view);
	p.nearPlane = c_nearPlane; p.farPlane = c_farPlane * 2.0f; p.step = c->step; p.largestep = largestep;
	p.light = make_float3(1, 1, -1.0f); p.ambient = make_float3(0.1f, 0.1f, 0.1f);   // constant_parameters.h:25-26
	dim3 block(RC_BX, RC_BY), grid((w + RC_BX - 1) / RC_BX, (h + RC_BY - 1) / RC_BY);
	k_render_volume<<<grid, block, 0, c->stream>>>(p);
	LAUNCHED(c);
	return finish_render(c, out, (size_t) n * 4);
}

int kfb_dump_volume(kfb_ctx* c, const char* path) {
	if (!path) return 0;                                           // cpp/kernels.cpp:1010
	CK(cudaSetDevice(c->device));
	printf("Dumping the volumetric representation on file: %s\n", path);
	// a z-slab context writes its slices at their place in the file (tsdf shorts, z order): the ranks of a group fill one dump
	// (created without truncation there: whichever rank comes first must not wipe what another has written)
	const int fd = open(path, c->world > 1 ? (O_CREAT | O_WRONLY) : (O_CREAT | O_WRONLY | O_TRUNC), 0644);
	FILE* f = fd >= 0 ? fdopen(fd, "wb") : nullptr;
	if (!f) { if (fd >= 0) close(fd); return set_err(KFB_E_ARG, "Error opening file: %s", path); }
	if (c->world > 1 && fseeko(f, (off_t) ((size_t) c->cfg.volume_res[0] * c->cfg.volume_res[1] * c->z0 * sizeof(short)), SEEK_SET) != 0) {
		fclose(f);
		return set_err(KFB_E_ARG, "cannot seek in %s", path);
	}
	// stream the slab out plane-group by plane-group: tsdf shorts only, x fastest
	const size_t plane = (size_t) c->cfg.volume_res[0] * c->cfg.volume_res[1];
	const uint32_t nz = c->z1 - c->z0;
	const uint32_t zb = (uint32_t) ((((size_t) 64 << 20) / (plane * sizeof(short))) ? (((size_t) 64 << 20) / (plane * sizeof(short))) : 1);
	short* d_tmp = nullptr;
	cudaError_t e = cudaMalloc(&d_tmp, plane * zb * sizeof(short));
	std::vector<short> h(plane * zb);
	for (uint32_t z = 0; e == cudaSuccess && z < nz; z += zb) {
		const uint32_t nzb = (z + zb <= nz) ? zb : nz - z;
		const size_t n = plane * nzb;
		k_extract_tsdf<<<148 * 4, 256, 0, c->stream>>>(d_tmp, c->d_vol + plane * z, n);
		LAUNCHED(c);
		e = cudaMemcpyAsync(h.data(), d_tmp, n * sizeof(short), cudaMemcpyDeviceToHost, c->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
		if (e == cudaSuccess) fwrite(h.data(), sizeof(short), n, f);
	}
	cudaFree(d_tmp);
	fclose(f);
	if (e != cudaSuccess) return set_err(KFB_E_CUDA, "kfb_dump_volume: %s", cudaGetErrorString(e));
	return 0;
}

// ------------------------------------------------------------------------ buffer access
static int resolve_buffer(kfb_ctx* c, int which, int level, void** ptr, size_t* bytes, bool* host) {
	*host = false;
	const size_t P = (size_t) c->cw * c->ch;
	if ((which == KFB_BUF_SCALEDDEPTH || which == KFB_BUF_INVERTEX || which == KFB_BUF_INNORMAL) && (level < 0 || level >= c->levels))
		return set_err(KFB_E_ARG, "level %d out of range", level);
	const size_t PL = (which == KFB_BUF_SCALEDDEPTH || which == KFB_BUF_INVERTEX || which == KFB_BUF_INNORMAL) ? (size_t) c->lw[level] * c->lh[level] : 0;
	switch (which) {
	case KFB_BUF_VOLUME: *ptr = c->d_vol; *bytes = c->slab_voxels * sizeof(short2); break;
	case KFB_BUF_VERTEX: *ptr = c->d_vertex; *bytes = P * 12; break;
	case KFB_BUF_NORMAL: *ptr = c->d_normal; *bytes = P * 12; break;
	case KFB_BUF_FLOATDEPTH: *ptr = c->d_floatDepth; *bytes = P * 4; break;
	case KFB_BUF_SCALEDDEPTH: *ptr = c->d_scaled[level]; *bytes = PL * 4; break;
	case KFB_BUF_INVERTEX: *ptr = c->d_inV[level]; *bytes = PL * 12; break;
	case KFB_BUF_INNORMAL: *ptr = c->d_inN[level]; *bytes = PL * 12; break;
	case KFB_BUF_REDUCTION: *ptr = c->reduction; *bytes = 32 * 4; *host = true; break;
	case KFB_BUF_TRACKSTATUS: *ptr = c->d_status; *bytes = P; break;
	case KFB_BUF_RAYCASTPOSE: *ptr = c->raycastPose; *bytes = 64; *host = true; break;
	case KFB_BUF_OLDPOSE: *ptr = c->oldPose; *bytes = 64; *host = true; break;
	case KFB_BUF_GAUSSIAN: *ptr = c->gaussian; *bytes = 20; *host = true; break;
	case KFB_BUF_INPUTDEPTH: *ptr = c->d_input; *bytes = c->input_bytes; break;
	case KFB_BUF_REDUCTION_DEV: *ptr = c->d_out32; *bytes = 32 * 4; break;
	case KFB_BUF_BRICKFLAGS:
		if (!c->brick.flag) return set_err(KFB_E_STATE, "brick flags are not maintained by this context");
		*ptr = c->brick.flag; *bytes = (size_t) c->brick.bnx * c->brick.bny * c->brick.bnz; break;
	case KFB_BUF_RAYTILECOST:
		*ptr = c->d_tile_cost; *bytes = (size_t) c->ray_tiles_cap * sizeof(unsigned int); break;
	case KFB_BUF_BRICKCLASS: *ptr = c->d_cls; *bytes = c->cls_bytes; break;
	default: return set_err(KFB_E_ARG, "unknown buffer %d", which);
	}
	return 0;
}
int kfb_buffer_bytes(kfb_ctx* c, int which, int level, size_t* bytes) {
	void* p; bool host;
	return resolve_buffer(c, which, level, &p, bytes, &host);
}
int kfb_device_ptr(kfb_ctx* c, int which, int level, void** dev_ptr) {
	size_t b; bool host;
	int rc = resolve_buffer(c, which, level, dev_ptr, &b, &host);
	if (rc) return rc;
	if (host) return set_err(KFB_E_ARG, "buffer %d lives on the host", which);
	return 0;
}
int kfb_read_buffer(kfb_ctx* c, int which, int level, void* dst, size_t bytes) {
	void* p; size_t b; bool host;
	int rc = resolve_buffer(c, which, level, &p, &b, &host);
	if (rc) return rc;
	if (bytes > b) return set_err(KFB_E_ARG, "read of %zu bytes from a %zu-byte buffer", bytes, b);
	CK(cudaSetDevice(c->device));
	if (host) { memcpy(dst, p, bytes); return 0; }
	CK(cudaStreamSynchronize(c->stream));
	CK(cudaMemcpy(dst, p, bytes, cudaMemcpyDeviceToHost));
	return 0;
}
int kfb_write_buffer(kfb_ctx* c, int which, int level, const void* src, size_t bytes) {
	void* p; size_t b; bool host;
	int rc = resolve_buffer(c, which, level, &p, &b, &host);
	if (rc) return rc;
	if (bytes > b) return set_err(KFB_E_ARG, "write of %zu bytes into a %zu-byte buffer", bytes, b);
	CK(cudaSetDevice(c->device));
	if (host) { memcpy(p, src, bytes); return 0; }
	if (which == KFB_BUF_FLOATDEPTH) { c->dmax_slot = -1; c->mip_valid[c->fd_cur] = false; }   // the cached max / pyramid no longer describe this image
	CK(cudaStreamSynchronize(c->stream));
	CK(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
	if (which == KFB_BUF_VOLUME && c->brick.flag) {   // the flags must describe the volume the raycaster will read
		CK(cudaMemsetAsync(c->brick.flag, 0, brick_bytes(c), c->stream));
		k_brick_rebuild<<<148 * 8, 256, 0, c->stream>>>(c->brick, c->d_vol, c->cfg.volume_res[0], c->cfg.volume_res[1], c->z1 - c->z0, c->z0);
		LAUNCHED(c);
		CK(cudaGetLastError());
	}
	return 0;
}

// -------------------------------------------------------------------------- measurement
int kfb_enable_timing(kfb_ctx* c, int mask) { c->timing = (uint32_t) mask; return 0; }
int kfb_reset_stats(kfb_ctx* c) {
	CK(cudaSetDevice(c->device));
	CK(cudaStreamSynchronize(c->stream));
	timer_resolve(c, c->t_pre); timer_resolve(c, c->t_track); timer_resolve(c, c->t_int); timer_resolve(c, c->t_ray);
	c->t_pre.total_ms = c->t_track.total_ms = c->t_int.total_ms = c->t_ray.total_ms = 0;
	c->t_pre.count = c->t_track.count = c->t_int.count = c->t_ray.count = 0;
	memset(&c->st, 0, sizeof c->st);
	c->integrate_count = 0;
	c->nupd_folded = 0;
	CK(cudaMemset(c->d_nupd, 0, NUPD_SLOTS * sizeof(unsigned long long)));
	return 0;
}
int kfb_get_stats(kfb_ctx* c, kfb_stats* out) {
	CK(cudaSetDevice(c->device));
	CK(cudaStreamSynchronize(c->stream));
	timer_resolve(c, c->t_pre); timer_resolve(c, c->t_track); timer_resolve(c, c->t_int); timer_resolve(c, c->t_ray);
	std::vector<unsigned long long> h(NUPD_SLOTS);
	CK(cudaMemcpy(h.data(), c->d_nupd, NUPD_SLOTS * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
	unsigned long long tot = 0;
	for (auto v : h) tot += v;
	c->st.voxels_updated_total = tot + c->nupd_folded;
	c->st.voxels_updated_last = c->integrate_count ? h[(c->integrate_count - 1) % NUPD_SLOTS] : 0;
	// totals over all calls since reset_stats (ms_* hold the SUM; divide by the call counts yourself)
	c->st.ms_preprocess = (float) c->t_pre.total_ms;
	c->st.ms_track = (float) c->t_track.total_ms;
	c->st.ms_integrate = (float) c->t_int.total_ms;
	c->st.ms_raycast = (float) c->t_ray.total_ms;
	*out = c->st;
	return 0;
}

// ---------------------------------------------------------------------------- multi-GPU
int kfb_slab_ipc_handle(kfb_ctx* c, uint8_t handle64[64]) {
	CK(cudaSetDevice(c->device));
	cudaIpcMemHandle_t h;
	CK(cudaIpcGetMemHandle(&h, c->d_vol));
	static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
	memcpy(handle64, &h, 64);
	return 0;
}
int kfb_slab_import(kfb_ctx* c, int rank, int world, const uint8_t* handles64, const uint32_t* z_begin) {
	if (world < 1 || world > KFB_MAX_SLABS || rank < 0 || rank >= world) return set_err(KFB_E_ARG, "bad rank/world %d/%d", rank, world);
	CK(cudaSetDevice(c->device));
	c->rank = rank; c->world = world;
	c->view_all.n_slabs = world;
	// z-slab mode interleaves collectives of the caller on this stream; its numbers were measured without the side-stream
	// overlap, and the overlap has not been validated on several GPUs: keep those contexts strictly serial
	if (world > 1) { c->overlap_enabled = false; c->overlap_ok = false; c->side_pending = false; }
	// each rank flags only what ITS slices touch: without the caller's merge the raycaster must not skip
	if (!(c->cfg.flags & KFB_FLAG_BRICKS_MERGED)) c->view_all.brick = nullptr;
	c->view_all.super = nullptr;   // the caller merges the brick flags only (KFB_BUF_BRICKFLAGS): no coarse leaps over peers' slabs
	for (int r = 0; r < world; ++r) {
		c->view_all.slab_z[r] = z_begin[r];
		if (r == rank) { c->view_all.slab_ptr[r] = c->d_vol; continue; }
		cudaIpcMemHandle_t h;
		memcpy(&h, handles64 + 64 * (size_t) r, 64);
		void* p = nullptr;
		CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
		c->peer_ptrs[r] = p;
		c->view_all.slab_ptr[r] = (const short2*) p;
	}
	c->view_all.slab_z[world] = c->cfg.volume_res[2];
	if (z_begin[rank] != c->z0) return set_err(KFB_E_ARG, "z_begin[%d]=%u does not match this context's slab start %u", rank, z_begin[rank], c->z0);
	return 0;
}
int kfb_ipc_export(kfb_ctx* c, kfb_ipc_handles* out) {
	if (!c || !out) return set_err(KFB_E_ARG, "null argument");
	CK(cudaSetDevice(c->device));
	memset(out, 0, sizeof *out);
	cudaIpcMemHandle_t h;
	static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
	CK(cudaIpcGetMemHandle(&h, c->d_vol)); memcpy(out->volume, &h, 64);
	CK(cudaIpcGetMemHandle(&h, c->d_vertex)); memcpy(out->vertex, &h, 64);
	CK(cudaIpcGetMemHandle(&h, c->d_normal)); memcpy(out->normal, &h, 64);
	CK(cudaIpcGetMemHandle(&h, c->d_sync)); memcpy(out->sync, &h, 64);
	if (c->brick.flag) { CK(cudaIpcGetMemHandle(&h, c->brick.flag)); memcpy(out->bricks, &h, 64); out->has_bricks = 1; }
	out->slab_z0 = c->z0; out->slab_z1 = c->z1;
	return 0;
}
int kfb_ipc_import(kfb_ctx* c, int rank, int world, const kfb_ipc_handles* all) {
	if (!c || !all) return set_err(KFB_E_ARG, "null argument");
	if (world < 1 || world > KFB_MAX_SLABS || rank < 0 || rank >= world) return set_err(KFB_E_ARG, "bad rank/world %d/%d", rank, world);
	if (all[rank].slab_z0 != c->z0 || all[rank].slab_z1 != c->z1) return set_err(KFB_E_ARG, "handles[%d] do not describe this context's slab", rank);
	for (int r = 0; r + 1 < world; ++r)
		if (all[r].slab_z1 != all[r + 1].slab_z0) return set_err(KFB_E_ARG, "slabs %d and %d are not contiguous in z", r, r + 1);
	if (all[0].slab_z0 != 0 || all[world - 1].slab_z1 != c->cfg.volume_res[2]) return set_err(KFB_E_ARG, "the slabs do not cover the volume");
	CK(cudaSetDevice(c->device));
	c->rank = rank; c->world = world;
	c->view_all.n_slabs = world;
	if (world > 1) { c->overlap_enabled = false; c->overlap_ok = false; c->side_pending = false; }
	bool all_bricks = c->brick.flag != nullptr;
	for (int r = 0; r < world; ++r) all_bricks = all_bricks && all[r].has_bricks;
	if (!all_bricks) { c->view_all.brick = nullptr; c->view_all.super = nullptr; }   // without every rank's flags the raycaster must not skip
	auto open = [&](const uint8_t* h64, void** out) -> int {
		cudaIpcMemHandle_t h;
		memcpy(&h, h64, 64);
		CK(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
		return 0;
	};
	for (int r = 0; r < world; ++r) {
		c->view_all.slab_z[r] = all[r].slab_z0;
		if (r == rank) {
			c->view_all.slab_ptr[r] = c->d_vol;
			c->sync_all.p[r] = c->d_sync;
			continue;
		}
		int rc;
		void* p = nullptr;
		if ((rc = open(all[r].volume, &p))) return rc;
		c->peer_ptrs[r] = p; c->view_all.slab_ptr[r] = (const short2*) p;
		if ((rc = open(all[r].vertex, &c->peer_open[r][0]))) return rc;
		if ((rc = open(all[r].normal, &c->peer_open[r][1]))) return rc;
		if (all_bricks && (rc = open(all[r].bricks, &c->peer_open[r][2]))) return rc;
		if ((rc = open(all[r].sync, &c->peer_open[r][3]))) return rc;
		c->peer_vertex[r] = (float*) c->peer_open[r][0]; c->peer_normal[r] = (float*) c->peer_open[r][1];
		c->peer_bricks[r] = (unsigned char*) c->peer_open[r][2];
		c->sync_all.p[r] = (PeerSync*) c->peer_open[r][3];
	}
	c->view_all.slab_z[world] = c->cfg.volume_res[2];
	c->peer_mode = world > 1;
	c->barrier_count = 0;
	return 0;
}
int kfb_peer_barrier(kfb_ctx* c) {
	if (!c) return set_err(KFB_E_ARG, "null ctx");
	if (!c->peer_mode) return 0;
	CK(cudaSetDevice(c->device));
	k_peer_barrier<<<1, KFB_MAX_SLABS, 0, c->stream>>>(c->sync_all, c->rank, c->world, ++c->barrier_count, (volatile unsigned int*) (c->h_out32_dev + 34));
	LAUNCHED(c);
	CK(cudaGetLastError());
	return 0;
}
int kfb_set_pixel_rows(kfb_ctx* c, uint32_t row0, uint32_t row1) {
	if (row0 == 0 && row1 == 0) row1 = c->ch;
	if (row1 > c->ch || row0 >= row1) return set_err(KFB_E_ARG, "bad pixel rows [%u,%u) for a %u-row image", row0, row1, c->ch);
	c->band0 = row0; c->band1 = row1;
	return 0;
}

}  // extern "C"
