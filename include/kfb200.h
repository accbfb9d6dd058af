/* kfb200 — C ABI of the B200-native KinectFusion backend (libkfb200.so).
 *
 * This is the drop-in boundary for the per-frame hot path of domantasjurkus/slambench
 * (`class Kfusion`, kfusion/include/kernels.h:83-195).  The reference resolves its
 * backend at LINK time: each backend is a library defining the same out-of-line
 * `Kfusion::` members (kfusion/CMakeLists.txt:40-75, CMakeLists.txt:53-54).  Our
 * backend glue (slambench_b200/csrc/kfusion_b200.cpp) defines those members and
 * forwards each one to the entry point below that cites it; INTEGRATION.md shows the
 * exact binding.  Everything here is plain C: opaque handle, POD arguments, int error
 * codes (0 = ok; see kfb_last_error()).  The library owns all device memory; callers
 * own the host buffers they pass.
 *
 * Matrices are 16 floats, row-major, camera->world — the memory image of the
 * reference's `Matrix4` (commons.h:317-319).  `k` is (fx, fy, cx, cy) = the
 * reference's `float4 k`.
 */
#ifndef KFB200_H
#define KFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KFB_MAX_LEVELS 8
#define KFB_ABI_VERSION 1

typedef struct kfb_ctx kfb_ctx;

/* flags for kfb_config.flags */
#define KFB_FLAG_ICP_HOST_SOLVE 0x1u   /* one k_track_reduce launch + host 6x6 solve per ICP iteration: the reference's control flow
                                         verbatim (A/B check); default = one persistent cooperative kernel per frame            */
#define KFB_FLAG_TRACK_STATUS   0x2u   /* keep the per-pixel ICP status plane that renderTrack visualises           */
#define KFB_FLAG_INTEGRATE_V1   0x4u   /* integrate with round 1's per-voxel-decision kernels instead of the brick-classified ones (A/B) */
#define KFB_FLAG_INTEGRATE_NO_CULL 0x8u /* integrate visits every voxel with the reference's full expression (A/B check) */
#define KFB_FLAG_RAYCAST_NO_SKIP 0x10u  /* raycast evaluates every sample (no brick flags) (A/B check)                    */
#define KFB_FLAG_BRICKS_MERGED 0x20u    /* z-slab mode: the caller merges (element-wise max over ranks) KFB_BUF_BRICKFLAGS between
                                          integrate and raycast, so the raycaster may skip over peers' free space too          */

typedef struct kfb_config {
	uint32_t compute_w, compute_h;      /* Kfusion ctor `inputSize` = computation size   kernels.h:99-101 */
	uint32_t volume_res[3];             /* volumeResolution                              kernels.h:104     */
	float volume_dim[3];                /* volumeDimensions (metres)                     kernels.h:103     */
	float init_pose[16];                /* initial camera->world pose                    kernels.h:105-109,122-127 */
	int32_t n_levels;                   /* pyramid.size()                                kernels.h:110-114 */
	int32_t iterations[KFB_MAX_LEVELS]; /* pyramid[level] ICP iterations, level 0 = full res */
	int32_t device;                     /* CUDA device ordinal                                               */
	uint32_t slab_z0, slab_z1;          /* z-slab owned by this context [z0, z1); 0,0 = whole volume        */
	uint32_t flags;
} kfb_config;

typedef struct kfb_stats {
	uint64_t kernel_launches;     /* kernels of THIS library launched since create/reset_stats               */
	uint64_t frames_integrated;
	uint64_t voxels_updated_last; /* N_upd of the last integrate: voxels for which the reference executes vol.set */
	uint64_t voxels_updated_total;
	uint64_t icp_iterations_last;
	uint64_t icp_iterations_total;
	uint64_t h2d_bytes, d2h_bytes;
	float ms_preprocess, ms_track, ms_integrate, ms_raycast; /* CUDA-event time of each stage SUMMED over all calls since
	                                                            kfb_reset_stats (needs kfb_enable_timing); divide by the call counts */
} kfb_stats;

/* buffers addressable through kfb_read_buffer / kfb_write_buffer (reference global of the same role,
 * kfusion/src/cpp/kernels.cpp:39-55) */
enum kfb_buffer {
	KFB_BUF_VOLUME = 0,       /* short2[N^3] (this context's slab), x fastest            `volume`           */
	KFB_BUF_VERTEX = 1,       /* float3[P] raycast vertex map, world frame                `vertex`           */
	KFB_BUF_NORMAL = 2,       /* float3[P]                                                `normal`           */
	KFB_BUF_FLOATDEPTH = 3,   /* float[P] raw metres                                      `floatDepth`       */
	KFB_BUF_SCALEDDEPTH = 4,  /* float[P >> 2*level] bilateral-filtered pyramid           `ScaledDepth[l]`   */
	KFB_BUF_INVERTEX = 5,     /* float3[P >> 2*level]                                     `inputVertex[l]`   */
	KFB_BUF_INNORMAL = 6,     /* float3[P >> 2*level]                                     `inputNormal[l]`   */
	KFB_BUF_REDUCTION = 7,    /* float[32]: row 0 of the reference's 8x32 `reductionoutput` after its final add */
	KFB_BUF_TRACKSTATUS = 8,  /* int8[P]: TrackData.result per pixel (only with KFB_FLAG_TRACK_STATUS)       */
	KFB_BUF_RAYCASTPOSE = 9,  /* float[16] (host)                                         `raycastPose`      */
	KFB_BUF_OLDPOSE = 10,     /* float[16] (host)                                         `oldPose`          */
	KFB_BUF_GAUSSIAN = 11,    /* float[5]                                                 `gaussian`         */
	KFB_BUF_INPUTDEPTH = 12,  /* uint16[in_w*in_h] device copy of the last sensor frame                      */
	KFB_BUF_REDUCTION_DEV = 13, /* float[32] DEVICE copy of the last track+reduce result (multi-GPU all-reduce operand) */
	KFB_BUF_BRICKFLAGS = 14,  /* uint8[ceil(N/8)^3] brick flags of the WHOLE volume (see KFB_FLAG_BRICKS_MERGED)            */
	KFB_BUF_BRICKCLASS = 16,  /* uint8[ceil(slab/8)][ceil(N/8)][ceil(N/8)] classes of the LAST integrate's bricks: 0 skip, 1 free (sdf == 1), 2 per-voxel */
	KFB_BUF_RAYTILECOST = 15  /* uint32[ceil(h/4)][ceil(w/8)] SM cycles per raycast tile, last launch                          */
};

int kfb_abi_version(void);
const char* kfb_last_error(void);

/* Kfusion::languageSpecificConstructor()  kernels.h:140, cpp/kernels.cpp:67-112 (allocations, gaussian, volume init) */
int kfb_create(const kfb_config* cfg, kfb_ctx** out);
/* Kfusion::~Kfusion()                      cpp/kernels.cpp:114-134 */
int kfb_destroy(kfb_ctx* ctx);
/* Kfusion::reset() -> initVolumeKernel     cpp/kernels.cpp:135-157 */
int kfb_reset(kfb_ctx* ctx);

/* Kfusion::preprocessing(const ushort*, uint2)   cpp/kernels.cpp:915-922.  `depth_mm` is a HOST buffer of the
 * sensor size; the H2D copy is part of the call.  A PAGEABLE buffer is copied into an internal pinned staging buffer
 * before the call returns (the caller may reuse it at once, like with the reference).  A PAGE-LOCKED buffer
 * (cudaHostAlloc / cudaHostRegister / kfb_register_host_buffer; asked from the driver on every call) is read by the copy
 * engine asynchronously: it may be reused after kfb_sync, or after the next kfb_track / kfb_compute_frame of this context
 * has returned (both wait for the copy, also on frames that are not tracked). */
int kfb_preprocess(kfb_ctx* ctx, const uint16_t* depth_mm, uint32_t in_w, uint32_t in_h);
/* Page-lock a host buffer the caller will pass to kfb_preprocess again and again (benchmark.cpp:103 mallocs ONE frame
 * buffer) so that its frames are DMA'd directly.  Opt-in: the library never registers a pointer behind the caller's back.
 * Unregister before freeing the buffer; kfb_destroy unregisters what is left. */
int kfb_register_host_buffer(kfb_ctx* ctx, const void* ptr, size_t bytes);
int kfb_unregister_host_buffer(kfb_ctx* ctx, const void* ptr);
/* same, with the sensor frame already resident in device memory (bench.py's HBM-resident arm) */
int kfb_preprocess_device(kfb_ctx* ctx, const uint16_t* dev_depth_mm, uint32_t in_w, uint32_t in_h);
/* Kfusion::tracking(float4 k, float icp_threshold, uint tracking_rate, uint frame)  cpp/kernels.cpp:924-971.
 * On return the host pose (kfb_get_pose) is final for this frame. */
int kfb_track(kfb_ctx* ctx, const float k[4], float icp_threshold, uint32_t tracking_rate, uint32_t frame, int* tracked);
/* Kfusion::integration(float4 k, uint integration_rate, float mu, uint frame)      cpp/kernels.cpp:988-1004 */
int kfb_integrate(kfb_ctx* ctx, const float k[4], uint32_t integration_rate, float mu, uint32_t frame, int* integrated);
/* Kfusion::raycasting(float4 k, float mu, uint frame)   cpp/kernels.cpp:973-986 (the reference always returns false) */
int kfb_raycast(kfb_ctx* ctx, const float k[4], float mu, uint32_t frame);
/* Kfusion::computeFrame(...)                            cpp/kernels.cpp:1048-1055.  The one entry point that knows the whole
 * frame up front: all four stages are enqueued before the host waits for anything (checkPose, inverse(pose) and
 * raycastPose * invK run in the ICP kernel's last CTA); it returns when pose / tracked / integrated are known, while
 * integrate and raycast may still be running.  KFB_NO_ASYNC=1 forces the staged calls instead (A/B). */
int kfb_compute_frame(kfb_ctx* ctx, const uint16_t* depth_mm, uint32_t in_w, uint32_t in_h, const float k[4],
		uint32_t integration_rate, uint32_t tracking_rate, float icp_threshold, float mu, uint32_t frame,
		int* tracked, int* integrated);

/* same, with the sensor frame already resident in device memory (bench.py's HBM-resident arm) */
int kfb_compute_frame_device(kfb_ctx* ctx, const uint16_t* dev_depth_mm, uint32_t in_w, uint32_t in_h, const float k[4],
		uint32_t integration_rate, uint32_t tracking_rate, float icp_threshold, float mu, uint32_t frame,
		int* tracked, int* integrated);

/* Kfusion::getPose() / the `pose` member          kernels.h:85-94,173-175 */
int kfb_get_pose(kfb_ctx* ctx, float pose[16]);
int kfb_set_pose(kfb_ctx* ctx, const float pose[16]);
/* synchroniseDevices()                             kernels.h:197, cuda/kernels.cu:950-952 */
int kfb_sync(kfb_ctx* ctx);

/* Kfusion::renderDepth / renderTrack / renderVolume   cpp/kernels.cpp:1032-1046 — `out_rgba` is a host uchar4[w*h] */
int kfb_render_depth(kfb_ctx* ctx, uint8_t* out_rgba, uint32_t w, uint32_t h);
int kfb_render_track(kfb_ctx* ctx, uint8_t* out_rgba, uint32_t w, uint32_t h);
int kfb_render_volume(kfb_ctx* ctx, uint8_t* out_rgba, uint32_t w, uint32_t h, int frame, int rate, const float k[4],
		float largestep, const float view_pose[16] /* NULL = current pose (kernels.h:176-181) */);
/* Kfusion::dumpVolume(const char*)                  cpp/kernels.cpp:1006-1030 (tsdf shorts only, x fastest) */
int kfb_dump_volume(kfb_ctx* ctx, const char* path);

/* ---- stage-level entry points: the reference's free kernel functions (kernels.h:18-69) run on this
 *      context's device buffers with explicit matrices — used for teacher-forced parity tests ---- */
/* halfSampleRobustImageKernel x (L-1), depth2vertexKernel + vertex2normalKernel x L   cpp/kernels.cpp:931-945 */
int kfb_k_pyramid(kfb_ctx* ctx, const float k[4]);
/* trackKernel + reduceKernel fused  cpp/kernels.cpp:956-961: out32 = row 0 of reductionoutput after :487-489 */
int kfb_k_track_reduce(kfb_ctx* ctx, int level, const float Ttrack[16], const float view[16], float dist_threshold,
		float normal_threshold, float out32[32]);
/* integrateKernel   cpp/kernels.cpp:628-673 */
int kfb_k_integrate(kfb_ctx* ctx, const float inv_track[16], const float K[16], float mu, float maxweight);
/* raycastKernel     cpp/kernels.cpp:726-757 */
int kfb_k_raycast(kfb_ctx* ctx, const float view[16], float near_plane, float far_plane, float step, float largestep);
/* updatePoseKernel / checkPoseKernel (host side of the product)   cpp/kernels.cpp:759-792 */
int kfb_k_update_pose(float pose[16], const float reduction32[32], float icp_threshold, int* converged);
int kfb_k_check_pose(float pose[16], const float old_pose[16], const float reduction32[32], uint32_t w, uint32_t h,
		float track_threshold, int* ok);
/* host 4x4 helpers the stage drivers use (commons.h:343-378) */
void kfb_inverse4(float out[16], const float in[16]);
void kfb_matmul4(float out[16], const float a[16], const float b[16]);
void kfb_camera_matrix(float out[16], const float k[4]);
void kfb_inverse_camera_matrix(float out[16], const float k[4]);

/* ---- buffer access (tests, teacher forcing, multi-GPU gather) ---- */
int kfb_buffer_bytes(kfb_ctx* ctx, int which, int level, size_t* bytes);
int kfb_read_buffer(kfb_ctx* ctx, int which, int level, void* host_dst, size_t bytes);
int kfb_write_buffer(kfb_ctx* ctx, int which, int level, const void* host_src, size_t bytes);
/* raw device pointer of a buffer (for zero-copy views from the Python harness) */
int kfb_device_ptr(kfb_ctx* ctx, int which, int level, void** dev_ptr);

/* ---- measurement ---- */
int kfb_get_stats(kfb_ctx* ctx, kfb_stats* out);
int kfb_reset_stats(kfb_ctx* ctx);
/* CUDA-event timing of the stages; mask bits: 1 preprocess, 2 track, 4 integrate, 8 raycast (0 = off) */
#define KFB_TIME_PREPROCESS 1
#define KFB_TIME_TRACK 2
#define KFB_TIME_INTEGRATE 4
#define KFB_TIME_RAYCAST 8
#define KFB_TIME_ALL 15
int kfb_enable_timing(kfb_ctx* ctx, int mask);
/* the CUDA stream all of this context's kernels are launched on (cudaStream_t as void*) */
int kfb_stream(kfb_ctx* ctx, void** stream);

/* ---- multi-GPU z-slab mode: one context per GPU/process ---- */
/* export this context's slab as a CUDA IPC handle (64 bytes) */
int kfb_slab_ipc_handle(kfb_ctx* ctx, uint8_t handle64[64]);
/* import the peers' slabs: handles[r] / z_begin[r] for r in [0, world); own rank's entry is ignored */
int kfb_slab_import(kfb_ctx* ctx, int rank, int world, const uint8_t* handles64, const uint32_t* z_begin);
/* The same over PEER MEMORY only (no NCCL on the data path): besides its slab, every context exports its raycast maps, its
 * brick-flag map and a barrier slot.  After kfb_ipc_import a context
 *   - stores the brick flags of its slab into every peer's map while it integrates (nothing to merge afterwards),
 *   - copies its band of the raycast vertex / normal maps into every peer's maps right behind k_raycast (the all-gather, as one
 *     kernel of 16-byte stores over NVLink; the band — whole rows — must start and end on 16 bytes: width * 12 % 16 == 0),
 *   - ends kfb_integrate and kfb_raycast with a stream-ordered barrier over the group (kfb_peer_barrier),
 * so a host only has to move the 64-byte handles between its processes once (files, pipes, MPI, torch.distributed ...)
 * and then drives every rank with the ordinary stage calls.  All ranks must issue the same calls in the same order. */
typedef struct kfb_ipc_handles {
	uint8_t volume[64], vertex[64], normal[64], bricks[64], sync[64];
	uint32_t has_bricks;          /* 0: this context keeps no brick flags (KFB_FLAG_RAYCAST_NO_SKIP) */
	uint32_t slab_z0, slab_z1;    /* the slab the handles belong to */
	uint32_t reserved;
} kfb_ipc_handles;
int kfb_ipc_export(kfb_ctx* ctx, kfb_ipc_handles* out);
int kfb_ipc_import(kfb_ctx* ctx, int rank, int world, const kfb_ipc_handles* all /* [world], rank order */);
/* stream-ordered barrier over the group (needed explicitly only around the stage-level kfb_k_* calls) */
int kfb_peer_barrier(kfb_ctx* ctx);
/* rows [row0, row1) of the computation image this context is responsible for in kfb_raycast / kfb_k_raycast and in
 * the stage-level kfb_k_track_reduce (level l uses [row0 >> l, row1 >> l)); (0, 0) = the whole image.  The host
 * harness all-gathers the raycast bands and all-reduces the 32 partial sums (KFB_BUF_REDUCTION_DEV) over NCCL. */
int kfb_set_pixel_rows(kfb_ctx* ctx, uint32_t row0, uint32_t row1);

#ifdef __cplusplus
}
#endif
#endif /* KFB200_H */
