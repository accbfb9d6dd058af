// TEST INFRASTRUCTURE — never linked into or called by the product path.
//
// extern "C" access to the UNMODIFIED reference C++ backend
// (/root/reference/kfusion/src/cpp/kernels.cpp, compiled where it lies by
// oracle/Makefile together with this file and the TooN stand-in) so that
//   * oracle/kfusion_oracle.c (our plain-C restatement) can be pinned against it,
//   * golden vectors under tests/golden/ can be generated from it,
//   * bench.py can time it as the CPU baseline (`cpu_baseline.kind = "reference"`).
// The exported names are the same `kfo_*` set that oracle/kfusion_oracle.c
// exports, so one Python wrapper drives either library.
//
// The reference keeps all state in file-scope globals (cpp/kernels.cpp:39-55),
// hence one Kfusion per process; kfo_kf_create() enforces that.
#include <kernels.h>
#include <cstring>
#include <vector>

// globals defined (with external linkage) in the reference translation unit
extern float* gaussian;
extern Volume volume;
extern float3* vertex;
extern float3* normal;
extern TrackData* trackingResult;
extern float* reductionoutput;
extern float** ScaledDepth;
extern float* floatDepth;
extern Matrix4 oldPose;
extern Matrix4 raycastPose;
extern float3** inputVertex;
extern float3** inputNormal;

static Matrix4 toM(const float* m) {
	Matrix4 r;
	std::memcpy(&r, m, sizeof(float) * 16);
	return r;
}
static Volume toV(short* data, const unsigned* size, const float* dim) {
	Volume v;
	v.size = make_uint3(size[0], size[1], size[2]);
	v.dim = make_float3(dim[0], dim[1], dim[2]);
	v.data = (short2*) data;
	return v;
}

static Kfusion* g_kf = NULL;
static std::vector<int> g_pyramid;
static uint2 g_csize;

extern "C" {

const char* kfo_impl_name() {
#ifdef _OPENMP
	return "reference-openmp";
#else
	return "reference-cpp";
#endif
}

// ------------------------------------------------------------- free kernels
void kfo_init_volume(short* data, const unsigned* size, const float* dim) {
	initVolumeKernel(toV(data, size, dim));
}
void kfo_mm2meters(float* out, unsigned ow, unsigned oh, const unsigned short* in, unsigned iw, unsigned ih) {
	mm2metersKernel(out, make_uint2(ow, oh), in, make_uint2(iw, ih));
}
void kfo_gaussian(float* out5) {
	// cpp/kernels.cpp:101-107 lives inside languageSpecificConstructor; same expression here
	for (unsigned int i = 0; i < (unsigned) (radius * 2 + 1); i++) {
		int x = i - 2;
		out5[i] = expf(-(x * x) / (2 * delta * delta));
	}
}
void kfo_bilateral(float* out, const float* in, unsigned w, unsigned h, const float* gauss, float e_d, int r) {
	bilateralFilterKernel(out, in, make_uint2(w, h), gauss, e_d, r);
}
void kfo_halfsample(float* out, const float* in, unsigned iw, unsigned ih, float e_d, int r) {
	halfSampleRobustImageKernel(out, in, make_uint2(iw, ih), e_d, r);
}
void kfo_depth2vertex(float* vtx, const float* depth, unsigned w, unsigned h, const float* invK) {
	depth2vertexKernel((float3*) vtx, depth, make_uint2(w, h), toM(invK));
}
void kfo_vertex2normal(float* out, const float* in, unsigned w, unsigned h) {
	vertex2normalKernel((float3*) out, (const float3*) in, make_uint2(w, h));
}
void kfo_track(void* trackdata, const float* inV, const float* inN, unsigned w, unsigned h, const float* refV,
		const float* refN, unsigned rw, unsigned rh, const float* Ttrack, const float* view, float dist_thr,
		float normal_thr) {
	trackKernel((TrackData*) trackdata, (const float3*) inV, (const float3*) inN, make_uint2(w, h),
			(const float3*) refV, (const float3*) refN, make_uint2(rw, rh), toM(Ttrack), toM(view), dist_thr,
			normal_thr);
}
void kfo_reduce(float* out8x32, void* trackdata, unsigned jw, unsigned jh, unsigned w, unsigned h) {
	reduceKernel(out8x32, (TrackData*) trackdata, make_uint2(jw, jh), make_uint2(w, h));
}
int kfo_update_pose(float* pose, const float* out8x32, float icp_threshold) {
	Matrix4 p = toM(pose);
	bool r = updatePoseKernel(p, out8x32, icp_threshold);
	std::memcpy(pose, &p, sizeof(float) * 16);
	return r;
}
int kfo_check_pose(float* pose, const float* old_pose, const float* out8x32, unsigned w, unsigned h, float thr) {
	Matrix4 p = toM(pose);
	bool r = checkPoseKernel(p, toM(old_pose), out8x32, make_uint2(w, h), thr);
	std::memcpy(pose, &p, sizeof(float) * 16);
	return r;
}
void kfo_integrate(short* data, const unsigned* size, const float* dim, const float* depth, unsigned w, unsigned h,
		const float* invTrack, const float* K, float mu, float maxw) {
	integrateKernel(toV(data, size, dim), depth, make_uint2(w, h), toM(invTrack), toM(K), mu, maxw);
}
void kfo_raycast(float* vtx, float* nrm, unsigned w, unsigned h, short* data, const unsigned* size, const float* dim,
		const float* view, float nearP, float farP, float step, float largestep) {
	raycastKernel((float3*) vtx, (float3*) nrm, make_uint2(w, h), toV(data, size, dim), toM(view), nearP, farP, step,
			largestep);
}
void kfo_render_depth(unsigned char* out, float* depth, unsigned w, unsigned h, float nearP, float farP) {
	renderDepthKernel((uchar4*) out, depth, make_uint2(w, h), nearP, farP);
}
void kfo_render_track(unsigned char* out, const void* trackdata, unsigned w, unsigned h) {
	renderTrackKernel((uchar4*) out, (const TrackData*) trackdata, make_uint2(w, h));
}
void kfo_render_volume(unsigned char* out, unsigned w, unsigned h, short* data, const unsigned* size, const float* dim,
		const float* view, float nearP, float farP, float step, float largestep) {
	renderVolumeKernel((uchar4*) out, make_uint2(w, h), toV(data, size, dim), toM(view), nearP, farP, step, largestep,
			light, ambient);
}

// ------------------------------------------------------------ host 4x4 math
void kfo_inverse(float* out, const float* in) {
	Matrix4 r = inverse(toM(in));
	std::memcpy(out, &r, sizeof(float) * 16);
}
void kfo_matmul(float* out, const float* a, const float* b) {
	Matrix4 r = toM(a) * toM(b);
	std::memcpy(out, &r, sizeof(float) * 16);
}
void kfo_camera_matrix(float* out, const float* k) {
	Matrix4 r = getCameraMatrix(make_float4(k[0], k[1], k[2], k[3]));
	std::memcpy(out, &r, sizeof(float) * 16);
}
void kfo_inverse_camera_matrix(float* out, const float* k) {
	Matrix4 r = getInverseCameraMatrix(make_float4(k[0], k[1], k[2], k[3]));
	std::memcpy(out, &r, sizeof(float) * 16);
}
void kfo_solve(double* x6, const float* vals27) {
	TooN::Vector<27, float> v;
	for (int i = 0; i < 27; ++i) v[i] = vals27[i];
	TooN::Vector<6> x = solve(v);
	for (int i = 0; i < 6; ++i) x6[i] = x[i];
}
void kfo_se3_exp(float* out16, const double* x6) {
	TooN::Vector<6> x;
	for (int i = 0; i < 6; ++i) x[i] = x6[i];
	Matrix4 r = toMatrix4(TooN::SE3<>(x));
	std::memcpy(out16, &r, sizeof(float) * 16);
}

// ------------------------------------------------- whole pipeline (class Kfusion)
int kfo_kf_create(unsigned cw, unsigned ch, const unsigned* vres, const float* vdim, const float* init_pos,
		const int* pyramid, int n_levels) {
	if (g_kf) return 1;  // the reference's globals allow one instance per process
	g_pyramid.assign(pyramid, pyramid + n_levels);
	g_csize = make_uint2(cw, ch);
	g_kf = new Kfusion(g_csize, make_uint3(vres[0], vres[1], vres[2]), make_float3(vdim[0], vdim[1], vdim[2]),
			make_float3(init_pos[0], init_pos[1], init_pos[2]), g_pyramid);
	return 0;
}
void kfo_kf_destroy() {
	delete g_kf;
	g_kf = NULL;
	// the reference never clears raycastPose/oldPose (zero-initialised statics); do it
	// here so a second instance in the same process starts like a fresh process
	std::memset(&raycastPose, 0, sizeof(raycastPose));
	std::memset(&oldPose, 0, sizeof(oldPose));
}
void kfo_kf_reset() { g_kf->reset(); }
int kfo_kf_preprocess(const unsigned short* depth, unsigned iw, unsigned ih) {
	return g_kf->preprocessing(depth, make_uint2(iw, ih));
}
int kfo_kf_track(const float* k, float icp_threshold, unsigned tracking_rate, unsigned frame) {
	return g_kf->tracking(make_float4(k[0], k[1], k[2], k[3]), icp_threshold, tracking_rate, frame);
}
int kfo_kf_integrate(const float* k, unsigned integration_rate, float mu, unsigned frame) {
	return g_kf->integration(make_float4(k[0], k[1], k[2], k[3]), integration_rate, mu, frame);
}
int kfo_kf_raycast(const float* k, float mu, unsigned frame) {
	return g_kf->raycasting(make_float4(k[0], k[1], k[2], k[3]), mu, frame);
}
void kfo_kf_get_pose(float* out16) {
	Matrix4 p = g_kf->getPose();
	std::memcpy(out16, &p, sizeof(float) * 16);
}
void kfo_kf_render_depth(unsigned char* out, unsigned w, unsigned h) { g_kf->renderDepth((uchar4*) out, make_uint2(w, h)); }
void kfo_kf_render_track(unsigned char* out, unsigned w, unsigned h) { g_kf->renderTrack((uchar4*) out, make_uint2(w, h)); }
void kfo_kf_render_volume(unsigned char* out, unsigned w, unsigned h, int frame, int rate, const float* k, float largestep) {
	g_kf->renderVolume((uchar4*) out, make_uint2(w, h), frame, rate, make_float4(k[0], k[1], k[2], k[3]), largestep);
}
void kfo_kf_dump_volume(const char* path) { g_kf->dumpVolume(path); }

// buffer access: returns a pointer into the backend's own storage (no copy)
// which: 0 volume(short2) 1 vertex 2 normal 3 floatDepth 4 ScaledDepth[l] 5 inputVertex[l]
//        6 inputNormal[l] 7 reductionoutput(8x32) 8 trackingResult 9 raycastPose 10 oldPose 11 gaussian
void* kfo_kf_buffer(int which, int level) {
	switch (which) {
	case 0: return volume.data;
	case 1: return vertex;
	case 2: return normal;
	case 3: return floatDepth;
	case 4: return ScaledDepth[level];
	case 5: return inputVertex[level];
	case 6: return inputNormal[level];
	case 7: return reductionoutput;
	case 8: return trackingResult;
	case 9: return &raycastPose;
	case 10: return &oldPose;
	case 11: return gaussian;
	}
	return NULL;
}

}  // extern "C"
