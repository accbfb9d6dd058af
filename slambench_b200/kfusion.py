"""Python host-side mirror of the reference's `class Kfusion`
(kfusion/include/kernels.h:83-195) over the C ABI of libkfb200.so (include/kfb200.h).

Same method names, argument meaning and return values as the reference:
`preprocessing / tracking / integration / raycasting / computeFrame / getPose / reset /
renderDepth / renderTrack / renderVolume / dumpVolume`.  This module is a thin ctypes
binding — all computation happens in the CUDA library; if the library or a CUDA device
is missing it raises, there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

KFB_MAX_LEVELS = 8
FLAG_ICP_HOST_SOLVE = 0x1
FLAG_INTEGRATE_V1 = 0x4
FLAG_TRACK_STATUS = 0x2
FLAG_INTEGRATE_NO_CULL = 0x8
FLAG_RAYCAST_NO_SKIP = 0x10
FLAG_BRICKS_MERGED = 0x20

(BUF_VOLUME, BUF_VERTEX, BUF_NORMAL, BUF_FLOATDEPTH, BUF_SCALEDDEPTH, BUF_INVERTEX, BUF_INNORMAL,
 BUF_REDUCTION, BUF_TRACKSTATUS, BUF_RAYCASTPOSE, BUF_OLDPOSE, BUF_GAUSSIAN, BUF_INPUTDEPTH, BUF_REDUCTION_DEV,
 BUF_BRICKFLAGS, BUF_RAYTILECOST, BUF_BRICKCLASS) = range(17)

# constant_parameters.h:15-23
E_DELTA, RADIUS, DIST_THRESHOLD, NORMAL_THRESHOLD, TRACK_THRESHOLD = 0.1, 2, 0.1, 0.8, 0.15
MAXWEIGHT, NEAR_PLANE, FAR_PLANE = 100.0, 0.4, 4.0


class KfbConfig(C.Structure):
    _fields_ = [
        ("compute_w", C.c_uint32), ("compute_h", C.c_uint32),
        ("volume_res", C.c_uint32 * 3), ("volume_dim", C.c_float * 3),
        ("init_pose", C.c_float * 16),
        ("n_levels", C.c_int32), ("iterations", C.c_int32 * KFB_MAX_LEVELS),
        ("device", C.c_int32),
        ("slab_z0", C.c_uint32), ("slab_z1", C.c_uint32),
        ("flags", C.c_uint32),
    ]


class KfbStats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_uint64), ("frames_integrated", C.c_uint64),
        ("voxels_updated_last", C.c_uint64), ("voxels_updated_total", C.c_uint64),
        ("icp_iterations_last", C.c_uint64), ("icp_iterations_total", C.c_uint64),
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
        ("ms_preprocess", C.c_float), ("ms_track", C.c_float), ("ms_integrate", C.c_float), ("ms_raycast", C.c_float),
    ]


class KfbIpcHandles(C.Structure):
    _fields_ = [("volume", C.c_uint8 * 64), ("vertex", C.c_uint8 * 64), ("normal", C.c_uint8 * 64), ("bricks", C.c_uint8 * 64),
                ("sync", C.c_uint8 * 64), ("has_bricks", C.c_uint32), ("slab_z0", C.c_uint32), ("slab_z1", C.c_uint32),
                ("reserved", C.c_uint32)]


class KfbError(RuntimeError):
    pass


_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """Load libkfb200.so (building it in-tree first if the sources are newer)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    if path is None and os.environ.get("KFB_LIB"):
        path = os.environ["KFB_LIB"]   # an alternative build of the same library (tuning experiments)
    if path is None:
        path = _build.LIB
        if not os.path.exists(path) or os.environ.get("KFB_REBUILD"):
            _build.build_lib()
    if not os.path.exists(path):
        raise KfbError(f"{path} is missing: the CUDA extension must be built (python -m slambench_b200.build); "
                       "there is no CPU fallback")
    lib = C.CDLL(path)
    lib.kfb_last_error.restype = C.c_char_p
    lib.kfb_abi_version.restype = C.c_int
    _lib = lib
    return lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a, n=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} floats, got {a.size}")
    return a


def identity_pose(t) -> np.ndarray:
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = np.asarray(t, np.float32)
    return m


class Kfusion:
    """B200 backend with the reference's `Kfusion` interface.

    Kfusion(inputSize, volumeResolution, volumeDimensions, initPose, pyramid)   kernels.h:99-138
    `inputSize` is the COMPUTATION size; `initPose` is either a 3-vector
    (translation, identity rotation — kernels.h:105-109) or a 4x4 matrix (kernels.h:122-127).
    """

    def __init__(self, inputSize, volumeResolution, volumeDimensions, initPose, pyramid=(10, 5, 4),
                 device: int = 0, flags: int = 0, slab=None, lib_path: str | None = None):
        self.lib = load_library(lib_path)
        cfg = KfbConfig()
        cfg.compute_w, cfg.compute_h = int(inputSize[0]), int(inputSize[1])
        vr = [int(volumeResolution)] * 3 if np.isscalar(volumeResolution) else [int(v) for v in volumeResolution]
        vd = [float(volumeDimensions)] * 3 if np.isscalar(volumeDimensions) else [float(v) for v in volumeDimensions]
        cfg.volume_res[:] = vr
        cfg.volume_dim[:] = vd
        ip = np.asarray(initPose, dtype=np.float32)
        pose = identity_pose(ip) if ip.size == 3 else ip.reshape(4, 4)
        cfg.init_pose[:] = [float(v) for v in pose.reshape(-1)]
        self._init_pose = pose[:3, 3].copy()
        if len(pyramid) > KFB_MAX_LEVELS:
            raise ValueError("too many pyramid levels")
        cfg.n_levels = len(pyramid)
        for i, it in enumerate(pyramid):
            cfg.iterations[i] = int(it)
        cfg.device = device
        if slab is not None:
            cfg.slab_z0, cfg.slab_z1 = int(slab[0]), int(slab[1])
        cfg.flags = flags
        self.cfg = cfg
        self.computationSize = (cfg.compute_w, cfg.compute_h)
        self.volumeResolution = tuple(vr)
        self.volumeDimensions = tuple(vd)
        self.slab = (cfg.slab_z0, cfg.slab_z1) if slab is not None else (0, vr[2])
        self.levels = len(pyramid)
        self._h = C.c_void_p()
        self._check(self.lib.kfb_create(C.byref(cfg), C.byref(self._h)))
        self._tracked = False
        self._integrated = False

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int):
        if rc != 0:
            raise KfbError(f"kfb error {rc}: {self.lib.kfb_last_error().decode()}")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.kfb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -------------------------------------------------------- the reference's API
    def reset(self):
        self._check(self.lib.kfb_reset(self._h))

    def preprocessing(self, inputDepth: np.ndarray, inputSize=None) -> bool:
        """`inputDepth`: host uint16 millimetres [h, w]; copied to the device inside the call."""
        d = inputDepth if (inputDepth.dtype == np.uint16 and inputDepth.flags.c_contiguous) else np.ascontiguousarray(inputDepth, np.uint16)
        h, w = d.shape if inputSize is None else (inputSize[1], inputSize[0])
        self._keep = d  # the copy is asynchronous
        self._check(self.lib.kfb_preprocess(self._h, _p(d), C.c_uint32(w), C.c_uint32(h)))
        return True

    def preprocessing_device(self, dev_ptr: int, inputSize) -> bool:
        """Same, for a uint16 frame already resident in device memory (raw pointer)."""
        self._check(self.lib.kfb_preprocess_device(self._h, C.c_void_p(dev_ptr), C.c_uint32(inputSize[0]), C.c_uint32(inputSize[1])))
        return True

    def tracking(self, k, icp_threshold: float, tracking_rate: int, frame: int) -> bool:
        t = C.c_int(0)
        self._check(self.lib.kfb_track(self._h, _p(_f32(k, 4)), C.c_float(icp_threshold), C.c_uint32(tracking_rate), C.c_uint32(frame), C.byref(t)))
        return bool(t.value)

    def integration(self, k, integration_rate: int, mu: float, frame: int) -> bool:
        t = C.c_int(0)
        self._check(self.lib.kfb_integrate(self._h, _p(_f32(k, 4)), C.c_uint32(integration_rate), C.c_float(mu), C.c_uint32(frame), C.byref(t)))
        return bool(t.value)

    def raycasting(self, k, mu: float, frame: int) -> bool:
        self._check(self.lib.kfb_raycast(self._h, _p(_f32(k, 4)), C.c_float(mu), C.c_uint32(frame)))
        return False  # the reference always returns false (cpp/kernels.cpp:975-984)

    def computeFrame(self, inputDepth, inputSize, k, integration_rate, tracking_rate, icp_threshold, mu, frame):
        """Kfusion::computeFrame (cpp/kernels.cpp:1048-1055) through the single C-ABI entry point."""
        d = inputDepth if (inputDepth.dtype == np.uint16 and inputDepth.flags.c_contiguous) else np.ascontiguousarray(inputDepth, np.uint16)
        h, w = d.shape if inputSize is None else (inputSize[1], inputSize[0])
        self._keep = d
        tr, it = C.c_int(0), C.c_int(0)
        self._check(self.lib.kfb_compute_frame(self._h, _p(d), C.c_uint32(w), C.c_uint32(h), _p(_f32(k, 4)), C.c_uint32(integration_rate),
                                               C.c_uint32(tracking_rate), C.c_float(icp_threshold), C.c_float(mu), C.c_uint32(frame),
                                               C.byref(tr), C.byref(it)))
        self._tracked, self._integrated = bool(tr.value), bool(it.value)

    def computeFrame_device(self, dev_ptr: int, inputSize, k, integration_rate, tracking_rate, icp_threshold, mu, frame):
        """computeFrame for a uint16 sensor frame already resident in device memory (raw pointer)."""
        tr, it = C.c_int(0), C.c_int(0)
        self._check(self.lib.kfb_compute_frame_device(self._h, C.c_void_p(dev_ptr), C.c_uint32(inputSize[0]), C.c_uint32(inputSize[1]),
                                                      _p(_f32(k, 4)), C.c_uint32(integration_rate), C.c_uint32(tracking_rate),
                                                      C.c_float(icp_threshold), C.c_float(mu), C.c_uint32(frame), C.byref(tr), C.byref(it)))
        self._tracked, self._integrated = bool(tr.value), bool(it.value)

    def getTracked(self) -> bool:
        return self._tracked

    def getIntegrated(self) -> bool:
        return self._integrated

    def getPose(self) -> np.ndarray:
        out = np.empty(16, np.float32)
        self._check(self.lib.kfb_get_pose(self._h, _p(out)))
        return out.reshape(4, 4)

    def setPose(self, pose):
        self._check(self.lib.kfb_set_pose(self._h, _p(_f32(pose, 16))))

    def getPosition(self) -> np.ndarray:
        return self.getPose()[:3, 3] - self._init_pose  # kernels.h:150-157

    def synchroniseDevices(self):
        self._check(self.lib.kfb_sync(self._h))

    def renderDepth(self) -> np.ndarray:
        w, h = self.computationSize
        out = np.zeros((h, w, 4), np.uint8)
        self._check(self.lib.kfb_render_depth(self._h, _p(out), C.c_uint32(w), C.c_uint32(h)))
        return out

    def renderTrack(self) -> np.ndarray:
        w, h = self.computationSize
        out = np.zeros((h, w, 4), np.uint8)
        self._check(self.lib.kfb_render_track(self._h, _p(out), C.c_uint32(w), C.c_uint32(h)))
        return out

    def renderVolume(self, frame: int, rate: int, k, largestep: float, viewPose=None) -> np.ndarray:
        w, h = self.computationSize
        out = np.zeros((h, w, 4), np.uint8)
        vp = None if viewPose is None else _p(_f32(viewPose, 16))
        self._check(self.lib.kfb_render_volume(self._h, _p(out), C.c_uint32(w), C.c_uint32(h), C.c_int(frame), C.c_int(rate),
                                               _p(_f32(k, 4)), C.c_float(largestep), vp))
        return out

    def dumpVolume(self, filename: str):
        self._check(self.lib.kfb_dump_volume(self._h, filename.encode()))

    # --------------------------------------- stage-level kernels (kernels.h:18-69)
    def pyramidKernels(self, k):
        self._check(self.lib.kfb_k_pyramid(self._h, _p(_f32(k, 4))))

    def trackReduceKernel(self, level, Ttrack, view, dist_threshold=DIST_THRESHOLD, normal_threshold=NORMAL_THRESHOLD) -> np.ndarray:
        out = np.empty(32, np.float32)
        self._check(self.lib.kfb_k_track_reduce(self._h, C.c_int(level), _p(_f32(Ttrack, 16)), _p(_f32(view, 16)),
                                                C.c_float(dist_threshold), C.c_float(normal_threshold), _p(out)))
        return out

    def integrateKernel(self, invTrack, K, mu: float, maxweight: float = MAXWEIGHT):
        self._check(self.lib.kfb_k_integrate(self._h, _p(_f32(invTrack, 16)), _p(_f32(K, 16)), C.c_float(mu), C.c_float(maxweight)))

    def raycastKernel(self, view, nearPlane=NEAR_PLANE, farPlane=FAR_PLANE, step=None, largestep=0.075):
        if step is None:
            step = float(np.float32(min(self.volumeDimensions)) / np.float32(max(self.volumeResolution)))
        self._check(self.lib.kfb_k_raycast(self._h, _p(_f32(view, 16)), C.c_float(nearPlane), C.c_float(farPlane), C.c_float(step), C.c_float(largestep)))

    def updatePoseKernel(self, pose, reduction32, icp_threshold):
        p = _f32(pose, 16).copy()
        conv = C.c_int(0)
        self._check(self.lib.kfb_k_update_pose(_p(p), _p(_f32(reduction32, 32)), C.c_float(icp_threshold), C.byref(conv)))
        return p.reshape(4, 4), bool(conv.value)

    def checkPoseKernel(self, pose, oldPose, reduction32, imageSize, track_threshold=TRACK_THRESHOLD):
        p = _f32(pose, 16).copy()
        ok = C.c_int(0)
        self._check(self.lib.kfb_k_check_pose(_p(p), _p(_f32(oldPose, 16)), _p(_f32(reduction32, 32)), C.c_uint32(imageSize[0]),
                                              C.c_uint32(imageSize[1]), C.c_float(track_threshold), C.byref(ok)))
        return p.reshape(4, 4), bool(ok.value)

    # host 4x4 helpers of the product (commons.h:343-378)
    def inverse(self, m):
        out = np.empty(16, np.float32)
        self.lib.kfb_inverse4(_p(out), _p(_f32(m, 16)))
        return out.reshape(4, 4)

    def matmul(self, a, b):
        out = np.empty(16, np.float32)
        self.lib.kfb_matmul4(_p(out), _p(_f32(a, 16)), _p(_f32(b, 16)))
        return out.reshape(4, 4)

    def cameraMatrix(self, k):
        out = np.empty(16, np.float32)
        self.lib.kfb_camera_matrix(_p(out), _p(_f32(k, 4)))
        return out.reshape(4, 4)

    def inverseCameraMatrix(self, k):
        out = np.empty(16, np.float32)
        self.lib.kfb_inverse_camera_matrix(_p(out), _p(_f32(k, 4)))
        return out.reshape(4, 4)

    # ------------------------------------------------------------ buffer access
    def _buf_shape(self, which, level):
        w, h = self.computationSize
        lw, lh = w >> level, h >> level
        vr = self.volumeResolution
        nz = self.slab[1] - self.slab[0]
        return {
            BUF_VOLUME: ((nz, vr[1], vr[0], 2), np.int16),
            BUF_VERTEX: ((h, w, 3), np.float32), BUF_NORMAL: ((h, w, 3), np.float32),
            BUF_FLOATDEPTH: ((h, w), np.float32), BUF_SCALEDDEPTH: ((lh, lw), np.float32),
            BUF_INVERTEX: ((lh, lw, 3), np.float32), BUF_INNORMAL: ((lh, lw, 3), np.float32),
            BUF_REDUCTION: ((32,), np.float32), BUF_TRACKSTATUS: ((h, w), np.int8),
            BUF_RAYCASTPOSE: ((4, 4), np.float32), BUF_OLDPOSE: ((4, 4), np.float32), BUF_GAUSSIAN: ((5,), np.float32),
            BUF_REDUCTION_DEV: ((32,), np.float32),
            BUF_BRICKFLAGS: (((vr[2] + 7) // 8, (vr[1] + 7) // 8, (vr[0] + 7) // 8), np.uint8),
            BUF_RAYTILECOST: (((h + 3) // 4, (w + 7) // 8), np.uint32),
            BUF_BRICKCLASS: (((nz + 7) // 8, (vr[1] + 7) // 8, (vr[0] + 7) // 8), np.uint8),
        }[which]

    def read(self, which: int, level: int = 0) -> np.ndarray:
        shape, dt = self._buf_shape(which, level)
        out = np.empty(shape, dt)
        self._check(self.lib.kfb_read_buffer(self._h, C.c_int(which), C.c_int(level), _p(out), C.c_size_t(out.nbytes)))
        return out

    def write(self, which: int, data: np.ndarray, level: int = 0):
        shape, dt = self._buf_shape(which, level)
        a = np.ascontiguousarray(data, dtype=dt)
        if a.shape != tuple(shape):
            raise ValueError(f"buffer {which}: expected shape {shape}, got {a.shape}")
        self._check(self.lib.kfb_write_buffer(self._h, C.c_int(which), C.c_int(level), _p(a), C.c_size_t(a.nbytes)))

    def device_ptr(self, which: int, level: int = 0) -> int:
        p = C.c_void_p()
        self._check(self.lib.kfb_device_ptr(self._h, C.c_int(which), C.c_int(level), C.byref(p)))
        return p.value

    # -------------------------------------------------------------- measurement
    def stream(self) -> int:
        s = C.c_void_p()
        self._check(self.lib.kfb_stream(self._h, C.byref(s)))
        return s.value or 0

    def torch_stream(self, torch):
        """This context's CUDA stream as a torch ExternalStream (collectives of the sharded mode are enqueued on it)."""
        return torch.cuda.ExternalStream(self.stream(), device=self.cfg.device)

    def enable_timing(self, on=True):
        """`on`: True/False for all stages, or a bitmask (1 preprocess, 2 track, 4 integrate, 8 raycast)."""
        mask = (15 if on else 0) if isinstance(on, bool) else int(on)
        self._check(self.lib.kfb_enable_timing(self._h, C.c_int(mask)))

    def reset_stats(self):
        self._check(self.lib.kfb_reset_stats(self._h))

    def stats(self) -> dict:
        st = KfbStats()
        self._check(self.lib.kfb_get_stats(self._h, C.byref(st)))
        return {f: getattr(st, f) for f, _ in KfbStats._fields_}

    # ---------------------------------------------------------------- multi-GPU
    def set_pixel_rows(self, row0: int, row1: int):
        self._check(self.lib.kfb_set_pixel_rows(self._h, C.c_uint32(row0), C.c_uint32(row1)))

    def ipc_export(self) -> bytes:
        """This context's slab, raycast maps, brick flags and barrier slot as CUDA IPC handles (kfb_ipc_handles, 336 bytes)."""
        h = KfbIpcHandles()
        self._check(self.lib.kfb_ipc_export(self._h, C.byref(h)))
        return bytes(h)

    def ipc_import(self, rank: int, world: int, handles: list[bytes]):
        """Enter the peer-memory z-slab mode: `handles[r]` = rank r's ipc_export()."""
        arr = (KfbIpcHandles * world).from_buffer_copy(b"".join(handles))
        self._check(self.lib.kfb_ipc_import(self._h, C.c_int(rank), C.c_int(world), arr))

    def peer_barrier(self):
        self._check(self.lib.kfb_peer_barrier(self._h))

    def slab_ipc_handle(self) -> bytes:
        buf = (C.c_uint8 * 64)()
        self._check(self.lib.kfb_slab_ipc_handle(self._h, buf))
        return bytes(buf)

    def slab_import(self, rank: int, world: int, handles: list[bytes], z_begin: list[int]):
        hb = (C.c_uint8 * (64 * world)).from_buffer_copy(b"".join(handles))
        zb = (C.c_uint32 * world)(*z_begin)
        self._check(self.lib.kfb_slab_import(self._h, C.c_int(rank), C.c_int(world), hb, zb))
