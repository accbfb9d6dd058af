"""ctypes driver for the CPU checkers (TEST INFRASTRUCTURE — see oracle/kfusion_oracle.c).

Loads either
  * oracle/libkfusion_oracle.so      — our plain-C restatement ("port"), or
  * oracle/_ref/libkfusion_ref.so    — the unmodified reference C++ backend ("reference"),
  * oracle/_ref/libkfusion_ref_omp.so— the same, built with -fopenmp,
which export the same `kfo_*` symbols.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(HERE, "libkfusion_oracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libkfusion_ref.so")
REF_OMP_LIB = os.path.join(HERE, "_ref", "libkfusion_ref_omp.so")
REFERENCE_ROOT = "/root/reference"

TRACKDATA = np.dtype([("result", np.int32), ("error", np.float32), ("J", np.float32, (6,))])
assert TRACKDATA.itemsize == 32

BUF_VOLUME, BUF_VERTEX, BUF_NORMAL, BUF_FLOATDEPTH, BUF_SCALEDDEPTH, BUF_INVERTEX, BUF_INNORMAL, \
    BUF_REDUCTION, BUF_TRACKDATA, BUF_RAYCASTPOSE, BUF_OLDPOSE, BUF_GAUSSIAN = range(12)


def build_port() -> str:
    """Compile the C restatement (gcc only; works on the GPU box too)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return PORT_LIB


def build_ref() -> str | None:
    """Compile the unmodified reference where it lies; only possible where /root/reference exists."""
    if not os.path.isdir(REFERENCE_ROOT):
        return REF_LIB if os.path.exists(REF_LIB) else None
    subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
    return REF_LIB


REF_CUDA_BIN = os.path.join(HERE, "_ref", "kfusion-benchmark-cuda")


def build_ref_cuda() -> str | None:
    """The reference's own CUDA backend, unmodified, compiled for sm_100a (GPU comparator); needs /root/reference + nvcc."""
    if os.path.isdir(REFERENCE_ROOT):
        try:
            subprocess.check_call(["make", "-s", "-C", HERE, "ref_cuda"])
        except subprocess.CalledProcessError:
            return None
    return REF_CUDA_BIN if os.path.exists(REF_CUDA_BIN) else None


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


class CpuKfusion:
    """Free kernels + the whole-pipeline `Kfusion` object of one CPU library."""

    def __init__(self, lib_path: str = PORT_LIB):
        if not os.path.exists(lib_path):
            if lib_path == PORT_LIB:
                build_port()
            else:
                raise FileNotFoundError(lib_path)
        self.lib = C.CDLL(lib_path)
        self.lib.kfo_impl_name.restype = C.c_char_p
        self.lib.kfo_kf_buffer.restype = C.c_void_p
        self.name = self.lib.kfo_impl_name().decode()
        self._cfg = None

    # ------------------------------------------------------------- 4x4 helpers
    def inverse(self, m):
        out = np.empty(16, np.float32)
        self.lib.kfo_inverse(_p(out), _p(_f32(m).reshape(16)))
        return out.reshape(4, 4)

    def matmul(self, a, b):
        out = np.empty(16, np.float32)
        self.lib.kfo_matmul(_p(out), _p(_f32(a).reshape(16)), _p(_f32(b).reshape(16)))
        return out.reshape(4, 4)

    def camera_matrix(self, k):
        out = np.empty(16, np.float32)
        self.lib.kfo_camera_matrix(_p(out), _p(_f32(k)))
        return out.reshape(4, 4)

    def inverse_camera_matrix(self, k):
        out = np.empty(16, np.float32)
        self.lib.kfo_inverse_camera_matrix(_p(out), _p(_f32(k)))
        return out.reshape(4, 4)

    def solve(self, vals27):
        x = np.empty(6, np.float64)
        self.lib.kfo_solve(_p(x), _p(_f32(vals27)))
        return x

    def se3_exp(self, x6):
        out = np.empty(16, np.float32)
        self.lib.kfo_se3_exp(_p(out), _p(np.ascontiguousarray(x6, dtype=np.float64)))
        return out.reshape(4, 4)

    # ------------------------------------------------------------ free kernels
    def init_volume(self, res):
        vol = np.empty((res[2], res[1], res[0], 2), np.int16)
        size = np.asarray(res, np.uint32)
        dim = np.ones(3, np.float32)
        self.lib.kfo_init_volume(_p(vol), _p(size), _p(dim))
        return vol

    def gaussian(self):
        g = np.empty(5, np.float32)
        self.lib.kfo_gaussian(_p(g))
        return g

    def mm2meters(self, depth_u16, out_wh):
        ih, iw = depth_u16.shape
        ow, oh = out_wh
        out = np.empty((oh, ow), np.float32)
        d = np.ascontiguousarray(depth_u16, dtype=np.uint16)
        self.lib.kfo_mm2meters(_p(out), C.c_uint(ow), C.c_uint(oh), _p(d), C.c_uint(iw), C.c_uint(ih))
        return out

    def bilateral(self, depth, gaussian, e_d=0.1, r=2):
        h, w = depth.shape
        out = np.empty((h, w), np.float32)
        self.lib.kfo_bilateral(_p(out), _p(_f32(depth)), C.c_uint(w), C.c_uint(h), _p(_f32(gaussian)), C.c_float(e_d), C.c_int(r))
        return out

    def halfsample(self, depth, e_d=0.3, r=1):
        h, w = depth.shape
        out = np.empty((h // 2, w // 2), np.float32)
        self.lib.kfo_halfsample(_p(out), _p(_f32(depth)), C.c_uint(w), C.c_uint(h), C.c_float(e_d), C.c_int(r))
        return out

    def depth2vertex(self, depth, invK):
        h, w = depth.shape
        out = np.empty((h, w, 3), np.float32)
        self.lib.kfo_depth2vertex(_p(out), _p(_f32(depth)), C.c_uint(w), C.c_uint(h), _p(_f32(invK).reshape(16)))
        return out

    def vertex2normal(self, vertex, out_init=None):
        h, w, _ = vertex.shape
        # invalid pixels only get .x written (cpp/kernels.cpp:240): start from the caller's buffer
        out = np.zeros((h, w, 3), np.float32) if out_init is None else np.array(out_init, np.float32, copy=True, order="C")
        self.lib.kfo_vertex2normal(_p(out), _p(_f32(vertex)), C.c_uint(w), C.c_uint(h))
        return out

    def track(self, inV, inN, refV, refN, Ttrack, view, dist_thr=0.1, normal_thr=0.8, out_init=None):
        h, w, _ = inV.shape
        rh, rw, _ = refV.shape
        td = np.zeros((rh, rw), TRACKDATA) if out_init is None else np.array(out_init, copy=True, order="C")
        self.lib.kfo_track(_p(td), _p(_f32(inV)), _p(_f32(inN)), C.c_uint(w), C.c_uint(h), _p(_f32(refV)), _p(_f32(refN)),
                           C.c_uint(rw), C.c_uint(rh), _p(_f32(Ttrack).reshape(16)), _p(_f32(view).reshape(16)),
                           C.c_float(dist_thr), C.c_float(normal_thr))
        return td

    def reduce(self, trackdata, size_wh):
        jh, jw = trackdata.shape
        out = np.zeros((8, 32), np.float32)
        td = np.ascontiguousarray(trackdata)
        self.lib.kfo_reduce(_p(out), _p(td), C.c_uint(jw), C.c_uint(jh), C.c_uint(size_wh[0]), C.c_uint(size_wh[1]))
        return out

    def update_pose(self, pose, reduction, icp_threshold=1e-5):
        p = _f32(pose).reshape(16).copy()
        conv = self.lib.kfo_update_pose(_p(p), _p(_f32(reduction).reshape(-1)), C.c_float(icp_threshold))
        return p.reshape(4, 4), bool(conv)

    def check_pose(self, pose, old_pose, reduction, size_wh, thr=0.15):
        p = _f32(pose).reshape(16).copy()
        ok = self.lib.kfo_check_pose(_p(p), _p(_f32(old_pose).reshape(16)), _p(_f32(reduction).reshape(-1)),
                                     C.c_uint(size_wh[0]), C.c_uint(size_wh[1]), C.c_float(thr))
        return p.reshape(4, 4), bool(ok)

    def integrate(self, vol, dim, depth, invTrack, K, mu=0.1, maxweight=100.0):
        """In place on `vol` (int16[z, y, x, 2])."""
        assert vol.dtype == np.int16 and vol.flags.c_contiguous
        size = np.asarray([vol.shape[2], vol.shape[1], vol.shape[0]], np.uint32)
        h, w = depth.shape
        self.lib.kfo_integrate(_p(vol), _p(size), _p(_f32(dim)), _p(_f32(depth)), C.c_uint(w), C.c_uint(h),
                               _p(_f32(invTrack).reshape(16)), _p(_f32(K).reshape(16)), C.c_float(mu), C.c_float(maxweight))
        return vol

    def raycast(self, vol, dim, size_wh, view, near=0.4, far=4.0, step=None, largestep=0.075, init=None):
        assert vol.dtype == np.int16 and vol.flags.c_contiguous
        size = np.asarray([vol.shape[2], vol.shape[1], vol.shape[0]], np.uint32)
        w, h = size_wh
        if step is None:
            step = float(np.float32(min(dim)) / np.float32(max(size)))
        if init is None:
            vtx = np.zeros((h, w, 3), np.float32)
            nrm = np.zeros((h, w, 3), np.float32)
        else:
            vtx, nrm = (np.array(a, np.float32, copy=True, order="C") for a in init)
        self.lib.kfo_raycast(_p(vtx), _p(nrm), C.c_uint(w), C.c_uint(h), _p(vol), _p(size), _p(_f32(dim)),
                             _p(_f32(view).reshape(16)), C.c_float(near), C.c_float(far), C.c_float(step), C.c_float(largestep))
        return vtx, nrm

    def render_depth(self, depth, near=0.4, far=4.0):
        h, w = depth.shape
        out = np.zeros((h, w, 4), np.uint8)
        self.lib.kfo_render_depth(_p(out), _p(_f32(depth)), C.c_uint(w), C.c_uint(h), C.c_float(near), C.c_float(far))
        return out

    def render_track(self, trackdata):
        h, w = trackdata.shape
        out = np.zeros((h, w, 4), np.uint8)
        self.lib.kfo_render_track(_p(out), _p(np.ascontiguousarray(trackdata)), C.c_uint(w), C.c_uint(h))
        return out

    def render_volume(self, vol, dim, size_wh, view, near=0.4, far=8.0, step=None, largestep=0.075):
        size = np.asarray([vol.shape[2], vol.shape[1], vol.shape[0]], np.uint32)
        w, h = size_wh
        if step is None:
            step = float(np.float32(min(dim)) / np.float32(max(size)))
        out = np.zeros((h, w, 4), np.uint8)
        self.lib.kfo_render_volume(_p(out), C.c_uint(w), C.c_uint(h), _p(vol), _p(size), _p(_f32(dim)),
                                   _p(_f32(view).reshape(16)), C.c_float(near), C.c_float(far), C.c_float(step), C.c_float(largestep))
        return out

    # ---------------------------------------------------------- whole pipeline
    def create(self, csize_wh, vres, vdim, init_pos, pyramid=(10, 5, 4)):
        vres = np.asarray([vres] * 3 if np.isscalar(vres) else vres, np.uint32)
        vdim = np.asarray([vdim] * 3 if np.isscalar(vdim) else vdim, np.float32)
        pyr = np.asarray(pyramid, np.int32)
        rc = self.lib.kfo_kf_create(C.c_uint(csize_wh[0]), C.c_uint(csize_wh[1]), _p(vres), _p(vdim),
                                    _p(_f32(init_pos)), _p(pyr), C.c_int(len(pyr)))
        if rc != 0:
            raise RuntimeError("one Kfusion per process/library (the reference keeps its state in globals)")
        self._cfg = dict(cw=csize_wh[0], ch=csize_wh[1], vres=vres, vdim=vdim, levels=len(pyr))

    def destroy(self):
        if self._cfg is not None:
            self.lib.kfo_kf_destroy()
            self._cfg = None

    def preprocessing(self, depth_u16):
        ih, iw = depth_u16.shape
        d = np.ascontiguousarray(depth_u16, dtype=np.uint16)
        return bool(self.lib.kfo_kf_preprocess(_p(d), C.c_uint(iw), C.c_uint(ih)))

    def tracking(self, k, icp_threshold, tracking_rate, frame):
        return bool(self.lib.kfo_kf_track(_p(_f32(k)), C.c_float(icp_threshold), C.c_uint(tracking_rate), C.c_uint(frame)))

    def integration(self, k, integration_rate, mu, frame):
        return bool(self.lib.kfo_kf_integrate(_p(_f32(k)), C.c_uint(integration_rate), C.c_float(mu), C.c_uint(frame)))

    def raycasting(self, k, mu, frame):
        return bool(self.lib.kfo_kf_raycast(_p(_f32(k)), C.c_float(mu), C.c_uint(frame)))

    def get_pose(self):
        out = np.empty(16, np.float32)
        self.lib.kfo_kf_get_pose(_p(out))
        return out.reshape(4, 4)

    def buffer(self, which, level=0) -> np.ndarray:
        """A live numpy VIEW of one of the backend's buffers."""
        c = self._cfg
        ptr = self.lib.kfo_kf_buffer(C.c_int(which), C.c_int(level))
        w, h = c["cw"] >> level, c["ch"] >> level
        if which == BUF_VOLUME:
            shape, dt = (int(c["vres"][2]), int(c["vres"][1]), int(c["vres"][0]), 2), np.int16
        elif which in (BUF_VERTEX, BUF_NORMAL):
            shape, dt = (c["ch"], c["cw"], 3), np.float32
        elif which == BUF_FLOATDEPTH:
            shape, dt = (c["ch"], c["cw"]), np.float32
        elif which == BUF_SCALEDDEPTH:
            shape, dt = (h, w), np.float32
        elif which in (BUF_INVERTEX, BUF_INNORMAL):
            shape, dt = (h, w, 3), np.float32
        elif which == BUF_REDUCTION:
            shape, dt = (8, 32), np.float32
        elif which == BUF_TRACKDATA:
            shape, dt = (c["ch"], c["cw"]), TRACKDATA
        elif which in (BUF_RAYCASTPOSE, BUF_OLDPOSE):
            shape, dt = (4, 4), np.float32
        elif which == BUF_GAUSSIAN:
            shape, dt = (5,), np.float32
        else:
            raise ValueError(which)
        n = int(np.prod(shape)) * np.dtype(dt).itemsize
        buf = (C.c_char * n).from_address(ptr)
        return np.frombuffer(buf, dtype=dt).reshape(shape)
