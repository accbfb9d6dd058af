"""Synthetic depth sequences for the KinectFusion hot path.

An analytic closed room strictly inside the [0, 4.8]^3 m volume cube, rendered as
PLANAR-z depth in millimetres exactly like the reference's dataset converter does
(kfusion/thirdparty/scene2raw.cpp:97-108), and written in the `.raw` container
that `RawDepthReader` parses (kfusion/include/interface.h:233-293; writer layout
scene2raw.cpp:170-176):

    per frame:  uint32 w, h ; uint16 depth_mm[w*h] ; uint32 w, h ; uint8 rgb[w*h*3]

Pose convention = the reference's: camera->world 4x4, identity rotation looks along
+z, image x = +x, image y = +y; pixel (u, v) back-projects to ((u-cx)/fx, (v-cy)/fy, 1)
(depth2vertexKernel, cpp/kernels.cpp:200-218).

Everything is deterministic (no RNG) so the GPU box regenerates bit-identical input.
"""
from __future__ import annotations

import numpy as np

W, H = 640, 480
# ICL-NUIM intrinsics, passed to the benchmark as `-k 481.2,480,320,240`
K_DEFAULT = (481.2, 480.0, 320.0, 240.0)
VOLUME_DIM = 4.8
# `-p 0.5,0.5,0.25` => t0 = p * volume_size
INIT_POS_FACTOR = (0.5, 0.5, 0.25)

ROOM_LO = np.array([0.3, 1.1, 0.2])
ROOM_HI = np.array([4.5, 3.7, 4.4])
SPHERE_C = np.array([2.6, 2.6, 3.0])
SPHERE_R = 0.5
BOXES = (
    (np.array([0.8, 2.5, 3.0]), np.array([1.6, 3.7, 4.4])),
    (np.array([3.4, 1.1, 2.4]), np.array([4.5, 2.0, 3.2])),
)


def rpy_to_R(r: float, p: float, y: float) -> np.ndarray:
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def trajectory_pose(frame: int, seed: int = 0, long_run: bool = False) -> np.ndarray:
    """Ground-truth camera->world pose of `frame`.  Frames 0-3 are static: the
    reference cannot track before its first raycast at frame 3."""
    k = max(0, frame - 3)
    t0 = np.array(INIT_POS_FACTOR) * VOLUME_DIM
    # eight distinct trajectories for the independent-sequence mode (config 3)
    sx = (1.0, -1.0)[seed & 1]
    sy = (1.0, -1.0)[(seed >> 1) & 1]
    sz = (1.0, 0.5)[(seed >> 2) & 1]
    if long_run:
        # low-frequency sinusoids keep a 1000-frame run inside the room
        ph = 2.0 * np.pi * k / 400.0
        t = t0 + np.array([0.45 * np.sin(ph) * sx, 0.20 * np.sin(2 * ph) * sy, 0.30 * (1 - np.cos(ph)) * sz])
        rpy = np.array([0.10 * np.sin(ph) * sy, 0.20 * np.sin(ph + 0.7) * sx, -0.12 * np.sin(2 * ph)])
    else:
        t = t0 + np.array([0.004 * sx, -0.002 * sy, 0.003 * sz]) * k
        rpy = np.array([0.001 * sy, 0.002 * sx, -0.0015]) * k
    T = np.eye(4)
    T[:3, :3] = rpy_to_R(*rpy)
    T[:3, 3] = t
    return T


def expected_pose(gt: np.ndarray, frame: int) -> np.ndarray:
    """Pose a tracker that starts at the reference's initial pose (identity rotation at t0, kernels.h:106-109) should
    report for `frame`: it builds its model in the frame of ITS first pose, so the ground truth is re-expressed there,
    P0 * inverse(gt[0]) * gt[frame].  Equal to gt[frame] whenever gt[0] is that initial pose (every short sequence)."""
    p0 = np.eye(4)
    p0[:3, 3] = gt[0][:3, 3]
    return p0 @ np.linalg.inv(gt[0]) @ gt[frame]


def _ray_aabb_exit(o, d, lo, hi):
    """Distance to the inside surface of a box that contains `o`."""
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (lo - o) / d
        t2 = (hi - o) / d
    tfar = np.maximum(t1, t2)
    tfar = np.where(np.isfinite(tfar), tfar, np.inf)
    return tfar.min(axis=-1)


def _ray_aabb_entry(o, d, lo, hi):
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (lo - o) / d
        t2 = (hi - o) / d
    tn = np.minimum(t1, t2)
    tf = np.maximum(t1, t2)
    tn = np.where(np.isnan(tn), -np.inf, tn).max(axis=-1)
    tf = np.where(np.isnan(tf), np.inf, tf).min(axis=-1)
    hit = (tn <= tf) & (tn > 0)
    return np.where(hit, tn, np.inf)


def _ray_sphere(o, d, c, r):
    oc = o - c
    a = (d * d).sum(-1)
    b = 2.0 * (d * oc).sum(-1)
    cc = (oc * oc).sum(-1) - r * r
    disc = b * b - 4 * a * cc
    ok = disc > 0
    sq = np.sqrt(np.where(ok, disc, 0.0))
    t = (-b - sq) / (2 * a)
    return np.where(ok & (t > 0), t, np.inf)


def render_depth_mm(pose: np.ndarray, k=K_DEFAULT, w: int = W, h: int = H) -> np.ndarray:
    """uint16[h, w] planar-z depth in mm (truncated), 0 = invalid (never produced here)."""
    fx, fy, cx, cy = k
    u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    d_cam = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], axis=-1)
    d = d_cam @ pose[:3, :3].T
    o = pose[:3, 3]
    t = _ray_aabb_exit(o, d, ROOM_LO, ROOM_HI)
    t = np.minimum(t, _ray_sphere(o, d, SPHERE_C, SPHERE_R))
    for lo, hi in BOXES:
        t = np.minimum(t, _ray_aabb_entry(o, d, lo, hi))
    # the camera-frame direction has z == 1, so the ray parameter IS the planar depth
    mm = np.floor(t * 1000.0)
    mm = np.where(np.isfinite(mm) & (mm < 65535), mm, 0)
    return mm.astype(np.uint16)


def make_sequence(n_frames: int = 100, seed: int = 0, long_run: bool | None = None,
                  k=K_DEFAULT, w: int = W, h: int = H):
    """Returns (depth uint16[n, h, w], gt_poses float64[n, 4, 4])."""
    if long_run is None:
        long_run = n_frames > 200
    depth = np.empty((n_frames, h, w), dtype=np.uint16)
    poses = np.empty((n_frames, 4, 4), dtype=np.float64)
    last_pose, last_img = None, None
    for f in range(n_frames):
        T = trajectory_pose(f, seed, long_run)
        poses[f] = T
        if last_pose is not None and np.array_equal(T, last_pose):
            depth[f] = last_img
        else:
            depth[f] = render_depth_mm(T, k, w, h)
        last_pose, last_img = T, depth[f]
    return depth, poses


def write_raw(path: str, depth: np.ndarray) -> None:
    """`.raw` writer with the layout RawDepthReader expects (interface.h:244-276)."""
    n, h, w = depth.shape
    hdr = np.array([w, h], dtype=np.uint32).tobytes()
    rgb = np.zeros(w * h * 3, dtype=np.uint8).tobytes()
    with open(path, "wb") as f:
        for i in range(n):
            f.write(hdr)
            f.write(np.ascontiguousarray(depth[i]).tobytes())
            f.write(hdr)
            f.write(rgb)


def read_raw(path: str):
    """Reader for the same container (used by tests of the writer)."""
    frames = []
    with open(path, "rb") as f:
        while True:
            hdr = f.read(8)
            if len(hdr) < 8:
                break
            w, h = np.frombuffer(hdr, dtype=np.uint32)
            d = np.frombuffer(f.read(int(w) * int(h) * 2), dtype=np.uint16).reshape(int(h), int(w))
            f.read(8)
            f.read(int(w) * int(h) * 3)
            frames.append(d)
    return np.stack(frames)


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser(description="write a synthetic .raw depth sequence")
    ap.add_argument("out")
    ap.add_argument("--frames", type=int, default=100)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    dep, _ = make_sequence(a.frames, a.seed)
    write_raw(a.out, dep)
    print(f"wrote {a.out}: {a.frames} frames {W}x{H}")
