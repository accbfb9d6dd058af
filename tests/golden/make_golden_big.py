"""Golden fixtures at the SHARDED sizes (BASELINE configs[3], [4]): tests/golden/big_{1024,2048}.npz.

    python tests/golden/make_golden_big.py [1024] [2048]        (build container: /root/reference mounted; ~64 GB RAM for 2048)

1024^3 is run by the UNMODIFIED reference (oracle/_ref/libkfusion_ref_omp.so).  2048^3 cannot be: the reference indexes
the volume with 32-bit products (commons.h:161-184, 306) — it is run by oracle/_ref/libkfusion_ref64_omp.so, the same
unmodified kernels.cpp compiled against a build-time copy of commons.h whose index products are widened to size_t
(oracle/Makefile target `ref64`).  A volume of this size is pinned by per-slice checksums (tests/volsum.py).

Per size:
  teacher-forced   fresh volume; integrateKernel of frames TF_FRAMES at their ground-truth poses, then raycastKernel from
                   the last of them: slice checksums after every integrate, N_upd of each, the raycast maps' checksums and
                   every 16th row of them in full
  free-running     the whole pipeline (preprocess -> track -> integrate -> raycast, benchmark.cpp's order) over the first
                   FR_FRAMES frames: pose / tracked / integrated per frame, reduction row 0, final slice checksums, final
                   raycast maps (checksums + every 16th row)
The depth frames are the synthetic sequence (slambench_b200/synth.py, deterministic); their SHA-256 is stored so that a test
on another machine can tell "synth differs here" from "the kernels differ".
"""
from __future__ import annotations

import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import cpu_backend as cb  # noqa: E402
from slambench_b200 import synth  # noqa: E402
import volsum  # noqa: E402

K = np.array(synth.K_DEFAULT, np.float32)
T0 = (np.array(synth.INIT_POS_FACTOR, np.float32) * np.float32(synth.VOLUME_DIM)).astype(np.float32)
DIM = np.array([4.8, 4.8, 4.8], np.float32)
MU = 0.1
TF_FRAMES = (4, 9, 14)
FR_FRAMES = 10


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make(n: int, lib_path: str) -> dict:
    be = cb.CpuKfusion(lib_path)
    depth, gt = synth.make_sequence(max(max(TF_FRAMES) + 1, FR_FRAMES))
    out = {"n": np.array([n]), "backend": np.array([be.name + (" (size_t-indexed commons.h)" if "ref64" in lib_path else "")]),
           "tf_frames": np.array(TF_FRAMES), "depth_sha256": np.array([sha(depth[f]) for f in range(len(depth))]),
           "gt_poses": gt.astype(np.float64), "mu": np.array([MU], np.float32)}
    # ---- teacher-forced integrate x3 + raycast
    t0 = time.time()
    vol = be.init_volume((n, n, n))
    Kmat = be.camera_matrix(K)
    sums, nupd = [], []
    for f in TF_FRAMES:
        pose = gt[f].astype(np.float32)
        raw = be.mm2meters(depth[f], (640, 480))
        w_before = vol[..., 1].sum(dtype=np.int64)
        be.integrate(vol, DIM, raw, be.inverse(pose), Kmat, MU)
        nupd.append(int(vol[..., 1].sum(dtype=np.int64) - w_before))   # every update bumps one weight (all < maxweight here)
        sums.append(volsum.slice_checksums_np(vol))
        print(f"  {n}^3 teacher-forced frame {f}: N_upd {nupd[-1]}  ({time.time() - t0:.0f} s)", flush=True)
    pose = gt[TF_FRAMES[-1]].astype(np.float32)
    view = be.matmul(pose, be.inverse_camera_matrix(K))
    vtx, nrm = be.raycast(vol, DIM, (640, 480), view, largestep=0.75 * MU)
    out.update(tf_slice_sums=np.stack(sums), tf_nupd=np.array(nupd, np.int64), tf_view=view,
               tf_vertex_sum=volsum.array_checksum(vtx), tf_normal_sum=volsum.array_checksum(nrm),
               tf_vertex_rows=vtx[::16].copy(), tf_normal_rows=nrm[::16].copy(),
               tf_hits=np.array([int((nrm[..., 0] != -2).sum())]))
    del vol
    # ---- free-running pipeline
    be.create((640, 480), n, 4.8, T0, (10, 5, 4))
    poses, flags, reds = [], [], []
    try:
        for f in range(FR_FRAMES):
            be.preprocessing(depth[f])
            tr = be.tracking(K, 1e-5, 1, f)
            it = be.integration(K, 1, MU, f)
            be.raycasting(K, MU, f)
            poses.append(be.get_pose().copy())
            flags.append((tr, it))
            reds.append(be.buffer(cb.BUF_REDUCTION)[0].copy())
            print(f"  {n}^3 free-running frame {f}: tracked {tr} integrated {it}  ({time.time() - t0:.0f} s)", flush=True)
        v, nm = be.buffer(cb.BUF_VERTEX), be.buffer(cb.BUF_NORMAL)
        out.update(fr_poses=np.stack(poses), fr_flags=np.array(flags, np.uint8), fr_reduction_row0=np.stack(reds),
                   fr_slice_sums=volsum.slice_checksums_np(be.buffer(cb.BUF_VOLUME)),
                   fr_vertex_sum=volsum.array_checksum(v), fr_normal_sum=volsum.array_checksum(nm),
                   fr_vertex_rows=v[::16].copy(), fr_normal_rows=nm[::16].copy())
    finally:
        be.destroy()
    return out


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [1024, 2048]
    cb.build_ref()
    import subprocess

    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref64"])
    for n in sizes:
        lib = cb.REF_OMP_LIB if n <= 1024 else os.path.join(os.path.dirname(cb.REF_LIB), "libkfusion_ref64_omp.so")
        t = time.time()
        d = make(n, lib)
        path = os.path.join(HERE, f"big_{n}.npz")
        np.savez_compressed(path, **d)
        print(f"{path}: {os.path.getsize(path)} bytes, {time.time() - t:.0f} s")


if __name__ == "__main__":
    main()
