"""Generate tests/golden/*.npz from the UNMODIFIED reference C++ backend compiled in place
(oracle/_ref/libkfusion_ref.so; oracle/Makefile `make ref`).  Run in the build container, where
/root/reference is mounted:

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors of its own, so these fixtures are what
pins the oracle restatement (tests/test_oracle_golden.py, CPU) and the CUDA path
(tests/test_gpu_golden.py, GPU) where /root/reference does not exist.  Inputs are deterministic
(seeded numpy / the analytic synthetic sequence) and are stored together with the outputs.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import cpu_backend as cb  # noqa: E402
from slambench_b200 import synth  # noqa: E402

K = np.array(synth.K_DEFAULT, np.float32)
T0 = (np.array(synth.INIT_POS_FACTOR, np.float32) * np.float32(synth.VOLUME_DIM)).astype(np.float32)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def crafted_depth(w, h, seed):
    rng = np.random.default_rng(seed)
    d = (1500 + 40 * np.sin(np.arange(w)[None, :] / 7.0) + 30 * np.cos(np.arange(h)[:, None] / 5.0)
         + 6 * rng.random((h, w))).astype(np.uint16)
    d[rng.random((h, w)) < 0.05] = 0           # invalid pixels
    d[h // 3: h // 3 + 6, w // 4: w // 4 + 9] += 900   # a depth edge > e_delta
    d[:2, :] = 0
    return d


def kernels(ref: cb.CpuKfusion) -> dict:
    """Per-kernel known-answer vectors at 64x48 (image) / 24^3 (volume)."""
    w, h = 64, 48
    out = {}
    d_mm = crafted_depth(2 * w, 2 * h, 1)
    out["in_depth_mm"] = d_mm
    raw = ref.mm2meters(d_mm, (w, h))                        # ratio 2
    out["mm2meters_r2"] = raw
    raw1 = ref.mm2meters(d_mm[:h, :w].copy(), (w, h))        # ratio 1
    out["mm2meters_r1"] = raw1
    g = ref.gaussian()
    out["gaussian"] = g
    filt = ref.bilateral(raw1, g)
    out["bilateral"] = filt
    hs1 = ref.halfsample(filt)
    hs2 = ref.halfsample(hs1)
    out["halfsample1"], out["halfsample2"] = hs1, hs2
    k = np.array([60.0, 60.0, 32.0, 24.0], np.float32)
    out["k"] = k
    invK = ref.inverse_camera_matrix(k)
    out["invK"] = invK
    vtx = ref.depth2vertex(filt, invK)
    nrm = ref.vertex2normal(vtx)
    out["vertex"], out["normal"] = vtx, nrm
    # integrate a smooth surface into a 24^3 volume of 3 m from a slightly rotated pose, three times
    pose = np.eye(4, dtype=np.float32)
    pose[:3, :3] = synth.rpy_to_R(0.02, -0.03, 0.01).astype(np.float32)
    pose[:3, 3] = (1.5, 1.5, 0.1)
    out["pose"] = pose
    Kmat = ref.camera_matrix(k)
    inv = ref.inverse(pose)
    out["Kmat"], out["inv_pose"] = Kmat, inv
    dim = np.array([3.0, 3.0, 3.0], np.float32)
    vol = ref.init_volume((24, 24, 24))
    for _ in range(3):
        ref.integrate(vol, dim, raw1, inv, Kmat, 0.2)
    out["volume_after_3_integrates"] = vol.copy()
    view = ref.matmul(pose, invK)
    out["view"] = view
    rv, rn = ref.raycast(vol, dim, (w, h), view, near=0.4, far=4.0, largestep=0.15)
    out["raycast_vertex"], out["raycast_normal"] = rv, rn
    # track + reduce against the raycast maps from a perturbed pose
    pose2 = pose.copy()
    pose2[:3, 3] += np.array([0.004, -0.003, 0.002], np.float32)
    proj = ref.matmul(Kmat, ref.inverse(pose))
    out["pose2"], out["projectReference"] = pose2, proj
    td = ref.track(vtx, nrm, rv, rn, pose2, proj)
    out["track_result"] = td["result"].copy()
    out["track_error"] = td["error"].copy()
    out["track_J"] = td["J"].copy()
    red = ref.reduce(td, (w, h))
    out["reduce_8x32"] = red
    p3, conv = ref.update_pose(pose2, red, 1e-5)
    out["update_pose"], out["update_pose_converged"] = p3, np.array([conv])
    p4, ok = ref.check_pose(p3, pose2, red, (w, h))
    out["check_pose"], out["check_pose_ok"] = p4, np.array([ok])
    out["solve_x"] = ref.solve(red[0, 1:28])
    out["se3_exp"] = ref.se3_exp(out["solve_x"])
    out["inverse_of_zero"] = ref.inverse(np.zeros((4, 4), np.float32))     # NaN start-up
    out["render_depth"] = ref.render_depth(raw1)
    out["render_track"] = ref.render_track(td)
    out["render_volume"] = ref.render_volume(vol, dim, (w, h), view, largestep=0.15)
    return out


def pipeline(ref: cb.CpuKfusion) -> dict:
    """Whole-pipeline run driven like benchmark.cpp:125-150: 640x480 synthetic frames, -c 4 (160x120
    computation size), 64^3 volume, 12 frames.  Poses/flags/reduction per frame, SHA-256 of the big
    buffers, and the last frame's raycast maps + a volume slice in full."""
    n, vres, ratio = 12, 64, 4
    cw, ch = 640 // ratio, 480 // ratio
    depth, gt = synth.make_sequence(n)
    k = (K / ratio).astype(np.float32)
    out = {"frames": np.array([n]), "vres": np.array([vres]), "ratio": np.array([ratio]), "k": k}
    ref.create((cw, ch), vres, 4.8, T0, (10, 5, 4))
    poses, flags, reds, vol_sha, vtx_sha = [], [], [], [], []
    try:
        for f in range(n):
            ref.preprocessing(depth[f])
            tr = ref.tracking(k, 1e-5, 1, f)
            it = ref.integration(k, 1, 0.1, f)
            ref.raycasting(k, 0.1, f)
            poses.append(ref.get_pose().copy())
            flags.append((tr, it))
            reds.append(ref.buffer(cb.BUF_REDUCTION)[0].copy())
            vol_sha.append(sha(ref.buffer(cb.BUF_VOLUME)))
            vtx_sha.append(sha(ref.buffer(cb.BUF_VERTEX)))
        out["last_vertex"] = ref.buffer(cb.BUF_VERTEX).copy()
        out["last_normal"] = ref.buffer(cb.BUF_NORMAL).copy()
        out["last_volume_z32"] = ref.buffer(cb.BUF_VOLUME)[32].copy()
        out["last_scaled_depth2"] = ref.buffer(cb.BUF_SCALEDDEPTH, 2).copy()
    finally:
        ref.destroy()
    out["poses"] = np.stack(poses)
    out["flags"] = np.array(flags, np.uint8)
    out["reduction_row0"] = np.stack(reds)
    out["volume_sha256"] = np.array(vol_sha)
    out["vertex_sha256"] = np.array(vtx_sha)
    out["gt_poses"] = gt.astype(np.float64)
    return out


def main():
    cb.build_ref()
    if not cb.have_ref():
        raise SystemExit("oracle/_ref/libkfusion_ref.so is missing and /root/reference is not mounted")
    ref = cb.CpuKfusion(cb.REF_LIB)
    assert ref.name == "reference-cpp"
    np.savez_compressed(os.path.join(HERE, "kernels_64x48.npz"), **kernels(ref))
    np.savez_compressed(os.path.join(HERE, "pipeline_c4_v64.npz"), **pipeline(ref))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
