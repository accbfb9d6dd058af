"""Teacher-forced parity of every CUDA stage against the oracle (SURVEY §8d "parity gates").

Each stage of the B200 backend is fed the ORACLE's inputs (same depth, pose, pre-state
volume, reference maps) through the C ABI and its output is compared with the oracle's:
  * preprocessing, pyramid, vertex/normal maps, integrate, raycast: expected bit-exact
    (hard gates from north_star: TSDF +-1 LSB, raycast 1e-4 m / 1e-3 rad);
  * fused track+reduce: integer counters exact, float sums within the stated tolerance
    (summation order differs by design).
All calls go through libkfb200.so's C ABI (slambench_b200.kfusion is a ctypes binding).
"""
from __future__ import annotations

import numpy as np
import pytest

from conftest import K, T0, run_cpu_pipeline
from oracle import cpu_backend as cb
from slambench_b200 import kfusion as kf

pytestmark = pytest.mark.gpu

VRES = 128
MU = 0.1
DIM = np.array([4.8, 4.8, 4.8], np.float32)
LARGESTEP = float(np.float32(0.75) * np.float32(MU))  # raycasting(): 0.75f * mu (cpp/kernels.cpp:981)


@pytest.fixture(scope="module")
def state(port, seq16):
    """Oracle state around frame 6 of a 128^3 run: everything a teacher-forced stage needs."""
    depth, _ = seq16
    snap = {}

    def grab(f, b):
        if f == 5:
            snap["vol5"] = b.buffer(cb.BUF_VOLUME).copy()
            snap["vertex5"] = b.buffer(cb.BUF_VERTEX).copy()
            snap["normal5"] = b.buffer(cb.BUF_NORMAL).copy()
            snap["pose5"] = b.get_pose().copy()
            snap["raycastPose5"] = b.buffer(cb.BUF_RAYCASTPOSE).copy()
        if f == 6:
            snap["vol6"] = b.buffer(cb.BUF_VOLUME).copy()
            snap["vertex6"] = b.buffer(cb.BUF_VERTEX).copy()
            snap["normal6"] = b.buffer(cb.BUF_NORMAL).copy()
            snap["pose6"] = b.get_pose().copy()
            snap["floatDepth6"] = b.buffer(cb.BUF_FLOATDEPTH).copy()
            snap["sd"] = [b.buffer(cb.BUF_SCALEDDEPTH, l).copy() for l in range(3)]
            snap["inV"] = [b.buffer(cb.BUF_INVERTEX, l).copy() for l in range(3)]
            snap["inN"] = [b.buffer(cb.BUF_INNORMAL, l).copy() for l in range(3)]

    poses, tracked, integrated = run_cpu_pipeline(port, depth, 7, VRES, MU, on_frame=grab)
    assert tracked == [False] * 4 + [True] * 3 and all(integrated)
    snap["depth"] = depth
    return snap


@pytest.fixture()
def gpu():
    g = kf.Kfusion((640, 480), VRES, 4.8, T0, (10, 5, 4))
    yield g
    g.close()


def test_preprocess_bit_exact(port, gpu, state):
    d = state["depth"][6]
    gpu.preprocessing(d)
    raw = port.mm2meters(d, (640, 480))
    filt = port.bilateral(raw, port.gaussian())
    assert np.array_equal(gpu.read(kf.BUF_FLOATDEPTH), raw)
    got = gpu.read(kf.BUF_SCALEDDEPTH, 0)
    assert np.array_equal(got.view(np.uint32), filt.view(np.uint32)), f"max diff {np.abs(got - filt).max()}"
    assert np.array_equal(gpu.read(kf.BUF_GAUSSIAN), port.gaussian())


@pytest.mark.parametrize("ratio", [2, 4])
def test_preprocess_compute_size_ratio(port, state, ratio):
    """-c 2 / -c 4: mm2meters subsamples (cpp/kernels.cpp:579-587)."""
    d = state["depth"][6]
    w, h = 640 // ratio, 480 // ratio
    with kf.Kfusion((w, h), 64, 4.8, T0, (10, 5, 4)) as g:
        g.preprocessing(d)
        raw = port.mm2meters(d, (w, h))
        assert np.array_equal(g.read(kf.BUF_FLOATDEPTH), raw)
        assert np.array_equal(g.read(kf.BUF_SCALEDDEPTH, 0), port.bilateral(raw, port.gaussian()))


def test_preprocess_invalid_and_edges(port):
    """Zero (invalid) depth pixels, ragged holes and image borders (clamped taps)."""
    rng = np.random.default_rng(0)
    d = (1000 + 2000 * rng.random((480, 640))).astype(np.uint16)
    d[rng.random((480, 640)) < 0.2] = 0
    d[:3, :] = 0
    d[:, -2:] = 0
    d[100:140, 200:260] += 1500  # a depth edge larger than e_delta
    with kf.Kfusion((640, 480), 32, 4.8, T0, (10, 5, 4)) as g:
        g.preprocessing(d)
        raw = port.mm2meters(d, (640, 480))
        filt = port.bilateral(raw, port.gaussian())
        assert np.array_equal(g.read(kf.BUF_FLOATDEPTH), raw)
        assert np.array_equal(g.read(kf.BUF_SCALEDDEPTH, 0).view(np.uint32), filt.view(np.uint32))
        # and the pyramid / normals on top of a holey map (NaN from 0/0 in halfSample must match too)
        g.pyramidKernels(K)
        sd1 = port.halfsample(filt)
        sd2 = port.halfsample(sd1)
        assert np.array_equal(g.read(kf.BUF_SCALEDDEPTH, 1).view(np.uint32), sd1.view(np.uint32))
        assert np.array_equal(g.read(kf.BUF_SCALEDDEPTH, 2).view(np.uint32), sd2.view(np.uint32))


def test_preprocess_invalid_ratio_is_an_error():
    with kf.Kfusion((640, 480), 32, 4.8, T0, (10, 5, 4)) as g:
        with pytest.raises(kf.KfbError, match="Invalid ratio"):
            g.preprocessing(np.zeros((100, 100), np.uint16))


def test_create_rejects_bad_slabs_and_frees_the_context():
    """A z-slab starts on a brick layer (the brick flags and the integrate items assume it), lies inside the volume, and a
    failed create leaves nothing behind: many failures in a row must not leak device memory."""
    import torch
    with pytest.raises(kf.KfbError, match="multiple of 8"):
        kf.Kfusion((640, 480), 64, 4.8, T0, (10, 5, 4), slab=(4, 64))
    with pytest.raises(kf.KfbError, match="bad z-slab"):
        kf.Kfusion((640, 480), 64, 4.8, T0, (10, 5, 4), slab=(8, 72))
    free0 = torch.cuda.mem_get_info(0)[0]
    for _ in range(20):
        with pytest.raises(kf.KfbError):
            kf.Kfusion((640, 480), 256, 4.8, T0, (10, 5, 4), slab=(4, 256))
    assert free0 - torch.cuda.mem_get_info(0)[0] < (8 << 20), "failed creates leak device memory"
    with kf.Kfusion((640, 480), 64, 4.8, T0, (10, 5, 4), slab=(8, 56)) as g:   # an aligned inner slab is fine
        g.synchroniseDevices()


def test_pyramid_vertex_normal_bit_exact(port, gpu, state):
    gpu.write(kf.BUF_SCALEDDEPTH, state["sd"][0], 0)
    gpu.pyramidKernels(K)
    for l in range(3):
        if l:
            assert np.array_equal(gpu.read(kf.BUF_SCALEDDEPTH, l).view(np.uint32), state["sd"][l].view(np.uint32)), f"depth level {l}"
        assert np.array_equal(gpu.read(kf.BUF_INVERTEX, l).view(np.uint32), state["inV"][l].view(np.uint32)), f"vertex level {l}"
        gn, on = gpu.read(kf.BUF_INNORMAL, l), state["inN"][l]
        assert np.array_equal(gn[..., 0], on[..., 0]), f"normal.x level {l}"
        valid = on[..., 0] != -2
        assert np.array_equal(gn[valid].view(np.uint32), on[valid].view(np.uint32)), f"normal level {l}"


def _exact_sums(td, w, h):
    """float64 reference of the 32 reduction outputs from the oracle's per-pixel TrackData."""
    t = td[:h, :w]
    ok = t["result"] == 1
    e = t["error"][ok].astype(np.float64)
    J = t["J"][ok].astype(np.float64)
    # products are formed in fp32 by the reference (row.error * row.J[i]); mirror that, sum in fp64
    e32, J32 = t["error"][ok], t["J"][ok]
    out = np.zeros(32)
    out[0] = (e32 * e32).astype(np.float64).sum()
    for i in range(6):
        out[1 + i] = (e32 * J32[:, i]).astype(np.float64).sum()
    q = 7
    for a in range(6):
        for b in range(a, 6):
            out[q] = (J32[:, a] * J32[:, b]).astype(np.float64).sum()
            q += 1
    out[28] = ok.sum()
    out[29] = (t["result"] == -4).sum()
    out[30] = (t["result"] == -5).sum()
    out[31] = ((t["result"] < 1) & (t["result"] > -4)).sum()
    del e, J
    return out


@pytest.mark.parametrize("level", [2, 1, 0])
def test_track_reduce(port, gpu, state, level):
    w, h = 640 >> level, 480 >> level
    Kmat = port.camera_matrix(K)
    view = port.matmul(Kmat, port.inverse(state["raycastPose5"]))
    pose = state["pose5"]
    td = port.track(state["inV"][level], state["inN"][level], state["vertex5"], state["normal5"], pose, view)
    red = port.reduce(td, (w, h))[0]
    exact = _exact_sums(td, w, h)

    for l in range(3):
        gpu.write(kf.BUF_INVERTEX, state["inV"][l], l)
        gpu.write(kf.BUF_INNORMAL, state["inN"][l], l)
    gpu.write(kf.BUF_VERTEX, state["vertex5"])
    gpu.write(kf.BUF_NORMAL, state["normal5"])
    got = gpu.trackReduceKernel(level, pose, view)

    # per-pixel decisions are bit-exact => the four counters match exactly
    assert np.array_equal(got[28:32], red[28:32]), (got[28:32], red[28:32])
    assert got[28] > 0.5 * w * h
    # sums: fp64 tree here vs serial fp32 in the reference; both must agree with the exact value,
    # ours at fp32 rounding (1e-6), the reference within its own accumulated error
    scale = np.maximum(np.abs(exact[:28]), 1e-3 * np.abs(exact[:28]).max())
    assert np.max(np.abs(got[:28] - exact[:28]) / scale) < 2e-6
    assert np.max(np.abs(got[:28] - red[:28]) / scale) < 2e-3
    # and the resulting pose update agrees with the reference's to far better than 1e-4 m / rad
    p_ref, c_ref = port.update_pose(pose, port.reduce(td, (w, h)))
    p_got, c_got = gpu.updatePoseKernel(pose, got, 1e-5)
    assert c_ref == c_got
    assert np.abs(p_ref - p_got).max() < 2e-6


def test_track_startup_nan_reference(port, gpu, state):
    """Frames 0-3: raycastPose is the zero matrix, inverse() is NaN, every valid pixel lands in -4."""
    zero = np.zeros((4, 4), np.float32)
    view = port.matmul(port.camera_matrix(K), port.inverse(zero))
    assert not np.isfinite(view).all()
    view_gpu = gpu.matmul(gpu.cameraMatrix(K), gpu.inverse(zero))
    assert np.array_equal(np.isnan(view_gpu), np.isnan(view))
    zeros = np.zeros((480, 640, 3), np.float32)
    pose = kf.identity_pose(T0)
    td = port.track(state["inV"][0], state["inN"][0], zeros, zeros, pose, view)
    red = port.reduce(td, (640, 480))[0]
    gpu.write(kf.BUF_INVERTEX, state["inV"][0], 0)
    gpu.write(kf.BUF_INNORMAL, state["inN"][0], 0)
    got = gpu.trackReduceKernel(0, pose, view_gpu)
    assert np.array_equal(got, red)
    assert got[28] == 0 and got[29] > 0
    p, ok = gpu.checkPoseKernel(pose, pose, got, (640, 480))
    p2, ok2 = port.check_pose(pose, pose, red, (640, 480))
    assert ok is False and ok2 is False


def test_integrate_bit_exact(port, gpu, state):
    pose = state["pose6"]
    inv, Kmat = port.inverse(pose), port.camera_matrix(K)
    want = state["vol5"].copy()
    port.integrate(want, DIM, state["floatDepth6"], inv, Kmat, MU)
    assert np.array_equal(want, state["vol6"])  # the oracle's free kernel reproduces its own pipeline
    gpu.write(kf.BUF_VOLUME, state["vol5"])
    gpu.write(kf.BUF_FLOATDEPTH, state["floatDepth6"])
    gpu.reset_stats()
    gpu.integrateKernel(gpu.inverse(pose), gpu.cameraMatrix(K), MU)
    got = gpu.read(kf.BUF_VOLUME)
    diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert diff.max() <= 1, f"TSDF parity gate (+-1 LSB) violated: max {diff.max()}"
    assert np.array_equal(got, want), f"{int((diff > 0).sum())} voxels differ (all within 1 LSB)"
    # N_upd is counted exactly: every updated voxel had its weight bumped (weights < maxweight here)
    n_upd = int((want[..., 1] != state["vol5"][..., 1]).sum())
    assert gpu.stats()["voxels_updated_last"] == n_upd


def test_integrate_weight_saturation_and_empty_depth(port):
    """maxweight clamp (w stays at 100) and an all-invalid depth map (no voxel may change)."""
    with kf.Kfusion((640, 480), 64, 4.8, T0, (10, 5, 4)) as g:
        pose = kf.identity_pose(T0)
        inv, Kmat = g.inverse(pose), g.cameraMatrix(K)
        vol = port.init_volume((64, 64, 64))
        vol[..., 1] = 99
        depth = np.full((480, 640), 2.5, np.float32)
        want = vol.copy()
        for _ in range(3):
            port.integrate(want, DIM, depth, port.inverse(pose), port.camera_matrix(K), MU)
        g.write(kf.BUF_VOLUME, vol)
        g.write(kf.BUF_FLOATDEPTH, depth)
        for _ in range(3):
            g.integrateKernel(inv, Kmat, MU)
        got = g.read(kf.BUF_VOLUME)
        assert np.array_equal(got, want)
        assert got[..., 1].max() == 100
        g.write(kf.BUF_FLOATDEPTH, np.zeros((480, 640), np.float32))
        g.integrateKernel(inv, Kmat, MU)
        assert np.array_equal(g.read(kf.BUF_VOLUME), want)
        assert g.stats()["voxels_updated_last"] == 0


def _angle(a, b):
    c = np.clip((a * b).sum(-1), -1, 1)
    return np.arccos(c)


def test_raycast_parity(port, gpu, state):
    pose = state["pose6"]
    view = port.matmul(pose, port.inverse_camera_matrix(K))
    want_v, want_n = port.raycast(state["vol6"], DIM, (640, 480), view, largestep=LARGESTEP,
                                  init=(state["vertex5"], state["normal5"]))
    assert np.array_equal(want_v, state["vertex6"])
    gpu.write(kf.BUF_VOLUME, state["vol6"])
    gpu.write(kf.BUF_VERTEX, state["vertex5"])
    gpu.write(kf.BUF_NORMAL, state["normal5"])
    gpu.raycastKernel(gpu.matmul(pose, gpu.inverseCameraMatrix(K)), largestep=LARGESTEP)
    got_v, got_n = gpu.read(kf.BUF_VERTEX), gpu.read(kf.BUF_NORMAL)
    hit_w, hit_g = want_n[..., 0] != -2, got_n[..., 0] != -2
    assert np.array_equal(hit_w, hit_g), f"{int((hit_w != hit_g).sum())} pixels differ in hit/miss"
    assert hit_w.mean() > 0.9
    assert np.abs(got_v - want_v).max() <= 1e-4, "raycast vertex gate (1e-4 m)"
    assert _angle(got_n[hit_w], want_n[hit_w]).max() <= 1e-3, "raycast normal gate (1e-3 rad)"
    # expected: bit-exact
    assert np.array_equal(got_v.view(np.uint32), want_v.view(np.uint32))
    assert np.array_equal(got_n[hit_w].view(np.uint32), want_n[hit_w].view(np.uint32))


def test_raycast_empty_volume_all_miss(port):
    with kf.Kfusion((640, 480), 32, 4.8, T0, (10, 5, 4)) as g:
        view = g.matmul(kf.identity_pose(T0), g.inverseCameraMatrix(K))
        g.raycastKernel(view)
        assert np.all(g.read(kf.BUF_VERTEX) == 0)
        n = g.read(kf.BUF_NORMAL)
        assert np.all(n[..., 0] == -2) and np.all(n[..., 1:] == 0)


def test_reset_restores_initial_volume(port):
    with kf.Kfusion((640, 480), 48, 4.8, T0, (10, 5, 4)) as g:
        init = port.init_volume((48, 48, 48))
        assert np.array_equal(g.read(kf.BUF_VOLUME), init)
        g.write(kf.BUF_FLOATDEPTH, np.full((480, 640), 2.0, np.float32))
        g.integrateKernel(g.inverse(kf.identity_pose(T0)), g.cameraMatrix(K), MU)
        assert not np.array_equal(g.read(kf.BUF_VOLUME), init)
        g.reset()
        assert np.array_equal(g.read(kf.BUF_VOLUME), init)


def test_render_kernels(port, gpu, state):
    gpu.write(kf.BUF_FLOATDEPTH, state["floatDepth6"])
    assert np.array_equal(gpu.renderDepth(), port.render_depth(state["floatDepth6"]))
    gpu.write(kf.BUF_VOLUME, state["vol6"])
    gpu.setPose(state["pose6"])
    view = port.matmul(state["pose6"], port.inverse_camera_matrix(K))
    want = port.render_volume(state["vol6"], DIM, (640, 480), view, largestep=LARGESTEP)
    got = gpu.renderVolume(0, 4, K, LARGESTEP)
    # shading goes through a float->uchar cast of a product; allow an off-by-one on a handful of pixels
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3
    # renderTrack: status plane of the last track launch vs the oracle's TrackData
    Kmat = port.camera_matrix(K)
    pr = port.matmul(Kmat, port.inverse(state["raycastPose5"]))
    td = port.track(state["inV"][0], state["inN"][0], state["vertex5"], state["normal5"], state["pose5"], pr)
    gpu.renderTrack()  # switches the status plane on
    gpu.write(kf.BUF_INVERTEX, state["inV"][0], 0)
    gpu.write(kf.BUF_INNORMAL, state["inN"][0], 0)
    gpu.write(kf.BUF_VERTEX, state["vertex5"])
    gpu.write(kf.BUF_NORMAL, state["normal5"])
    gpu.trackReduceKernel(0, state["pose5"], pr)
    assert np.array_equal(gpu.read(kf.BUF_TRACKSTATUS), td["result"].astype(np.int8))
    assert np.array_equal(gpu.renderTrack(), port.render_track(td))


def _random_pose(rng, scale=1.0):
    from slambench_b200 import synth

    pose = np.eye(4, dtype=np.float32)
    pose[:3, :3] = synth.rpy_to_R(*(rng.uniform(-1, 1, 3) * scale)).astype(np.float32)
    pose[:3, 3] = rng.uniform(-0.5, 5.3, 3)   # sometimes outside the volume cube
    return pose


@pytest.mark.parametrize("vres", [96, 256])
def test_integrate_culling_is_exact(port, seq16, vres):
    """The conservative z-interval + approximate-division fast paths must not change one voxel:
    compare with the same kernel visiting every voxel with the reference's full expression
    (KFB_FLAG_INTEGRATE_NO_CULL), over random poses (camera inside/outside the cube, large rotations),
    holey depth maps and accumulated weights; at 96^3 also against the oracle."""
    depth, _ = seq16
    rng = np.random.default_rng(vres)
    d = depth[7].copy()
    d[rng.random(d.shape) < 0.05] = 0
    d[200:260, 300:420] = 700
    with kf.Kfusion((640, 480), vres, 4.8, T0, (10, 5, 4)) as a, \
            kf.Kfusion((640, 480), vres, 4.8, T0, (10, 5, 4), flags=kf.FLAG_INTEGRATE_NO_CULL) as b:
        a.preprocessing(d)      # also produces max(depth) for the far cull
        b.preprocessing(d)
        want = port.init_volume((vres,) * 3) if vres <= 96 else None
        raw = port.mm2meters(d, (640, 480))
        poses = [kf.identity_pose(T0)] + [_random_pose(rng, s) for s in (0.05, 0.05, 0.3, 0.3, 1.0, 1.0, 3.0)]
        total = 0
        for i, pose in enumerate(poses):
            mu = (0.1, 0.05, 0.3)[i % 3]
            a.reset_stats(); b.reset_stats()
            a.integrateKernel(a.inverse(pose), a.cameraMatrix(K), mu)
            b.integrateKernel(b.inverse(pose), b.cameraMatrix(K), mu)
            na, nb = a.stats()["voxels_updated_last"], b.stats()["voxels_updated_last"]
            assert na == nb, f"pose {i}: N_upd {na} != {nb}"
            total += na
            va, vb = a.read(kf.BUF_VOLUME), b.read(kf.BUF_VOLUME)
            assert np.array_equal(va, vb), f"pose {i}: {int((va != vb).any(-1).sum())} voxels differ"
            if want is not None:
                port.integrate(want, DIM, raw, port.inverse(pose), port.camera_matrix(K), mu)
                assert np.array_equal(va, want), f"pose {i} vs oracle"
        assert total > 0.05 * vres ** 3


def test_integrate_culling_exact_with_teacher_forced_depth(port, gpu, state):
    """Depth written from the host (no cached max(depth) => no far cull) and a non-cubic volume."""
    rng = np.random.default_rng(9)
    res, dim = (80, 48, 112), np.array([4.8, 2.9, 4.8], np.float32)
    with kf.Kfusion((640, 480), res, dim, T0, (10, 5, 4)) as g:
        want = port.init_volume(res)
        g.write(kf.BUF_FLOATDEPTH, state["floatDepth6"])
        for s in (0.0, 0.1, 0.6):
            pose = kf.identity_pose(T0) if s == 0 else _random_pose(rng, s)
            port.integrate(want, dim, state["floatDepth6"], port.inverse(pose), port.camera_matrix(K), MU)
            g.integrateKernel(g.inverse(pose), g.cameraMatrix(K), MU)
            assert np.array_equal(g.read(kf.BUF_VOLUME), want)
