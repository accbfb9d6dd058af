"""Free-running parity of the whole per-frame pipeline (preprocess -> track -> integrate ->
raycast) against the oracle on the same synthetic frames, driven exactly like
kfusion/src/benchmark.cpp:125-150 drives `Kfusion`.

Hard gates (north_star): per-frame pose within 1e-4 m / 1e-4 rad, identical tracked and
integrated flags.  TSDF / raycast maps are reported as the fraction inside tolerance:
a ~1e-7 pose difference legitimately flips `(uint)pixel` for voxels that project onto a
pixel edge.
"""
from __future__ import annotations

import numpy as np
import pytest

from conftest import K, T0, run_cpu_pipeline
from oracle import cpu_backend as cb
from slambench_b200 import kfusion as kf

pytestmark = pytest.mark.gpu


def rot_angle(Ra, Rb):
    """Rotation angle between two rotation matrices.  arccos((tr-1)/2) cannot resolve angles below
    ~5e-4 rad on float32 matrices (the trace carries ~1e-7 of rounding), so use the skew part:
    sin(theta) = |vee(M - M^T)| / 2."""
    M = Ra.astype(np.float64).T @ Rb.astype(np.float64)
    s = 0.5 * np.sqrt((M[2, 1] - M[1, 2]) ** 2 + (M[0, 2] - M[2, 0]) ** 2 + (M[1, 0] - M[0, 1]) ** 2)
    return float(np.arcsin(min(1.0, s)))


def run_gpu_pipeline(depth, n_frames, vres, mu=0.1, csize=(640, 480), pyramid=(10, 5, 4), flags=0, on_frame=None):
    poses, tracked, integrated = [], [], []
    with kf.Kfusion(csize, vres, 4.8, T0, pyramid, flags=flags) as g:
        for f in range(n_frames):
            g.preprocessing(depth[f])
            tr = g.tracking(K, 1e-5, 1, f)
            it = g.integration(K, 1, mu, f)
            g.raycasting(K, mu, f)
            poses.append(g.getPose().copy())
            tracked.append(tr)
            integrated.append(it)
            if on_frame is not None:
                on_frame(f, g)
    return np.stack(poses), tracked, integrated


@pytest.mark.parametrize("vres,n_frames", [(128, 16), (256, 10)])
def test_free_running_pose_and_flags(port, seq16, vres, n_frames):
    depth, gt = seq16
    cpu_snap, gpu_snap = {}, {}

    def grab_cpu(f, b):
        if f == n_frames - 1:
            cpu_snap["vol"] = b.buffer(cb.BUF_VOLUME).copy()
            cpu_snap["vertex"] = b.buffer(cb.BUF_VERTEX).copy()
            cpu_snap["normal"] = b.buffer(cb.BUF_NORMAL).copy()

    def grab_gpu(f, g):
        if f == n_frames - 1:
            gpu_snap["vol"] = g.read(kf.BUF_VOLUME)
            gpu_snap["vertex"] = g.read(kf.BUF_VERTEX)
            gpu_snap["normal"] = g.read(kf.BUF_NORMAL)

    p_cpu, t_cpu, i_cpu = run_cpu_pipeline(port, depth, n_frames, vres, on_frame=grab_cpu)
    p_gpu, t_gpu, i_gpu = run_gpu_pipeline(depth, n_frames, vres, on_frame=grab_gpu)

    assert t_gpu == t_cpu, "tracked flags differ"
    assert i_gpu == i_cpu, "integrated flags differ"
    assert t_cpu[:4] == [False] * 4 and all(t_cpu[4:]), "start-up: frames 0-3 untracked, then tracked"
    assert all(i_cpu)
    for f in range(n_frames):
        dt = np.abs(p_gpu[f][:3, 3] - p_cpu[f][:3, 3]).max()
        dr = rot_angle(p_gpu[f][:3, :3], p_cpu[f][:3, :3])
        assert dt <= 1e-4 and dr <= 1e-4, f"frame {f}: pose differs by {dt} m / {dr} rad"
    # and both follow the ground truth of the synthetic trajectory
    err = np.abs(p_gpu[-1][:3, 3] - gt[n_frames - 1][:3, 3]).max()
    assert err < 5e-3, f"drift from ground truth {err} m"

    dv = np.abs(gpu_snap["vol"].astype(np.int32) - cpu_snap["vol"].astype(np.int32))
    frac_vol = float((dv.max(axis=-1) <= 1).mean())
    hit = cpu_snap["normal"][..., 0] != -2
    same_mask = float(((gpu_snap["normal"][..., 0] != -2) == hit).mean())
    both = hit & (gpu_snap["normal"][..., 0] != -2)
    frac_vtx = float((np.abs(gpu_snap["vertex"] - cpu_snap["vertex"]).max(-1)[both] <= 1e-4).mean())
    print(f"\n[free-running {vres}^3 x{n_frames}] max pose diff "
          f"{max(np.abs(p_gpu[f][:3, 3] - p_cpu[f][:3, 3]).max() for f in range(n_frames)):.2e} m; "
          f"TSDF within 1 LSB: {frac_vol:.6f}; hit mask equal: {same_mask:.6f}; vertex within 1e-4: {frac_vtx:.6f}")
    assert frac_vol > 0.999 and same_mask > 0.999 and frac_vtx > 0.99


def test_tracking_rate_and_integration_rate_gates(port, seq16):
    """-t 2 / -r 2: frame % rate gates (cpp/kernels.cpp:927, 994)."""
    depth, _ = seq16
    n = 10
    flags_cpu, flags_gpu = [], []
    port.create((640, 480), 64, 4.8, T0, (10, 5, 4))
    try:
        for f in range(n):
            port.preprocessing(depth[f])
            tr = port.tracking(K, 1e-5, 2, f)
            it = port.integration(K, 2, 0.2, f)
            port.raycasting(K, 0.2, f)
            flags_cpu.append((tr, it))
    finally:
        port.destroy()
    with kf.Kfusion((640, 480), 64, 4.8, T0, (10, 5, 4)) as g:
        for f in range(n):
            g.preprocessing(depth[f])
            tr = g.tracking(K, 1e-5, 2, f)
            it = g.integration(K, 2, 0.2, f)
            g.raycasting(K, 0.2, f)
            flags_gpu.append((tr, it))
    assert flags_gpu == flags_cpu


def test_tracking_loss_restores_pose(port, seq16):
    """A frame that matches nothing in the model must be rejected and the pose rolled back
    (checkPoseKernel, cpp/kernels.cpp:777-792); the next good frame tracks again."""
    depth, _ = seq16
    bad = np.full_like(depth[0], 600)  # a wall 0.6 m in front of the camera: no pixel within dist_threshold
    frames = [depth[f] for f in range(7)] + [bad] + [depth[7], depth[8]]
    res_cpu, res_gpu = [], []
    port.create((640, 480), 128, 4.8, T0, (10, 5, 4))
    try:
        for f, d in enumerate(frames):
            port.preprocessing(d)
            tr = port.tracking(K, 1e-5, 1, f)
            it = port.integration(K, 1, 0.1, f)
            port.raycasting(K, 0.1, f)
            res_cpu.append((tr, it, port.get_pose().copy()))
    finally:
        port.destroy()
    with kf.Kfusion((640, 480), 128, 4.8, T0, (10, 5, 4)) as g:
        for f, d in enumerate(frames):
            g.preprocessing(d)
            tr = g.tracking(K, 1e-5, 1, f)
            it = g.integration(K, 1, 0.1, f)
            g.raycasting(K, 0.1, f)
            res_gpu.append((tr, it, g.getPose().copy()))
    assert [r[:2] for r in res_gpu] == [r[:2] for r in res_cpu]
    assert res_cpu[7][0] is False and res_cpu[7][1] is False, "the bad frame must be rejected by the reference logic"
    assert np.array_equal(res_gpu[7][2], res_gpu[6][2]), "pose must be restored to the previous frame's"
    for a, b in zip(res_gpu, res_cpu):
        assert np.abs(a[2] - b[2]).max() <= 1e-4


def test_full_size_512_integrate_properties(port, seq16):
    """BASELINE config 2 size (512^3): one teacher-forced integrate vs the oracle (bit-exact), plus
    size-independent properties: weights grow by exactly one where (and only where) the voxel was
    updated, N_upd equals that count, and a second identical integrate updates the same voxel set."""
    depth, _ = seq16
    N = 512
    pose = kf.identity_pose(T0)
    raw = port.mm2meters(depth[0], (640, 480))
    want = port.init_volume((N, N, N))
    port.integrate(want, np.array([4.8] * 3, np.float32), raw, port.inverse(pose), port.camera_matrix(K), 0.1)
    with kf.Kfusion((640, 480), N, 4.8, T0, (10, 5, 4)) as g:
        g.preprocessing(depth[0])
        g.reset_stats()
        g.integrateKernel(g.inverse(pose), g.cameraMatrix(K), 0.1)
        n1 = g.stats()["voxels_updated_last"]
        got = g.read(kf.BUF_VOLUME)
        assert np.array_equal(got, want)
        assert n1 == int((got[..., 1] == 1).sum())
        assert 0.02 * N**3 < n1 < 0.5 * N**3
        g.integrateKernel(g.inverse(pose), g.cameraMatrix(K), 0.1)
        st = g.stats()
        assert st["voxels_updated_last"] == n1 and st["voxels_updated_total"] == 2 * n1
        got2 = g.read(kf.BUF_VOLUME)
        assert int((got2[..., 1] == 2).sum()) == n1 and int((got2[..., 1] == 1).sum()) == 0


def test_device_resident_icp_matches_host_loop(seq16):
    """Default: the whole ICP schedule runs on the device (solve + SE3 exp + per-level break in the last CTA,
    certified-Cholesky fast path).  KFB_FLAG_ICP_HOST_SOLVE: one launch + host Jacobi solve per iteration,
    the reference's control flow verbatim.  Same poses (<= 2e-6), same flags, same iteration counts."""
    depth, _ = seq16
    n = 14
    its = {}

    def count(tag):
        def f(fr, g):
            its.setdefault(tag, []).append(g.stats()["icp_iterations_last"])
        return f

    p_dev, t_dev, i_dev = run_gpu_pipeline(depth, n, 128, on_frame=count("dev"))
    p_host, t_host, i_host = run_gpu_pipeline(depth, n, 128, flags=kf.FLAG_ICP_HOST_SOLVE, on_frame=count("host"))
    assert t_dev == t_host and i_dev == i_host
    assert np.abs(p_dev - p_host).max() <= 2e-6
    assert its["dev"] == its["host"], (its["dev"], its["host"])
    assert its["dev"][:4] == [3, 3, 3, 3] and max(its["dev"]) <= 19


def test_zero_iteration_pyramid_levels(port, seq16):
    """-y 0,0,4 style schedules: levels with zero iterations are skipped (cpp/kernels.cpp:952)."""
    depth, _ = seq16
    pyr = (3, 0, 2)
    p_cpu, t_cpu, i_cpu = run_cpu_pipeline(port, depth, 8, 64, pyramid=pyr)
    p_gpu, t_gpu, i_gpu = run_gpu_pipeline(depth, 8, 64, pyramid=pyr)
    assert t_gpu == t_cpu and i_gpu == i_cpu
    assert np.abs(p_gpu - p_cpu).max() <= 1e-4


def test_raycast_brick_skipping_is_exact(seq16):
    """Brick flags let the raycaster step over samples proven >= 0.82 without reading the volume.  The maps, poses
    and flags of a whole run must be bit-identical to the run that evaluates every sample (KFB_FLAG_RAYCAST_NO_SKIP),
    also for the 2x-far-plane volume render and after the volume is replaced from the host (flags rebuilt)."""
    depth, _ = seq16
    n = 12
    snaps = {}

    def grab(tag):
        def f(fr, g):
            if fr >= 3:
                snaps.setdefault(tag, []).append((g.read(kf.BUF_VERTEX).copy(), g.read(kf.BUF_NORMAL).copy()))
            if fr == n - 1:
                snaps[tag + "_render"] = g.renderVolume(0, 1, K, 0.075)
                snaps[tag + "_vol"] = g.read(kf.BUF_VOLUME)
        return f

    pa, ta, ia = run_gpu_pipeline(depth, n, 256, on_frame=grab("skip"))
    pb, tb, ib = run_gpu_pipeline(depth, n, 256, flags=kf.FLAG_RAYCAST_NO_SKIP, on_frame=grab("full"))
    assert ta == tb and ia == ib and np.array_equal(pa, pb)
    for (va, na), (vb, nb) in zip(snaps["skip"], snaps["full"]):
        assert np.array_equal(va.view(np.uint32), vb.view(np.uint32)) and np.array_equal(na.view(np.uint32), nb.view(np.uint32))
    assert np.array_equal(snaps["skip_render"], snaps["full_render"])
    assert np.array_equal(snaps["skip_vol"], snaps["full_vol"])
    # host-written volume: the flags are rebuilt from it
    with kf.Kfusion((640, 480), 256, 4.8, T0, (10, 5, 4)) as a, kf.Kfusion((640, 480), 256, 4.8, T0, (10, 5, 4), flags=kf.FLAG_RAYCAST_NO_SKIP) as b:
        view = a.matmul(pa[-1], a.inverseCameraMatrix(K))
        for g in (a, b):
            g.write(kf.BUF_VOLUME, snaps["full_vol"])
            g.raycastKernel(view)
        assert np.array_equal(a.read(kf.BUF_VERTEX).view(np.uint32), b.read(kf.BUF_VERTEX).view(np.uint32))
        assert np.array_equal(a.read(kf.BUF_NORMAL).view(np.uint32), b.read(kf.BUF_NORMAL).view(np.uint32))
        assert (a.read(kf.BUF_NORMAL)[..., 0] != -2).mean() > 0.9


@pytest.mark.parametrize("res,dim,csize,pyr,pose_tol", [
    ((100, 60, 76), (4.8, 2.9, 3.7), (640, 480), (10, 5, 4), 1e-3),   # ragged volume: not multiples of 8 / 32, non-cubic, 5 cm voxels
    ((64, 64, 64), (4.8, 4.8, 4.8), (320, 240), (6, 3), 1e-4),        # -c 2 computation size, two pyramid levels
    ((72, 72, 40), (4.8, 4.8, 4.8), (160, 120), (5,), 1e-4),          # -c 4, a single level
])
def test_odd_configurations(port, seq16, res, dim, csize, pyr, pose_tol):
    """Whole pipeline on awkward sizes (ragged volumes exercise the brick-flag and work-list edges; smaller
    computation sizes the mm2meters ratios; short pyramids the ICP schedule).
    (1) A/B: the culled / brick-skipping kernels against the same library visiting every voxel and every ray
        sample — bit-identical poses, volume and maps;
    (2) against the oracle: identical flags, pose within tolerance (the 5 cm-voxel case is ill-conditioned: a
        1e-7 difference in the ICP sums is amplified, so its bound is looser)."""
    depth, _ = seq16
    n = 9
    ratio = 640 // csize[0]
    k = (K / ratio).astype(np.float32)
    port.create(csize, res, dim, T0, pyr)
    try:
        pc = []
        for f in range(n):
            port.preprocessing(depth[f])
            tr = port.tracking(k, 1e-5, 1, f)
            it = port.integration(k, 1, 0.1, f)
            port.raycasting(k, 0.1, f)
            pc.append((tr, it, port.get_pose().copy()))
        vol_c = port.buffer(cb.BUF_VOLUME).copy()
        nrm_c = port.buffer(cb.BUF_NORMAL).copy()
    finally:
        port.destroy()
    runs = {}
    for tag, flags in (("fast", 0), ("full", kf.FLAG_RAYCAST_NO_SKIP | kf.FLAG_INTEGRATE_NO_CULL)):
        with kf.Kfusion(csize, res, dim, T0, pyr, flags=flags) as g:
            poses, fl = [], []
            for f in range(n):
                g.preprocessing(depth[f])
                tr = g.tracking(k, 1e-5, 1, f)
                it = g.integration(k, 1, 0.1, f)
                g.raycasting(k, 0.1, f)
                poses.append(g.getPose().copy())
                fl.append((tr, it))
            runs[tag] = (np.stack(poses), fl, g.read(kf.BUF_VOLUME), g.read(kf.BUF_VERTEX), g.read(kf.BUF_NORMAL))
    a, b = runs["fast"], runs["full"]
    assert a[1] == b[1] and np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    assert np.array_equal(a[3].view(np.uint32), b[3].view(np.uint32)) and np.array_equal(a[4].view(np.uint32), b[4].view(np.uint32))
    assert a[1] == [t[:2] for t in pc]
    for f in range(n):
        assert np.abs(a[0][f] - pc[f][2]).max() <= pose_tol, f"frame {f}"
    assert a[2].shape == vol_c.shape
    d = np.abs(a[2].astype(np.int32) - vol_c.astype(np.int32)).max(-1)
    assert (d <= 1).mean() > 0.99
    assert ((a[4][..., 0] != -2) == (nrm_c[..., 0] != -2)).mean() > 0.99


@pytest.mark.parametrize("vres", [128, 200])
def test_raycast_bulk_staging_is_exact(seq16, vres, monkeypatch):
    """The bulk-async staging experiment (KFB_RAY_BULK=1: flagged bricks copied into shared memory by cp.async.bulk + mbarrier,
    taps read from there) must produce the same maps, poses and volume as the __ldg gather path, bit for bit
    (200^3: bricks cut by the volume's edge)."""
    depth, _ = seq16
    res = []
    for bulk in ("0", "1"):
        monkeypatch.setenv("KFB_RAY_BULK", bulk)
        with kf.Kfusion((640, 480), vres, 4.8, T0, (10, 5, 4)) as g:
            for f in range(10):
                g.computeFrame(depth[f], None, K, 1, 1, 1e-5, 0.1, f)
            res.append((g.getPose().copy(), g.read(kf.BUF_VERTEX), g.read(kf.BUF_NORMAL), g.read(kf.BUF_VOLUME)))
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a, b.view(np.uint32) if b.dtype == np.float32 else b)
    assert (res[1][2][..., 0] != -2).mean() > 0.5, "nothing was hit: the staged path was not exercised"


@pytest.mark.parametrize("rates", [(1, 1), (2, 3)])
def test_compute_frame_entry_point(seq16, rates):
    """kfb_compute_frame == preprocessing + tracking + integration + raycasting (cpp/kernels.cpp:1048-1055), bit for bit:
    the whole-frame enqueue (checkPose / inverse(pose) / raycastPose * invK evaluated by the ICP kernel's last CTA, integrate
    gated on the device) must leave the same poses, flags, volume and raycast maps as the staged calls, including frames
    where tracking or integration is skipped by its rate and a frame where tracking is lost."""
    depth, _ = seq16
    irate, trate = rates
    n = 10
    lost = depth[7].copy()
    lost[:] = 600                                  # a wall 0.6 m away: ICP cannot explain it -> tracking lost, pose restored
    frames = [lost if f == 7 else depth[f] for f in range(n)]
    res = []
    for one_call in (False, True):
        with kf.Kfusion((640, 480), 96, 4.8, T0, (10, 5, 4)) as g:
            out = []
            for f in range(n):
                if one_call:
                    g.computeFrame(frames[f], None, K, irate, trate, 1e-5, 0.1, f)
                    tr, it = g.getTracked(), g.getIntegrated()
                else:
                    g.preprocessing(frames[f])
                    tr = g.tracking(K, 1e-5, trate, f)
                    it = g.integration(K, irate, 0.1, f)
                    g.raycasting(K, 0.1, f)
                out.append((tr, it, g.getPose().copy()))
            res.append((out, g.read(kf.BUF_VOLUME), g.read(kf.BUF_VERTEX), g.read(kf.BUF_NORMAL), g.read(kf.BUF_RAYCASTPOSE),
                        g.stats()["frames_integrated"]))
            if one_call:
                assert np.array_equal(g.getPosition(), g.getPose()[:3, 3] - T0)
    (a, va, xa, na, ra, ca), (b, vb, xb, nb, rb, cb_) = res
    for f in range(n):
        assert a[f][:2] == b[f][:2], f"frame {f}: flags {b[f][:2]} != staged {a[f][:2]}"
        assert np.array_equal(a[f][2], b[f][2]), f"frame {f}: pose differs"
    if rates == (1, 1):
        assert a[7][0] is False and a[6][0] is True, "the tracking-loss frame was not exercised"
    assert np.array_equal(va, vb) and np.array_equal(xa, xb) and np.array_equal(na, nb) and np.array_equal(ra, rb)
    assert ca == cb_ == sum(x[1] for x in a)


@pytest.mark.parametrize("pyr", [(4, 3, 3, 2, 2), (3, 0, 2, 2)])
def test_more_than_three_pyramid_levels(port, seq16, pyr):
    """The reference accepts any `-y` list (default_parameters.h:394-396): 4 and 5 levels, and a level with no iterations."""
    depth, _ = seq16
    n = 8
    pc, tc, ic = run_cpu_pipeline(port, depth, n, 64, pyramid=pyr)
    pg, tg, ig = run_gpu_pipeline(depth, n, 64, pyramid=pyr)
    assert tg == tc and ig == ic
    assert any(tc), "nothing tracked: the configuration does not exercise the deeper levels"
    for f in range(n):
        assert np.abs(pg[f] - pc[f]).max() <= 1e-4, f"frame {f}"
    # the deepest level's maps, teacher-forced: bit-exact
    with kf.Kfusion((640, 480), 64, 4.8, T0, pyr) as g:
        g.preprocessing(depth[5])
        g.pyramidKernels(K)
        lvl = len(pyr) - 1
        raw = port.mm2meters(depth[5], (640, 480))
        d = port.bilateral(raw, port.gaussian())
        for _ in range(lvl):
            d = port.halfsample(d)
        assert np.array_equal(g.read(kf.BUF_SCALEDDEPTH, lvl), d)
        v = port.depth2vertex(d, port.inverse_camera_matrix(K / np.float32(1 << lvl)))
        assert np.array_equal(g.read(kf.BUF_INVERTEX, lvl), v)
        nrm = port.vertex2normal(v)
        got = g.read(kf.BUF_INNORMAL, lvl)
        valid = nrm[..., 0] != -2
        assert np.array_equal(got[..., 0] == -2, ~valid) and np.array_equal(got[valid], nrm[valid])


def test_compute_size_ratio_8(port, seq16):
    """`-c 8` (default_parameters.h:274-277): 80x60 computation size from the 640x480 sensor frame, whole pipeline."""
    depth, _ = seq16
    n, k8 = 8, (K / np.float32(8)).astype(np.float32)
    pc, tc, ic = run_cpu_pipeline(port, depth, n, 64, csize=(80, 60), k=k8)
    poses, tr, it = [], [], []
    with kf.Kfusion((80, 60), 64, 4.8, T0, (10, 5, 4)) as g:
        for f in range(n):
            g.preprocessing(depth[f])
            tr.append(g.tracking(k8, 1e-5, 1, f))
            it.append(g.integration(k8, 1, 0.1, f))
            g.raycasting(k8, 0.1, f)
            poses.append(g.getPose().copy())
        raw = port.mm2meters(depth[n - 1], (80, 60))
        assert np.array_equal(g.read(kf.BUF_FLOATDEPTH), raw)
    assert tr == tc and it == ic
    for f in range(n):
        assert np.abs(poses[f] - pc[f]).max() <= 1e-4, f"frame {f}"
