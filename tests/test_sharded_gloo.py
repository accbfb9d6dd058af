"""Host-side logic of the z-slab sharded mode (slambench_b200/sharded.py) on the CPU: world_size 2,
`gloo` backend, a numpy stand-in for the per-GPU context.  Checks the partitioning, the band
all-gather, the stream-ordered barrier calls, the all-reduced ICP iteration loop (with the product's
real host pose algebra from libkfb200.so) and the slab gather."""
from __future__ import annotations

import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from slambench_b200 import kfusion as kf
from slambench_b200 import sharded

W, H, N = 64, 48, 16


def test_partitions():
    assert sharded.slab_bounds(1024, 8) == [(i * 128, (i + 1) * 128) for i in range(8)]
    s = sharded.slab_bounds(10, 4)
    assert s == [(0, 3), (3, 6), (6, 8), (8, 10)] and s[-1][1] == 10
    assert sharded.row_bands(480, 8)[3] == (180, 240)
    with pytest.raises(ValueError):
        sharded.row_bands(480, 7)
    with pytest.raises(ValueError):
        sharded.slab_bounds(3, 4)
    # load-aware boundaries: contiguous cover, aligned, heavier regions get thinner slabs
    w = np.concatenate([np.zeros(64), np.linspace(1, 50, 192)])
    for world in (2, 4, 8):
        sl = sharded.slab_bounds(256, world, w, align=8)
        assert sl[0][0] == 0 and sl[-1][1] == 256 and len(sl) == world
        assert all(a[1] == b[0] for a, b in zip(sl, sl[1:])) and all(z0 % 8 == 0 and z1 > z0 for z0, z1 in sl)
        assert sl[0][1] - sl[0][0] > sl[-1][1] - sl[-1][0]
    assert sharded.slab_bounds(64, 4, np.ones(64), align=8)[0][0] == 0
    # worlds that do not divide the volume (256^3 on 3 ranks used to start slabs at z = 86, 171): every
    # partition the sharded host can choose starts its slabs on brick layers, which kfb_create also insists on
    from slambench_b200 import synth
    Kc = np.array(synth.K_DEFAULT, np.float32)
    pose = kf.identity_pose(np.array(synth.INIT_POS_FACTOR, np.float32) * np.float32(4.8))
    for n_z, world in ((256, 3), (256, 5), (256, 7), (1024, 6), (100, 3)):
        parts = dict(sharded.candidate_slabs(n_z, world, 4.8, pose, Kc, (640, 480), far=4.0))
        parts["default"] = sharded.slab_bounds(n_z, world, align=8)
        for name, sl in parts.items():
            assert len(sl) == world and sl[0][0] == 0 and sl[-1][1] == n_z, (name, sl)
            assert all(a[1] == b[0] for a, b in zip(sl, sl[1:])), (name, sl)
            assert all(z0 % 8 == 0 and z1 > z0 for z0, z1 in sl), (name, sl)
    K = np.array(synth.K_DEFAULT, np.float32)
    fw = sharded.frustum_slice_weights(64, 4.8, kf.identity_pose([2.4, 2.4, 1.2]), K, (640, 480), far=3.2)
    assert fw[:16].sum() == 0 and fw[20:56].min() > 0 and np.all(np.diff(fw[20:56]) >= 0)    # behind the camera: nothing; then growing


def synthetic_track_data(level):
    """Deterministic per-pixel (error, J) for a level: the 'image' every rank reduces a band of."""
    w, h = W >> level, H >> level
    rng = np.random.default_rng(100 + level)
    J = rng.normal(size=(h, w, 6)).astype(np.float32)
    e = (rng.normal(size=(h, w)) * 1e-3).astype(np.float32)
    return e, J


def reduce_rows(level, r0, r1):
    e, J = synthetic_track_data(level)
    e, J = e[r0:r1].reshape(-1).astype(np.float64), J[r0:r1].reshape(-1, 6).astype(np.float64)
    out = np.zeros(32)
    out[0] = (e * e).sum()
    out[1:7] = J.T @ e
    out[7:28] = (J.T @ J)[np.triu_indices(6)]
    out[28] = e.size
    return out.astype(np.float32)


class FakeLocal:
    """numpy/torch-CPU stand-in with the subset of `Kfusion` that ShardedKfusion drives."""

    def __init__(self, device, slab, flags):
        self.slab, self.rows = slab, None
        self.lib = kf.load_library()
        self.t = {kf.BUF_VERTEX: torch.zeros(H, W, 3), kf.BUF_NORMAL: torch.zeros(H, W, 3), kf.BUF_REDUCTION_DEV: torch.zeros(32),
                  kf.BUF_BRICKFLAGS: torch.zeros(2, 2, 2, dtype=torch.uint8)}
        assert flags & kf.FLAG_BRICKS_MERGED
        self.pose = np.eye(4, dtype=np.float32)
        self.host = {}
        self.calls = []

    def slab_ipc_handle(self):
        return bytes([self.slab[0] % 251]) * 64

    def slab_import(self, rank, world, handles, z_begin):
        assert len(handles) == world and all(len(h) == 64 for h in handles)
        assert handles[rank] == self.slab_ipc_handle() and z_begin[rank] == self.slab[0]
        self.rank = rank

    def set_pixel_rows(self, r0, r1):
        self.rows = (r0, r1)

    def tensor(self, which):
        return self.t[which]

    def preprocessing(self, d, size=None):
        return True

    def pyramidKernels(self, k):
        self.calls.append("pyramid")

    def getPose(self):
        return self.pose.copy()

    def setPose(self, p):
        self.pose = np.array(p, np.float32)

    def read(self, which):
        return np.eye(4, dtype=np.float32)

    def write(self, which, data):
        self.host[which] = np.array(data)

    def cameraMatrix(self, k):
        return np.eye(4, dtype=np.float32)

    def inverse(self, m):
        return np.linalg.inv(m).astype(np.float32)

    def matmul(self, a, b):
        return (a @ b).astype(np.float32)

    def trackReduceKernel(self, level, pose, view):
        r0, r1 = self.rows[0] >> level, self.rows[1] >> level
        self.t[kf.BUF_REDUCTION_DEV].copy_(torch.from_numpy(reduce_rows(level, r0, r1)))

    def updatePoseKernel(self, pose, red, thr):
        p = np.ascontiguousarray(pose, np.float32).reshape(16).copy()
        conv = C.c_int(0)
        self.lib.kfb_k_update_pose(p.ctypes.data_as(C.c_void_p), np.ascontiguousarray(red, np.float32).ctypes.data_as(C.c_void_p), C.c_float(thr), C.byref(conv))
        return p.reshape(4, 4), bool(conv.value)

    def checkPoseKernel(self, pose, old, red, size):
        p = np.ascontiguousarray(pose, np.float32).reshape(16).copy()
        ok = C.c_int(0)
        self.lib.kfb_k_check_pose(p.ctypes.data_as(C.c_void_p), np.ascontiguousarray(old, np.float32).ctypes.data_as(C.c_void_p),
                                  np.ascontiguousarray(red, np.float32).ctypes.data_as(C.c_void_p), C.c_uint32(size[0]), C.c_uint32(size[1]), C.c_float(0.15), C.byref(ok))
        return p.reshape(4, 4), bool(ok.value)

    def integration(self, k, rate, mu, frame):
        self.calls.append("integrate")
        self.t[kf.BUF_BRICKFLAGS].view(-1)[self.rank] = 1      # each rank flags what its own slices touch
        return True

    def raycasting(self, k, mu, frame):
        r0, r1 = self.rows
        rows = torch.arange(r0, r1, dtype=torch.float32)[:, None, None]
        self.t[kf.BUF_VERTEX][r0:r1] = rows + 1000 * (self.rank + 1)
        self.t[kf.BUF_NORMAL][r0:r1] = -rows - 1000 * (self.rank + 1)
        return False

    def read_volume(self):
        return np.full((self.slab[1] - self.slab[0], 2, 2, 2), self.rank, np.int16)

    def close(self):
        pass


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        holder = {}

        def factory(**kw):
            holder["l"] = FakeLocal(**kw)
            return holder["l"]

        s = sharded.ShardedKfusion((W, H), N, 4.8, np.zeros(3, np.float32), (3, 2, 2), rank=rank, world=world, icp_mode="allreduce",
                                   dist=dist, local_factory=factory)
        loc = holder["l"]
        assert loc.slab == sharded.slab_bounds(N, world)[rank] and loc.rows == sharded.row_bands(H, world)[rank]
        # raycast bands are all-gathered: every rank ends with the full maps
        s.raycasting(None, 0.1, frame=5)
        want = np.zeros((H, W, 3), np.float32)
        for r, (r0, r1) in enumerate(sharded.row_bands(H, world)):
            want[r0:r1] = np.arange(r0, r1, dtype=np.float32)[:, None, None] + 1000 * (r + 1)
        assert np.array_equal(loc.t[kf.BUF_VERTEX].numpy(), want) and np.array_equal(loc.t[kf.BUF_NORMAL].numpy(), -want)
        # frames <= 2: no raycast yet, no collective (cpp/kernels.cpp:977)
        loc.t[kf.BUF_VERTEX].zero_()
        s.raycasting(None, 0.1, frame=2)
        assert float(loc.t[kf.BUF_VERTEX][sharded.row_bands(H, world)[1 - rank][0]].abs().sum()) == 0
        # all-reduced ICP loop == the same loop on the whole image
        tracked = s.tracking(np.zeros(4, np.float32), 1e-5, 1, frame=7)
        pose = s.getPose()
        ref_pose = np.eye(4, dtype=np.float32)
        last = None
        for level in (2, 1, 0):
            for _ in range((3, 2, 2)[level]):
                parts = [reduce_rows(level, b0 >> level, b1 >> level) for b0, b1 in sharded.row_bands(H, world)]
                last = parts[0] + parts[1]
                ref_pose, conv = loc.updatePoseKernel(ref_pose, last, 1e-5)
                if conv:
                    break
        assert np.array_equal(pose, ref_pose), "pose after the all-reduced iterations"
        assert np.array_equal(loc.host[kf.BUF_REDUCTION], last)
        assert s.tracking(np.zeros(4, np.float32), 1e-5, 2, frame=7) is False       # frame % tracking_rate gate
        assert s.integration(None, 1, 0.1, 7) is True
        assert loc.t[kf.BUF_BRICKFLAGS].view(-1)[:3].tolist() == [1, 1, 0], "brick flags must be the union over ranks"
        q.put((rank, pose.copy(), bool(tracked)))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_orchestration_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda t: t[0])
    assert np.array_equal(res[0][1], res[1][1]), "every rank must hold the same pose"
    assert res[0][2] == res[1][2]


def _render_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench

        depth, gt = bench.render_sequence_distributed(7, rank, world, torch, dist, "cpu")
        q.put((rank, depth, gt))
    finally:
        dist.destroy_process_group()


def test_distributed_frame_rendering_world2_gloo():
    """Long sharded runs render frame f on rank f % world and all-gather the frames (bench.py): every rank must end
    up with exactly the sequence synth.make_sequence() produces."""
    from slambench_b200 import synth

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_render_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref, gt = synth.make_sequence(7, long_run=False)
    for _, depth, g in res:
        assert depth.dtype == np.uint16 and np.array_equal(depth, ref) and np.allclose(g, gt)
