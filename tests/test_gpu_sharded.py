"""z-slab sharded volume on 2 GPUs (CUDA-IPC peer memory; NCCL only for the optional ICP all-reduce) against
  * the same run on one GPU: bit-identical volume (integrate is exact whatever the slab), identical raycast maps,
    identical poses and flags — both ICP modes, both transports;
  * the REFERENCE at 1024^3 (tests/golden/big_1024.npz, made by the unmodified reference): teacher-forced integrate +
    raycast bit-exact per slab, free-running pose / flags within the north_star gates.
Needs >= 2 GPUs (gpurun --gpus 2)."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N_FRAMES, VRES = 9, 128


def _worker(rank, world, port, mode, transport, q):
    import torch.distributed as dist

    from slambench_b200 import kfusion as kf
    from slambench_b200 import sharded, synth

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        K = np.array(synth.K_DEFAULT, np.float32)
        T0 = (np.array(synth.INIT_POS_FACTOR, np.float32) * np.float32(4.8)).astype(np.float32)
        depth, _ = synth.make_sequence(N_FRAMES)
        out = {"poses": [], "flags": []}
        with sharded.ShardedKfusion((640, 480), VRES, 4.8, T0, (10, 5, 4), rank=rank, world=world, device=rank, icp_mode=mode,
                                    transport=transport, balance_k=K if mode == "replicated" else None) as s:   # load-aware and even slabs
            assert s.transport == transport
            for f in range(N_FRAMES):
                s.preprocessing(depth[f])
                tr = s.tracking(K, 1e-5, 1, f)
                it = s.integration(K, 1, 0.1, f)
                s.raycasting(K, 0.1, f)
                out["poses"].append(s.getPose().copy())
                out["flags"].append((tr, it))
            s.synchroniseDevices()
            out["vertex"] = s.local.read(kf.BUF_VERTEX)
            out["normal"] = s.local.read(kf.BUF_NORMAL)
            vol = s.gather_volume()
        if rank == 0:
            # the same sequence on ONE GPU, unsharded
            ref = {"poses": [], "flags": []}
            flags = kf.FLAG_ICP_HOST_SOLVE if mode == "allreduce" else 0
            with kf.Kfusion((640, 480), VRES, 4.8, T0, (10, 5, 4), device=0, flags=flags) as g:
                for f in range(N_FRAMES):
                    g.preprocessing(depth[f])
                    tr = g.tracking(K, 1e-5, 1, f)
                    it = g.integration(K, 1, 0.1, f)
                    g.raycasting(K, 0.1, f)
                    ref["poses"].append(g.getPose().copy())
                    ref["flags"].append((tr, it))
                ref["vertex"], ref["normal"], ref["vol"] = g.read(kf.BUF_VERTEX), g.read(kf.BUF_NORMAL), g.read(kf.BUF_VOLUME)
            q.put((rank, out, vol, ref))
        else:
            q.put((rank, out, None, None))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(target, args):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, 2, port, *args, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=900) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return res


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode,transport", [("replicated", "peer"), ("replicated", "nccl"), ("allreduce", "peer")])
def test_two_gpu_slabs_match_one_gpu(mode, transport):
    res = _run(_worker, (mode, transport))
    (_, out0, vol, ref), (_, out1, _, _) = res
    assert out0["flags"] == out1["flags"] == ref["flags"]
    assert [f[0] for f in ref["flags"]] == [False] * 4 + [True] * (N_FRAMES - 4)
    p0, p1, pr = np.stack(out0["poses"]), np.stack(out1["poses"]), np.stack(ref["poses"])
    assert np.array_equal(p0, p1), "ranks disagree on the pose"
    if mode == "replicated":
        assert np.array_equal(p0, pr), "sharded pose differs from the single-GPU pose"
        assert np.array_equal(vol, ref["vol"]), "sharded volume differs from the single-GPU volume"
        for key in ("vertex", "normal"):
            assert np.array_equal(out0[key].view(np.uint32), ref[key].view(np.uint32)) and np.array_equal(out1[key].view(np.uint32), ref[key].view(np.uint32))
    else:
        # band-split sums are added in a different order: poses agree to fp32 rounding
        assert np.abs(p0 - pr).max() <= 2e-6
        d = np.abs(vol.astype(np.int32) - ref["vol"].astype(np.int32)).max(-1)
        assert (d <= 1).mean() > 0.9999


def _worker_golden(rank, world, port, q):
    """1024^3 over two slabs against the reference's fixture: teacher-forced kernels per slab, then the free-running pipeline."""
    import hashlib
    import sys

    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import volsum
    from slambench_b200 import kfusion as kf
    from slambench_b200 import sharded, synth
    from slambench_b200.sharded import _DevArray

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        gold = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "big_1024.npz")))
        n = int(gold["n"][0])
        K = np.array(synth.K_DEFAULT, np.float32)
        T0 = (np.array(synth.INIT_POS_FACTOR, np.float32) * np.float32(4.8)).astype(np.float32)
        depth, _ = synth.make_sequence(len(gold["depth_sha256"]))
        for f in range(len(depth)):
            assert hashlib.sha256(depth[f].tobytes()).hexdigest() == str(gold["depth_sha256"][f])
        mu, gt = float(gold["mu"][0]), gold["gt_poses"]
        out = {}

        def slab_sums(g):
            nz = g.slab[1] - g.slab[0]
            g.synchroniseDevices()
            t = torch.as_tensor(_DevArray(g.device_ptr(kf.BUF_VOLUME), (nz, n * n), "<i4"), device=f"cuda:{rank}")
            return volsum.slice_checksums_torch(t)

        with sharded.ShardedKfusion((640, 480), n, 4.8, T0, (10, 5, 4), rank=rank, world=world, device=rank, balance_k=K) as s:
            g = s.local
            z0, z1 = g.slab
            out["slab"] = (z0, z1)
            # teacher-forced: integrateKernel on each slab, raycastKernel over both slabs (bands stored into the peer)
            bad = []
            for i, f in enumerate(gold["tf_frames"]):
                pose = gt[f].astype(np.float32)
                g.preprocessing(depth[f])
                g.integrateKernel(g.inverse(pose), g.cameraMatrix(K), mu)
                bad.append(int((slab_sums(g) != gold["tf_slice_sums"][i][z0:z1]).any(axis=1).sum()))
            out["tf_bad_slices"] = bad
            g.peer_barrier()
            g.raycastKernel(gold["tf_view"], largestep=0.75 * mu)
            g.peer_barrier()
            g.synchroniseDevices()
            v, nm = g.read(kf.BUF_VERTEX), g.read(kf.BUF_NORMAL)
            out["tf_maps_equal"] = bool(np.array_equal(volsum.array_checksum(v), gold["tf_vertex_sum"])
                                        and np.array_equal(volsum.array_checksum(nm), gold["tf_normal_sum"])
                                        and np.array_equal(v[::16], gold["tf_vertex_rows"]))
            dist.barrier()
            # free-running, from a fresh volume
            g.reset()
            g.setPose(kf.identity_pose(T0))
            g.write(kf.BUF_RAYCASTPOSE, np.zeros((4, 4), np.float32))
            g.write(kf.BUF_OLDPOSE, np.zeros((4, 4), np.float32))
            g.write(kf.BUF_REDUCTION, np.zeros(32, np.float32))
            g.synchroniseDevices()
            dist.barrier()
            poses, flags = [], []
            for f in range(len(gold["fr_poses"])):
                s.preprocessing(depth[f])
                tr = s.tracking(K, 1e-5, 1, f)
                it = s.integration(K, 1, mu, f)
                s.raycasting(K, mu, f)
                poses.append(s.getPose().copy())
                flags.append((tr, it))
            s.synchroniseDevices()
            out["fr_poses"], out["fr_flags"] = np.stack(poses), flags
            out["fr_same_slices"] = float((slab_sums(g) == gold["fr_slice_sums"][z0:z1]).all(axis=1).mean())
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_slabs_match_the_reference_at_1024():
    from test_gpu_pipeline import rot_angle

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "big_1024.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/big_1024.npz was not generated")
    gold = dict(np.load(path))
    (_, a), (_, b) = _run(_worker_golden, ())
    assert a["slab"][0] == 0 and a["slab"][1] == b["slab"][0] and b["slab"][1] == 1024
    assert a["tf_bad_slices"] == [0, 0, 0] and b["tf_bad_slices"] == [0, 0, 0], "teacher-forced slabs differ from the reference"
    assert a["tf_maps_equal"] and b["tf_maps_equal"], "raycast over peer slabs differs from the reference"
    assert np.array_equal(a["fr_poses"], b["fr_poses"]) and a["fr_flags"] == b["fr_flags"]
    assert a["fr_flags"] == [(bool(t), bool(i)) for t, i in gold["fr_flags"]]
    for f in range(len(gold["fr_poses"])):
        assert np.abs(a["fr_poses"][f][:3, 3] - gold["fr_poses"][f][:3, 3]).max() <= 1e-4, f"frame {f}: position"
        assert rot_angle(a["fr_poses"][f][:3, :3], gold["fr_poses"][f][:3, :3]) <= 1e-4, f"frame {f}: rotation"
    print(f"1024^3 on 2 GPUs, free-running: slices bit-identical to the reference: {a['fr_same_slices']:.4f} / {b['fr_same_slices']:.4f}")
    # free-running: the ICP sums are added in a different order than the reference's, the pose differs by ~1e-7 and the
    # truncated pixel of a voxel on a pixel edge flips: the fraction is reported, only a collapse is an error
    assert min(a["fr_same_slices"], b["fr_same_slices"]) > 0.2
