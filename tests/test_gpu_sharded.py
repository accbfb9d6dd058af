"""z-slab sharded volume on 2 GPUs (NCCL, CUDA-IPC peer slabs) against the same run on one GPU:
bit-identical volume (integrate is exact whatever the slab), identical raycast maps, identical
poses and flags — in both ICP modes.  Needs >= 2 GPUs (gpurun --gpus 2)."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N_FRAMES, VRES = 9, 128


def _worker(rank, world, port, mode, q):
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
                                    balance_k=K if mode == "replicated" else None) as s:   # load-aware and even slabs
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


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["replicated", "allreduce"])
def test_two_gpu_slabs_match_one_gpu(mode):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
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
