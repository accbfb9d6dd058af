"""z-slab sharded KinectFusion over several GPUs (BASELINE configs[3], [4]).

One process per GPU (`torch.distributed`, NCCL over NVLink/NVSwitch).  Only the TSDF volume is
partitioned — the reference has no multi-device mode, so this is the host-side orchestration of
the single-GPU C ABI (include/kfb200.h):

  integrate   fully local: rank r owns the slices z in [z_r, z_{r+1}) of the reference layout
              (`kfb_config.slab_z0/z1`); the kernel replays the reference's additions from z = 0, so
              every voxel is bit-identical to the unsharded volume.  No exchange.
  raycast     pixels are partitioned (row bands): rank r marches ITS rays through the WHOLE volume,
              reading the other ranks' slabs through CUDA-IPC peer pointers (NVLink P2P loads inside
              k_raycast); the bands of the vertex / normal maps are then all-gathered (NCCL) so that
              every rank holds the full ICP reference.
  track       "replicated" (default): every rank runs the whole persistent ICP kernel on all pixels —
              bitwise identical poses on every rank, zero collectives (ICP is ~0.15 ms; a per-iteration
              all-reduce costs more than it saves).
              "allreduce": rank r tracks its row band; the 32 partial sums are combined with an NCCL
              all-reduce per ICP iteration and every rank runs the same host solve (the reference's
              control flow, cpp/kernels.cpp:950-969).
  preprocess  replicated (614 KB of input per rank; cheaper than exchanging).

Ordering between ranks is stream-ordered: a 1-element all-reduce after integrate (nobody raycasts a
peer's slab before that peer has integrated the frame) and the all-gather after raycast (nobody
integrates the next frame while a peer still reads its slab).  All collectives are enqueued on the
context's own CUDA stream.
"""
from __future__ import annotations

import numpy as np

from . import kfusion as kf


def slab_bounds(n_z: int, world: int) -> list[tuple[int, int]]:
    """Contiguous z-ranges, sizes differing by at most one slice, in rank order."""
    if world < 1 or n_z < world:
        raise ValueError(f"cannot cut {n_z} slices into {world} slabs")
    base, extra = divmod(n_z, world)
    out, z = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((z, z + n))
        z += n
    return out


def row_bands(h: int, world: int) -> list[tuple[int, int]]:
    """Equal row bands (the all-gather needs equal counts)."""
    if h % world != 0:
        raise ValueError(f"{h} image rows are not divisible by {world} ranks")
    n = h // world
    return [(r * n, (r + 1) * n) for r in range(world)]


class _DevArray:
    """Minimal __cuda_array_interface__ carrier: a torch view of a raw device pointer."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}


class ShardedKfusion:
    """`Kfusion` (kernels.h:83-195) over `world` GPUs; same method names and return values."""

    def __init__(self, inputSize, volumeResolution, volumeDimensions, initPose, pyramid=(10, 5, 4), *, rank: int, world: int,
                 device: int = 0, icp_mode: str = "replicated", dist=None, local_factory=None, flags: int = 0):
        if dist is None:
            import torch.distributed as dist  # noqa: PLC0415
        import torch  # noqa: PLC0415

        self.torch, self.dist = torch, dist
        self.rank, self.world, self.device = rank, world, device
        if icp_mode not in ("replicated", "allreduce"):
            raise ValueError(icp_mode)
        self.icp_mode = icp_mode
        vr = [int(volumeResolution)] * 3 if np.isscalar(volumeResolution) else [int(v) for v in volumeResolution]
        self.slabs = slab_bounds(vr[2], world)
        self.bands = row_bands(int(inputSize[1]), world)
        self.pyramid = tuple(int(i) for i in pyramid)
        make = local_factory or (lambda **kw: kf.Kfusion(inputSize, vr, volumeDimensions, initPose, self.pyramid, **kw))
        self.local = make(device=device, slab=self.slabs[rank], flags=flags)
        self.computationSize = (int(inputSize[0]), int(inputSize[1]))
        # peer slabs: CUDA IPC handles travel through the (CPU) object collective
        handles = [None] * world
        dist.all_gather_object(handles, self.local.slab_ipc_handle())
        self.local.slab_import(rank, world, handles, [z[0] for z in self.slabs])
        self.local.set_pixel_rows(*self.bands[rank])
        self._stream = self.local.torch_stream(torch) if hasattr(self.local, "torch_stream") else None
        w, h = self.computationSize
        self._vertex = self._view(kf.BUF_VERTEX, (h, w, 3))
        self._normal = self._view(kf.BUF_NORMAL, (h, w, 3))
        self._red = self._view(kf.BUF_REDUCTION_DEV, (32,))
        self._token = torch.zeros(1, device=self._vertex.device)
        dist.barrier()

    # ------------------------------------------------------------------ plumbing
    def _view(self, which, shape):
        if hasattr(self.local, "tensor"):           # CPU stand-in used by the gloo tests
            return self.local.tensor(which)
        return self.torch.as_tensor(_DevArray(self.local.device_ptr(which), shape, "<f4"), device=f"cuda:{self.device}")

    def _on_stream(self):
        if self._stream is None:
            import contextlib  # noqa: PLC0415

            return contextlib.nullcontext()
        return self.torch.cuda.stream(self._stream)

    def close(self):
        self.dist.barrier()       # nobody unmaps a slab a peer may still read
        self.local.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------- the reference's API
    def preprocessing(self, inputDepth, inputSize=None) -> bool:
        return self.local.preprocessing(inputDepth, inputSize)

    def tracking(self, k, icp_threshold: float, tracking_rate: int, frame: int) -> bool:
        if self.icp_mode == "replicated":
            return self.local.tracking(k, icp_threshold, tracking_rate, frame)
        return self._tracking_allreduce(k, icp_threshold, tracking_rate, frame)

    def _tracking_allreduce(self, k, icp_threshold, tracking_rate, frame) -> bool:
        """Kfusion::tracking (cpp/kernels.cpp:924-971) with the reduction split over ranks."""
        g = self.local
        if frame % tracking_rate != 0:
            return False
        g.pyramidKernels(k)
        pose = g.getPose()
        old_pose = pose.copy()
        view = g.matmul(g.cameraMatrix(k), g.inverse(g.read(kf.BUF_RAYCASTPOSE)))      # projectReference (:948)
        red = np.zeros(32, np.float32)
        for level in range(len(self.pyramid) - 1, -1, -1):
            for _ in range(self.pyramid[level]):
                g.trackReduceKernel(level, pose, view)          # this rank's band -> KFB_BUF_REDUCTION_DEV
                with self._on_stream():
                    self.dist.all_reduce(self._red)             # sum over ranks: same 32 floats everywhere
                    red = self._red.cpu().numpy().copy()
                pose, converged = g.updatePoseKernel(pose, red, icp_threshold)
                if converged:
                    break
        pose, ok = g.checkPoseKernel(pose, old_pose, red, self.computationSize)
        g.setPose(pose)
        g.write(kf.BUF_OLDPOSE, old_pose)
        g.write(kf.BUF_REDUCTION, red)
        return ok

    def integration(self, k, integration_rate: int, mu: float, frame: int) -> bool:
        done = self.local.integration(k, integration_rate, mu, frame)
        with self._on_stream():
            self.dist.all_reduce(self._token)       # stream-ordered barrier: every slab holds this frame before any peer reads it
        return done

    def raycasting(self, k, mu: float, frame: int) -> bool:
        self.local.raycasting(k, mu, frame)          # this rank's row band, through all slabs (peer loads)
        if frame > 2:
            r0, r1 = self.bands[self.rank]
            with self._on_stream():
                self.dist.all_gather_into_tensor(self._vertex, self._vertex[r0:r1])
                self.dist.all_gather_into_tensor(self._normal, self._normal[r0:r1])
        return False

    def computeFrame(self, inputDepth, inputSize, k, integration_rate, tracking_rate, icp_threshold, mu, frame):
        self.preprocessing(inputDepth, inputSize)
        tr = self.tracking(k, icp_threshold, tracking_rate, frame)
        it = self.integration(k, integration_rate, mu, frame)
        self.raycasting(k, mu, frame)
        return tr, it

    def getPose(self):
        return self.local.getPose()

    def synchroniseDevices(self):
        self.local.synchroniseDevices()

    def stats(self):
        return self.local.stats()

    def gather_volume(self):
        """The whole volume on rank 0 (slabs concatenated in z order), None elsewhere."""
        mine = self.local.read(kf.BUF_VOLUME)
        parts = [None] * self.world if self.rank == 0 else None
        self.dist.gather_object(mine, parts, dst=0)
        return np.concatenate(parts, axis=0) if self.rank == 0 else None
