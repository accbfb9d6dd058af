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
              every rank holds the full ICP reference.  The brick flags that let the raycaster step over
              free space are merged over the ranks (all-reduce MAX, 2 MB at 1024^3) right after integrate,
              so most samples never touch a peer's memory.
  track       "replicated" (default): every rank runs the whole persistent ICP kernel on all pixels —
              bitwise identical poses on every rank, zero collectives (ICP is ~0.15 ms; a per-iteration
              all-reduce costs more than it saves).
              "allreduce": rank r tracks its row band; the 32 partial sums are combined with an NCCL
              all-reduce per ICP iteration and every rank runs the same host solve (the reference's
              control flow, cpp/kernels.cpp:950-969).
  preprocess  replicated (614 KB of input per rank; cheaper than exchanging).

Transport.  "peer" (default): after one exchange of CUDA-IPC handles the library moves every per-frame
byte itself over NVLink peer memory — k_integrate_run2 stores the brick flags of its slab into every peer's
map, k_raycast stores its band of the vertex / normal maps into every peer's maps (the all-gather, fused into
the kernel), and kfb_integrate / kfb_raycast end with a stream-ordered barrier over peer memory
(k_peer_barrier): no NCCL collective per frame in the replicated ICP mode.  "nccl": round 1's data path
(flags all-reduced, bands all-gathered), kept for A/B.  Ordering: nobody raycasts a peer's slab before that
peer has integrated the frame; nobody integrates the next frame while a peer still reads its slab.
"""
from __future__ import annotations

import numpy as np

from . import kfusion as kf


def slab_bounds(n_z: int, world: int, weights=None, align: int = 1) -> list[tuple[int, int]]:
    """Contiguous z-ranges in rank order.  Without `weights`: sizes differing by at most one `align`-slice layer.  With
    per-slice `weights` (expected integrate work, see frustum_slice_weights): boundaries (multiples of `align`)
    that equalise the summed weight, every slab keeping at least `align` slices."""
    if world < 1 or n_z < world * align:
        raise ValueError(f"cannot cut {n_z} slices into {world} slabs")
    if weights is None:
        units = (n_z + align - 1) // align              # whole `align`-slice layers; the last one may be partial
        base, extra = divmod(units, world)
        out, u = [], 0
        for r in range(world):
            n = base + (1 if r < extra else 0)
            out.append((u * align, min((u + n) * align, n_z)))
            u += n
        return out
    w = np.asarray(weights, np.float64)
    if w.shape != (n_z,) or not np.all(w >= 0):
        raise ValueError("weights must be one non-negative number per slice")
    w = w + w.sum() * 0.02 / n_z + 1e-12          # a floor: empty regions still cost the raycaster's peer reads

    # cost of a slab [a, b): the work of its slices (frustum_slice_weights already weighs per-voxel bricks against
    # free-space ones) plus, for slabs that hold per-voxel work, the replay of the reference's additions from z = 0 up to
    # the slab (once per column half: a small term since round 2)
    def cost(a, b):
        return w[a:b].sum() + 0.002 * a * w[a:b].max()

    # minimise the largest slab cost: bisection on the bound, greedy cuts from the far end
    cand = list(range(0, n_z + 1, align))
    if cand[-1] != n_z:
        cand.append(n_z)

    def cuts_for(bound):
        cuts, hi = [n_z], n_z
        for _ in range(world - 1):
            lo_ok = None
            for a in reversed([c for c in cand if c < hi]):
                if cost(a, hi) <= bound:
                    lo_ok = a
                else:
                    break
            if lo_ok is None or lo_ok == 0:
                break
            cuts.append(lo_ok)
            hi = lo_ok
        return cuts, cost(0, hi) <= bound

    lo, hi_b = 0.0, cost(0, n_z)
    for _ in range(40):
        mid = 0.5 * (lo + hi_b)
        if cuts_for(mid)[1]:
            hi_b = mid
        else:
            lo = mid
    cuts = sorted(set(cuts_for(hi_b)[0]) | {0})
    # exactly `world` slabs: split the widest ones if the greedy pass needed fewer
    while len(cuts) - 1 < world:
        widths = [(cuts[i + 1] - cuts[i], i) for i in range(len(cuts) - 1)]
        wd, i = max(widths)
        mid = (cuts[i] + cuts[i + 1]) // 2 // align * align
        if mid <= cuts[i] or mid >= cuts[i + 1]:
            raise ValueError("cannot place the slab boundaries")
        cuts.insert(i + 1, mid)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def equalise_slabs(density, world: int, align: int = 8) -> list[tuple[int, int]]:
    """Boundaries (multiples of `align`, at least `align` slices per slab) that give every slab the same share of `density`."""
    n_z = len(density)
    cum = np.concatenate([[0.0], np.cumsum(np.asarray(density, np.float64))])
    cuts = [0]
    for r in range(1, world):
        z = int(np.searchsorted(cum, cum[-1] * r / world))
        z = int(round(z / align)) * align
        z = min(max(z, cuts[-1] + align), n_z - (world - r) * align)
        cuts.append(z)
    cuts.append(n_z)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def refit_density(density, slabs, times):
    """One step of iterative proportional fitting: inside every slab the per-slice cost density keeps its shape and is
    scaled so that it sums to the slab's MEASURED integrate time.  Starting from the a-priori model
    (frustum_slice_weights) a few rounds of measure -> refit -> equalise_slabs balance the slabs on what the camera really
    sees; a volume does not depend on how it is cut."""
    d = np.array(density, np.float64) + 1e-12
    for (a, b), t in zip(slabs, times):
        d[a:b] *= max(float(t), 1e-9) / d[a:b].sum()
    return d


def rebalance_slabs(slabs, times, align: int = 8) -> list[tuple[int, int]]:
    """New boundaries from ONE measurement with a flat density inside every slab (see refit_density for the iterated form)."""
    return equalise_slabs(refit_density(np.ones(slabs[-1][1]), slabs, times), len(slabs), align)


def frustum_slice_weights(n_z: int, volume_dim: float, pose, k, image_wh, far: float = 4.0, mu: float = 0.1, band_mu: float = 2.5,
                          band_weight: float = 3.0):
    """Expected integrate work per z-slice for a camera at `pose` that sees surfaces around depth `far`: the area of the
    slice that projects into the image, weighted by what integrate does there (cpp/kernels.cpp:647-661 as k_integrate_run2
    executes it): free space in front of the surface is a streaming update (weight 1), the band around the surface is
    decided voxel by voxel (`band_weight`, over a band of `band_mu` * mu either side of `far`), behind it nothing happens (0).  Used once, at
    set-up, to place the slab boundaries; it is only a load-balance heuristic: any partition gives the same voxels."""
    pose = np.asarray(pose, np.float64).reshape(4, 4)
    fx, fy, cx, cy = [float(v) for v in k]
    w, h = image_wh
    n = 48                                                     # coarse grid over the slice
    g = (np.arange(n) + 0.5) / n * volume_dim
    X, Y = np.meshgrid(g, g)
    Rinv, t = pose[:3, :3].T, pose[:3, 3]
    band = band_mu * mu                                        # mu plus the slack of brick-granular decisions
    out = np.zeros(n_z)
    for z in range(n_z):
        P = np.stack([X - t[0], Y - t[1], np.full_like(X, (z + 0.5) / n_z * volume_dim - t[2])], -1) @ Rinv.T
        d = P[..., 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            u, v = fx * P[..., 0] / d + cx, fy * P[..., 1] / d + cy
        inside = (d > 1e-4) & (u >= 0) & (u <= w - 1) & (v >= 0) & (v <= h - 1)
        out[z] = np.count_nonzero(inside & (d < far - band)) + band_weight * np.count_nonzero(inside & (np.abs(d - far) <= band))
    return out


def candidate_slabs(n_z: int, world: int, volume_dim: float, pose, k, image_wh, far: float) -> dict:
    """A handful of a-priori partitions for `bench.py`'s set-up to try.  What a slab costs is plan + replay (per slab that
    holds part of the surface band, hardly divisible) + streaming of its free space + per-voxel work of its part of the
    band; no closed form ranks the partitions reliably on 2, 4 and 8 GPUs alike (measured, profiles/r2_summary.md), so the
    harness times a few frames on each candidate and keeps the fastest."""
    out = {"even": slab_bounds(n_z, world, align=8)}
    for name, (bm, bw) in {"band2.5x3": (2.5, 3.0), "band1.7x3.5": (1.7, 3.5), "band4x1.5": (4.0, 1.5), "band2.5x1": (2.5, 1.0), "band2.5x0.5": (2.5, 0.5)}.items():
        w = frustum_slice_weights(n_z, volume_dim, pose, k, image_wh, far=far, band_mu=bm, band_weight=bw)
        out[name] = slab_bounds(n_z, world, w, align=8)
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
                 device: int = 0, icp_mode: str = "replicated", dist=None, local_factory=None, flags: int = 0,
                 balance_k=None, balance_far: float = 4.0, transport: str = "peer", slabs=None):
        if dist is None:
            import torch.distributed as dist  # noqa: PLC0415
        import torch  # noqa: PLC0415

        self.torch, self.dist = torch, dist
        self.rank, self.world, self.device = rank, world, device
        if icp_mode not in ("replicated", "allreduce"):
            raise ValueError(icp_mode)
        if transport not in ("peer", "nccl"):
            raise ValueError(transport)
        self.icp_mode = icp_mode
        vr = [int(volumeResolution)] * 3 if np.isscalar(volumeResolution) else [int(v) for v in volumeResolution]
        weights = None
        if balance_k is not None:
            # load-aware slab boundaries from the initial pose (identical on every rank: pure function of the arguments)
            ip = np.asarray(initPose, np.float32)
            pose0 = kf.identity_pose(ip) if ip.size == 3 else ip.reshape(4, 4)
            vd = float(volumeDimensions) if np.isscalar(volumeDimensions) else float(volumeDimensions[2])
            weights = frustum_slice_weights(vr[2], vd, pose0, balance_k, (int(inputSize[0]), int(inputSize[1])), far=balance_far)
        # every slab starts on a brick layer (8 slices): integrate classifies and the raycaster skips per 8^3 brick
        self.slabs = [tuple(int(v) for v in z) for z in slabs] if slabs is not None else slab_bounds(vr[2], world, weights, align=8)
        self.bands = row_bands(int(inputSize[1]), world)
        self.pyramid = tuple(int(i) for i in pyramid)
        make = local_factory or (lambda **kw: kf.Kfusion(inputSize, vr, volumeDimensions, initPose, self.pyramid, **kw))
        self.local = make(device=device, slab=self.slabs[rank], flags=flags | kf.FLAG_BRICKS_MERGED)
        self.computationSize = (int(inputSize[0]), int(inputSize[1]))
        # CUDA IPC handles travel through the (CPU) object collective, once.  "peer" (default): slabs, raycast maps, brick
        # flags and barrier slots are all exchanged — the library then moves every per-frame byte itself over NVLink peer
        # memory (flags and map bands stored straight into the peers, barriers in kfb_integrate / kfb_raycast) and this
        # class issues NO collective per frame in the replicated ICP mode.  "nccl": round 1's data path (peer loads of
        # the slabs only; flags all-reduced and bands all-gathered by NCCL), kept for A/B runs.
        self.transport = transport if hasattr(self.local, "ipc_export") else "nccl"
        handles = [None] * world
        if self.transport == "peer":
            dist.all_gather_object(handles, self.local.ipc_export())
            self.local.ipc_import(rank, world, handles)
        else:
            dist.all_gather_object(handles, self.local.slab_ipc_handle())
            self.local.slab_import(rank, world, handles, [z[0] for z in self.slabs])
        self.local.set_pixel_rows(*self.bands[rank])
        self._stream = self.local.torch_stream(torch) if hasattr(self.local, "torch_stream") else None
        w, h = self.computationSize
        self._vertex = self._view(kf.BUF_VERTEX, (h, w, 3))
        self._normal = self._view(kf.BUF_NORMAL, (h, w, 3))
        self._red = self._view(kf.BUF_REDUCTION_DEV, (32,))
        self._token = torch.zeros(1, device=self._vertex.device)
        nb = [(v + 7) // 8 for v in vr]
        self._bricks = None if (flags & kf.FLAG_RAYCAST_NO_SKIP) else self._view(kf.BUF_BRICKFLAGS, (nb[2], nb[1], nb[0]), "|u1")
        dist.barrier()

    # ------------------------------------------------------------------ plumbing
    def _view(self, which, shape, typestr="<f4"):
        if hasattr(self.local, "tensor"):           # CPU stand-in used by the gloo tests
            return self.local.tensor(which)
        return self.torch.as_tensor(_DevArray(self.local.device_ptr(which), shape, typestr), device=f"cuda:{self.device}")

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
        if self.transport == "peer":
            return done                                  # flags already stored into the peers; barrier inside kfb_integrate
        with self._on_stream():
            # stream-ordered barrier (every slab holds this frame before any peer reads it) that also merges the
            # brick flags: a rank flags only the bricks ITS slices touch; the raycaster needs the union
            if self._bricks is not None:
                self.dist.all_reduce(self._bricks, op=self.dist.ReduceOp.MAX)
            else:
                self.dist.all_reduce(self._token)
        return done

    def raycasting(self, k, mu: float, frame: int) -> bool:
        self.local.raycasting(k, mu, frame)          # this rank's row band, through all slabs (peer loads)
        if frame > 2 and self.transport != "peer":       # "peer": bands already stored into the peers; barrier inside kfb_raycast
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
