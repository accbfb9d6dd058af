"""Position-sensitive checksums of a TSDF volume, one pair per z-slice — the size-independent way the big-volume
parity tests (1024^3, 2048^3) compare a volume with the reference's: a 2048^3 volume is 34 GB, its 2048 x 2 sums are 32 KB.

For slice z with voxels v_i (the short2 {tsdf, weight} read as one little-endian uint32, i = x + y * N):
    s1 = sum_i v_i                        mod 2^64
    s2 = sum_i v_i * ((i mod 65521) + 1)  mod 2^64
Two volumes with equal (s1, s2) in every slice are bit-identical with overwhelming probability; a single differing voxel,
two swapped voxels, or a shifted row all change s2.  numpy (CPU reference volumes) and torch (device volumes) versions
compute the same numbers."""
from __future__ import annotations

import numpy as np

MOD = 65521


def slice_checksums_np(vol: np.ndarray, chunk: int = 8) -> np.ndarray:
    """vol: int16[z, y, x, 2] -> uint64[z, 2]."""
    nz = vol.shape[0]
    n = vol.shape[1] * vol.shape[2]
    v32 = vol.reshape(nz, n, 2).view(np.uint32).reshape(nz, n)
    wgt = (np.arange(n, dtype=np.uint64) % np.uint64(MOD)) + np.uint64(1)
    out = np.empty((nz, 2), np.uint64)
    for z0 in range(0, nz, chunk):
        v = v32[z0:z0 + chunk].astype(np.uint64)
        out[z0:z0 + chunk, 0] = v.sum(axis=1, dtype=np.uint64)
        out[z0:z0 + chunk, 1] = (v * wgt).sum(axis=1, dtype=np.uint64)
    return out


def slice_checksums_torch(vol_i32, chunk: int = 8) -> np.ndarray:
    """vol_i32: torch int32 tensor [z, y * x] viewing the device volume -> uint64[z, 2] (on the host)."""
    import torch

    nz, n = vol_i32.shape
    wgt = (torch.arange(n, dtype=torch.int64, device=vol_i32.device) % MOD) + 1
    out = torch.empty((nz, 2), dtype=torch.int64, device=vol_i32.device)
    for z0 in range(0, nz, chunk):
        v = vol_i32[z0:z0 + chunk].to(torch.int64) & 0xFFFFFFFF
        out[z0:z0 + chunk, 0] = v.sum(dim=1)                 # int64 arithmetic wraps mod 2^64: the same bits as uint64
        out[z0:z0 + chunk, 1] = (v * wgt).sum(dim=1)
    return out.cpu().numpy().view(np.uint64)


def array_checksum(a: np.ndarray) -> np.ndarray:
    """uint64[2] checksum of any array's bytes (padded to 4): the same (s1, s2) over its uint32 words."""
    b = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    if b.size % 4:
        b = np.concatenate([b, np.zeros(4 - b.size % 4, np.uint8)])
    v = b.view(np.uint32).astype(np.uint64)
    wgt = (np.arange(v.size, dtype=np.uint64) % np.uint64(MOD)) + np.uint64(1)
    return np.array([v.sum(dtype=np.uint64), (v * wgt).sum(dtype=np.uint64)], np.uint64)
