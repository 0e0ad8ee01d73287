"""View-batch sharding: independent camera views of a fly-through are partitioned across the GPUs of one box.

A view reads only immutable data (map, textures, lightmaps) and writes its own framebuffer, so there is NO exchange step and
no collective on the hot path: every rank holds a full replica of the static set and renders a contiguous block of the batch
(contiguous blocks keep the host-side visibility coherent from one pose to the next).  torch.distributed is used only for the
plumbing around the path: the timing barrier, max-over-ranks of the elapsed time and the gather of per-view digests in tests.

The reference has no multi-device path at all (one rayon pool on one host, src/square_wheel/host.rs:62-81); the unit of
parallel work there is a screen rect (rect_splitting.rs:23), here it is a whole view.
"""
from __future__ import annotations

import hashlib
from typing import Iterable, List, Sequence

import numpy as np


def partition_views(num_views: int, world_size: int, rank: int) -> range:
    """View i belongs to rank floor(i * world_size / num_views): contiguous blocks whose sizes differ by at most one."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    if num_views < 0:
        raise ValueError("negative view count")
    # first index owned by rank r = ceil(r * num_views / world_size)
    start = -((-rank * num_views) // world_size)
    stop = -((-(rank + 1) * num_views) // world_size)
    return range(start, stop)


def owner_of_view(view_index: int, num_views: int, world_size: int) -> int:
    return view_index * world_size // num_views


def resample_camera_path(path: np.ndarray, factor: int, seed: int = 0x5157, jitter: float = 4.0) -> np.ndarray:
    """BASELINE config 5: the fixed pose path resampled x`factor` with seeded jitter.  `path` is [n, 5] =
    position xyz, azimuth, elevation; consecutive poses are interpolated linearly (the path is a closed loop) and every pose
    gets a small seeded offset so that no two views of the batch are identical."""
    path = np.asarray(path, np.float32)
    n = len(path)
    if n == 0 or factor <= 0:
        return path.copy()
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty((n * factor, 5), np.float32)
    for i in range(n):
        a, b = path[i], path[(i + 1) % n]
        d_az = (b[3] - a[3] + np.pi) % (2 * np.pi) - np.pi  # shortest way round
        for k in range(factor):
            t = k / factor
            p = a + (b - a) * t
            p[3] = a[3] + d_az * t
            out[i * factor + k] = p
    out[:, :3] += rng.uniform(-jitter, jitter, size=(n * factor, 3)).astype(np.float32)
    out[:, 3] += rng.uniform(-0.02, 0.02, size=n * factor).astype(np.float32)
    return out


def frame_digest(jobs: np.ndarray, cmds: np.ndarray, polygon_points: np.ndarray) -> str:
    """Digest of everything the host hands to the C-ABI for one view (inputs of S1 and S2)."""
    h = hashlib.sha1()
    for a in (jobs, cmds, polygon_points):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def picture_digest(pixels: np.ndarray) -> str:
    return hashlib.sha1(np.ascontiguousarray(pixels).tobytes()).hexdigest()


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Elapsed time of a multi-rank run = the slowest rank's (device-timed) elapsed time."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_digests(local: Sequence[str], view_indices: Iterable[int], num_views: int, dist=None) -> List[str]:
    """All ranks' per-view digests in view order (test / verification plumbing; not on the hot path)."""
    pairs = list(zip(view_indices, local))
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        gathered = [pairs]
    else:
        gathered = [None] * dist.get_world_size()
        dist.all_gather_object(gathered, pairs)
    out = [None] * num_views
    for part in gathered:
        for idx, d in part:
            if out[idx] is not None:
                raise RuntimeError(f"view {idx} rendered twice")
            out[idx] = d
    missing = [i for i, d in enumerate(out) if d is None]
    if missing:
        raise RuntimeError(f"views not rendered by any rank: {missing[:8]}")
    return out
