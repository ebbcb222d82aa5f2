"""Ray-sharded rendering across the GPUs of one box (one process per GPU, torch.distributed).

The reference has no working distributed path (its DDP flag is dead code); the
hot path shards trivially because rays are independent units:

  * the ray batch is split into `world` contiguous bands (row bands of the image);
  * the encoding volume, the three source images and the packed MLP weights are replicated -- each
    rank builds them itself from the same inputs (bit-identical, no communication);
  * every rank renders its band with the fused kernel, then ONE all-gather of rgb (+ depth)
    assembles the frame on every rank (491 520 + 163 840 bytes per rank at 512x640 over 8 ranks).

There is no exchange step inside the render, so no collective sits on the data path.  Because the
kernel's pixels do not depend on how rays are grouped (tests/test_gpu_parity.py checks partitions
bit-for-bit), the gathered frame equals the single-GPU frame exactly.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced band [lo, hi) of `n` rays for `rank` (bands differ by at most one ray)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_bands(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather per-rank bands of a [n_local, ...] tensor into [n_total, ...] on every rank.

    Uses a single all_gather_into_tensor when the bands are equal (every BASELINE config), and a
    padded gather otherwise."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    n_max = max(hi - lo for lo, hi in sizes)
    tail = tuple(local.shape[1:])
    if all(hi - lo == n_max for lo, hi in sizes):
        out = torch.empty((n_total,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    padded = torch.zeros((n_max,) + tail, dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    buf = torch.empty((world * n_max,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    parts = [buf[r * n_max: r * n_max + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, 0)


def render_rays_sharded(rays: torch.Tensor, render_fn, group=None):
    """Render `rays` [N, 8] (replicated on every rank) with `render_fn(rays_band) -> (rgb, depth)` on
    this rank's band and return the assembled (rgb [N,3], depth [N]) on every rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return render_fn(rays)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = rays.shape[0]
    lo, hi = shard_bounds(n, rank, world)
    rgb, depth = render_fn(rays[lo:hi])
    return gather_bands(rgb, n, group), gather_bands(depth, n, group)
