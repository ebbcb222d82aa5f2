"""Ray-sharded rendering across the GPUs of one box (one process per GPU, torch.distributed).

The reference has no working distributed path (its DDP flag is dead code); the
hot path shards trivially because rays are independent units:

  * the ray batch is split into `world` contiguous bands (row bands of the image);
  * the encoding volume, the three source images and the packed MLP weights are replicated -- each
    rank builds them itself from the same inputs (bit-identical, no communication);
  * every rank renders its band with the fused kernel and the frame is assembled on every rank.

Two ways to assemble the frame, same result bit for bit:

  PeerFrame (NVLink peer stores, the B200-native path).  Every rank owns one [n_pixels, 4] (r, g, b, depth)
  copy of the frame in exportable device memory (CUDA IPC through the C ABI, `mvsn_peer_buffer_*`); the render
  kernel's compositing epilogue stores each finished pixel into ALL copies (16 bytes per pixel per peer over
  NVLink / NVSwitch).  There is no gather pass and no staging buffer: when the kernels have completed the frame
  is complete everywhere; `PeerFrame.complete()` is the one-element stream-level barrier that says so.

  gather_bands (NCCL).  ONE `all_gather_into_tensor` of the packed [n_band, 4] pixels (rgb and depth travel
  in the same buffer; round 1 issued two collectives).  Fallback for boxes without peer access, and the
  baseline the peer-store path is measured against in bench.py.

There is no exchange step inside the render, so no collective sits on the data path.  Because the
kernel's pixels do not depend on how rays are grouped (tests/test_gpu_parity.py checks partitions
bit-for-bit), the assembled frame equals the single-GPU frame exactly.
"""
from __future__ import annotations

import ctypes as C

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


def pack_pixels(rgb: torch.Tensor, depth: torch.Tensor) -> torch.Tensor:
    """[n,3] + [n] -> [n,4] (r, g, b, depth): the frame texel both assembly paths move."""
    return torch.cat([rgb, depth.unsqueeze(-1)], -1)


class _RawCuda:
    """A raw device pointer as a __cuda_array_interface__ object (for torch.as_tensor)."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class PeerFrame:
    """`n_buffers` copies-per-rank of a [n_pixels, 4] frame that every rank's render kernel writes into directly.

        frame = PeerFrame(n_pixels)                        # collective: allocates, exchanges IPC handles, maps peers
        lo, hi = shard_bounds(n_pixels, rank, world)
        backend.render_rays(rays[lo:hi], ..., sink=frame.sink(lo))
        frame.complete()                                   # all ranks' kernels done => frame complete everywhere
        rgb, depth = frame.pixels()                        # views of this rank's copy
        frame.rotate()                                     # next step writes the other buffer

    A buffer may be written again only after every rank has consumed it; with n_buffers = 2 the `complete()` of the
    following step provides that order for stream-ordered consumers."""

    def __init__(self, n_pixels: int, group=None, n_buffers: int = 2, device=None):
        from . import lib as _lib
        self._lib = _lib.load()
        self._libmod = _lib
        self.group = group
        self.n_pixels = int(n_pixels)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > _lib.MAX_PEERS:
            raise RuntimeError(f"PeerFrame: {self.world} ranks > MVSN_MAX_PEERS = {_lib.MAX_PEERS}")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._own, self._peers, self._tensors = [], [], []
        nbytes = self.n_pixels * 16
        with torch.cuda.device(self.device):
            for _ in range(n_buffers):
                ptr = C.c_void_p()
                handle = C.create_string_buffer(_lib.PEER_HANDLE_BYTES)
                _lib.check(self._lib.mvsn_peer_buffer_create(nbytes, C.byref(ptr), handle), "mvsn_peer_buffer_create")
                self._own.append(ptr.value)
                handles = [None] * self.world
                if self.world > 1:
                    dist.all_gather_object(handles, handle.raw, group=group)
                ptrs = []
                for r in range(self.world):
                    if r == self.rank:
                        ptrs.append(ptr.value)
                        continue
                    pp = C.c_void_p()
                    _lib.check(self._lib.mvsn_peer_buffer_open(handles[r], C.byref(pp)), "mvsn_peer_buffer_open")
                    ptrs.append(pp.value)
                self._peers.append(ptrs)
                t = torch.as_tensor(_RawCuda(ptr.value, self.n_pixels * 4), device=self.device).view(self.n_pixels, 4)
                t.zero_()
                self._tensors.append(t)
        self._flag = torch.zeros(1, device=self.device)
        self._cur = 0
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=group)

    def sink(self, first_pixel: int):
        s = self._libmod.PeerSink()
        for r, p in enumerate(self._peers[self._cur]):
            s.frame[r] = p
        s.n_peers, s.first_pixel = self.world, int(first_pixel)
        return s

    def complete(self):
        """Stream-level barrier: returns (asynchronously) once every rank's preceding work on its stream -- the render
        launch with its peer stores -- has completed, i.e. the current buffer holds the whole frame on every rank."""
        if self.world > 1:
            dist.all_reduce(self._flag, group=self.group)

    def pixels(self):
        t = self._tensors[self._cur]
        return t[:, :3], t[:, 3]

    def tensor(self):
        return self._tensors[self._cur]

    def rotate(self):
        self._cur = (self._cur + 1) % len(self._tensors)

    def close(self):
        if not self._own:
            return
        torch.cuda.synchronize(self.device)
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)              # nobody unmaps while a peer may still be storing
        self._tensors = []
        with torch.cuda.device(self.device):
            for ptrs, own in zip(self._peers, self._own):
                for r, p in enumerate(ptrs):
                    if r != self.rank:
                        self._lib.mvsn_peer_buffer_close(C.c_void_p(p))
            if self.world > 1 and dist.is_initialized():
                dist.barrier(group=self.group)          # every mapping closed before the owners free
            for own in self._own:
                self._lib.mvsn_peer_buffer_destroy(C.c_void_p(own))
        self._own, self._peers = [], []

    def __del__(self):
        try:
            if self._own and not (self.world > 1):
                self.close()
        except Exception:
            pass


def render_rays_sharded(rays: torch.Tensor, render_fn, group=None, frame: PeerFrame | None = None):
    """Render `rays` [N, 8] (replicated on every rank) on this rank's band and return the assembled
    (rgb [N,3], depth [N]) on every rank.

    frame is None : `render_fn(rays_band) -> (rgb, depth)`; ONE NCCL all-gather of the packed [n_band, 4] pixels.
    frame given   : `render_fn(rays_band, sink=...)` must route its pixels into the sink (backend.render_rays does);
                    no gather pass -- the returned tensors are views of this rank's copy of the frame, valid in
                    stream order after the barrier enqueued here."""
    if not (dist.is_available() and dist.is_initialized()):
        if frame is None:
            return render_fn(rays)
        render_fn(rays, sink=frame.sink(0))
        return frame.pixels()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = rays.shape[0]
    lo, hi = shard_bounds(n, rank, world)
    if frame is not None:
        if frame.n_pixels != n:
            raise RuntimeError(f"PeerFrame holds {frame.n_pixels} pixels, the batch has {n} rays")
        render_fn(rays[lo:hi], sink=frame.sink(lo))
        frame.complete()
        return frame.pixels()
    rgb, depth = render_fn(rays[lo:hi])
    px = gather_bands(pack_pixels(rgb, depth), n, group)
    return px[:, :3], px[:, 3]
