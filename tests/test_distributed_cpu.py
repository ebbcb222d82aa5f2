"""CPU, gloo, world_size 2 (and 3 for ragged bands): the ray-sharding host logic of the multi-GPU path.
The per-band render is replaced by a deterministic per-ray function, so the test checks exactly what
the distributed layer owns: band bounds, gather order, equality with the single-process result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mvsnerf_b200.distributed import gather_bands, render_rays_sharded, shard_bounds


def fake_render(rays):
    rgb = torch.stack([rays[:, 0] * 2 + rays[:, 3], rays[:, 1] - rays[:, 4], rays[:, 2] * rays[:, 5]], -1)
    return rgb, rays[:, 6] + rays[:, 7]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rays = torch.randn(n, 8, generator=torch.Generator().manual_seed(0))
    rgb, depth = render_rays_sharded(rays, fake_render)
    ref_rgb, ref_depth = fake_render(rays)
    ok = torch.equal(rgb, ref_rgb) and torch.equal(depth, ref_depth)
    g = gather_bands(torch.full((shard_bounds(n, rank, world)[1] - shard_bounds(n, rank, world)[0],), float(rank)), n)
    ok = ok and all(float(g[i]) == r for r in range(world) for i in range(*shard_bounds(n, r, world)))
    out[rank] = int(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 4096), (2, 4097), (3, 1000)])
def test_sharded_render_matches_single_process(world, n):
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), n, out), nprocs=world, join=True)
    assert [out[r] for r in range(world)] == [1] * world


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 4096, 327680, 327681):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_passthrough():
    rays = torch.randn(10, 8)
    rgb, depth = render_rays_sharded(rays, fake_render)
    assert torch.equal(rgb, fake_render(rays)[0])
