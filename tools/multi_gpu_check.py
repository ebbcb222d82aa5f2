#!/usr/bin/env python
"""Multi-GPU check of the shipped sharding API on real NCCL / NVLink (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py

A small scene (64x64, pad 8) and two ray batches (a 4096-ray frame and a ragged 1001-ray batch) are rendered
  (a) by this rank alone,
  (b) sharded through distributed.render_rays_sharded with ONE NCCL all-gather of packed pixels,
  (c) sharded with kernel-epilogue NVLink peer stores into a distributed.PeerFrame,
and (b), (c) must equal (a) BIT FOR BIT on every rank, in every MLP mode.  Prints one JSON line on rank 0; exit code
1 on any mismatch.  tests/test_gpu_seams.py launches this when two GPUs are visible.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from mvsnerf_b200 import backend, lib, synthetic
    from mvsnerf_b200 import distributed as mdist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    fn, mvs = backend.MVSNeRF().to(dev), backend.MVSNet().to(dev).train()
    backend.load_weights_npz(fn, mvs, os.path.join(ROOT, "tests", "golden", "mvsnerf_v0_weights.npz"))
    sc = synthetic.make_scene(64, 64, pad=8, seed=3)
    d = sc.to(dev)
    report, bad = {"world": world}, 0
    with torch.no_grad():
        vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
        # replicated builds must be bit-identical (fixed-point BatchNorm statistics): compare with rank 0's copy
        vol0 = vol.contiguous().clone()
        dist.broadcast(vol0, 0)
        same = torch.tensor([float(torch.equal(vol, vol0))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        report["replicated_volume_bit_identical"] = bool(same.item() > 0)
        bad += int(same.item() < 1)
        rays_all = synthetic.scene_rays(sc).to(dev)
        for name, rays in (("frame_4096", rays_all), ("ragged_1001", rays_all[:1001].contiguous())):
            n = rays.shape[0]
            frame = mdist.PeerFrame(n, n_buffers=2)
            for mname, mode in (("fp32", lib.MLP_FP32), ("half", lib.MLP_TC_HALF), ("split", lib.MLP_TC_SPLIT),
                                ("pair", lib.MLP_TC_PAIR)):
                def render(r, sink=None, m=mode):
                    return backend.render_rays(r, vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad),
                                               N_samples=32, mlp_mode=m, sink=sink)
                rgb1, dep1 = render(rays)
                rgb_n, dep_n = mdist.render_rays_sharded(rays, render)
                rgb_p, dep_p = mdist.render_rays_sharded(rays, render, frame=frame)
                torch.cuda.synchronize()
                ok_n = torch.equal(rgb_n, rgb1) and torch.equal(dep_n, dep1)
                ok_p = torch.equal(rgb_p, rgb1) and torch.equal(dep_p, dep1)
                frame.rotate()
                flags = torch.tensor([float(ok_n), float(ok_p)], device=dev)
                dist.all_reduce(flags, op=dist.ReduceOp.MIN)
                report[f"{name}/{mname}"] = {"nccl_bit_equal": bool(flags[0] > 0), "peer_bit_equal": bool(flags[1] > 0)}
                bad += int(flags.min().item() < 1)
            frame.close()
    if rank == 0:
        report["ok"] = bad == 0
        print(json.dumps(report))
    dist.destroy_process_group()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
