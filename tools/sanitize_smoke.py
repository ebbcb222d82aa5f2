#!/usr/bin/env python
"""Small-shape pass through EVERY kernel of the library, for compute-sanitizer (memcheck / initcheck / racecheck):

    compute-sanitizer --tool initcheck python tools/sanitize_smoke.py

K-F FeatureNet, K-A cost volume (the F5 zero border matters for initcheck), K-B CostRegNet (train + eval BN), the four
K-C render kernels (both entries), the peer sink, the backward kernel + reduce + both Adam kernels, layout helpers."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mvsnerf_b200 import backend, lib, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
fn, mvs = backend.MVSNeRF().to(dev), backend.MVSNet().to(dev).train()
backend.load_weights_npz(fn, mvs, os.path.join(ROOT, "tests", "golden", "mvsnerf_v0_weights.npz"))
sc = synthetic.make_scene(32, 32, pad=4, seed=0)
d = sc.to(dev)
only = set(sys.argv[1:])          # optionally restrict: encoder render backward
with torch.no_grad():
    vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad, return_color=True)
    if not only or "encoder" in only:
        mvs.eval()(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
        mvs.train()
    rays = synthetic.scene_rays(sc)[::3].contiguous().to(dev)
    if not only or "render" in only:
        for mode in (lib.MLP_FP32, lib.MLP_TC_HALF, lib.MLP_TC_SPLIT, lib.MLP_TC_PAIR):
            frame = torch.zeros(rays.shape[0] + 8, 4, device=dev)
            sink = lib.PeerSink()
            sink.frame[0], sink.n_peers, sink.first_pixel = frame.data_ptr(), 1, 5
            backend.render_rays(rays, vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad), N_samples=24, mlp_mode=mode,
                                out=(torch.empty(rays.shape[0], 3, device=dev), torch.empty(rays.shape[0], device=dev)), sink=sink)
            xyz, _, rd, z = backend.ray_marcher(rays[:100], N_samples=24)
            ndc = backend.get_ndc_coordinate(d.pose_source["w2cs"][0], d.pose_source["intrinsics"][0], xyz,
                                             torch.tensor([sc.W - 1.0, sc.H - 1.0], device=dev), near=sc.near_far[0],
                                             far=sc.near_far[1], pad=sc.pad)

            class A:
                use_color_volume = False
            backend.rendering(A(), d.pose_source, xyz, ndc, z, None, rd, volume_feature=backend.RefVolume(vol.contiguous().clone()),
                              imgs=d.imgs_raw, network_fn=fn, mlp_mode=mode)
    if not only or "render" in only:
        dirs = torch.randn(16, 24, 3, device=dev)
        backend.get_rays(dirs, d.pose_source["c2ws"][0], sc.near_far[0], sc.near_far[1])          # mvsn_make_rays
if not only or "backward" in only:
    # channels-last volume (what MVSNet.forward returns: element-wise Adam kernel), then checkpoint layout (planar kernel)
    for volume in (backend.RefVolume(vol.detach().clone()), backend.RefVolume(vol.detach().contiguous().clone())):
        tuner = backend.FineTuner(fn, volume, d.imgs_raw, d.pose_source, lr=1e-4)
        print("sanitize_smoke: FineTuner planar =", tuner.planar)
        for S in (32, 48, 128):
            xyz, _, rd, z = backend.ray_marcher(rays[:37], N_samples=S, perturb=1.0)
            ndc = backend.get_ndc_coordinate(d.pose_source["w2cs"][0], d.pose_source["intrinsics"][0], xyz,
                                             torch.tensor([sc.W - 1.0, sc.H - 1.0], device=dev), near=sc.near_far[0],
                                             far=sc.near_far[1], pad=sc.pad)
            tuner.step(xyz, ndc, z, rd, torch.rand(37, 3, device=dev), want_forward=True)
torch.cuda.synchronize()
print("sanitize_smoke: done")
