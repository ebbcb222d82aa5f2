"""BASELINE.json configs[1] at FULL size (512x640, pad 24, D=128, N_samples=128) against the oracle.

On the CPU the oracle needs minutes per frame, so the other GPU tests only sample a few hundred rays
at this size.  The oracle follows the device of its inputs, so here it runs on cuda:0 -- i.e. the
reference's own PyTorch-GPU path (same library primitives, fp32, TF32 off) -- and every one of the
327 680 pixels and 4.7 M voxels is compared.  The same run records that path's speed, the denominator of
the north star's ">= 10x the reference's single-GPU PyTorch rays/s" (written to
gpurun_out/torch_gpu_baseline.json; nothing is asserted on speed).
"""
import json
import math
import os

import pytest
import torch

from conftest import GOLDEN, ROOT
from oracle import mvsnerf_oracle as orc
from mvsnerf_b200 import backend, lib, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"


def psnr(a, b):
    return -10.0 * math.log10(float(((a - b) ** 2).mean()))      # utils.py:12-15


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)


@pytest.fixture(scope="module")
def full():
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    fn = backend.MVSNeRF().to(DEV)
    mvs = backend.MVSNet().to(DEV)
    backend.load_weights_npz(fn, mvs, os.path.join(GOLDEN, "mvsnerf_v0_weights.npz"))
    mvs.train()
    w = {k: v.to(DEV) for k, v in orc.load_weights_npz(os.path.join(GOLDEN, "mvsnerf_v0_weights.npz")).items()}
    sc = synthetic.make_scene(512, 640, pad=24, seed=0)
    d = sc.to(DEV)
    rays = synthetic.scene_rays(sc).to(DEV)
    record = {"config": "512x640, 3 views, pad 24, D=128, N_samples=128, fp32, torch %s" % torch.__version__}
    with torch.no_grad():
        orc.encode_volume(d.imgs_norm, d.proj_mats, sc.near_far, sc.pad, w)                      # warm-up (cuDNN plans)
        vol_ref, ms = timed(lambda: orc.encode_volume(d.imgs_norm, d.proj_mats, sc.near_far, sc.pad, w))
        record["volume_build_torch_gpu_ms"] = ms
        mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
        (vol, _, _), ms = timed(lambda: mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad))
        record["volume_build_ours_ms"] = ms

        def ref_frame():
            return orc.render_rays(rays, vol_ref, d.imgs_raw, d.pose_source, w, sc.H, sc.W, sc.near_far,
                                   float(sc.pad), n_samples=128)                                  # 5120-ray chunks
        orc.render_rays(rays[:5120], vol_ref, d.imgs_raw, d.pose_source, w, sc.H, sc.W, sc.near_far, float(sc.pad))
        (rgb_ref, depth_ref), ms = timed(ref_frame)
        record["render_torch_gpu_ms_per_frame"] = ms
        record["render_torch_gpu_rays_per_s"] = rays.shape[0] / ms * 1e3
        torch.backends.cuda.matmul.allow_tf32 = True                                              # torch-1.10 default
        _, ms = timed(ref_frame)
        torch.backends.cuda.matmul.allow_tf32 = False
        record["render_torch_gpu_tf32_rays_per_s"] = rays.shape[0] / ms * 1e3
    yield dict(sc=sc, d=d, rays=rays, fn=fn, vol=vol, vol_ref=vol_ref, rgb_ref=rgb_ref, depth_ref=depth_ref,
               record=record)
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "torch_gpu_baseline.json"), "w") as f:
        json.dump(record, f, indent=1)
    print("\n[torch-gpu baseline]", json.dumps(record))
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def test_full_volume_every_voxel(full):
    """K-A + K-B (+ FeatureNet) vs the oracle on all 8 x 128 x 176 x 208 values (SURVEY 8(d) gate)."""
    vol, ref = full["vol"], full["vol_ref"]
    assert vol.shape == ref.shape == (1, 8, 128, 176, 208)
    err = (vol - ref).abs().max().item()
    scale = ref.abs().max().item()
    full["record"]["volume_linf"] = err
    full["record"]["volume_absmax"] = scale
    assert err <= 1e-4 * scale + 1e-5, (err, scale)


@pytest.mark.parametrize("mode,tol", [(lib.MLP_FP32, 1e-4), (lib.MLP_TC_SPLIT, 1e-4), (lib.MLP_TC_HALF, 5e-3)])
def test_full_frame_every_pixel(full, mode, tol):
    """K-C on the oracle's own volume: RGB Linf over all 327 680 rays, depth, and the PSNR gate
    (|PSNR(kernel, target) - PSNR(oracle, target)| <= 0.05 dB against a fixed target image)."""
    sc, d = full["sc"], full["d"]
    with torch.no_grad():
        rgb, depth = backend.render_rays(full["rays"], full["vol_ref"], d.imgs_raw, d.pose_source, full["fn"],
                                         sc.near_far, float(sc.pad), N_samples=128, mlp_mode=mode)
    e_rgb = (rgb - full["rgb_ref"]).abs().max().item()
    e_depth = (depth - full["depth_ref"]).abs().max().item()
    target = d.imgs_raw[0, 0].permute(1, 2, 0).reshape(-1, 3)                       # reference-view photo as "GT"
    dpsnr = abs(psnr(rgb, target) - psnr(full["rgb_ref"], target))
    full["record"]["mode%d" % mode] = {"rgb_linf": e_rgb, "depth_linf": e_depth, "dpsnr_db": dpsnr,
                                       "psnr_vs_oracle_db": psnr(rgb, full["rgb_ref"]) if e_rgb > 0 else float("inf")}
    assert e_rgb <= tol, e_rgb
    assert e_depth <= (1e-3 if tol <= 1e-4 else 2e-2), e_depth
    assert dpsnr <= 0.05, dpsnr


def test_end_to_end_own_volume(full):
    """Whole path (our volume -> our render) against the oracle's whole path, fp32 mode."""
    sc, d = full["sc"], full["d"]
    with torch.no_grad():
        rgb, depth = backend.render_rays(full["rays"], full["vol"], d.imgs_raw, d.pose_source, full["fn"],
                                         sc.near_far, float(sc.pad), N_samples=128)
    e = (rgb - full["rgb_ref"]).abs().max().item()
    full["record"]["end_to_end_rgb_linf"] = e
    assert e <= 1e-4, e
    assert (depth - full["depth_ref"]).abs().max().item() <= 1e-3
