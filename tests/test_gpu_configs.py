"""BASELINE.json configs 3 and 5 at FULL size against the oracle (config 2 is tests/test_gpu_fullsize_oracle.py,
config 4's sharded batch is tests/test_gpu_seams.py + bench.py's strong-scaling entries):

  C3  Blender-shaped 800x800, pad 0 (README.md:90), near_far [2, 6] (data/blender.py:143), white_bkgd
  C5  LLFF-shaped 960x640 (the reference's legal size, data/llff.py:168 / SURVEY F9), pad 24

For each: the encoding volume (FeatureNet + cost volume + CostRegNet) on every voxel and the rendered frame on
every pixel against the oracle running on cuda:0 (the reference's own PyTorch-GPU path, TF32 off), plus a few
hundred sampled rays against the oracle on the CPU so that one comparison never touches library GPU kernels.
"""
import math
import os

import pytest
import torch

from conftest import GOLDEN
from oracle import mvsnerf_oracle as orc
from mvsnerf_b200 import backend, lib, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"

CONFIGS = {
    "c3_blender_800x800": dict(H=800, W=800, pad=0, near_far=(2.0, 6.0), white_bkgd=True, vol=(1, 8, 128, 200, 200)),
    "c5_llff_960x640": dict(H=640, W=960, pad=24, near_far=(2.125, 4.525), white_bkgd=False, vol=(1, 8, 128, 208, 288)),
}


def psnr(a, b):
    return -10.0 * math.log10(float(((a - b) ** 2).mean()))


@pytest.fixture(scope="module", params=list(CONFIGS))
def cfg(request):
    c = CONFIGS[request.param]
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    wpath = os.path.join(GOLDEN, "mvsnerf_v0_weights.npz")
    fn, mvs = backend.MVSNeRF().to(DEV), backend.MVSNet().to(DEV).train()
    backend.load_weights_npz(fn, mvs, wpath)
    w_cpu = orc.load_weights_npz(wpath)
    w = {k: v.to(DEV) for k, v in w_cpu.items()}
    sc = synthetic.make_scene(c["H"], c["W"], pad=c["pad"], seed=11, near_far=c["near_far"])
    d = sc.to(DEV)
    rays = synthetic.scene_rays(sc).to(DEV)
    with torch.no_grad():
        vol_ref = orc.encode_volume(d.imgs_norm, d.proj_mats, sc.near_far, sc.pad, w)
        vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
        rgb_ref, depth_ref = orc.render_rays(rays, vol_ref, d.imgs_raw, d.pose_source, w, sc.H, sc.W, sc.near_far,
                                             float(sc.pad), n_samples=128, white_bkgd=c["white_bkgd"])
    yield dict(c=c, name=request.param, sc=sc, d=d, rays=rays, fn=fn, vol=vol, vol_ref=vol_ref, rgb_ref=rgb_ref,
               depth_ref=depth_ref, w_cpu=w_cpu)
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    del vol, vol_ref
    torch.cuda.empty_cache()


def test_volume_every_voxel(cfg):
    vol, ref = cfg["vol"], cfg["vol_ref"]
    assert tuple(vol.shape) == tuple(ref.shape) == cfg["c"]["vol"]
    err, scale = (vol - ref).abs().max().item(), ref.abs().max().item()
    assert err <= 1e-4 * scale + 1e-5, (err, scale)


@pytest.mark.parametrize("mode,tol", [(lib.MLP_TC_SPLIT, 1e-4), (lib.MLP_TC_PAIR, 5e-3), (lib.MLP_FP32, 1e-4)])
def test_frame_every_pixel(cfg, mode, tol):
    sc, d, c = cfg["sc"], cfg["d"], cfg["c"]
    with torch.no_grad():
        rgb, depth = backend.render_rays(cfg["rays"], cfg["vol_ref"], d.imgs_raw, d.pose_source, cfg["fn"], sc.near_far,
                                         float(sc.pad), N_samples=128, white_bkgd=c["white_bkgd"], mlp_mode=mode)
    assert rgb.shape == (c["H"] * c["W"], 3)
    e_rgb = (rgb - cfg["rgb_ref"]).abs().max().item()
    e_depth = (depth - cfg["depth_ref"]).abs().max().item()
    target = d.imgs_raw[0, 0].permute(1, 2, 0).reshape(-1, 3)
    dpsnr = abs(psnr(rgb, target) - psnr(cfg["rgb_ref"], target))
    print(f"\n[{cfg['name']} mode {mode}] rgb Linf {e_rgb:.3e} depth Linf {e_depth:.3e} dPSNR {dpsnr:.2e} dB")
    assert e_rgb <= tol, e_rgb
    assert e_depth <= (1e-3 if tol <= 1e-4 else 2e-2) * (sc.near_far[1] / 4.525), e_depth
    assert dpsnr <= 0.05, dpsnr


def test_whole_path_own_volume_and_cpu_oracle_sample(cfg):
    """our volume -> our render (default mode) on every pixel, and 256 sampled rays against the oracle on the CPU."""
    sc, d, c = cfg["sc"], cfg["d"], cfg["c"]
    with torch.no_grad():
        rgb, depth = backend.render_rays(cfg["rays"], cfg["vol"], d.imgs_raw, d.pose_source, cfg["fn"], sc.near_far,
                                         float(sc.pad), N_samples=128, white_bkgd=c["white_bkgd"])
    assert (rgb - cfg["rgb_ref"]).abs().max().item() <= 1e-4
    idx = torch.randperm(cfg["rays"].shape[0], generator=torch.Generator().manual_seed(5))[:256]
    r_cpu = cfg["rays"][idx.to(DEV)].cpu().contiguous()
    ref_rgb, ref_depth = orc.render_rays(r_cpu, cfg["vol_ref"].cpu().contiguous(), sc.imgs_raw, sc.pose_source, cfg["w_cpu"],
                                         sc.H, sc.W, sc.near_far, float(sc.pad), n_samples=128, white_bkgd=c["white_bkgd"])
    with torch.no_grad():
        rgb_s, depth_s = backend.render_rays(r_cpu.to(DEV), cfg["vol_ref"], d.imgs_raw, d.pose_source, cfg["fn"],
                                             sc.near_far, float(sc.pad), N_samples=128, white_bkgd=c["white_bkgd"])
    assert (rgb_s.cpu() - ref_rgb).abs().max().item() <= 1e-4
    assert (depth_s.cpu() - ref_depth).abs().max().item() <= 1e-3 * (sc.near_far[1] / 4.525)
