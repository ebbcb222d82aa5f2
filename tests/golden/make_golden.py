"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference through oracle/ref_shim.py (stubs for the absent third-party
modules, F5 border pinned to zero, MVSNet in train mode as every shipped caller uses it)
and records, for seeded synthetic scenes, the inputs and the reference's outputs at every
stage boundary of the hot path.  Large tensors are recorded at a seeded subset of voxels.
The fixtures travel to the GPU box; the reference does not.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from mvsnerf_b200 import synthetic  # noqa: E402


def export_weights(R):
    out = {}
    for k, v in R.render_kwargs["network_fn"].state_dict().items():
        out["mlp/" + k] = v.detach().cpu().numpy()
    for k, v in R.mvsnet.state_dict().items():
        out["mvs/" + k] = v.detach().cpu().numpy()
    np.savez(os.path.join(HERE, "mvsnerf_v0_weights.npz"), **out)
    print("weights:", len(out), "tensors")


def scene_inputs(sc):
    return dict(imgs_raw=sc.imgs_raw.numpy(), imgs_norm=sc.imgs_norm.numpy(),
                proj_mats=sc.proj_mats.numpy(), w2cs=sc.pose_source["w2cs"].numpy(),
                c2ws=sc.pose_source["c2ws"].numpy(), intrinsics=sc.pose_source["intrinsics"].numpy(),
                near_far=np.array(sc.near_far, dtype=np.float64), c2w_target=sc.c2w_target.numpy(),
                HW_pad=np.array([sc.H, sc.W, sc.pad]))


def run_render(R, sc, rays, vol, S, white_bkgd=False, lindisp=False):
    ref = R.ref
    kw = dict(R.render_kwargs)
    kw["white_bkgd"] = white_bkgd
    with torch.no_grad():
        xyz, ro, rd, z = ref.ray_utils.ray_marcher(rays, N_samples=S, lindisp=lindisp)
        inv_scale = torch.tensor([sc.W - 1, sc.H - 1])
        ndc = ref.utils.get_ndc_coordinate(sc.pose_source["w2cs"][0], sc.pose_source["intrinsics"][0].clone(),
                                           xyz, inv_scale, near=sc.near_far[0], far=sc.near_far[1],
                                           pad=sc.pad * 1.0, lindisp=lindisp)
        rgb, feat, weights, depth, alpha, _ = ref.renderer.rendering(
            R.args, sc.pose_source, xyz, ndc, z, ro, rd, vol, sc.imgs_raw, **kw)
    return dict(xyz=xyz, ndc=ndc, z=z, rgb=rgb, feat=feat, weights=weights, depth=depth, alpha=alpha)


def stage_outputs(R, sc, lindisp=False):
    """MVSNet.forward with every intermediate the oracle/kernels are checked against."""
    m = R.mvsnet
    with torch.no_grad():
        imgs = sc.imgs_norm
        B, V, _, H, W = imgs.shape
        feats = m.feature(imgs.reshape(B * V, 3, H, W))
        vol, feats_l, depth_values = m(imgs, sc.proj_mats, sc.near_far, pad=sc.pad, lindisp=lindisp)
        cv, masks = m.build_volume_costvar_img(imgs, feats.view(B, V, *feats.shape[1:]), sc.proj_mats,
                                               depth_values, pad=sc.pad)
        c0 = m.cost_reg_2.conv0(cv)
        c1 = m.cost_reg_2.conv1(c0)
    return dict(feats=feats, depth_values=depth_values[0], cost_volume=cv[0], in_masks=masks[0],
                conv0=c0[0], conv1=c1[0], volume=vol)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    R = ref_shim.build_reference()
    export_weights(R)

    # ---- G1 "tiny": 32x32 image, pad 4 -> 8x8 features, 16x16x128 volume; full tensors ----
    sc = synthetic.make_scene(32, 32, pad=4, seed=1)
    st = stage_outputs(R, sc)
    rays = synthetic.scene_rays(sc)
    r32 = run_render(R, sc, rays, st["volume"], 32)
    g = torch.Generator().manual_seed(7)
    nvox = st["cost_volume"][0].numel()
    vox_idx = torch.randperm(nvox, generator=g)[:4096]
    out = scene_inputs(sc)
    out.update(rays=rays.numpy(), feats=st["feats"].numpy(), depth_values=st["depth_values"].numpy(),
               vox_idx=vox_idx.numpy(),
               cost_volume_sub=st["cost_volume"].reshape(41, -1)[:, vox_idx].numpy(),
               in_masks_sub=st["in_masks"].reshape(3, -1)[:, vox_idx].numpy(),
               cost_volume_chsum=st["cost_volume"].double().sum((1, 2, 3)).numpy(),
               conv0_sub=st["conv0"].reshape(8, -1)[:, vox_idx].numpy(),
               conv1=st["conv1"].numpy(),
               volume=st["volume"].numpy(),
               ndc=r32["ndc"].numpy(), z=r32["z"].numpy(), rgb=r32["rgb"].numpy(),
               depth=r32["depth"].numpy(), weights=r32["weights"].numpy(), alpha=r32["alpha"].numpy(),
               feat_first128=r32["feat"][:128].numpy())
    # variants on the same volume: white background; lindisp sampling (S=16)
    rw = run_render(R, sc, rays[:256], st["volume"], 32, white_bkgd=True)
    out.update(rgb_white256=rw["rgb"].numpy())
    rl = run_render(R, sc, rays[:256], st["volume"], 16, lindisp=True)
    out.update(rgb_lindisp256=rl["rgb"].numpy(), depth_lindisp256=rl["depth"].numpy(),
               ndc_lindisp256=rl["ndc"].numpy())
    np.savez_compressed(os.path.join(HERE, "tiny_32x32_pad4.npz"), **out)
    print("tiny: rgb mean", r32["rgb"].mean(0), "alpha mean", r32["alpha"].mean().item())

    # ---- G1b: lindisp depth planes in the volume build (pad 0: 8x8x128 volume) ----
    sc0 = synthetic.make_scene(32, 32, pad=0, seed=2)
    st0 = stage_outputs(R, sc0, lindisp=True)
    out = scene_inputs(sc0)
    out.update(depth_values=st0["depth_values"].numpy(), volume=st0["volume"].numpy(),
               feats=st0["feats"].numpy(),
               cost_volume_chsum=st0["cost_volume"].double().sum((1, 2, 3)).numpy())
    np.savez_compressed(os.path.join(HERE, "tiny_32x32_pad0_lindisp.npz"), **out)

    # ---- G2 = BASELINE config 1: 64x64 crop, 3 views, pad 24, N_samples 32 (+ a 128-sample slice) ----
    sc = synthetic.make_scene(64, 64, pad=24, seed=0)
    st = stage_outputs(R, sc)
    rays = synthetic.scene_rays(sc)
    r32 = run_render(R, sc, rays, st["volume"], 32)
    r128 = run_render(R, sc, rays[1024:1280], st["volume"], 128)
    nvox = st["volume"][0, 0].numel()
    vox_idx = torch.randperm(nvox, generator=g)[:8192]
    out = scene_inputs(sc)
    out.update(rays=rays.numpy(), vox_idx=vox_idx.numpy(),
               volume_sub=st["volume"][0].reshape(8, -1)[:, vox_idx].numpy(),
               volume_chsum=st["volume"][0].double().sum((1, 2, 3)).numpy(),
               cost_volume_sub=st["cost_volume"].reshape(41, -1)[:, vox_idx].numpy(),
               in_masks_sub=st["in_masks"].reshape(3, -1)[:, vox_idx].numpy(),
               rgb=r32["rgb"].numpy(), depth=r32["depth"].numpy(),
               rgb128=r128["rgb"].numpy(), depth128=r128["depth"].numpy(),
               weights128=r128["weights"].numpy())
    np.savez_compressed(os.path.join(HERE, "c1_64x64_pad24.npz"), **out)
    print("c1: rgb mean", r32["rgb"].mean(0), "vol absmax", st["volume"].abs().max().item())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
