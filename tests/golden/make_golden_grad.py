"""Generate tests/golden/grad_tiny_32x32_pad4.npz from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_grad.py

The fine-tuning step of the reference (train_mvs_nerf_finetuning_pl.py:140-189) differentiates
`renderer.rendering` with respect to the MLP parameters and the encoding volume (an nn.Parameter there) under
loss = img2mse(rgb, target).  This records that loss and those gradients, computed by the reference's own autograd on
the CPU, for two ray batches on the volume of tiny_32x32_pad4.npz: N_samples 32, and N_samples 128 with white_bkgd.
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


def one_case(R, sc, vol0, rays, S, white, target):
    ref = R.ref
    kw = dict(R.render_kwargs)
    kw["white_bkgd"] = white
    fn = R.render_kwargs["network_fn"]
    for p in fn.parameters():
        p.grad = None
    vol = vol0.clone().requires_grad_(True)
    xyz, ro, rd, z = ref.ray_utils.ray_marcher(rays, N_samples=S)
    inv_scale = torch.tensor([sc.W - 1, sc.H - 1])
    ndc = ref.utils.get_ndc_coordinate(sc.pose_source["w2cs"][0], sc.pose_source["intrinsics"][0].clone(), xyz, inv_scale,
                                       near=sc.near_far[0], far=sc.near_far[1], pad=sc.pad * 1.0)
    rgb, feat, weights, depth, alpha, _ = ref.renderer.rendering(R.args, sc.pose_source, xyz, ndc, z, ro, rd, vol,
                                                                sc.imgs_raw, **kw)
    loss = ref.utils.img2mse(rgb, target)
    loss.backward()
    out = dict(xyz=xyz.detach().numpy(), ndc=ndc.detach().numpy(), z=z.detach().numpy(), rays=rays.numpy(),
               target=target.numpy(), rgb=rgb.detach().numpy(), loss=np.float64(loss.item()),
               grad_volume=vol.grad.numpy())
    for k, p in fn.named_parameters():
        out["grad_mlp/" + k] = p.grad.detach().clone().numpy()
    return out


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    R = ref_shim.build_reference()
    sc = synthetic.make_scene(32, 32, pad=4, seed=1)                        # the scene of tiny_32x32_pad4.npz
    tiny = np.load(os.path.join(HERE, "tiny_32x32_pad4.npz"))
    vol0 = torch.from_numpy(tiny["volume"])
    assert np.array_equal(tiny["imgs_raw"], sc.imgs_raw.numpy())
    rays_all = synthetic.scene_rays(sc)
    g = torch.Generator().manual_seed(21)
    out = {}
    for tag, S, white, n in (("s32", 32, False, 192), ("s128w", 128, True, 96)):
        rays = rays_all[torch.randperm(rays_all.shape[0], generator=g)[:n]].contiguous()
        target = torch.rand(n, 3, generator=g)
        c = one_case(R, sc, vol0, rays, S, white, target)
        print(tag, "loss", c["loss"], "|g_vol|max", np.abs(c["grad_volume"]).max(),
              "nonzero voxels", int((c["grad_volume"] != 0).sum()))
        # the volume gradient is sparse (only voxels the rays touch): store its nonzero entries
        gv = c.pop("grad_volume").reshape(-1)
        nz = np.nonzero(gv)[0]
        c["grad_volume_idx"] = nz.astype(np.int64)
        c["grad_volume_val"] = gv[nz]
        c["grad_volume_shape"] = np.array(vol0.shape)
        for k, v in c.items():
            out[tag + "/" + k] = v
    path = os.path.join(HERE, "grad_tiny_32x32_pad4.npz")
    np.savez_compressed(path, **out)
    print(os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
