"""GPU parity tests: the CUDA path (through the C ABI, via mvsnerf_b200.backend) against
 (a) golden vectors from the unmodified reference (tests/golden/*.npz) and
 (b) the CPU oracle (oracle/mvsnerf_oracle.py) on seeded synthetic scenes.
Tolerances are the north-star gates: RGB Linf <= 1e-4 for the fp32 mode."""
import os

import pytest
import torch

from conftest import GOLDEN
from oracle import mvsnerf_oracle as orc
from mvsnerf_b200 import backend, synthetic

pytestmark = pytest.mark.gpu

RGB_TOL = 1e-4          # north-star fp32 gate
DEPTH_TOL = 1e-3        # depth in scene units (near/far ~ 2..5)
DEV = "cuda"


def pose_of(g, dev=DEV):
    return {k: g[k].to(dev) for k in ("w2cs", "c2ws", "intrinsics")}


def dims(g):
    return tuple(int(v) for v in g["HW_pad"])


@pytest.fixture(scope="module")
def nets():
    fn = backend.MVSNeRF().to(DEV)
    mvs = backend.MVSNet().to(DEV)
    backend.load_weights_npz(fn, mvs, os.path.join(GOLDEN, "mvsnerf_v0_weights.npz"))
    mvs.train()
    return fn, mvs


class Args:
    feat_dim = 20
    img_downscale = 1.0
    use_color_volume = False
    net_type = "v0"


# ------------------------------------------------------------------------------------------------
# render kernel vs golden (reference outputs)
# ------------------------------------------------------------------------------------------------
def test_rendering_signature_path_vs_golden(golden_tiny, nets):
    g = golden_tiny
    fn, _ = nets
    rays = g["rays"]
    pts, z = orc.march_rays(rays, 32)
    out = backend.rendering(Args, pose_of(g), pts.to(DEV), g["ndc"].to(DEV), z.to(DEV), rays[:, :3].to(DEV),
                            rays[:, 3:6].to(DEV), g["volume"].to(DEV), g["imgs_raw"].to(DEV), network_fn=fn,
                            perturb=0.0, N_importance=0, network_fine=None, use_viewdirs=True, raw_noise_std=0.0)
    rgb, feat, wts, depth, alpha, extra = out
    assert extra == {}
    assert (rgb.cpu() - g["rgb"]).abs().max() < RGB_TOL
    assert (depth.cpu() - g["depth"]).abs().max() < DEPTH_TOL
    assert (wts.cpu() - g["weights"]).abs().max() < RGB_TOL
    assert (alpha.cpu() - g["alpha"]).abs().max() < RGB_TOL
    assert (feat[:128].cpu() - g["feat_first128"]).abs().max() < 1e-4
    # tighter than the gate: the fp32 kernel should sit at rounding level
    assert (rgb.cpu() - g["rgb"]).abs().max() < 2e-5


def test_render_rays_fused_path_vs_golden(golden_tiny, nets):
    g = golden_tiny
    fn, _ = nets
    H, W, pad = dims(g)
    nf = g["near_far"].tolist()
    rgb, depth = backend.render_rays(g["rays"].to(DEV), g["volume"].to(DEV), g["imgs_raw"].to(DEV), pose_of(g), fn,
                                     nf, float(pad), N_samples=32)
    assert (rgb.cpu() - g["rgb"]).abs().max() < RGB_TOL
    assert (depth.cpu() - g["depth"]).abs().max() < DEPTH_TOL
    rgb, _ = backend.render_rays(g["rays"][:256].to(DEV), g["volume"].to(DEV), g["imgs_raw"].to(DEV), pose_of(g),
                                 fn, nf, float(pad), N_samples=32, white_bkgd=True)
    assert (rgb.cpu() - g["rgb_white256"]).abs().max() < RGB_TOL
    rgb, depth = backend.render_rays(g["rays"][:256].to(DEV), g["volume"].to(DEV), g["imgs_raw"].to(DEV), pose_of(g),
                                     fn, nf, float(pad), N_samples=16, lindisp=True)
    assert (rgb.cpu() - g["rgb_lindisp256"]).abs().max() < RGB_TOL
    assert (depth.cpu() - g["depth_lindisp256"]).abs().max() < DEPTH_TOL


# ------------------------------------------------------------------------------------------------
# encoding volume vs golden
# ------------------------------------------------------------------------------------------------
def test_cost_volume_vs_golden(golden_tiny, nets):
    g = golden_tiny
    _, mvs = nets
    H, W, pad = dims(g)
    cost, masks = mvs.build_volume_costvar_img(g["imgs_norm"].to(DEV), g["feats"][None].to(DEV),
                                               g["proj_mats"].to(DEV), g["depth_values"][None].to(DEV), pad=pad)
    idx = g["vox_idx"]
    assert torch.equal(masks[0].reshape(3, -1)[:, idx].cpu(), g["in_masks_sub"])
    assert torch.allclose(cost[0].reshape(41, -1)[:, idx].cpu(), g["cost_volume_sub"], rtol=1e-5, atol=1e-4)
    assert torch.allclose(cost[0].double().sum((1, 2, 3)).cpu(), g["cost_volume_chsum"], rtol=1e-5, atol=1e-2)
    assert cost[0, :3, :, :pad].abs().max() == 0        # F5: border of the ref-RGB channels is zero


def test_featurenet_vs_golden(golden_tiny, golden_c1, nets):
    """K-F (mvsn_featurenet_forward) against the unmodified reference's FeatureNet outputs."""
    _, mvs = nets
    g = golden_tiny
    feats = mvs.feature(g["imgs_norm"][0].to(DEV))
    assert feats.shape == g["feats"].shape
    assert (feats.cpu() - g["feats"]).abs().max() < 1e-4 * g["feats"].abs().max() + 1e-5


@pytest.mark.parametrize("V,H,W", [(3, 128, 160), (3, 36, 52), (2, 30, 34), (1, 17, 23), (4, 8, 12), (3, 4, 4)])
def test_featurenet_vs_oracle_shapes(nets, weights, V, H, W):
    """Ragged sizes: widths that are not multiples of 4 at some level (scalar row path), odd sizes
    (k5 s2 p2 output = ceil(n/2)), a single view, a single strip."""
    _, mvs = nets
    x = torch.randn(V, 3, H, W, generator=torch.Generator().manual_seed(V * 100 + H))
    ref = orc.feature_net(x, weights)
    out = mvs.feature(x.to(DEV))
    assert out.shape == ref.shape
    assert (out.cpu() - ref).abs().max() < 1e-4 * ref.abs().max() + 1e-5


def test_costreg_vs_golden(golden_tiny, nets):
    g = golden_tiny
    _, mvs = nets
    H, W, pad = dims(g)
    cost, _ = orc.cost_volume(g["imgs_norm"][0], g["feats"], g["proj_mats"][0], g["depth_values"], pad)
    vol = mvs.cost_reg_2(cost[None].to(DEV))
    assert vol.shape == g["volume"].shape
    assert (vol.cpu() - g["volume"]).abs().max() < 5e-4     # |volume| ~ 10; fp32 BN statistics noise
    # the returned tensor is a channels-last view: rendering consumes it without a copy
    assert vol[0].permute(1, 2, 3, 0).is_contiguous()


def test_mvsnet_forward_vs_golden(golden_tiny, golden_tiny_lindisp, golden_c1, nets):
    _, mvs = nets
    for g, lindisp in ((golden_tiny, False), (golden_tiny_lindisp, True)):
        H, W, pad = dims(g)
        vol, feats, depth_values = mvs(g["imgs_norm"].to(DEV), g["proj_mats"].to(DEV), g["near_far"].tolist(),
                                       pad=pad, lindisp=lindisp)
        assert (feats[0].cpu() - g["feats"]).abs().max() < 1e-3
        assert torch.allclose(depth_values[0].cpu(), g["depth_values"], atol=1e-6)
        assert (vol.cpu() - g["volume"]).abs().max() < 2e-3
    g = golden_c1                                          # BASELINE config 1
    H, W, pad = dims(g)
    vol, _, _ = mvs(g["imgs_norm"].to(DEV), g["proj_mats"].to(DEV), g["near_far"].tolist(), pad=pad)
    idx = g["vox_idx"]
    assert (vol[0].reshape(8, -1)[:, idx].cpu() - g["volume_sub"]).abs().max() < 2e-3


def test_config1_end_to_end_vs_golden(golden_c1, nets):
    """BASELINE config 1 through the CUDA path end to end: volume build + render, vs reference RGB."""
    g = golden_c1
    fn, mvs = nets
    H, W, pad = dims(g)
    nf = g["near_far"].tolist()
    vol, _, _ = mvs(g["imgs_norm"].to(DEV), g["proj_mats"].to(DEV), nf, pad=pad)
    rgb, depth = backend.render_rays(g["rays"].to(DEV), vol, g["imgs_raw"].to(DEV), pose_of(g), fn, nf, float(pad),
                                     N_samples=32)
    assert (rgb.cpu() - g["rgb"]).abs().max() < RGB_TOL
    rgb, depth = backend.render_rays(g["rays"][1024:1280].to(DEV), vol, g["imgs_raw"].to(DEV), pose_of(g), fn, nf,
                                     float(pad), N_samples=128)
    assert (rgb.cpu() - g["rgb128"]).abs().max() < RGB_TOL
    assert (depth.cpu() - g["depth128"]).abs().max() < DEPTH_TOL


# ------------------------------------------------------------------------------------------------
# vs the CPU oracle on fresh seeded scenes (sizes the oracle finishes in seconds)
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def mid_scene(weights, nets):
    sc = synthetic.make_scene(128, 160, pad=8, seed=5)          # h=32, w=40 -> 48 x 56 padded
    vol = orc.encode_volume(sc.imgs_norm, sc.proj_mats, sc.near_far, sc.pad, weights)
    return sc, vol


def test_volume_vs_oracle_mid(mid_scene, nets):
    sc, vol_ref = mid_scene
    _, mvs = nets
    d = sc.to(DEV)
    vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
    err = (vol.cpu() - vol_ref).abs().max()
    assert err < 1e-4 * vol_ref.abs().max() + 1e-3, err


@pytest.mark.parametrize("S,nrays", [(128, 2048), (32, 1000), (24, 333), (192, 130), (1, 64), (128, 1)])
def test_render_vs_oracle_shapes(mid_scene, nets, weights, S, nrays):
    """Ragged / edge shapes: S not dividing 128, S > 128 (transmittance carry), N not a tile multiple."""
    sc, vol_ref = mid_scene
    fn, _ = nets
    rays = synthetic.scene_rays(sc)
    g = torch.Generator().manual_seed(S * 1000 + nrays)
    rays = rays[torch.randperm(rays.shape[0], generator=g)[:nrays]]
    rgb_ref, depth_ref = orc.render_rays(rays, vol_ref, sc.imgs_raw, sc.pose_source, weights, sc.H, sc.W,
                                         sc.near_far, float(sc.pad), n_samples=S)
    d = sc.to(DEV)
    rgb, depth = backend.render_rays(rays.to(DEV), vol_ref.to(DEV), d.imgs_raw, d.pose_source, fn, sc.near_far,
                                     float(sc.pad), N_samples=S)
    assert (rgb.cpu() - rgb_ref).abs().max() < RGB_TOL
    assert (depth.cpu() - depth_ref).abs().max() < DEPTH_TOL


def test_render_empty_and_errors(mid_scene, nets):
    sc, vol_ref = mid_scene
    fn, _ = nets
    d = sc.to(DEV)
    rgb, depth = backend.render_rays(torch.empty(0, 8, device=DEV), vol_ref.to(DEV), d.imgs_raw, d.pose_source, fn,
                                     sc.near_far, float(sc.pad), N_samples=16)
    assert rgb.shape == (0, 3) and depth.shape == (0,)
    with pytest.raises(RuntimeError):                      # CPU tensors are rejected, never silently handled
        backend.render_rays(torch.zeros(4, 8), vol_ref.to(DEV), d.imgs_raw, d.pose_source, fn, sc.near_far, 8.0)
    with pytest.raises(RuntimeError):                      # illegal volume dims (F9)
        backend.MVSNet().to(DEV).train().cost_reg_2(torch.zeros(1, 41, 128, 20, 24, device=DEV))


def test_rays_outside_the_volume(mid_scene, nets, weights):
    """Zero padding of the volume lookup, border padding + masks of the colour gather."""
    sc, vol_ref = mid_scene
    fn, _ = nets
    rays = synthetic.scene_rays(sc)[::37].clone()
    rays[:, 3] += 0.35           # swing the directions so many samples leave every frustum
    rays[:, 6] = 0.5             # start in front of the near plane (ndc z < 0)
    rays[:, 7] = 8.0
    rgb_ref, depth_ref = orc.render_rays(rays, vol_ref, sc.imgs_raw, sc.pose_source, weights, sc.H, sc.W,
                                         sc.near_far, float(sc.pad), n_samples=64)
    d = sc.to(DEV)
    rgb, depth = backend.render_rays(rays.to(DEV), vol_ref.to(DEV), d.imgs_raw, d.pose_source, fn, sc.near_far,
                                     float(sc.pad), N_samples=64)
    assert (rgb.cpu() - rgb_ref).abs().max() < RGB_TOL
    assert (depth.cpu() - depth_ref).abs().max() < 5 * DEPTH_TOL


# ------------------------------------------------------------------------------------------------
# BASELINE config 2 size (512 x 640, pad 24, 128 samples): size-independent properties + a sampled
# oracle comparison
# ------------------------------------------------------------------------------------------------
def test_full_size_properties(nets, weights):
    fn, mvs = nets
    sc = synthetic.make_scene(512, 640, pad=24, seed=0)
    d = sc.to(DEV)
    vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
    assert vol.shape == (1, 8, 128, 176, 208)
    assert torch.isfinite(vol).all()
    rays = synthetic.scene_rays(sc).to(DEV)
    N = rays.shape[0]
    assert N == 512 * 640
    rgb, depth = backend.render_rays(rays, vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad))
    assert torch.isfinite(rgb).all() and rgb.min() >= 0 and rgb.max() <= 1 + 1e-5
    assert depth.min() >= 0 and depth.max() <= sc.near_far[1] + 1e-3
    # ray independence: any partition of the rays gives bit-identical pixels (what makes 8-GPU sharding exact)
    cut = 100003
    a, da = backend.render_rays(rays[:cut], vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad))
    b, db = backend.render_rays(rays[cut:], vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad))
    assert torch.equal(torch.cat([a, b]), rgb) and torch.equal(torch.cat([da, db]), depth)
    # permutation equivariance
    perm = torch.randperm(N, device=DEV, generator=torch.Generator(DEV).manual_seed(1))[:50000]
    p, _ = backend.render_rays(rays[perm].contiguous(), vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad))
    assert torch.equal(p, rgb[perm])
    # sampled oracle check on the GPU-built volume
    idx = torch.arange(0, N, 641)[:384]
    rgb_ref, depth_ref = orc.render_rays(rays[idx].cpu(), vol.cpu().contiguous(), sc.imgs_raw, sc.pose_source, weights,
                                         sc.H, sc.W, sc.near_far, float(sc.pad), n_samples=128)
    assert (rgb[idx].cpu() - rgb_ref).abs().max() < RGB_TOL
    assert (depth[idx].cpu() - depth_ref).abs().max() < DEPTH_TOL
    # PSNR of the kernel image vs the oracle image on the sampled rays (gate: within 0.05 dB => mse tiny)
    mse = ((rgb[idx].cpu() - rgb_ref) ** 2).mean()
    assert mse < 1e-9


# ------------------------------------------------------------------------------------------------
# tensor-core mode (tcgen05, fp16 operands / fp32 accumulate): north-star 16-bit gate 5e-3
# ------------------------------------------------------------------------------------------------
TC_RGB_TOL = 5e-3


@pytest.mark.parametrize("S,nrays", [(128, 2048), (128, 1), (128, 3), (32, 1000), (64, 515), (16, 77), (1, 64), (256, 130),
                                     (24, 333), (200, 50), (128, 20000)])
def test_render_tc_half_vs_oracle(mid_scene, nets, weights, S, nrays):
    from mvsnerf_b200 import lib
    sc, vol_ref = mid_scene
    fn, _ = nets
    rays = synthetic.scene_rays(sc)
    g = torch.Generator().manual_seed(S * 1000 + nrays)
    rays = rays[torch.randperm(rays.shape[0], generator=g)[:nrays]]
    rgb_ref, depth_ref = orc.render_rays(rays, vol_ref, sc.imgs_raw, sc.pose_source, weights, sc.H, sc.W,
                                         sc.near_far, float(sc.pad), n_samples=S)
    d = sc.to(DEV)
    rgb, depth = backend.render_rays(rays.to(DEV), vol_ref.to(DEV), d.imgs_raw, d.pose_source, fn, sc.near_far,
                                     float(sc.pad), N_samples=S, mlp_mode=lib.MLP_TC_HALF)
    torch.cuda.synchronize()
    e = (rgb.cpu() - rgb_ref).abs().max().item()
    assert e < TC_RGB_TOL, e
    assert (depth.cpu() - depth_ref).abs().max() < 2e-2
    # the mode-independent parts (gather, compositing) are exact: compare against the fp32 CUDA kernel too
    rgb32, depth32 = backend.render_rays(rays.to(DEV), vol_ref.to(DEV), d.imgs_raw, d.pose_source, fn, sc.near_far,
                                         float(sc.pad), N_samples=S, mlp_mode=lib.MLP_FP32)
    assert (rgb - rgb32).abs().max() < TC_RGB_TOL


def test_render_tc_half_signature_path_and_aux(golden_tiny, nets):
    from mvsnerf_b200 import lib
    g = golden_tiny
    fn, _ = nets
    rays = g["rays"]
    pts, z = orc.march_rays(rays, 32)
    rgb, feat, wts, depth, alpha, _ = backend.rendering(
        Args, pose_of(g), pts.to(DEV), g["ndc"].to(DEV), z.to(DEV), rays[:, :3].to(DEV), rays[:, 3:6].to(DEV),
        g["volume"].to(DEV), g["imgs_raw"].to(DEV), network_fn=fn, mlp_mode=lib.MLP_TC_HALF)
    assert (rgb.cpu() - g["rgb"]).abs().max() < TC_RGB_TOL
    assert (wts.cpu() - g["weights"]).abs().max() < TC_RGB_TOL
    assert (alpha.cpu() - g["alpha"]).abs().max() < TC_RGB_TOL
    assert (feat[:128].cpu() - g["feat_first128"]).abs().max() < 1e-4      # gather is fp32 in every mode


def test_full_frame_tc_half(nets, weights):
    from mvsnerf_b200 import lib
    fn, mvs = nets
    sc = synthetic.make_scene(512, 640, pad=24, seed=0)
    d = sc.to(DEV)
    vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
    rays = synthetic.scene_rays(sc).to(DEV)
    rgb, depth = backend.render_rays(rays, vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad),
                                     mlp_mode=lib.MLP_TC_HALF)
    rgb32, depth32 = backend.render_rays(rays, vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad))
    assert torch.isfinite(rgb).all()
    err = (rgb - rgb32).abs()
    assert err.max() < TC_RGB_TOL, err.max()
    mse = (err ** 2).mean().item()
    # PSNR of each image against any ground truth differs by < 0.05 dB when the mutual MSE is this small
    assert mse < 1e-7, mse
    # determinism + ray independence in the tensor-core mode as well
    a, _ = backend.render_rays(rays[:100003], vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad),
                               mlp_mode=lib.MLP_TC_HALF)
    assert torch.equal(a, rgb[:100003])


# ------------------------------------------------------------------------------------------------
# fp32-grade tensor-core mode (tcgen05, 2-term fp16 operand split, 3 MMAs per K-step): fp32 gate 1e-4
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("S,nrays", [(128, 2048), (128, 3), (32, 1000), (24, 333), (200, 50), (128, 20000)])
def test_render_tc_split_vs_oracle(mid_scene, nets, weights, S, nrays):
    from mvsnerf_b200 import lib
    sc, vol_ref = mid_scene
    fn, _ = nets
    rays = synthetic.scene_rays(sc)
    g = torch.Generator().manual_seed(S * 1000 + nrays)
    rays = rays[torch.randperm(rays.shape[0], generator=g)[:nrays]]
    rgb_ref, depth_ref = orc.render_rays(rays, vol_ref, sc.imgs_raw, sc.pose_source, weights, sc.H, sc.W,
                                         sc.near_far, float(sc.pad), n_samples=S)
    d = sc.to(DEV)
    rgb, depth = backend.render_rays(rays.to(DEV), vol_ref.to(DEV), d.imgs_raw, d.pose_source, fn, sc.near_far,
                                     float(sc.pad), N_samples=S, mlp_mode=lib.MLP_TC_SPLIT)
    torch.cuda.synchronize()
    e = (rgb.cpu() - rgb_ref).abs().max().item()
    assert e < RGB_TOL, e
    assert (depth.cpu() - depth_ref).abs().max() < DEPTH_TOL


def test_render_tc_split_vs_golden(golden_tiny, nets):
    from mvsnerf_b200 import lib
    g = golden_tiny
    fn, _ = nets
    rays = g["rays"]
    pts, z = orc.march_rays(rays, 32)
    rgb, feat, wts, depth, alpha, _ = backend.rendering(
        Args, pose_of(g), pts.to(DEV), g["ndc"].to(DEV), z.to(DEV), rays[:, :3].to(DEV), rays[:, 3:6].to(DEV),
        g["volume"].to(DEV), g["imgs_raw"].to(DEV), network_fn=fn, mlp_mode=lib.MLP_TC_SPLIT)
    assert (rgb.cpu() - g["rgb"]).abs().max() < RGB_TOL
    assert (wts.cpu() - g["weights"]).abs().max() < RGB_TOL
    assert (alpha.cpu() - g["alpha"]).abs().max() < RGB_TOL
    assert (depth.cpu() - g["depth"]).abs().max() < DEPTH_TOL


# ------------------------------------------------------------------------------------------------
# training step (config 3, fine-tuning): gradients of `rendering` w.r.t. the MLP and the encoding volume
# ------------------------------------------------------------------------------------------------
def _oracle_loss(weights_t, vol_t, sc, rays, S, target):
    pts, z = orc.march_rays(rays, S)
    ndc = orc.ndc_coords(sc.pose_source["w2cs"][0], sc.pose_source["intrinsics"][0], pts, sc.H, sc.W,
                         sc.near_far[0], sc.near_far[1], float(sc.pad))
    rgb, _, w, depth, _ = orc.render_samples(pts, ndc, z, rays[:, 3:6], vol_t, sc.imgs_raw, sc.pose_source, weights_t)
    return ((rgb - target) ** 2).mean() + 0.05 * depth.mean() + 0.01 * w.sum(-1).mean(), (pts, ndc, z)


@pytest.mark.parametrize("mode", ["fp32", "half"])
def test_rendering_gradients_vs_oracle_autograd(mid_scene, weights, mode):
    """loss.backward() through backend.rendering (CUDA forward) gives the oracle's autograd gradients for all 22
    MLP tensors and for RefVolume.feat_volume (train_mvs_nerf_finetuning_pl.py:140-189)."""
    from mvsnerf_b200 import lib
    sc, vol_ref = mid_scene
    S, n = 32, 384
    rays = synthetic.scene_rays(sc)
    rays = rays[torch.randperm(rays.shape[0], generator=torch.Generator().manual_seed(3))[:n]]
    target = torch.rand(n, 3, generator=torch.Generator().manual_seed(4))
    # oracle side (CPU autograd)
    wt = {k: v.clone().requires_grad_(k.startswith("mlp/")) for k, v in weights.items()}
    vt = vol_ref.clone().requires_grad_(True)
    loss_ref, (pts, ndc, z) = _oracle_loss(wt, vt, sc, rays, S, target)
    loss_ref.backward()
    # our side
    fn = backend.MVSNeRF().to(DEV)
    backend.load_weights_npz(fn, None, os.path.join(GOLDEN, "mvsnerf_v0_weights.npz"))
    volume = backend.RefVolume(vol_ref.clone().to(DEV)).to(DEV)
    d = sc.to(DEV)
    rgb, _, w, depth, _, _ = backend.rendering(
        Args(), d.pose_source, pts.to(DEV), ndc.to(DEV), z.to(DEV), rays[:, :3].to(DEV), rays[:, 3:6].to(DEV),
        volume_feature=volume, imgs=d.imgs_raw, network_fn=fn,
        mlp_mode=lib.MLP_FP32 if mode == "fp32" else lib.MLP_TC_HALF)
    assert rgb.requires_grad and depth.requires_grad
    loss = ((rgb - target.to(DEV)) ** 2).mean() + 0.05 * depth.mean() + 0.01 * w.sum(-1).mean()
    assert abs(loss.item() - loss_ref.item()) < (1e-5 if mode == "fp32" else 2e-3)
    loss.backward()
    # the upstream gradient 2 (rgb - target) / n carries the forward kernel's own rounding: fp32 ~1e-6, half ~1e-3
    rel = 2e-4 if mode == "fp32" else 5e-3
    names = dict(fn.named_parameters())
    for k, p in names.items():
        g_ref = wt["mlp/" + k].grad
        assert p.grad is not None, k
        err = (p.grad.cpu() - g_ref).abs().max().item()
        assert err <= rel * g_ref.abs().max().item() + 1e-7, (k, err, g_ref.abs().max().item())
    gv = volume.feat_volume.grad
    assert gv is not None and gv.shape == vt.grad.shape
    assert (gv.cpu() - vt.grad).abs().max().item() <= rel * vt.grad.abs().max().item() + 1e-9
    # validation renders under no_grad stay graph-free
    with torch.no_grad():
        out = backend.rendering(Args(), d.pose_source, pts.to(DEV), ndc.to(DEV), z.to(DEV), rays[:, :3].to(DEV),
                                rays[:, 3:6].to(DEV), volume_feature=volume, imgs=d.imgs_raw, network_fn=fn)
    assert not out[0].requires_grad


def test_finetune_steps_reduce_loss(mid_scene):
    """Adam steps on (MLP, volume) with the fine-tuning script's batch shape (1024 rays x 128 samples,
    train_mvs_nerf_finetuning_pl.py:140-189): the kernel re-packs the updated parameters every step, the loss
    goes down; the step time is recorded (gpurun_out/finetune_step.json), nothing is asserted on it."""
    import json
    import time
    from conftest import ROOT
    sc, vol_ref = mid_scene
    S, n = 128, 1024
    fn = backend.MVSNeRF().to(DEV)
    backend.load_weights_npz(fn, None, os.path.join(GOLDEN, "mvsnerf_v0_weights.npz"))
    volume = backend.RefVolume(vol_ref.clone().to(DEV)).to(DEV)
    d = sc.to(DEV)
    rays = synthetic.scene_rays(sc)
    rays = rays[torch.randperm(rays.shape[0], generator=torch.Generator().manual_seed(7))[:n]].to(DEV)
    pts, z = orc.march_rays(rays, S)
    ndc = orc.ndc_coords(d.pose_source["w2cs"][0], d.pose_source["intrinsics"][0], pts, sc.H, sc.W,
                         sc.near_far[0], sc.near_far[1], float(sc.pad))
    target = torch.full((n, 3), 0.25, device=DEV)
    opt = torch.optim.Adam(list(fn.parameters()) + list(volume.parameters()), lr=5e-4, betas=(0.9, 0.999))
    losses, t0 = [], None
    for it in range(14):
        if it == 4:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        rgb = backend.rendering(Args(), d.pose_source, pts, ndc, z, rays[:, :3], rays[:, 3:6], volume_feature=volume,
                                imgs=d.imgs_raw, network_fn=fn)[0]
        loss = ((rgb - target) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 10 * 1e3
    assert losses[-1] < 0.9 * losses[0], losses
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "finetune_step.json"), "w") as f:
        json.dump({"rays": n, "samples": S, "ms_per_step": ms, "losses": losses,
                   "note": "forward = fused kernel (fp32 mode), backward = PyTorch recompute, Adam on MLP + volume"}, f)
