"""CPU: the oracle restatement against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  This is what pins oracle/mvsnerf_oracle.py."""
import pytest
import torch

from oracle import mvsnerf_oracle as orc

TOL_VOL = 2e-4      # |volume| <= ~12; fp32 summation-order noise of train-mode BN is ~6e-6
TOL_RGB = 2e-6


def pose_of(g):
    return {"w2cs": g["w2cs"], "c2ws": g["c2ws"], "intrinsics": g["intrinsics"]}


def dims(g):
    H, W, pad = (int(v) for v in g["HW_pad"])
    return H, W, pad


def test_feature_net(golden_tiny, weights):
    g = golden_tiny
    f = orc.feature_net(g["imgs_norm"][0], weights)
    assert f.shape == g["feats"].shape
    assert (f - g["feats"]).abs().max() < 1e-4


def test_depth_planes(golden_tiny, golden_tiny_lindisp):
    for g, lin in ((golden_tiny, False), (golden_tiny_lindisp, True)):
        d = orc.depth_planes(float(g["near_far"][0]), float(g["near_far"][1]), lindisp=lin)
        assert torch.allclose(d, g["depth_values"], rtol=0, atol=1e-6)


def test_cost_volume(golden_tiny, weights):
    g = golden_tiny
    H, W, pad = dims(g)
    cv, masks = orc.cost_volume(g["imgs_norm"][0], g["feats"], g["proj_mats"][0], g["depth_values"], pad)
    idx = g["vox_idx"]
    assert torch.equal(masks.reshape(3, -1)[:, idx], g["in_masks_sub"])
    assert torch.allclose(cv.reshape(41, -1)[:, idx], g["cost_volume_sub"], rtol=1e-5, atol=1e-5)
    assert torch.allclose(cv.double().sum((1, 2, 3)), g["cost_volume_chsum"], rtol=1e-5, atol=1e-2)
    # F5: the never-written border of channels 0:3 is zero
    assert cv[:3, :, :pad].abs().max() == 0 and cv[:3, :, :, :pad].abs().max() == 0


def test_cost_reg_net_and_volume(golden_tiny, weights):
    g = golden_tiny
    H, W, pad = dims(g)
    vol, st = orc.encode_volume(g["imgs_norm"], g["proj_mats"], g["near_far"].tolist(), pad, weights,
                                return_stages=True)
    assert vol.shape == g["volume"].shape
    assert (vol - g["volume"]).abs().max() < TOL_VOL


def test_volume_lindisp(golden_tiny_lindisp, weights):
    g = golden_tiny_lindisp
    H, W, pad = dims(g)
    vol = orc.encode_volume(g["imgs_norm"], g["proj_mats"], g["near_far"].tolist(), pad, weights, lindisp=True)
    assert (vol - g["volume"]).abs().max() < TOL_VOL


def test_ray_march_and_ndc(golden_tiny):
    g = golden_tiny
    H, W, pad = dims(g)
    pts, z = orc.march_rays(g["rays"], 32)
    assert torch.equal(z, g["z"])
    ndc = orc.ndc_coords(g["w2cs"][0], g["intrinsics"][0], pts, H, W, float(g["near_far"][0]),
                         float(g["near_far"][1]), float(pad))
    assert (ndc - g["ndc"]).abs().max() < 1e-6
    pts, z = orc.march_rays(g["rays"][:256], 16, lindisp=True)
    ndc = orc.ndc_coords(g["w2cs"][0], g["intrinsics"][0], pts, H, W, float(g["near_far"][0]),
                         float(g["near_far"][1]), float(pad), lindisp=True)
    assert (ndc - g["ndc_lindisp256"]).abs().max() < 1e-5


def test_render_tiny(golden_tiny, weights):
    g = golden_tiny
    H, W, pad = dims(g)
    rays = g["rays"]
    pts, z = orc.march_rays(rays, 32)
    rgb, feat, wts, depth, alpha = orc.render_samples(pts, g["ndc"], z, rays[:, 3:6], g["volume"],
                                                      g["imgs_raw"], pose_of(g), weights)
    assert (feat[:128] - g["feat_first128"]).abs().max() < 1e-5
    assert (rgb - g["rgb"]).abs().max() < TOL_RGB
    assert (depth - g["depth"]).abs().max() < 1e-5
    assert (wts - g["weights"]).abs().max() < TOL_RGB
    assert (alpha - g["alpha"]).abs().max() < TOL_RGB


def test_render_variants(golden_tiny, weights):
    g = golden_tiny
    H, W, pad = dims(g)
    nf = g["near_far"].tolist()
    rgb, _ = orc.render_rays(g["rays"][:256], g["volume"], g["imgs_raw"], pose_of(g), weights, H, W, nf,
                             float(pad), n_samples=32, white_bkgd=True)
    assert (rgb - g["rgb_white256"]).abs().max() < TOL_RGB
    rgb, depth = orc.render_rays(g["rays"][:256], g["volume"], g["imgs_raw"], pose_of(g), weights, H, W, nf,
                                 float(pad), n_samples=16, lindisp=True)
    assert (rgb - g["rgb_lindisp256"]).abs().max() < TOL_RGB
    assert (depth - g["depth_lindisp256"]).abs().max() < 1e-5


def test_config1_end_to_end(golden_c1, weights):
    """BASELINE config 1: 64x64 crop, 3 views, pad 24, N_samples 32, CPU forward from the v0 weights."""
    g = golden_c1
    H, W, pad = dims(g)
    nf = g["near_far"].tolist()
    vol, st = orc.encode_volume(g["imgs_norm"], g["proj_mats"], nf, pad, weights, return_stages=True)
    idx = g["vox_idx"]
    assert torch.equal(st["in_masks"].reshape(3, -1)[:, idx], g["in_masks_sub"])
    # variance channels reach ~2e2 (E[x^2]-E[x]^2 cancellation): tolerance is relative to that scale
    assert torch.allclose(st["cost_volume"].reshape(41, -1)[:, idx], g["cost_volume_sub"], rtol=1e-5, atol=1e-4)
    assert (vol[0].reshape(8, -1)[:, idx] - g["volume_sub"]).abs().max() < TOL_VOL
    assert torch.allclose(vol[0].double().sum((1, 2, 3)), g["volume_chsum"], rtol=1e-4, atol=0.5)
    rgb, depth = orc.render_rays(g["rays"], vol, g["imgs_raw"], pose_of(g), weights, H, W, nf, float(pad),
                                 n_samples=32)
    assert (rgb - g["rgb"]).abs().max() < 2e-5      # volume noise (<=2e-4) propagates ~0.1x
    assert (depth - g["depth"]).abs().max() < 2e-4
    rgb, depth = orc.render_rays(g["rays"][1024:1280], vol, g["imgs_raw"], pose_of(g), weights, H, W, nf,
                                 float(pad), n_samples=128)
    assert (rgb - g["rgb128"]).abs().max() < 2e-5


def test_bn_modes_against_reference_fixture(golden_bn_modes, weights):
    """Both behaviours of the reference's BatchNorm on one scene (tests/golden/make_golden_bn.py): the running statistics
    a train-mode forward leaves behind (momentum 0.1, unbiased variance; InPlaceABN, models.py:661-685), and the
    eval-mode forward that then uses them."""
    g = golden_bn_modes
    H, W, pad = dims(g)
    nf = g["near_far"].tolist()
    idx = g["vox_idx"]
    rec = {}
    vol = orc.encode_volume(g["imgs_norm"], g["proj_mats"], nf, pad, weights, record=rec)
    assert (vol[0].reshape(8, -1)[:, idx] - g["volume_train_sub"]).abs().max() < TOL_VOL
    assert len(rec) == 18
    after = dict(weights)
    for name, (mean, var_unbiased) in rec.items():
        key = "stats/" + name[len("mvs/"):]
        want_m, want_v = g[key + ".running_mean"], g[key + ".running_var"]
        got_m = 0.9 * weights[name + ".running_mean"] + 0.1 * mean
        got_v = 0.9 * weights[name + ".running_var"] + 0.1 * var_unbiased
        assert (got_m - want_m).abs().max() <= 1e-5 * max(1.0, float(want_m.abs().max())), name
        assert (got_v - want_v).abs().max() <= 1e-5 * max(1.0, float(want_v.abs().max())), name
        assert int(g[key + ".num_batches_tracked"]) == int(weights[name + ".num_batches_tracked"]) + 1
        after[name + ".running_mean"], after[name + ".running_var"] = want_m, want_v
    vol_e = orc.encode_volume(g["imgs_norm"], g["proj_mats"], nf, pad, after, eval_mode=True)
    scale = max(1.0, float(g["volume_eval_sub"].abs().max()))
    assert (vol_e[0].reshape(8, -1)[:, idx] - g["volume_eval_sub"]).abs().max() < 2e-4 * scale
    assert torch.allclose(vol_e[0].double().sum((1, 2, 3)), g["volume_eval_chsum"], rtol=1e-4, atol=0.5)
    f = orc.feature_net(g["imgs_norm"][0], after, eval_mode=True)
    assert (f - g["feats_eval"]).abs().max() <= 1e-5 * float(g["feats_eval"].abs().max())


@pytest.mark.parametrize("tag,white", [("s32", False), ("s128w", True)])
def test_finetuning_gradients_against_reference_fixture(golden_grad, golden_tiny, weights, tag, white):
    """d img2mse / d (MLP parameters, encoding volume) computed by the reference's own autograd through
    renderer.rendering (tests/golden/make_golden_grad.py) vs autograd through the oracle's restatement."""
    g = {k[len(tag) + 1:]: v for k, v in golden_grad.items() if k.startswith(tag + "/")}
    t = golden_tiny
    wt = {k: v.clone().requires_grad_(k.startswith("mlp/")) for k, v in weights.items()}
    vol = t["volume"].clone().requires_grad_(True)
    rgb, _, _, _, _ = orc.render_samples(g["xyz"], g["ndc"], g["z"], g["rays"][:, 3:6], vol, t["imgs_raw"], pose_of(t), wt,
                                         white_bkgd=white)
    loss = ((rgb - g["target"]) ** 2).mean()
    loss.backward()
    assert (rgb.detach() - g["rgb"]).abs().max() < 2e-6
    assert abs(loss.item() - float(g["loss"])) < 1e-6
    n = 0
    for k, v in g.items():
        if k.startswith("grad_mlp/"):
            got = wt["mlp/" + k[len("grad_mlp/"):]].grad
            assert (got - v).abs().max() <= 2e-5 * float(v.abs().max()) + 1e-9, k
            n += 1
    assert n == 22
    ref_v = torch.zeros(vol.numel())
    ref_v[g["grad_volume_idx"]] = g["grad_volume_val"]
    assert (vol.grad.reshape(-1) - ref_v).abs().max() <= 2e-5 * float(ref_v.abs().max())


def test_manual_samplers_match_grid_sample():
    """Pins SURVEY.md App. A1 tap arithmetic (what the CUDA kernels implement) to F.grid_sample."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    src = torch.rand(5, 9, 11, generator=g)
    grid = torch.rand(1, 40, 30, 2, generator=g) * 2.6 - 1.3
    grid[0, 0, 0] = torch.tensor([-1.0, -1.0]); grid[0, 0, 1] = torch.tensor([1.0, 1.0])
    for border in (False, True):
        ref = F.grid_sample(src[None], grid, mode="bilinear", align_corners=True,
                            padding_mode="border" if border else "zeros")[0]
        got = orc.bilinear_manual(src, grid[0, ..., 0], grid[0, ..., 1], border)
        assert (ref - got).abs().max() < 1e-5
    vol = torch.rand(1, 4, 6, 7, 8, generator=g)
    n = torch.rand(50, 9, 3, generator=g) * 1.4 - 0.2
    ref = orc.lookup_volume(vol, n)
    got = orc.trilinear_manual(vol[0], n)
    assert (ref - got).abs().max() < 1e-5
