"""CPU, build container only: the oracle restatement next to the LIVE unmodified reference
(imported through oracle/ref_shim.py) on a fresh seeded scene.  Skipped where /root/reference
is absent (the GPU box) -- the committed golden vectors cover that case."""
import pytest
import torch

from oracle import mvsnerf_oracle as orc
from oracle import ref_shim
from mvsnerf_b200 import synthetic

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def live():
    return ref_shim.build_reference(N_samples=24)


def test_weights_fixture_matches_checkpoint(live, weights):
    sd = live.render_kwargs["network_fn"].state_dict()
    for k, v in sd.items():
        assert torch.equal(v.cpu(), weights["mlp/" + k]), k
    n = sum(1 for k in weights if k.startswith("mvs/"))
    assert n == len(live.mvsnet.state_dict()) == 110


def test_live_reference_vs_oracle(live, weights):
    sc = synthetic.make_scene(64, 96, pad=4, seed=11)     # h=16,w=24 -> 24x32 padded: legal
    ref = live.ref
    with torch.no_grad():
        vol_ref, _, _ = live.mvsnet(sc.imgs_norm, sc.proj_mats, sc.near_far, pad=sc.pad)
    vol = orc.encode_volume(sc.imgs_norm, sc.proj_mats, sc.near_far, sc.pad, weights)
    assert (vol - vol_ref).abs().max() < 2e-4
    rays = synthetic.scene_rays(sc)[::5]
    with torch.no_grad():
        xyz, ro, rd, z = ref.ray_utils.ray_marcher(rays, N_samples=24)
        ndc = ref.utils.get_ndc_coordinate(sc.pose_source["w2cs"][0], sc.pose_source["intrinsics"][0].clone(),
                                           xyz, torch.tensor([sc.W - 1, sc.H - 1]), near=sc.near_far[0],
                                           far=sc.near_far[1], pad=sc.pad * 1.0)
        rgb_ref, feat_ref, w_ref, depth_ref, alpha_ref, _ = ref.renderer.rendering(
            live.args, sc.pose_source, xyz, ndc, z, ro, rd, vol_ref, sc.imgs_raw, **live.render_kwargs)
    rgb, depth = orc.render_rays(rays, vol_ref, sc.imgs_raw, sc.pose_source, weights, sc.H, sc.W,
                                 sc.near_far, float(sc.pad), n_samples=24)
    assert (rgb - rgb_ref).abs().max() < 2e-6
    assert (depth - depth_ref).abs().max() < 1e-5


def test_live_reference_eval_mode_vs_oracle(live, weights):
    """MVSNet.eval() (running-statistics BatchNorm, models.py:661-685 through the InPlaceABN stub) vs the oracle's eval_mode."""
    sc = synthetic.make_scene(64, 96, pad=4, seed=12)
    live.mvsnet.eval()
    try:
        with torch.no_grad():
            vol_ref, _, _ = live.mvsnet(sc.imgs_norm, sc.proj_mats, sc.near_far, pad=sc.pad)
    finally:
        live.mvsnet.train()
    # the live module's running statistics, not the checkpoint's: every train-mode forward of the reference (the other
    # tests of this module ran some) updates them in place -- the side effect MVSN_BN_BATCH_UPDATE reproduces
    w = dict(weights)
    w.update({"mvs/" + k: v.detach().clone() for k, v in live.mvsnet.state_dict().items()})
    vol = orc.encode_volume(sc.imgs_norm, sc.proj_mats, sc.near_far, sc.pad, w, eval_mode=True)
    assert (vol - vol_ref).abs().max() < 2e-4 * max(1.0, float(vol_ref.abs().max()))
    vol_train = orc.encode_volume(sc.imgs_norm, sc.proj_mats, sc.near_far, sc.pad, weights)
    assert (vol - vol_train).abs().max() > 0.1          # the two modes differ grossly with this checkpoint


def test_host_mirrors_of_ray_marcher_and_ndc_vs_live_reference(live):
    """mvsnerf_b200.backend.ray_marcher / get_ndc_coordinate (what the fine-tuning step calls before `rendering`) against
    the reference's own data/ray_utils.ray_marcher and utils.get_ndc_coordinate, bit for bit."""
    from mvsnerf_b200 import backend
    sc = synthetic.make_scene(64, 96, pad=4, seed=13)
    rays = synthetic.scene_rays(sc)[::9].contiguous()
    ref = live.ref
    for lindisp in (False, True):
        xyz_r, ro_r, rd_r, z_r = ref.ray_utils.ray_marcher(rays, N_samples=20, lindisp=lindisp)
        xyz, ro, rd, z = backend.ray_marcher(rays, N_samples=20, lindisp=lindisp)
        assert torch.equal(xyz, xyz_r) and torch.equal(z, z_r) and torch.equal(rd, rd_r)
        inv = torch.tensor([sc.W - 1, sc.H - 1])
        a = ref.utils.get_ndc_coordinate(sc.pose_source["w2cs"][0], sc.pose_source["intrinsics"][0].clone(), xyz_r, inv,
                                         near=sc.near_far[0], far=sc.near_far[1], pad=4.0, lindisp=lindisp)
        b = backend.get_ndc_coordinate(sc.pose_source["w2cs"][0], sc.pose_source["intrinsics"][0].clone(), xyz, inv,
                                       near=sc.near_far[0], far=sc.near_far[1], pad=4.0, lindisp=lindisp)
        assert torch.equal(a, b)
