"""GPU: the seams around the kernels that the parity tests do not cross -- the factory (`create_nerf_mvs`,
models.py:569-654), the host-buffer entry (`HostFrameRenderer`), `return_color`, the scene caches, fine-tuned
checkpoints / frame writer (scene_io) and the sharding API on a real NCCL group."""
import os
import warnings
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import mvsnerf_oracle as orc
from mvsnerf_b200 import backend, distributed, lib, scene_io, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"
WPATH = os.path.join(GOLDEN, "mvsnerf_v0_weights.npz")


def make_args(ckpt):
    """The argument namespace the notebooks build."""
    return SimpleNamespace(multires=10, i_embed=0, pts_dim=3, multires_views=4, dir_dim=3, netdepth=6, netwidth=128,
                           feat_dim=20, net_type="v0", N_importance=0, netchunk=1024, ckpt=ckpt, perturb=1.0,
                           N_samples=128, use_viewdirs=True, white_bkgd=False, raw_noise_std=0.0, img_downscale=1.0,
                           use_color_volume=False)


@pytest.fixture(scope="module")
def ckpt_tar(tmp_path_factory):
    """A checkpoint in the schema of ckpts/mvsnerf-v0.tar rebuilt from the committed weight export."""
    z = np.load(WPATH)
    ck = {"global_step": 181104,
          "network_fn_state_dict": {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("mlp/")},
          "network_mvs_state_dict": {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("mvs/")}}
    p = tmp_path_factory.mktemp("ckpt") / "mvsnerf-v0.tar"
    torch.save(ck, p)
    return str(p)


@pytest.fixture(scope="module")
def scene():
    sc = synthetic.make_scene(64, 96, pad=8, seed=4)
    return sc, sc.to(DEV)


def test_create_nerf_mvs_factory(ckpt_tar, scene):
    sc, d = scene
    args = make_args(ckpt_tar)
    train, test, start, grad_vars = backend.create_nerf_mvs(args, use_mvs=True, dir_embedder=False, pts_embedder=True)
    # models.py:637-648: the dict schema the Lightning modules / notebooks index into
    keys = {"network_query_fn", "perturb", "N_importance", "network_fine", "N_samples", "network_fn", "network_mvs",
            "use_viewdirs", "white_bkgd", "raw_noise_std"}
    assert set(train) == keys and set(test) == keys
    assert test["perturb"] is False and train["perturb"] == 1.0 and start == 0
    fn, mvs = train["network_fn"], train["network_mvs"]
    assert test["network_fn"] is fn
    assert len(fn.state_dict()) == 22 and len(mvs.state_dict()) == 110
    assert all(p.is_cuda for p in fn.parameters())
    # the MLP is trainable, the encoder is forward-only and therefore NOT handed to the optimiser
    assert {id(p) for p in grad_vars} == {id(p) for p in fn.parameters()}
    # loaded strictly from the checkpoint: same numbers as the direct loader
    fn2, mvs2 = backend.MVSNeRF().to(DEV), backend.MVSNet().to(DEV)
    backend.load_weights_npz(fn2, mvs2, WPATH)
    for (k, a), (_, b) in zip(sorted(fn.state_dict().items()), sorted(fn2.state_dict().items())):
        assert torch.equal(a, b), k
    for (k, a), (_, b) in zip(sorted(mvs.state_dict().items()), sorted(mvs2.state_dict().items())):
        assert torch.equal(a, b), k
    # the pair works end to end exactly like the direct modules
    mvs.train()
    with torch.no_grad():
        vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
        vol2, _, _ = mvs2.train()(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
    # same weights, same inputs: bit-identical (BatchNorm statistics are accumulated in fixed point, conv_common.cuh)
    assert torch.equal(vol, vol2)
    # alpha-only query function (renderer.py:42-63) still answers
    x = torch.rand(7, 5, 3, device=DEV)
    out = train["network_query_fn"](x, None, torch.rand(7, 5, 20, device=DEV), fn)
    assert out.shape == (7, 5, 1)
    with pytest.raises(RuntimeError):
        backend.create_nerf_mvs(args, use_mvs=True, dir_embedder=True, pts_embedder=True)


def test_mvsnet_warns_under_autograd(scene):
    sc, d = scene
    mvs = backend.MVSNet().to(DEV).train()
    backend.load_weights_npz(None, mvs, WPATH)
    with pytest.warns(backend.FrozenEncoderWarning):
        vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
    assert not vol.requires_grad
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        with torch.no_grad():
            mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)


def test_return_color(scene, weights):
    """MVSNet.forward(return_color=True) (models.py:924-926): the second output becomes [B,V,4,D,h',w'] =
    warped source colours + in-bounds masks."""
    sc, d = scene
    mvs = backend.MVSNet().to(DEV).train()
    backend.load_weights_npz(None, mvs, WPATH)
    with torch.no_grad():
        vol, feats, depth_values = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad, return_color=True)
    hp, wp = sc.H // 4 + 2 * sc.pad, sc.W // 4 + 2 * sc.pad
    assert feats.shape == (1, 3, 4, 128, hp, wp) and depth_values.shape == (1, 128)
    f = orc.feature_net(sc.imgs_norm[0], weights)
    cost, masks = orc.cost_volume(sc.imgs_norm[0], f, sc.proj_mats[0], orc.depth_planes(*sc.near_far), sc.pad)
    ref = torch.cat((cost[:9].view(1, 3, 3, *cost.shape[1:]), masks[None].unsqueeze(2)), dim=2)
    assert torch.equal(feats[:, :, 3].cpu(), ref[:, :, 3])                      # masks exactly
    assert (feats[:, :, :3].cpu() - ref[:, :, :3]).abs().max() < 1e-4


def test_host_frame_renderer_matches_device_call(scene):
    sc, d = scene
    fn, mvs = backend.MVSNeRF().to(DEV), backend.MVSNet().to(DEV).train()
    backend.load_weights_npz(fn, mvs, WPATH)
    with torch.no_grad():
        vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
        rays = synthetic.scene_rays(sc)
        hfr = backend.HostFrameRenderer(rays.shape[0], DEV)
        assert hfr.h2d_bytes == rays.numel() * 4 and hfr.d2h_bytes == rays.shape[0] * 16
        for mode in (lib.MLP_TC_SPLIT, lib.MLP_TC_PAIR):
            rgb_h, depth_h = hfr.render(rays.pin_memory(), vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad),
                                        N_samples=48, mlp_mode=mode)
            rgb, depth = backend.render_rays(rays.to(DEV), vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad),
                                             N_samples=48, mlp_mode=mode)
            assert not rgb_h.is_cuda and rgb_h.is_pinned()
            assert torch.equal(rgb_h, rgb.cpu()) and torch.equal(depth_h, depth.cpu())
        with pytest.raises(RuntimeError):
            hfr.render(rays.to(DEV), vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad))


def test_volume_cache_hits_for_planar_volumes(scene):
    """A planar [1,8,D,H,W] volume (RefVolume / checkpoint tensor) is transposed to channels-last ONCE."""
    sc, d = scene
    fn, mvs = backend.MVSNeRF().to(DEV), backend.MVSNet().to(DEV).train()
    backend.load_weights_npz(fn, mvs, WPATH)
    with torch.no_grad():
        vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
        ref_vol = backend.RefVolume(vol.contiguous().clone()).to(DEV)               # planar memory, as a checkpoint holds it
        rays = synthetic.scene_rays(sc)[::5].contiguous().to(DEV)
        backend.clear_cache()
        h0, m0 = backend.cache_stats["hit"], backend.cache_stats["miss"]
        outs = [backend.render_rays(rays, ref_vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad), N_samples=32)
                for _ in range(3)]
        # volume + images: one miss each, then hits
        assert backend.cache_stats["miss"] - m0 == 2 and backend.cache_stats["hit"] - h0 == 4
        direct = backend.render_rays(rays, vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad), N_samples=32)
        assert torch.equal(outs[0][0], direct[0]) and torch.equal(outs[2][0], direct[0])
        ref_vol.feat_volume.mul_(0.5)                                             # in-place update, as an optimiser step does it (under no_grad)
        changed = backend.render_rays(rays, ref_vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad), N_samples=32)
        assert not torch.equal(changed[0], direct[0])


def test_finetuned_ckpt_roundtrip_and_video_loop(scene, tmp_path):
    sc, d = scene
    fn, mvs = backend.MVSNeRF().to(DEV), backend.MVSNet().to(DEV).train()
    backend.load_weights_npz(fn, mvs, WPATH)
    with torch.no_grad():
        vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
    volume = backend.RefVolume(vol.detach())
    path = scene_io.save_finetuned_ckpt(str(tmp_path / "ft" / "latest.tar"), fn, volume, network_mvs=mvs, global_step=77)
    raw = torch.load(path, map_location="cpu", weights_only=False)
    # the reference notebook's own read: torch.load(args.ckpt)['volume']['feat_volume'] (renderer_video.ipynb)
    assert set(raw) == {"global_step", "network_fn_state_dict", "volume", "network_mvs_state_dict"}
    assert raw["volume"]["feat_volume"].shape == vol.shape and raw["volume"]["feat_volume"].is_contiguous()
    fn2 = backend.MVSNeRF().to(DEV)
    vol2, step = scene_io.load_finetuned_ckpt(path, network_fn=fn2, device=DEV)
    assert step == 77 and isinstance(vol2, backend.RefVolume) and torch.equal(vol2.feat_volume, vol)
    # free-viewpoint loop from the loaded checkpoint == direct rendering from the live volume
    c2ws = synthetic.spiral_path(sc, 3).to(DEV)
    with torch.no_grad():
        frames = scene_io.render_video(c2ws, d.directions, vol2, d.imgs_raw, d.pose_source, fn2, sc.near_far, float(sc.pad),
                                       N_samples=32)
        for (rgb, depth), c2w in zip(frames, c2ws):
            rays = synthetic.camera_rays(d.directions, c2w, *sc.near_far)
            assert torch.equal(scene_io.camera_rays(d.directions, c2w, *sc.near_far), rays)
            r2, d2 = backend.render_rays(rays, vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad), N_samples=32)
            assert torch.equal(rgb, r2) and torch.equal(depth, d2)
        with scene_io.FrameWriter(str(tmp_path / "frames"), sc.H, sc.W, fmt="ppm") as wr:
            n = scene_io.render_video(c2ws, d.directions, vol2, d.imgs_raw, d.pose_source, fn2, sc.near_far, float(sc.pad),
                                      writer=wr, N_samples=32)
    assert n == 3 and wr.frames_written == 3
    for i, (rgb, depth) in enumerate(frames):
        blob = open(tmp_path / "frames" / f"rgb_{i:05d}.ppm", "rb").read()
        header = f"P6\n{sc.W} {sc.H}\n255\n".encode()
        assert blob.startswith(header)
        img = np.frombuffer(blob[len(header):], dtype=np.uint8).reshape(sc.H, sc.W, 3)
        want = (rgb.view(sc.H, sc.W, 3).clamp(0, 1) * 255).round().to(torch.uint8).cpu().numpy()
        assert np.array_equal(img, want)
        dsaved = np.load(tmp_path / "frames" / f"depth_{i:05d}.npy")
        assert np.array_equal(dsaved, depth.view(sc.H, sc.W).cpu().numpy())


def test_sharded_api_on_a_one_rank_nccl_group(scene):
    """distributed.render_rays_sharded on a real NCCL process group (world 1 here; bench.py runs the same call at
    2/4/8 ranks): the assembled frame equals the direct call bit for bit."""
    import torch.distributed as dist
    sc, d = scene
    fn, mvs = backend.MVSNeRF().to(DEV), backend.MVSNet().to(DEV).train()
    backend.load_weights_npz(fn, mvs, WPATH)
    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29577", rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
        created = True
    try:
        with torch.no_grad():
            vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
            rays = synthetic.scene_rays(sc).to(DEV)
            render = lambda r: backend.render_rays(r, vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad), N_samples=32)
            rgb, depth = distributed.render_rays_sharded(rays, render)
            r2, d2 = render(rays)
        assert torch.equal(rgb, r2) and torch.equal(depth, d2)
    finally:
        if created:
            dist.destroy_process_group()


@pytest.mark.parametrize("mode", [lib.MLP_FP32, lib.MLP_TC_HALF, lib.MLP_TC_SPLIT, lib.MLP_TC_PAIR])
def test_peer_sink_stores_equal_the_plain_outputs(scene, mode):
    """mvsn_render_rays_to_peers: the (r,g,b,depth) texels the compositing epilogue stores into the frame copies
    (two 'peers' here, both on this GPU, at a non-zero first_pixel) are bit-identical to the rgb / depth arrays,
    pixels outside the band stay untouched, and the sink-only form (rgb = depth = NULL) writes the same frame."""
    sc, d = scene
    fn, mvs = backend.MVSNeRF().to(DEV), backend.MVSNet().to(DEV).train()
    backend.load_weights_npz(fn, mvs, WPATH)
    with torch.no_grad():
        vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
        rays = synthetic.scene_rays(sc)[:1501].contiguous().to(DEV)
        n, first = rays.shape[0], 37
        frames = [torch.full((n + 100, 4), -7.0, device=DEV) for _ in range(2)]
        sink = lib.PeerSink()
        sink.frame[0], sink.frame[1], sink.n_peers, sink.first_pixel = frames[0].data_ptr(), frames[1].data_ptr(), 2, first
        rgb, depth = backend.render_rays(rays, vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad), N_samples=32,
                                         mlp_mode=mode, out=(torch.empty(n, 3, device=DEV), torch.empty(n, device=DEV)), sink=sink)
        plain = backend.render_rays(rays, vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad), N_samples=32, mlp_mode=mode)
        assert torch.equal(rgb, plain[0]) and torch.equal(depth, plain[1])
        for f in frames:
            assert torch.equal(f[first:first + n, :3], rgb) and torch.equal(f[first:first + n, 3], depth)
            assert bool((f[:first] == -7.0).all()) and bool((f[first + n:] == -7.0).all())
        only = torch.full((n + 100, 4), -7.0, device=DEV)
        s1 = lib.PeerSink()
        s1.frame[0], s1.n_peers, s1.first_pixel = only.data_ptr(), 1, first
        r = backend.render_rays(rays, vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad), N_samples=32, mlp_mode=mode, sink=s1)
        assert r == (None, None) and torch.equal(only, frames[0])


def test_peer_frame_single_process(scene):
    """distributed.PeerFrame without a process group: exportable buffer, tensor view over it, render through the sink."""
    sc, d = scene
    fn, mvs = backend.MVSNeRF().to(DEV), backend.MVSNet().to(DEV).train()
    backend.load_weights_npz(fn, mvs, WPATH)
    with torch.no_grad():
        vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
        rays = synthetic.scene_rays(sc).to(DEV)
        frame = distributed.PeerFrame(rays.shape[0], n_buffers=2)
        render = lambda r, sink=None: backend.render_rays(r, vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad),
                                                          N_samples=32, sink=sink)
        for _ in range(3):
            rgb, depth = distributed.render_rays_sharded(rays, render, frame=frame)
            r2, d2 = render(rays)
            assert torch.equal(rgb, r2) and torch.equal(depth, d2)
            frame.rotate()
        frame.close()


def test_two_gpu_sharding_nccl_and_peer_stores():
    """tools/multi_gpu_check.py under torchrun on 2 GPUs: NCCL all-gather and NVLink peer-store assembly both
    bit-equal to the single-GPU frame in every MLP mode.  Skipped on single-GPU boxes (the driver's GPU test tier);
    profiles/r02_multi_gpu_check_2gpu.json is the committed record of a 2-GPU run."""
    import subprocess
    import sys
    from conftest import ROOT
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])


def test_get_rays_on_device_and_camera_level_host_renderer(scene):
    """mvsn_make_rays (data/ray_utils.get_rays on the device) against the host construction of the same rays, and
    HostFrameRenderer.render_camera (host input = the camera pose) against the rays-level call."""
    sc, d = scene
    fn, mvs = backend.MVSNeRF().to(DEV), backend.MVSNet().to(DEV).train()
    backend.load_weights_npz(fn, mvs, WPATH)
    c2ws = synthetic.spiral_path(sc, 3)
    with torch.no_grad():
        vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
        hfr = backend.HostFrameRenderer(sc.H * sc.W, DEV)
        for c2w in c2ws:
            want = synthetic.camera_rays(sc.directions, c2w, *sc.near_far)                       # host, torch matmul
            for m in (c2w.to(DEV), c2w[:3].contiguous().to(DEV)):                                # [4,4] and [3,4]
                got = backend.get_rays(d.directions, m, *sc.near_far)
                assert got.shape == want.shape
                assert torch.equal(got[:, :3].cpu(), want[:, :3]) and torch.equal(got[:, 6:].cpu(), want[:, 6:])
                assert (got[:, 3:6].cpu() - want[:, 3:6]).abs().max().item() <= 2e-7 * want[:, 3:6].abs().max().item()
            rgb_h, depth_h = hfr.render_camera(c2w.pin_memory(), d.directions, vol, d.imgs_raw, d.pose_source, fn, sc.near_far,
                                               float(sc.pad), N_samples=32)
            rgb, depth = backend.render_rays(backend.get_rays(d.directions, c2w.to(DEV), *sc.near_far), vol, d.imgs_raw,
                                             d.pose_source, fn, sc.near_far, float(sc.pad), N_samples=32)
            assert torch.equal(rgb_h, rgb.cpu()) and torch.equal(depth_h, depth.cpu())
            ref_rgb, _ = backend.render_rays(want.to(DEV), vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad), N_samples=32)
            assert (rgb - ref_rgb).abs().max().item() < 1e-5
