"""Bring-up / A-B of the CTA-pair render kernel (csrc/render_tc2.cu, MVSN_MLP_TC_PAIR) against the round-1
tensor-core kernel and the CPU oracle.  NOT a test (tests/test_gpu_parity.py covers the mode); run under a timeout:

    gpurun -- 'timeout 300 python tools/pair_bringup.py'

Every mbarrier wait traps after a few seconds instead of hanging."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvsnerf_b200 import backend, lib, synthetic
from oracle import mvsnerf_oracle as orc            # checker only

dev = torch.device("cuda:0")
W = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mvsnerf_v0_weights.npz")
weights = orc.load_weights_npz(W)
fn, mvs = backend.MVSNeRF().to(dev), backend.MVSNet().to(dev).train()
backend.load_weights_npz(fn, mvs, W)
MODES = (("half_v1", lib.MLP_TC_HALF), ("pair", lib.MLP_TC_PAIR))

sc = synthetic.make_scene(128, 160, pad=8, seed=5)
d = sc.to(dev)
vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
rays = synthetic.scene_rays(sc)
bad = False
for S, n in ((128, 1), (128, 4 * 32), (128, 2048), (32, 1000), (24, 333), (128, 20480)):
    r = rays[torch.randperm(rays.shape[0], generator=torch.Generator().manual_seed(S + n))[:n]].contiguous()
    ref_rgb, ref_depth = orc.render_rays(r, vol.cpu().contiguous(), sc.imgs_raw, sc.pose_source, weights, sc.H, sc.W,
                                         sc.near_far, float(sc.pad), n_samples=S)
    out = {}
    for name, mode in MODES:
        rgb, depth = backend.render_rays(r.to(dev), vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad),
                                         N_samples=S, mlp_mode=mode)
        torch.cuda.synchronize()
        out[name] = (rgb.cpu(), depth.cpu())
        e = (rgb.cpu() - ref_rgb).abs().max().item()
        print(f"S={S:4d} n={n:6d} {name:8s}: rgb Linf vs oracle {e:.3e}  depth {(depth.cpu() - ref_depth).abs().max():.3e}", flush=True)
        bad |= not (e < 5e-3)
    print(f"                     pair vs half_v1: rgb {(out['pair'][0] - out['half_v1'][0]).abs().max():.3e}", flush=True)

# signature-compatible entry (rays_pts / rays_ndc given) incl. the optional per-sample outputs
r = rays[::37][:1500].contiguous()
pts, z = orc.march_rays(r, 64)
ndc = orc.ndc_coords(sc.pose_source["w2cs"][0], sc.pose_source["intrinsics"][0], pts, sc.H, sc.W, sc.near_far[0], sc.near_far[1], pad=float(sc.pad))
if ndc is not None:
    class A: use_color_volume = False
    res = {}
    for name, mode in MODES:
        o = backend.rendering(A, {k: v for k, v in d.pose_source.items()}, pts.to(dev), ndc.to(dev), z.to(dev), r[:, :3].to(dev),
                              r[:, 3:6].to(dev), vol, d.imgs_raw, network_fn=fn, mlp_mode=mode)
        torch.cuda.synchronize()
        res[name] = [t.cpu() for t in o[:5]]
    for i, nm in enumerate(("rgb", "feat", "weights", "depth", "alpha")):
        print(f"rendering() {nm}: pair vs half_v1 {(res['pair'][i] - res['half_v1'][i]).abs().max():.3e}", flush=True)

sc = synthetic.make_scene(512, 640, pad=24, seed=0)
d = sc.to(dev)
vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
rays = synthetic.scene_rays(sc).to(dev)
frames = {}
for name, mode in MODES + (("split", lib.MLP_TC_SPLIT),):
    for _ in range(3):
        rgb, _ = backend.render_rays(rays, vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad), mlp_mode=mode)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        backend.render_rays(rays, vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad), mlp_mode=mode)
    b.record(); torch.cuda.synchronize(); ms = a.elapsed_time(b) / 5
    frames[name] = rgb
    print(f"512x640 frame, {name:8s}: {ms:.2f} ms = {rays.shape[0] / ms / 1e3:.2f} M rays/s", flush=True)
for name in ("half_v1", "pair"):
    print(f"512x640 {name} vs split: rgb Linf {(frames[name] - frames['split']).abs().max():.3e}")
print("BRINGUP", "FAIL" if bad else "OK")
sys.exit(1 if bad else 0)
