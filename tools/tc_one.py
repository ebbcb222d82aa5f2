import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvsnerf_b200 import backend, synthetic, lib
dev = torch.device("cuda:0")
fn, mvs = backend.MVSNeRF().to(dev), backend.MVSNet().to(dev).train()
backend.load_weights_npz(fn, mvs, "tests/golden/mvsnerf_v0_weights.npz")
sc = synthetic.make_scene(64, 64, pad=8, seed=0); d = sc.to(dev)
with torch.no_grad():
    vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=8)
rays = synthetic.scene_rays(sc).to(dev)
for n in (32, 64, 4096):
    try:
        rgb, depth = backend.render_rays(rays[:n].contiguous(), vol, d.imgs_raw, d.pose_source, fn, sc.near_far, 8.0, N_samples=128, mlp_mode=lib.MLP_TC_HALF)
        torch.cuda.synchronize()
        r32, _ = backend.render_rays(rays[:n].contiguous(), vol, d.imgs_raw, d.pose_source, fn, sc.near_far, 8.0, N_samples=128)
        print(n, "ok", (rgb - r32).abs().max().item())
    except Exception as e:
        print(n, "ERR", repr(e)[:300]); break
