"""Round-2 bring-up of the CTA-pair render kernel (csrc/wip/render_tc_pair.cu, mlp_mode 3).  NOT a test.

    python -m mvsnerf_b200.build --wip                       # here (nvcc cross-compiles)
    gpurun -- 'MVSN_LIB=mvsnerf_b200/libmvsnerf_b200_wip.so timeout 120 python tools/pair_bringup.py'

Renders a small scene with the validated modes and with mode 3, prints RGB/depth L-inf against the CPU oracle and
against MVSN_MLP_TC_HALF, then times a 512x640 frame.  Every mbarrier wait traps after ~4 s instead of hanging."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvsnerf_b200 import backend, lib, synthetic
from oracle import mvsnerf_oracle as orc            # checker only

assert "wip" in os.environ.get("MVSN_LIB", ""), "point MVSN_LIB at libmvsnerf_b200_wip.so (python -m mvsnerf_b200.build --wip)"
dev = torch.device("cuda:0")
W = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mvsnerf_v0_weights.npz")
weights = orc.load_weights_npz(W)
fn, mvs = backend.MVSNeRF().to(dev), backend.MVSNet().to(dev).train()
backend.load_weights_npz(fn, mvs, W)

sc = synthetic.make_scene(128, 160, pad=8, seed=5)
d = sc.to(dev)
vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
rays = synthetic.scene_rays(sc)
for S, n in ((128, 2048), (32, 1000), (24, 333), (128, 6 * 32), (128, 1)):
    r = rays[torch.randperm(rays.shape[0], generator=torch.Generator().manual_seed(S + n))[:n]].contiguous()
    ref_rgb, ref_depth = orc.render_rays(r, vol.cpu().contiguous(), sc.imgs_raw, sc.pose_source, weights, sc.H, sc.W,
                                         sc.near_far, float(sc.pad), n_samples=S)
    out = {}
    for name, mode in (("half", lib.MLP_TC_HALF), ("pair", lib.MLP_TC_PAIR_WIP)):
        rgb, depth = backend.render_rays(r.to(dev), vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad),
                                         N_samples=S, mlp_mode=mode)
        torch.cuda.synchronize()
        out[name] = (rgb.cpu(), depth.cpu())
        print(f"S={S:4d} n={n:5d} {name}: rgb Linf vs oracle {(rgb.cpu() - ref_rgb).abs().max():.3e}  "
              f"depth {(depth.cpu() - ref_depth).abs().max():.3e}")
    print(f"             pair vs half: rgb {(out['pair'][0] - out['half'][0]).abs().max():.3e}")

sc = synthetic.make_scene(512, 640, pad=24, seed=0)
d = sc.to(dev)
vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=sc.pad)
rays = synthetic.scene_rays(sc).to(dev)
for name, mode in (("half", lib.MLP_TC_HALF), ("pair", lib.MLP_TC_PAIR_WIP)):
    for _ in range(3):
        backend.render_rays(rays, vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad), mlp_mode=mode)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        backend.render_rays(rays, vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(sc.pad), mlp_mode=mode)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 5 * 1e3
    print(f"512x640 frame, {name}: {ms:.2f} ms = {rays.shape[0] / ms / 1e3:.2f} M rays/s")
