"""Dump the pipeline timeline of CTA 0 of the tensor-core render kernel (debug aid)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvsnerf_b200 import backend, synthetic, lib
dev = torch.device("cuda:0")
fn, mvs = backend.MVSNeRF().to(dev), backend.MVSNet().to(dev).train()
backend.load_weights_npz(fn, mvs, "tests/golden/mvsnerf_v0_weights.npz")
sc = synthetic.make_scene(512, 640, pad=24, seed=0); d = sc.to(dev)
with torch.no_grad():
    vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=24)
rays = synthetic.scene_rays(sc).to(dev)
L = lib.load()
for _ in range(2):
    backend.render_rays(rays, vol, d.imgs_raw, d.pose_source, fn, sc.near_far, 24.0, mlp_mode=lib.MLP_TC_HALF)
buf = torch.zeros(8 * 1024, dtype=torch.int64, device=dev)
L.mvsn_debug_set_trace(lib.ptr(buf))
backend.render_rays(rays, vol, d.imgs_raw, d.pose_source, fn, sc.near_far, 24.0, mlp_mode=lib.MLP_TC_HALF)
torch.cuda.synchronize()
L.mvsn_debug_set_trace(None)
b = buf.cpu().view(8, 1024)
names = {0: "s0p0", 1: "s0p1", 2: "s1p0", 3: "s1p1", 4: "mma", 5: "load"}
ev = []
for r in range(6):
    for v in b[r].tolist():
        if v: ev.append((v >> 8, names[r], v & 255))
ev.sort()
t0 = ev[0][0]
# steady state: show passes 3 and 4
marks = [t for t, n, e in ev if n == "mma" and e == 200]
lo, hi = marks[2], marks[4]
print("pass durations (cycles):", [marks[i + 1] - marks[i] for i in range(min(8, len(marks) - 1))])
for t, n, e in ev:
    if lo <= t <= hi:
        print(f"{t - lo:7d} {n:5s} {e}")
