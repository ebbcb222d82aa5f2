"""Per-role cycle totals of CTA 0 of conv0_tc_kernel (trace build):
    python -m mvsnerf_b200.build --trace
    MVSN_LIB=mvsnerf_b200/libmvsnerf_b200_trace.so python tools/conv0_profile.py
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvsnerf_b200 import backend, synthetic, lib
dev = torch.device("cuda:0")
mvs = backend.MVSNet().to(dev).train()
backend.load_weights_npz(None, mvs, "tests/golden/mvsnerf_v0_weights.npz")
sc = synthetic.make_scene(512, 640, pad=24, seed=0); d = sc.to(dev)
L = lib.load()
with torch.no_grad():
    feats = mvs.feature(d.imgs_norm.reshape(3, 3, 512, 640)).view(1, 3, 32, 128, 160)
    dv = torch.linspace(sc.near_far[0], sc.near_far[1], 128, device=dev)[None]
    cost, _ = mvs.build_volume_costvar_img(d.imgs_norm, feats, d.proj_mats, dv, pad=24)
    mvs.cost_reg_2(cost)
    buf = torch.zeros(8 * 1024, dtype=torch.int64, device=dev)
    L.mvsn_debug_set_trace(lib.ptr(buf))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); mvs.cost_reg_2(cost); b.record(); torch.cuda.synchronize()
    L.mvsn_debug_set_trace(None)
v = buf.cpu().view(-1, 8)[:4].tolist()
tiles = 25 * 24
print(f"costreg {a.elapsed_time(b):.3f} ms; CTA 0 handled ~{tiles} tiles; cycles per tile:")
names = {0: ("epilogue ch 0-1 (warp 0)", ["wait acc_ready", "process tile", "brick write-out"]),
         1: ("epilogue ch 2-3 (warp 4)", ["wait acc_ready", "process tile", "brick write-out"]),
         2: ("producer (warp 16)", ["wait a_free", "split+store", "wait stage_full (TMA)"]),
         3: ("issuer", ["wait a_full", "wait acc_free", "issue"])}
for r, (n, slots) in names.items():
    print(" ", n, {s: round(v[r][i] / tiles) for i, s in enumerate(slots)})
