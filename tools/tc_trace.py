"""Dump the pipeline timeline of CTA 0 of a tensor-core render kernel (debug aid; needs the trace build):

    python -m mvsnerf_b200.build --trace
    MVSN_LIB=mvsnerf_b200/libmvsnerf_b200_trace.so python tools/tc_trace.py [half|pair] > profiles/..._timeline.txt

pair (render_tc2.cu) roles: e0/e1 = epilogue warp 0 of slot 0/1, p0/p1 = producer warp 0 of slot 0/1, mma = issuer.
  epilogue events: 10+op accumulator seen complete, 30+op epilogue written + about to signal, 53 compositing done
  producer events: 1 tile start, 4 gathers done, 5 encoding done, 2 operand tiles handed over
  issuer events:   100+2op+s waiting for the slot, 140+2op+s slot granted, 80+2op+s MMAs issued (before commit), 200 tile end
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvsnerf_b200 import backend, synthetic, lib
which = sys.argv[1] if len(sys.argv) > 1 else "half"
mode = {"half": lib.MLP_TC_HALF, "pair": lib.MLP_TC_PAIR}[which]
dev = torch.device("cuda:0")
fn, mvs = backend.MVSNeRF().to(dev), backend.MVSNet().to(dev).train()
backend.load_weights_npz(fn, mvs, "tests/golden/mvsnerf_v0_weights.npz")
sc = synthetic.make_scene(512, 640, pad=24, seed=0); d = sc.to(dev)
with torch.no_grad():
    vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=24)
rays = synthetic.scene_rays(sc).to(dev)
L = lib.load()
for _ in range(2):
    backend.render_rays(rays, vol, d.imgs_raw, d.pose_source, fn, sc.near_far, 24.0, mlp_mode=mode)
buf = torch.zeros(8 * 1024, dtype=torch.int64, device=dev)
L.mvsn_debug_set_trace(lib.ptr(buf))
backend.render_rays(rays, vol, d.imgs_raw, d.pose_source, fn, sc.near_far, 24.0, mlp_mode=mode)
torch.cuda.synchronize()
L.mvsn_debug_set_trace(None)
b = buf.cpu().view(8, 1024)
names = {0: "s0p0", 1: "s0p1", 2: "s1p0", 3: "s1p1", 4: "mma", 5: "load"} if which == "half" else \
        {0: "e0", 1: "e1", 2: "p0", 3: "p1", 4: "mma"}
ev = []
for r in names:
    for v in b[r].tolist():
        if v: ev.append((v >> 8, names[r], v & 255))
ev.sort()
marks = [t for t, n, e in ev if n == "mma" and e == 200]
print("tile-pass durations (cycles):", [marks[i + 1] - marks[i] for i in range(min(10, len(marks) - 1))])
if which == "pair":
    # per-op latency breakdown of slot 0 in steady state: grant -> issued -> acc seen -> epilogue done -> next grant
    import collections
    seq = [(t, n, e) for t, n, e in ev if marks[2] <= t <= marks[6]]
    agg = collections.defaultdict(list)
    last = {}
    for t, n, e in seq:
        if n == "mma" and 140 <= e < 160 and (e - 140) % 2 == 0: last["grant", (e - 140) // 2] = t
        if n == "mma" and 80 <= e < 100 and (e - 80) % 2 == 0:
            op = (e - 80) // 2
            if ("grant", op) in last: agg[f"op{op} grant->issued"].append(t - last["grant", op])
            last["issued", op] = t
        if n == "e0" and 10 <= e < 18:
            op = e - 10
            if ("issued", op) in last: agg[f"op{op} issued->acc_seen"].append(t - last["issued", op])
            last["seen", op] = t
        if n == "e0" and 30 <= e < 38:
            op = e - 30
            if ("seen", op) in last: agg[f"op{op} acc_seen->epilogue_done"].append(t - last["seen", op])
            last["done", op] = t
        if n == "mma" and 140 <= e < 160 and (e - 140) % 2 == 0:
            op = (e - 140) // 2
            if op >= 1 and ("done", op - 1) in last: agg[f"op{op - 1} epilogue_done->next_grant"].append(t - last["done", op - 1])
    print("slot 0 latency chain (median cycles over the window):")
    for k in sorted(agg):
        v = sorted(agg[k]); print(f"  {k:34s} {v[len(v) // 2]:6d}  (n={len(v)})")
lo, hi = marks[2], marks[4]
for t, n, e in ev:
    if lo <= t <= hi:
        print(f"{t - lo:7d} {n:5s} {e}")
