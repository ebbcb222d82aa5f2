"""Localise the mask disagreement between K-A and the oracle running on cuda:0 at feature height 160 (BASELINE config 5;
r02b: the cost volume's variance channels and in_masks differ on ~4 000 voxels while the warped RGB channels agree).
Compares THREE statements of the plane-sweep grid: oracle on the CPU, oracle on cuda:0, K-A's masks."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import mvsnerf_oracle as orc
from mvsnerf_b200 import backend, synthetic

DEV = "cuda"
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
wpath = os.path.join(ROOT, "tests", "golden", "mvsnerf_v0_weights.npz")
out = {}
for name, (H, W, pad) in {"tall": (640, 640, 24), "c2": (512, 640, 24)}.items():
    mvs = backend.MVSNet().to(DEV).train()
    backend.load_weights_npz(None, mvs, wpath)
    sc = synthetic.make_scene(H, W, pad=pad, seed=11)
    d = sc.to(DEV)
    h, w = H // 4, W // 4
    r = {}
    with torch.no_grad():
        dv_c = orc.depth_planes(float(sc.near_far[0]), float(sc.near_far[1]), False, device="cpu")
        dv_g = dv_c.to(DEV)
        feats = torch.randn(3, 32, h, w, generator=torch.Generator().manual_seed(0))
        _, m_k = mvs.build_volume_costvar_img(d.imgs_norm, feats.to(DEV)[None], d.proj_mats, dv_g[None], pad=pad)
        m_k = m_k[0].cpu()
        for v in (1, 2):
            g_c = orc.plane_sweep_grid(sc.proj_mats[0, v], dv_c, h, w, pad)
            g_g = orc.plane_sweep_grid(d.proj_mats[0, v], dv_g, h, w, pad).cpu()
            mc, mg = orc._inside(g_c), orc._inside(g_g)
            r[f"v{v}_grid_cpu_vs_gpu_linf"] = float((g_c - g_g).abs().max())
            r[f"v{v}_mask_cpu_vs_gpu"] = int((mc != mg).sum())
            r[f"v{v}_mask_kernel_vs_cpu"] = int((m_k[v] != mc).sum())
            r[f"v{v}_mask_kernel_vs_gpu"] = int((m_k[v] != mg).sum())
            bad = torch.nonzero(m_k[v] != mg)
            ex = []
            for idx in bad[:6].tolist():
                dd, y, x = idx
                ex.append({"idx": idx, "g_cpu": [float(t) for t in g_c[dd, y, x]], "g_gpu": [float(t) for t in g_g[dd, y, x]],
                           "m_kernel": float(m_k[v][dd, y, x]), "m_cpu": float(mc[dd, y, x]), "m_gpu": float(mg[dd, y, x])})
            r[f"v{v}_examples"] = ex
            # pieces of the GPU grid: matmul vs an explicit fp32 formula
            P = d.proj_mats[0, v]
            hp, wp = h + 2 * pad, w + 2 * pad
            ys, xs = torch.meshgrid(torch.arange(hp, dtype=torch.float32, device=DEV) - pad,
                                    torch.arange(wp, dtype=torch.float32, device=DEV) - pad, indexing="ij")
            pix = torch.stack([xs, ys, torch.ones_like(xs)], 0).reshape(3, -1)
            rot_mm = P[:, :3] @ pix
            rot_el = P[:, 0:1] * pix[0:1] + P[:, 1:2] * pix[1:2] + P[:, 2:3] * pix[2:3]
            rot_cpu = sc.proj_mats[0, v][:, :3] @ pix.cpu()
            r[f"v{v}_rot_matmul_vs_elementwise_gpu"] = float((rot_mm - rot_el).abs().max())
            r[f"v{v}_rot_matmul_gpu_vs_cpu"] = float((rot_mm.cpu() - rot_cpu).abs().max())
            r[f"v{v}_rot_scale"] = float(rot_cpu.abs().max())
    out[name] = r
    print(name, json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "diag_c5.json"), "w"), indent=1)
