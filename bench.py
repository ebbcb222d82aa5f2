#!/usr/bin/env python
"""Benchmark of the MVSNeRF render hot path (BASELINE.json metric: rays/s @ 128 samples, DTU 512x640).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--mode fp32]

A "step" is one pass of the hot path over one batch of synthetic rays: one full 512x640 frame
(327 680 rays x 128 samples) per GPU rendered against a resident encoding volume -- the
reference's own frame loop (renderer_video.ipynb "DTU video rendering": the volume is built once per
scene, frames are then rendered from it).  Every step renders a different target camera of a
seeded spiral path.  The once-per-scene volume build (FeatureNet + cost volume + CostRegNet) is
timed separately and reported under "volume_build".

value      : whole-job rays/s with the rays already resident in HBM (kernel launch only).
e2e        : the same metric through the host-buffer call (pinned host rays -> H2D -> kernel ->
             D2H of rgb+depth), copies inside the timed region.
roofline   : the render kernel's algorithmic FLOPs (32 178 176 per ray) over its
             CUDA-event duration, against the measured bf16 tensor peak (MEASURED_PEAKS.json).
cpu_baseline / --impl reference : the CPU restatement of the reference path (oracle/, PyTorch CPU,
             all host threads) on a bounded sample of the same workload.

Multi-GPU (torchrun, one rank per GPU): rays shard across ranks, the volume is replicated, one
NCCL all-gather of rgb+depth per step sits inside the timed region; weak scaling (one frame per
GPU per step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

H, W, PAD, S = 512, 640, 24, 128
N_RAYS = H * W
FLOP_PER_RAY = 32_178_176          # 128 x 251 392 MLP FLOPs           
BYTES_PER_RAY = 51_248             # 128 x 400 B gather + 32 B in + 16 B out
WEIGHTS = os.path.join(ROOT, "tests", "golden", "mvsnerf_v0_weights.npz")
METRIC = "rays/sec @ 128 samples, DTU 512x640"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU reference arm (the oracle port on the host cores)
# ---------------------------------------------------------------------------------------------
def cpu_reference_setup():
    from oracle import mvsnerf_oracle as orc
    from mvsnerf_b200 import synthetic
    torch.set_num_threads(min(32, os.cpu_count() or 1))   # big hosts: the intra-op pool regresses past ~32 threads
    weights = orc.load_weights_npz(WEIGHTS)
    sc = synthetic.make_scene(H, W, pad=PAD, seed=0)
    return orc, synthetic, weights, sc


def cpu_pick_threads(orc, weights, sc, rays, volume):
    """Give the CPU arm its best configuration: PyTorch's intra-op pool stops scaling (and then
    regresses) well below the core count of a big host, so try a few pool sizes on a small sample
    and keep the fastest.  Returns the thread count used from here on."""
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        cpu_render_sample(orc, weights, sc, rays[:256], volume)
        t = cpu_render_sample(orc, weights, sc, rays[:1024], volume)
        if t < best_t:
            best, best_t = c, t
    torch.set_num_threads(best)
    return best


def cpu_render_sample(orc, weights, sc, rays, volume):
    t0 = time.perf_counter()
    with torch.no_grad():
        orc.render_rays(rays, volume, sc.imgs_raw, sc.pose_source, weights, H, W, sc.near_far, float(PAD),
                        n_samples=S, chunk=5120)
    return time.perf_counter() - t0


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference
    is pure Python/PyTorch and /root/reference does not exist on the GPU box), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    orc, synthetic, weights, sc = cpu_reference_setup()
    cores = os.cpu_count() or 1
    sample = 2048
    with torch.no_grad():
        t0 = time.perf_counter()
        volume = orc.encode_volume(sc.imgs_norm, sc.proj_mats, sc.near_far, PAD, weights)
        t_vol = time.perf_counter() - t0
    path = synthetic.spiral_path(sc, max(args.steps + args.warmup, 2))
    threads = cpu_pick_threads(orc, weights, sc, synthetic.scene_rays(sc)[::37][:2048].contiguous(), volume)
    times = []
    for i in range(args.warmup + args.steps):
        rays = synthetic.scene_rays(sc, path[i % len(path)])
        g = torch.Generator().manual_seed(i)
        rays = rays[torch.randperm(N_RAYS, generator=g)[:sample]].contiguous()
        dt = cpu_render_sample(orc, weights, sc, rays, volume)
        if i >= args.warmup:
            times.append(dt)
    t = sum(times) / len(times)
    value = sample / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "DTU-shaped 512x640 frame, 3 source views, pad 24, N_samples=128, volume resident",
                   "sample": f"{sample} random rays of the frame per step (bounded sample; rays/s is per-ray linear)"},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "host_cpus": cores, "kind": "port",
                         "sample": f"{sample} rays x 128 samples per step, chunk 5120, torch {torch.__version__} CPU, "
                                   f"{torch.get_num_threads()} threads; volume build once: {t_vol:.1f} s"},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "volume_build": {"ms": t_vol * 1e3},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from mvsnerf_b200 import backend, synthetic, lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback; "
                         "use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mode = {"fp32": lib.MLP_FP32, "half": lib.MLP_TC_HALF, "split": lib.MLP_TC_SPLIT}[args.mode]

    fn, mvs = backend.MVSNeRF().to(dev), backend.MVSNet().to(dev).train()
    backend.load_weights_npz(fn, mvs, WEIGHTS)
    sc = synthetic.make_scene(H, W, pad=PAD, seed=0)
    d = sc.to(dev)

    def ev_time(f, reps):
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); f(); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b))
        return ts

    # ---- once-per-scene encoding volume (reported, not part of the step) -----------------------
    with torch.no_grad():
        vol, _, _ = mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=PAD)            # warm-up + result
        torch.cuda.synchronize()
        t_build = min(ev_time(lambda: mvs(d.imgs_norm, d.proj_mats, sc.near_far, pad=PAD), 3))
        t_feat = min(ev_time(lambda: mvs.feature(d.imgs_norm.reshape(3, 3, H, W)), 3))
        feats = mvs.feature(d.imgs_norm.reshape(3, 3, H, W)).view(1, 3, 32, H // 4, W // 4)
        dv = torch.linspace(sc.near_far[0], sc.near_far[1], 128, device=dev)[None]
        t_cost = min(ev_time(lambda: mvs.build_volume_costvar_img(d.imgs_norm, feats, d.proj_mats, dv, pad=PAD), 3))
        cost, _ = mvs.build_volume_costvar_img(d.imgs_norm, feats, d.proj_mats, dv, pad=PAD)
        t_reg = min(ev_time(lambda: mvs.cost_reg_2(cost), 3))
        del cost, feats
    pk = peaks()
    nvox = 128 * (H // 4 + 2 * PAD) * (W // 4 + 2 * PAD)

    # ---- frames: a spiral of target cameras, each rank renders its own frame of every step ------
    n_frames = args.warmup + args.steps
    n_distinct = min(n_frames, 16)                     # distinct target cameras kept resident (cycled for long runs)
    path = synthetic.spiral_path(sc, max(n_distinct * world, 2))
    rays_host = [synthetic.scene_rays(sc, path[(i * world + rank) % len(path)]).pin_memory() for i in range(n_distinct)]
    rays_dev = [r.to(dev) for r in rays_host]
    rgb = torch.empty(N_RAYS, 3, device=dev)
    depth = torch.empty(N_RAYS, device=dev)
    rgb_all = torch.empty(world * N_RAYS, 3, device=dev) if world > 1 else None
    depth_all = torch.empty(world * N_RAYS, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)                  # > 126 MB L2
    launches = [0]

    def step(i):
        backend.render_rays(rays_dev[i % n_distinct], vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(PAD),
                            N_samples=S, mlp_mode=mode, out=(rgb, depth))
        launches[0] += 1
        if world > 1:
            dist.all_gather_into_tensor(rgb_all, rgb)
            dist.all_gather_into_tensor(depth_all, depth)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for i in range(args.warmup):
            step(i)
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        launches[0] = 0
        evs = []
        t_wall0 = time.perf_counter()
        for i in range(args.warmup, n_frames):
            flush.zero_()                                                          # L2 flush, untimed
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(i); b.record()
            evs.append((a, b))
        barrier()
        t_wall = time.perf_counter() - t_wall0
        clocks = sampler.stop() if rank == 0 else None
        n_launch = launches[0]
        step_ms = [a.elapsed_time(b) for a, b in evs]
        total_ms = torch.tensor([sum(step_ms)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        ms_per_step = float(total_ms.item()) / args.steps
        value = world * N_RAYS / (ms_per_step * 1e-3)

        # kernel-only duration (no collective, no flush) for the roofline: CUDA events around the launch
        kern_ms = ev_time(lambda: backend.render_rays(rays_dev[-1], vol, d.imgs_raw, d.pose_source, fn, sc.near_far,
                                                      float(PAD), N_samples=S, mlp_mode=mode, out=(rgb, depth)), 3)
        kern = sum(kern_ms) / len(kern_ms)

        # ---- e2e: host rays -> H2D -> kernel -> D2H, through the host-buffer call ------------------
        hfr = backend.HostFrameRenderer(N_RAYS, dev)
        for i in range(min(args.warmup, 2)):
            hfr.render(rays_host[i % n_distinct], vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(PAD), N_samples=S,
                       mlp_mode=mode)
        barrier()
        e2e_steps = args.steps
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            hfr.render(rays_host[(args.warmup + i) % n_distinct], vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(PAD),
                       N_samples=S, mlp_mode=mode)
            if world > 1:
                dist.barrier()
        barrier()
        e2e_t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e_value = world * N_RAYS * e2e_steps / float(e2e_t.item())

        # parity of this mode against the fp32 CUDA kernel on the last frame (the oracle-gated reference, tests/)
        other = {}
        if rank == 0 and mode != lib.MLP_FP32:
            r32 = torch.empty_like(rgb); d32 = torch.empty_like(depth)
            t32 = min(ev_time(lambda: backend.render_rays(rays_dev[-1], vol, d.imgs_raw, d.pose_source, fn, sc.near_far,
                                                          float(PAD), N_samples=S, mlp_mode=lib.MLP_FP32, out=(r32, d32)), 2))
            backend.render_rays(rays_dev[-1], vol, d.imgs_raw, d.pose_source, fn, sc.near_far, float(PAD), N_samples=S,
                                mlp_mode=mode, out=(rgb, depth))
            err = (rgb - r32).abs()
            other = {"fp32_mode": {"value": N_RAYS / (t32 * 1e-3), "unit": "rays/s", "ms_per_frame": t32,
                                   "note": "MVSN_MLP_FP32 kernel (FFMA), the 1e-4 parity mode"},
                     "parity_vs_fp32_kernel": {"rgb_linf": float(err.max()), "rgb_mse": float((err ** 2).mean()),
                                               "gate": 5e-3 if mode == lib.MLP_TC_HALF else 1e-4}}
            if mode == lib.MLP_TC_HALF:            # also report the fp32-grade tensor-core mode on the same frame
                rs = torch.empty_like(rgb); ds = torch.empty_like(depth)
                ts = min(ev_time(lambda: backend.render_rays(rays_dev[-1], vol, d.imgs_raw, d.pose_source, fn, sc.near_far,
                                                             float(PAD), N_samples=S, mlp_mode=lib.MLP_TC_SPLIT, out=(rs, ds)), 3))
                es = (rs - r32).abs()
                other["fp32_grade_tensor_mode"] = {
                    "value": N_RAYS / (ts * 1e-3), "unit": "rays/s", "ms_per_frame": ts,
                    "rgb_linf_vs_fp32_kernel": float(es.max()), "gate": 1e-4,
                    "executed_tensor_TFLOPs": 3 * N_RAYS * FLOP_PER_RAY / (ts * 1e-3) / 1e12,
                    "note": "MVSN_MLP_TC_SPLIT: 2-term fp16 operand split, 3 tcgen05 MMAs per K-step, fp32 accumulate"}

    if rank == 0:
        tflops = N_RAYS * FLOP_PER_RAY / (kern * 1e-3) / 1e12
        gbs = N_RAYS * BYTES_PER_RAY / (kern * 1e-3) / 1e9
        prof = {}
        pj = os.path.join(ROOT, "profiles", "render_kernel_ncu.json")
        if os.path.exists(pj):
            prof = json.load(open(pj)).get(args.mode, {})
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp32": "fp32", "half": "fp16 operands / fp32 accumulate",
                                           "split": "2x fp16 split operands / fp32 accumulate"}[args.mode],
            "data": "synthetic",
            "config": {"workload": "DTU-shaped 512x640 frame per GPU per step (configs[1]): 3 source views, pad 24, "
                                   "D=128 volume 8x128x176x208 resident, N_samples=128, ckpt mvsnerf-v0 weights",
                       "rays_per_gpu_per_step": N_RAYS, "mlp_mode": args.mode,
                       "parallelism": f"ray-sharded x{world}, volume replicated, all-gather of rgb+depth per step"
                                      if world > 1 else "single GPU",
                       "l2": "flushed between timed steps (256 MiB write) and inputs (150 MB volume) exceed L2"},
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": hfr.h2d_bytes * world,
                    "d2h_bytes_per_step": hfr.d2h_bytes * world, "steps": e2e_steps,
                    "api": "mvsnerf_b200.backend.HostFrameRenderer.render -> mvsn_render_rays (C ABI)"},
            "gpu_launches": n_launch,
            "roofline": {"bound": "tensor", "achieved": tflops, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": tflops / pk["bf16_tflops"], "traffic": prof.get("dram_bytes_per_launch"),
                         "kernel": prof.get("kernel", "render kernel"), "kernel_ms": kern, "peak_source": pk["source"],
                         "hbm_gather_GBs": gbs, "hbm_frac": gbs / pk["hbm_gbs"],
                         "note": "algorithmic MLP FLOPs (32 178 176 / ray) over the CUDA-event kernel time, vs the "
                                 "measured cuBLAS bf16 burst peak; hbm_* is the 51 248 B/ray gather definition"},
            "volume_build": {"ms": t_build, "featurenet_ms": t_feat, "cost_volume_ms": t_cost, "costreg_ms": t_reg,
                             "cost_volume_GBs": (176.0 * nvox + 8.6e6) / (t_cost * 1e-3) / 1e9,
                             "cost_volume_hbm_frac": (176.0 * nvox + 8.6e6) / (t_cost * 1e-3) / 1e9 / pk["hbm_gbs"],
                             "costreg_TFLOPs": 111.3e9 / (t_reg * 1e-3) / 1e12,
                             "note": "once per scene (K-F FeatureNet + K-A cost volume + K-B CostRegNet, all hand-written kernels), "
                                     "not inside the step"},
            "clocks": clocks, "wall_s_timed_region": t_wall,
        }
        line.update(other)
        tg = os.path.join(ROOT, "profiles", "r01_torch_gpu_baseline_and_fullsize_parity.json")
        if os.path.exists(tg):                              # recorded by tests/test_gpu_fullsize_oracle.py, not re-timed here
            rec = json.load(open(tg))
            line["reference_pytorch_gpu"] = {
                "value": rec["render_torch_gpu_rays_per_s"], "tf32_value": rec["render_torch_gpu_tf32_rays_per_s"],
                "unit": "rays/s", "volume_build_ms": rec["volume_build_torch_gpu_ms"],
                "note": "the oracle's PyTorch modules on one B200 (fp32, 5120-ray chunks), recorded by "
                        "tests/test_gpu_fullsize_oracle.py into profiles/; the north star's >=10x denominator"}
        if world == 1 and not args.no_cpu_baseline:
            orc, _, weights, sc_cpu = cpu_reference_setup()
            sample = args.cpu_sample
            rays = rays_host[-1][torch.randperm(N_RAYS, generator=torch.Generator().manual_seed(0))[:sample]].contiguous()
            vol_cpu = vol.detach().cpu().contiguous()
            threads = cpu_pick_threads(orc, weights, sc_cpu, rays, vol_cpu)         # also the warm-up
            dt = cpu_render_sample(orc, weights, sc_cpu, rays, vol_cpu)
            line["cpu_baseline"] = {"value": sample / dt, "unit": "rays/s", "cores": threads,
                                    "host_cpus": os.cpu_count(), "kind": "port",
                                    "sample": f"{sample} random rays of one frame x 128 samples ({dt:.1f} s), oracle "
                                              f"port of the reference path, torch CPU, best of several pool sizes"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=os.environ.get("MVSN_BENCH_MODE", "half"), choices=["fp32", "half", "split"])
    ap.add_argument("--cpu-sample", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
