#!/bin/bash
# One GPU round under gpurun: tests, bench (both arms), ncu launch list, ncu full captures.
# usage: tools/gpu_round.sh <tag> [mode]
TAG=${1:-r01}
MODE=${2:-half}
OUT=gpurun_out/$TAG
mkdir -p $OUT
KREGEX='regex:render_|conv|toplayer|cost_volume|finalize_volume|downsample_images|feats_to_quads|pack_|vol_'
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log
cp gpurun_out/torch_gpu_baseline.json gpurun_out/finetune_step.json $OUT/ 2>/dev/null
timeout 600 python bench.py --steps 10 --warmup 3 --mode $MODE > $OUT/bench_$MODE.json 2> $OUT/bench_$MODE.err; echo "bench rc=$?"
python - <<PY
import json; d=json.load(open("$OUT/bench_$MODE.json")); print(json.dumps({k:d.get(k) for k in ("value","ms_per_step","volume_build","e2e","roofline","fp32_mode","fp32_grade_tensor_mode","cpu_baseline")})[:2500])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"; tail -c 600 $OUT/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 400 --csv --log-file $OUT/launches_$MODE.csv \
    python bench.py --steps 2 --warmup 3 --mode $MODE --no-cpu-baseline > $OUT/ncu_launch_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_tc_kernel -s 3 -c 1 -o $OUT/render_$MODE \
    python bench.py --steps 1 --warmup 3 --mode $MODE --no-cpu-baseline > $OUT/ncu_full_run.log 2>&1
# volume-build kernels: K-A, K-B (generic conv, sub-pixel deconv), K-F (5x5 s2 layer)
timeout 600 ncu --set full --clock-control none --import-source on \
    -k 'regex:cost_volume_kernel|deconv3d_subpixel|conv3d_k3_kernel|conv2d_kernel' -s 18 -c 18 -o $OUT/volume_kernels \
    python bench.py --steps 1 --warmup 3 --mode $MODE --no-cpu-baseline > $OUT/ncu_volume_run.log 2>&1
ls -la $OUT
