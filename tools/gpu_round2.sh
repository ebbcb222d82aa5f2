#!/bin/bash
# One GPU round under gpurun (round 2): tests, bench (both arms), ncu launch list, ncu full capture of the render kernel.
# usage: tools/gpu_round2.sh <tag> [mode] [skip-tests]
TAG=${1:-r02}
MODE=${2:-pair}
OUT=gpurun_out/$TAG
mkdir -p $OUT
declare -A KNAME=( [pair]=render_tc2_kernel [half]=render_tc_kernel [split]=render_tcs_kernel [fp32]=render_fp32_kernel )
KREGEX='regex:render_|conv|toplayer|cost_volume|finalize_volume|downsample_images|feats_to_quads|pack_|vol_|adam|bwd'
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
if [ -z "$3" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  tail -8 $OUT/pytest_gpu.log
fi
timeout 900 python bench.py --steps 10 --warmup 3 --mode $MODE > $OUT/bench_$MODE.json 2> $OUT/bench_$MODE.err; echo "bench rc=$?"; tail -3 $OUT/bench_$MODE.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$MODE.json")); print(json.dumps({k:d.get(k) for k in ("value","ms_per_step","e2e","roofline","fp32_tier","reference_pytorch_gpu","strong_scaling","finetune_step","parity_vs_fp32_kernel","volume_build","cpu_baseline")})[:4000])
except Exception as e: print("bench parse failed", e)
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"; tail -c 400 $OUT/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 400 --csv --log-file $OUT/launches_$MODE.csv \
    python bench.py --steps 2 --warmup 3 --mode $MODE --no-cpu-baseline --no-fp32-tier --no-torch-gpu --no-finetune > $OUT/ncu_launch_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${KNAME[$MODE]} -s 3 -c 1 -o $OUT/render_$MODE \
    python bench.py --steps 1 --warmup 3 --mode $MODE --no-cpu-baseline --no-fp32-tier --no-torch-gpu --no-finetune > $OUT/ncu_full_run.log 2>&1
ls -la $OUT
