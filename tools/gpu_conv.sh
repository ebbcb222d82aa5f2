#!/bin/bash
# GPU check of the volume-build kernels (K-F/K-A/K-B): tests, bench line, launch list of the builds.
TAG=${1:-r01g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
cp gpurun_out/torch_gpu_baseline.json gpurun_out/finetune_step.json $OUT/ 2>/dev/null
timeout 600 python bench.py --steps 10 --warmup 3 --mode half > $OUT/bench_half.json 2> $OUT/bench_half.err; echo "bench rc=$?"
python - <<PY
import json; d=json.load(open("$OUT/bench_half.json")); print(json.dumps({k:d[k] for k in ("value","ms_per_step","volume_build","e2e")}))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:conv|toplayer|cost_volume|finalize_volume|downsample' -c 120 --csv \
    --log-file $OUT/launches_volume.csv python bench.py --steps 1 --warmup 3 --mode half --no-cpu-baseline > $OUT/ncu_launch_run.log 2>&1
ls -la $OUT
