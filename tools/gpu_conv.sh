#!/bin/bash
# GPU check of the volume-build kernels (K-A/K-B): tests, bench line, launch list of one build, ncu full of conv0.
TAG=${1:-r01f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --mode half > $OUT/bench_half.json 2> $OUT/bench_half.err; echo "bench rc=$?"
tail -c 3000 $OUT/bench_half.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:conv|cost_volume|finalize_volume|downsample' -c 60 --csv \
    --log-file $OUT/launches_volume.csv python bench.py --steps 1 --warmup 3 --mode half --no-cpu-baseline > $OUT/ncu_launch_run.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv0_k3 -s 1 -c 1 -o $OUT/conv0 \
    python bench.py --steps 1 --warmup 3 --mode half --no-cpu-baseline > $OUT/ncu_conv0_run.log 2>&1
ls -la $OUT
