#!/bin/bash
# Run on the GPU box via gpurun: tests, bench (both arms), ncu launch list, ncu full captures.
# usage: tools/gpu_round.sh <tag> [mode]
TAG=${1:-r01}
MODE=${2:-half}
OUT=gpurun_out/$TAG
mkdir -p $OUT
KREGEX='regex:render_|conv3d_|deconv3d_|cost_volume|finalize_volume|downsample_images|pack_|vol_'
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --mode $MODE > $OUT/bench_$MODE.json 2> $OUT/bench_$MODE.err; echo "bench rc=$?"
tail -c 3800 $OUT/bench_$MODE.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 400 --csv --log-file $OUT/launches_$MODE.csv \
    python bench.py --steps 2 --warmup 3 --mode $MODE --no-cpu-baseline > $OUT/ncu_launch_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_ -s 3 -c 1 -o $OUT/render_$MODE \
    python bench.py --steps 1 --warmup 3 --mode $MODE --no-cpu-baseline > $OUT/ncu_full_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_tcs -c 1 -o $OUT/render_split \
    python bench.py --steps 1 --warmup 3 --mode split --no-cpu-baseline > $OUT/ncu_split_run.log 2>&1
ls -la $OUT
