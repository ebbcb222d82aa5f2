#!/bin/bash
mkdir -p gpurun_out/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_tc -s 3 -c 1 -o gpurun_out/ncu/render_half_cur \
    python bench.py --steps 1 --warmup 3 --mode half --no-cpu-baseline > gpurun_out/ncu/run.log 2>&1
ls -la gpurun_out/ncu
