#!/bin/bash
mkdir -p gpurun_out/tc
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -k "tc_half_vs_oracle" 2>&1 | tail -4
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python bench.py --steps 10 --warmup 3 --mode half --no-cpu-baseline 2>&1 | tail -3 | cut -c1-330
MVSN_LIB=$PWD/mvsnerf_b200/libmvsnerf_b200_trace.so timeout 100 python tools/tc_trace.py > gpurun_out/trace_latest.txt 2>&1; head -1 gpurun_out/trace_latest.txt
