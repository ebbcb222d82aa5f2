#!/bin/bash
mkdir -p gpurun_out/tc
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_umma.py -x -q -k "tc or umma" 2>&1 | tail -15
timeout 300 python bench.py --steps 10 --warmup 3 --mode half --no-cpu-baseline 2>&1 | tail -3 | cut -c1-1500
