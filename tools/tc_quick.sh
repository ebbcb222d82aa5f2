#!/bin/bash
mkdir -p gpurun_out/tc
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --mode half --no-cpu-baseline 2>&1 | tail -3 | cut -c1-330
