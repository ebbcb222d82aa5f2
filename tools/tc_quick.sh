#!/bin/bash
mkdir -p gpurun_out/tc
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "tc" 2>&1 | tail -25
