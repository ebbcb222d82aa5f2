#!/bin/bash
# build here (no GPU needed): nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tma_probe/tma_probe tools/tma_probe/tma_probe.cu
for v in 6 7 2 3; do timeout 60 tools/tma_probe/tma_probe $v 2>&1 | tail -1; done
