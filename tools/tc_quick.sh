#!/bin/bash
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --steps 5 --warmup 3 --mode split --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_split.json; cut -c1-330 gpurun_out/bench_split.json; python -c "
import json; d=json.load(open('gpurun_out/bench_split.json')); print('SPLIT', d['ms_per_step'], d['value'], d['e2e']['value'], d.get('parity_vs_fp32_kernel'))"
