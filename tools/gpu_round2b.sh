#!/bin/bash
# Round 2, second half: full GPU test tier, full bench line, reference arm, launch list, ncu --set full of conv0_tc, sanitizer
# on the encoder (new TMA / tcgen05 conv0).   usage: tools/gpu_round2b.sh <tag>
TAG=${1:-r02z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_pair.json 2> $OUT/bench_pair.err; echo "bench rc=$?"; tail -2 $OUT/bench_pair.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/bench_pair.json") if l.startswith("{")][-1]); print(json.dumps({k:d.get(k) for k in ("value","ms_per_step","e2e","volume_build","finetune_step","fp32_tier")})[:2500])
except Exception as e: print("bench parse failed", e)
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:render_|conv|toplayer|cost_volume|finalize_volume|downsample_images|feats_to_quads|pack_|vol_|adam|bwd|reduce|bn_update' -c 500 --csv --log-file $OUT/launches_pair.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp32-tier --no-torch-gpu > $OUT/ncu_launch_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv0_tc_kernel -s 1 -c 1 -o $OUT/conv0_tc \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-fp32-tier --no-torch-gpu --no-finetune > $OUT/ncu_conv0_run.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --log-file $OUT/sanitizer_memcheck_encoder.log python tools/sanitize_smoke.py encoder > $OUT/san1.out 2>&1; tail -2 $OUT/sanitizer_memcheck_encoder.log
timeout 600 compute-sanitizer --tool racecheck --log-file $OUT/sanitizer_racecheck_encoder.log python tools/sanitize_smoke.py encoder > $OUT/san2.out 2>&1; tail -2 $OUT/sanitizer_racecheck_encoder.log
ls $OUT
