mkdir -p gpurun_out/r02b
timeout 300 python tools/diag_c5.py > gpurun_out/r02b/diag_c5.log 2>&1; echo "diag rc=$?"; tail -8 gpurun_out/r02b/diag_c5.log | cut -c1-900
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02b/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/r02b/pytest_gpu.log | cut -c1-400
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-torch-gpu --no-fp32-tier > gpurun_out/r02b/bench.json 2> gpurun_out/r02b/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02b/bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02b/bench.json')); print(json.dumps({k:d.get(k) for k in ('value','finetune_step','strong_scaling')}))"
