mkdir -p gpurun_out/r02j
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bn_modes.py tests/test_gpu_seams.py tests/test_gpu_configs.py -m gpu -q -x > gpurun_out/r02j/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02j/pytest.log | cut -c1-200
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-torch-gpu --no-fp32-tier --no-finetune > gpurun_out/r02j/bench.json 2> gpurun_out/r02j/bench.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02j/bench.json') if l.startswith('{')][-1]); print('volume_build', json.dumps(d['volume_build'])[:300]); print(d['value'])"
