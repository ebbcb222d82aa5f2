mkdir -p gpurun_out/r02c
timeout 600 python tools/diag_c5.py > gpurun_out/r02c/diag_c5.log 2>&1; echo "diag rc=$?"; tail -4 gpurun_out/r02c/diag_c5.log | cut -c1-2500
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02c/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02c/pytest_gpu.log | cut -c1-300
for tool in memcheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/r02c/sanitizer_$tool.log python tools/sanitize_smoke.py > gpurun_out/r02c/sanitizer_$tool.out 2>&1; echo "$tool rc=$?"; tail -3 gpurun_out/r02c/sanitizer_$tool.log
done
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r02c/sanitizer_racecheck.log python tools/sanitize_smoke.py encoder backward > gpurun_out/r02c/sanitizer_racecheck.out 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r02c/sanitizer_racecheck.log
