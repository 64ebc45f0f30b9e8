mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_misc.py tests/test_gpu_golden.py -m gpu -x -q) > gpurun_out/s15_tests.log 2>&1; tail -3 gpurun_out/s15_tests.log
B="timeout 300 python bench.py --no-cpu-baseline --no-e2e --samples 40 --steps 2 --warmup 2"
for v in "KMX_HIST_NOFUSE=1" "KMX_HIST_CAP=4" "KMX_HIST_CAP=3" "KMX_HIST_CAP=4 KMX_HR_TILE=512" "KMX_HIST_CAP=4 KMX_HR_TILE=1024"; do
  env $v $B > gpurun_out/s15_x.log 2>&1; echo "$v"; grep -o '"ms_per_step": [0-9.]*\|"kernel_ms_per_step": {[^}]*}\|"ms_per_step_1lane": [0-9.]*' gpurun_out/s15_x.log | tr '\n' ' '; echo
done
