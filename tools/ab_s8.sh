mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_misc.py -m gpu -x -q) > gpurun_out/s8_tests.log 2>&1; tail -5 gpurun_out/s8_tests.log
B="python bench.py --no-cpu-baseline --no-e2e --samples 40 --steps 2 --warmup 2"
for sc in 1 2 3; do
  KMX_SWEEP_CAP=$sc timeout 300 $B > gpurun_out/s8_s${sc}.log 2>&1; echo "sweep_cap=$sc"; grep -o '"ms_per_step": [0-9.]*\|"kernel_ms_per_step": {[^}]*}\|"ms_per_step_1lane": [0-9.]*' gpurun_out/s8_s${sc}.log | tr '\n' ' '; echo
done
