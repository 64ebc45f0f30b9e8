mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/s13_tests.log 2>&1; tail -3 gpurun_out/s13_tests.log
B="python bench.py --no-cpu-baseline --no-e2e --samples 40 --steps 2 --warmup 2"
timeout 300 $B > gpurun_out/s13_a.log 2>&1; grep -o '"ms_per_step": [0-9.]*\|"kernel_ms_per_step": {[^}]*}\|"ms_per_step_1lane": [0-9.]*' gpurun_out/s13_a.log | tr '\n' ' '; echo
