# quick regression: all GPU tests + a 40-sample bench line (kernel spans)
mkdir -p gpurun_out
T=${1:-q}
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e --samples 40 --steps 2 --warmup 2 > gpurun_out/${T}_b40.log 2>&1
grep -o '"ms_per_step": [0-9.]*\|"kernel_ms_per_step": {[^}]*}\|"ms_per_step_1lane": [0-9.]*' gpurun_out/${T}_b40.log | tr '\n' ' '; echo
