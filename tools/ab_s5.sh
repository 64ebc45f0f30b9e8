mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_misc.py -m gpu -x -q) > gpurun_out/s5_tests.log 2>&1; tail -3 gpurun_out/s5_tests.log
B="python bench.py --no-cpu-baseline --no-e2e --samples 40 --steps 2 --warmup 2"
for hc in 2 3 4 8; do for sc in 2 4; do
  KMX_HIST_CAP=$hc KMX_SWEEP_CAP=$sc $B > gpurun_out/s5_h${hc}_s${sc}.log 2>&1; echo "hist_cap=$hc sweep_cap=$sc"; grep -o '"ms_per_step": [0-9.]*\|"kernel_ms_per_step": {[^}]*}\|"ms_per_step_1lane": [0-9.]*' gpurun_out/s5_h${hc}_s${sc}.log | tr '\n' ' '; echo
done; done
for l in 2 6 8; do KMX_HIST_CAP=3 KMX_SWEEP_CAP=2 $B --lanes $l > gpurun_out/s5_l$l.log 2>&1; echo "lanes=$l h3 s2"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/s5_l$l.log | head -1; done
