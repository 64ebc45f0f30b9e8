mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/s12_tests.log 2>&1; tail -3 gpurun_out/s12_tests.log
B="python bench.py --no-cpu-baseline --no-e2e --samples 40 --steps 2 --warmup 2"
timeout 300 $B > gpurun_out/s12_a.log 2>&1; grep -o '"ms_per_step": [0-9.]*\|"kernel_ms_per_step": {[^}]*}\|"ms_per_step_1lane": [0-9.]*' gpurun_out/s12_a.log | tr '\n' ' '; echo
ncu --set full --clock-control none --import-source on -k regex:"s1_superk" -s 1 -c 1 -o gpurun_out/prof_s12 python bench.py --samples 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --lanes 1 > gpurun_out/s12_ncu.log 2>&1
ls -la gpurun_out/prof_s12*
