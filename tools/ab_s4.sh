mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_misc.py -m gpu -x -q) > gpurun_out/s4_tests.log 2>&1; tail -3 gpurun_out/s4_tests.log
B="python bench.py --no-cpu-baseline --no-e2e --samples 40 --steps 2 --warmup 2"
for t in 256 512 1024; do KMX_HR_TILE=$t $B > gpurun_out/s4_tile$t.log 2>&1; echo tile$t; grep -o '"ms_per_step": [0-9.]*\|"kernel_ms_per_step": {[^}]*}' gpurun_out/s4_tile$t.log; done
KMX_HIST_NOROLL=1 $B > gpurun_out/s4_noroll.log 2>&1; echo noroll; grep -o '"ms_per_step": [0-9.]*\|"kernel_ms_per_step": {[^}]*}' gpurun_out/s4_noroll.log
ncu --set full --clock-control none --import-source on -k regex:"hash_hist_roll|hash_sweep|fq_index_lines" -s 6 -c 3 -o gpurun_out/prof_s4 python bench.py --samples 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --lanes 1 > gpurun_out/s4_ncu.log 2>&1
KMX_HIST_NOROLL=1 ncu --set full --clock-control none --import-source on -k regex:"hash_hist_kernel" -s 2 -c 1 -o gpurun_out/prof_s4_noroll python bench.py --samples 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --lanes 1 >> gpurun_out/s4_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
