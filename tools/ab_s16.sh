mkdir -p gpurun_out
B="timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 2"
$B --samples 20 --mode kmer:count:bin > gpurun_out/s16_kmer.log 2>&1; echo kmer:count k31 20 samples; grep -o '"ms_per_step": [0-9.]*\|"kernel_ms_per_step": {[^}]*}\|"ms_per_step_1lane": [0-9.]*\|"host_wall_ms_per_step": {[^}]*}' gpurun_out/s16_kmer.log | tr '\n' ' '; echo
$B --samples 20 --mode kmer:pa:bin --kmer-size 63 > gpurun_out/s16_k63.log 2>&1; echo kmer:pa k63 20 samples; grep -o '"ms_per_step": [0-9.]*\|"kernel_ms_per_step": {[^}]*}\|"ms_per_step_1lane": [0-9.]*\|"host_wall_ms_per_step": {[^}]*}' gpurun_out/s16_k63.log | tr '\n' ' '; echo; tail -c 300 gpurun_out/s16_k63.log
$B --samples 20 --mode hash:bft:bin > gpurun_out/s16_bft.log 2>&1; echo hash:bft 20 samples; grep -o '"ms_per_step": [0-9.]*\|"kernel_ms_per_step": {[^}]*}\|"ms_per_step_1lane": [0-9.]*\|"host_wall_ms_per_step": {[^}]*}' gpurun_out/s16_bft.log | tr '\n' ' '; echo
