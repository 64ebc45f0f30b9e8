mkdir -p gpurun_out
N="ncu --set full --clock-control none -k regex:hash_hist_roll -s 2 -c 1"
B="python bench.py --samples 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --lanes 1"
KMX_HIST_CAP=4 $N -o gpurun_out/prof_s6_c4 $B > gpurun_out/s6.log 2>&1
KMX_HIST_CAP=2 $N -o gpurun_out/prof_s6_c2 $B >> gpurun_out/s6.log 2>&1
KMX_HIST_CAP=4 KMX_HR_TILE=1024 $N -o gpurun_out/prof_s6_c4_t1024 $B >> gpurun_out/s6.log 2>&1
KMX_HIST_CAP=8 KMX_HR_TILE=1024 $N -o gpurun_out/prof_s6_c8_t1024 $B >> gpurun_out/s6.log 2>&1
ls -la gpurun_out/prof_s6*
