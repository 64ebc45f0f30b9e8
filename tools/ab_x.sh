mkdir -p gpurun_out
B="timeout 300 python bench.py --no-cpu-baseline --no-e2e --samples 40 --steps 2 --warmup 2"
for v in "KMX_COMPACT_CAP=0" "KMX_COMPACT_CAP=5" "KMX_COMPACT_CAP=4" "KMX_COMPACT_CAP=3"; do
  env $v $B > gpurun_out/x.log 2>&1; echo "$v"; grep -o '"ms_per_step": [0-9.]*\|"kernel_ms_per_step": {[^}]*}\|"ms_per_step_1lane": [0-9.]*' gpurun_out/x.log | tr '\n' ' '; echo
done
for l in 2 3 6 8; do $B --lanes $l > gpurun_out/x.log 2>&1; echo "lanes=$l"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/x.log | head -1; done
