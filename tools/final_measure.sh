# round-end measurement set: GPU tests, default bench, ncu launch list, ncu --set full of the top kernels
mkdir -p gpurun_out
T=${1:-r1c}
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
(time python bench.py) > gpurun_out/${T}_bench.log 2>&1; tail -c 2600 gpurun_out/${T}_bench.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${T}.csv python bench.py --samples 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --lanes 1 > gpurun_out/${T}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"s1_superk|hash_hist_roll|hash_compact|hash_copy|hash_scan|fq_index_lines|fq_count_newlines|fq_cta_pos" -s 7 -c 7 -o gpurun_out/prof_${T} python bench.py --samples 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --lanes 1 > gpurun_out/${T}_ncu_full.log 2>&1
ls -la gpurun_out/prof_${T}* gpurun_out/launches_${T}.csv
