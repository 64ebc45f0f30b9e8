#!/usr/bin/env bash
# Builds libkmx_sm100.so (CUDA kernels + C ABI) in-tree for sm_100a.
set -euo pipefail
cd "$(dirname "$0")"
SRC=kmtricks_b200/csrc
OUT=kmtricks_b200/libkmx_sm100.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall ${KMX_NVCC_EXTRA:-}"
mkdir -p kmtricks_b200/_build
pids=()
for f in s1_superk s1_v5 s2_hash s2_bin s2_sort s2_ht s3_merge s4_bits synth kmx_api; do
  [ -f $SRC/$f.cu ] || continue
  if [ ! -f kmtricks_b200/_build/$f.o ] || [ -n "$(find $SRC include -newer kmtricks_b200/_build/$f.o -type f | head -1)" ]; then
    $NVCC $FLAGS -c $SRC/$f.cu -o kmtricks_b200/_build/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT kmtricks_b200/_build/*.o -lcudart -ldl
echo "built $OUT"
# C++ host (CLI + run-dir + IMergePlugin host) over the C ABI
mkdir -p kmtricks_b200/bin
g++ -std=c++17 -O2 -Wall -Iinclude -o kmtricks_b200/bin/kmx kmtricks_b200/csrc/host/kmx_main.cpp -Lkmtricks_b200 -lkmx_sm100 -ldl -lz -lpthread -Wl,-rpath,'$ORIGIN/..' -Wl,-rpath,/usr/local/cuda/lib64
echo "built kmtricks_b200/bin/kmx"
