#!/usr/bin/env bash
# usage (on a GPU box): tools/ubench/run.sh   -- builds here or there; prints the throughput table
set -e
cd "$(dirname "$0")"
[ -x atoms ] || /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o atoms atoms.cu
./atoms
