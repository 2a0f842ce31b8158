#!/bin/bash
# Runs every variant of build/mma_rate (scripts/mma_rate.cu) in its own process, on 2 CTAs and on all 148 SMs.
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/mma_rate scripts/mma_rate.cu
out=${1:-gpurun_out/mma_rate.txt}
: > "$out"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> "$out"
for grid in ${GRIDS:-2 148}; do
  for v in ${VARIANTS:-0 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15 16 17 18 19 20 21 22 23 24 25 26}; do
    timeout 20 build/mma_rate $v $grid >> "$out" 2>&1 || echo "variant $v grid $grid: exit $?" >> "$out"
  done
done
cat "$out"
