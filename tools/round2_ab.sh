#!/usr/bin/env bash
# First GPU call of the next round: A/B of experiments that are compiled out by default.
#   RS_VALS_ASYNC   radix pass: values prefetched to shared memory with cp.async (radix_sort.cuh)
# Build here (nvcc cross-compiles):   bash tools/round2_ab.sh build
# Run on the GPU box:                 gpurun -- 'bash tools/round2_ab.sh run'
set -e
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo"
mkdir -p build/sb
if [ "${1:-build}" = build ]; then
  $NV tools/sortbench.cu -o build/sb/sb_base
  $NV -DRS_VALS_ASYNC tools/sortbench.cu -o build/sb/sb_vals_async
  $NV -DRS_VALS_ASYNC -DRS_IPT64_CFG=12 -DRS_MIN_CTAS_CFG=4 tools/sortbench.cu -o build/sb/sb_vals_async_12x4 || true
  ls -la build/sb
else
  mkdir -p gpurun_out/ab
  for b in build/sb/sb_*; do for lg in 28 30; do timeout 120 $b $lg 48 | sed "s#^#$(basename $b): #" | tee -a gpurun_out/ab/sortbench.txt; done; done
fi
