#!/bin/bash
# full ncu capture of one steady-state raycast_dda launch (and optionally other kernels), product library
set -u
TAG=${1:-r2a}; shift
mkdir -p gpurun_out
for k in ${KERNELS:-raycast_dda}; do
  ncu --set full --clock-control none --import-source on -k regex:"$k" -s 4 -c 1 -o gpurun_out/prof_${TAG}_$k \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_full_${TAG}_$k.log 2>&1
done
ls -la gpurun_out | tail -5
