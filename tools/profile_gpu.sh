#!/bin/bash
# Runs on the GPU box (under gpurun). Produces under gpurun_out/:
#   launches_<tag>.csv   every kernel launch of a short bench run with its device time (ncu, cold cache, serialised)
#   prof_<tag>.ncu-rep   ncu --set full captures of the hot kernels (steady-state scans)
#   bench_<tag>.json     the bench line of the same build (NOT under the profiler)
set -u
TAG=${1:-r1}
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"raycast_dda|apply_update|resolve_leaves|compact_leaves|prep_rays" -s 30 -c 5 \
    -o gpurun_out/prof_${TAG} python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
