#!/bin/bash
# Runs on the GPU box (under gpurun). Produces under gpurun_out/:
#   bench_<tag>.json / bench_ref_<tag>.json   the bench lines of this build (NOT under the profiler)
#   launches_<tag>.csv   every kernel launch of a short bench run with its device time (ncu, cold cache, serialised)
#   prof_<tag>.ncu-rep   ncu --set full captures of the hot kernels on STEADY-STATE scans (each kernel is captured after
#                        skipping its first 6 launches = the scans that create the map; bench runs 2 legs x 6 scans)
set -u
TAG=${1:-r1}
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong-block --no-shim-block > gpurun_out/ncu_launches_${TAG}.log 2>&1
# (VDBM_OVERLAP=0: under the profiler kernels are serialised anyway, and with the two-stream pipeline the DRAM counters of
#  one kernel would include the write-back of the update kernel that ran on the other stream just before it)
for k in raycast_dda apply_update resolve_leaves compact_leaves prep_rays merge_near; do
  VDBM_OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 4 -c 1 -o gpurun_out/prof_${TAG}_$k \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong-block --no-shim-block > gpurun_out/ncu_full_${TAG}_$k.log 2>&1
done
ls -la gpurun_out | tail -12
