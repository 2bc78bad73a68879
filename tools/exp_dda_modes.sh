#!/bin/bash
# Quick A/B of the scan step on the GPU box: per-kernel times of a short bench run.
# The DDA marking experiments (VDBM_DDA_MODE=1: no mask writes, 2: plain stores) exist only in a library built with
# -DVDBM_EXPERIMENTS (add it to NVCC_FLAGS in vdb_mapping_b200/build.py); the product build ignores the variable.
for mode in ${MODES:-0}; do
  VDBM_DDA_MODE=$mode python bench.py --steps 20 --warmup 3 --no-cpu-baseline | python -c "
import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['by_kernel']; print('mode $mode', 'dda_ms', round(k['raycast_dda_kernel']['ms'],3), 'upd_ms', round(k['apply_update_kernel']['ms'],3), 'ms/step', round(d['ms_per_step'],3))"
done
