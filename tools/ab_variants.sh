#!/bin/bash
# GPU-box script: A/B of library variants built by tools/build_variant.sh against the product library (short bench runs).
set -u
TAG=${1:-ab}
mkdir -p gpurun_out
ab() {
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-strong-block --no-shim-block $2 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['by_kernel']; print('$1', 'dda_ms', round(k['raycast_dda_kernel']['ms'],3), 'upd_ms', round(k['apply_update_kernel']['ms'],3), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3))" | tee -a gpurun_out/ab_${TAG}.txt
}
ab product ""
cp vdb_mapping_b200/libvdbm_b200.so /tmp/product.so
for v in ${VARIANTS:-}; do
  cp tools/build/$v/libvdbm_b200.so vdb_mapping_b200/libvdbm_b200.so
  ab "$v" ""
  for w in ${WORKLOADS:-}; do ab "${v}_$w" "--workload $w"; done
done
cp /tmp/product.so vdb_mapping_b200/libvdbm_b200.so
for w in ${WORKLOADS:-}; do ab "product_$w" "--workload $w"; done
