#!/bin/bash
# GPU-box script (round 2): parity suite, bench line, then A/B of library variants built by tools/build_variant.sh.
set -u
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_${TAG}.log
tail -5 gpurun_out/pytest_${TAG}.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
ab() {
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-strong-block $2 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['by_kernel']; print('$1', 'dda_ms', round(k['raycast_dda_kernel']['ms'],3), 'upd_ms', round(k['apply_update_kernel']['ms'],3), 'ms/step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3))" | tee -a gpurun_out/ab_${TAG}.txt
}
ab product ""
cp vdb_mapping_b200/libvdbm_b200.so /tmp/product.so
for v in ${VARIANTS:-}; do
  cp tools/build/$v/libvdbm_b200.so vdb_mapping_b200/libvdbm_b200.so
  if [ "$v" = exp ]; then
    for mode in 1 2; do VDBM_DDA_MODE=$mode ab "exp_mode$mode" ""; done
  else
    ab "$v" ""
  fi
done
cp /tmp/product.so vdb_mapping_b200/libvdbm_b200.so
for w in cfg1 cfg3; do ab "product_$w" "--workload $w"; done
