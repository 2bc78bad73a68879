#!/bin/bash
# GPU-box script (N GPUs): 2-GPU parity tests of the sharded map, then the bench line(s) with the strong_cfg4 block.
set -u
N=${1:-2}; TAG=${2:-r2m}
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != 1 ]; then
  timeout 900 python -m pytest tests/test_dist_gloo.py -m gpu -x -q > gpurun_out/pytest_dist_${TAG}.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_dist_${TAG}.log
  tail -4 gpurun_out/pytest_dist_${TAG}.log
fi
for n in ${NS:-$N}; do
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err
  else
    VDBM_DIST_PROFILE=${PROFILE:-} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_n$n.json 2> gpurun_out/bench_${TAG}_n$n.err
  fi
  echo "bench n=$n rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_n$n.json").read().strip().splitlines()[-1])
    s=d.get("strong_cfg4") or {}
    print("n=$n weak ms/step", round(d["ms_per_step"],3), "rays/s", round(d["value"]/1e6,1), "M identical", d.get("sharded_map_identical"))
    print("   strong ms/step", s.get("ms_per_step"), "identical", s.get("sharded_map_identical"), "phases", json.dumps(s.get("per_rank_ms")))
    print("   touched", s.get("leaves_touched_per_rank"), "owned", s.get("leaves_owned_per_rank"))
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/bench_${TAG}_n$n.err").read()[-3000:])
PY
done
