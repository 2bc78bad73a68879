#!/bin/bash
# 8-GPU box: multi-GPU parity tests (kept as evidence), the single-process group bench, the torchrun bench line.
set -u
N=${1:-8}
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_group.py tests/test_dist_gloo.py -m gpu -v --tb=short 2>&1 | tail -40 > gpurun_out/pytest_${N}gpu.log; tail -25 gpurun_out/pytest_${N}gpu.log
for n in 4 $N; do timeout 300 python tools/bench_group.py --gpus $n --workload cfg4 --steps 8 --warmup 4 2>&1 | tail -1 > gpurun_out/bench_group_cfg4_n$n.json; cut -c1-700 gpurun_out/bench_group_cfg4_n$n.json; done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().splitlines()[-1])
s=d['strong_cfg4']
print('weak ms/step', d['ms_per_step'], 'rays/s', d['value'], 'e2e', d['e2e']['value'], 'identical', d.get('sharded_map_identical'))
print('strong ms', s['ms_per_step'], 'identical', s['sharded_map_identical'], s['per_rank_ms'])
PY
