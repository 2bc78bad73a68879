#!/bin/bash
# 2-GPU box: multi-GPU parity tests + the torchrun bench line on the final code of the round
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_group.py tests/test_dist_gloo.py -m gpu -v --tb=short 2>&1 | tail -40 > gpurun_out/pytest_2gpu_r2q.log; tail -22 gpurun_out/pytest_2gpu_r2q.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_r2q_2gpu.json 2> gpurun_out/bench_r2q_2gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2q_2gpu.json').read().strip().splitlines()[-1])
s=d['strong_cfg4']
print('weak ms/step', d['ms_per_step'], 'rays/s', d['value'], 'e2e', d['e2e']['value'], 'identical', d.get('sharded_map_identical'))
print('strong ms', s['ms_per_step'], 'identical', s['sharded_map_identical'], s['per_rank_ms'])
PY
timeout 300 python tools/bench_group.py --gpus 2 --workload cfg4 --steps 8 --warmup 4 2>&1 | tail -1 > gpurun_out/bench_group_cfg4_n2_r2q.json; cut -c1-500 gpurun_out/bench_group_cfg4_n2_r2q.json
