#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_1gpu_r2q.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_1gpu_r2q.log
tail -4 gpurun_out/pytest_1gpu_r2q.log
{
echo "compute-sanitizer on tools/sanitize_smoke.py (scan path sync + queued/overlapped, fast_mode, raytrace, sections, import, host mirror stream, vdbm_integrate_from, 2-shard group), B200, final code of round 2"
for tool in memcheck initcheck synccheck; do
  echo "== $tool"
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_smoke.py 2>&1 | grep -v "^=========  \|^$" | tail -8
done
} > gpurun_out/sanitizer_r2q.txt 2>&1
cat gpurun_out/sanitizer_r2q.txt
python __graft_entry__.py smoke 2>&1 | tail -2
