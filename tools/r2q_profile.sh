#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_1gpu_r2q.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_1gpu_r2q.log
tail -6 gpurun_out/pytest_1gpu_r2q.log
bash tools/profile_gpu.sh r2q > gpurun_out/profile_r2q.log 2>&1
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2q.json"))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], d["roofline"]["frac_of_step_time"])
print("by_kernel", {k: round(v["ms"], 4) for k, v in d["roofline"]["by_kernel"].items()})
print("shim", json.dumps({k: v for k, v in d.get("shim", {}).items() if k not in ("api", "modes")}))
print("strong", d["strong_cfg4"]["ms_per_step"] if d.get("strong_cfg4") else None)
r = json.load(open("gpurun_out/bench_ref_r2q.json"))
print("ref", r["value"], r["ms_per_step"], r["config"].get("sample"))
PY
python bench.py --workload cfg5 --steps 100 --warmup 10 > gpurun_out/bench_cfg5_r2q.json 2>> gpurun_out/bench_r2q.err; echo "cfg5 rc=$?"
for w in cfg1 cfg3 cfg4; do python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_r2q.json 2>> gpurun_out/bench_r2q.err; echo "$w rc=$?"; done
python tools/bench_shim.py 30 > gpurun_out/bench_shim_r2q.txt 2>&1; cat gpurun_out/bench_shim_r2q.txt
ls -la gpurun_out | tail -30
