#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_1gpu_r2r.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_1gpu_r2r.log
tail -4 gpurun_out/pytest_1gpu_r2r.log
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_r2r.json 2> gpurun_out/bench_r2r.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2r.json"))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], d["roofline"]["frac_of_step_time"])
print("by_kernel", {k: round(v["ms"], 4) for k, v in d["roofline"]["by_kernel"].items()})
print("shim", json.dumps({k: v for k, v in d.get("shim", {}).items() if k not in ("api", "modes")}))
print("strong", d["strong_cfg4"]["ms_per_step"] if d.get("strong_cfg4") else None)
PY
VDBM_MIRROR_PROFILE=1 python tools/bench_shim.py 40 2>&1 | tail -12
for w in cfg4 cfg5; do python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_r2r.json 2>> gpurun_out/bench_r2r.err; python -c "
import json;d=json.load(open('gpurun_out/bench_${w}_r2r.json'));print('$w', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), 'upd', d['roofline']['by_kernel']['apply_update_kernel'])"; done
