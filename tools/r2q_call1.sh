#!/bin/bash
# GPU-box script (round 2, session 4): full GPU test suite, the default bench line (with the new shim block), mirror chunk sweep
# through the C++ class, A/B of the predicated-flush DDA variant.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_1gpu_r2q.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_1gpu_r2q.log
tail -5 gpurun_out/pytest_1gpu_r2q.log
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_r2q.json 2> gpurun_out/bench_r2q.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2q.json"))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], d["roofline"]["frac_of_step_time"])
print("shim", json.dumps(d.get("shim"))[:1500])
PY
# mirror chunk sweep (eager mode through the C++ class)
python - <<'PY'
import os, struct, subprocess, sys
import numpy as np
sys.path.insert(0, ".")
from vdb_mapping_b200 import scans
c = scans.CONFIGS[2]
n = 16
with open("/tmp/scans.bin", "wb") as f:
    f.write(struct.pack("7d", c.resolution, c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max, n))
    for k in range(n):
        pts, origin = scans.make_scan(2, k)
        p16 = np.ones((pts.shape[0], 4), dtype=np.float32); p16[:, :3] = pts[:, :3]
        f.write(struct.pack("3d", *origin)); f.write(struct.pack("I", p16.shape[0])); f.write(p16.tobytes())
for chunk in (4096, 8192, 16384, 32768, 65536):
    env = dict(os.environ, VDBM_MIRROR_CHUNK=str(chunk))
    p = subprocess.run(["tools/build/bench_shim", "/tmp/scans.bin", "eager", "4"], capture_output=True, text=True, env=env)
    print("chunk", chunk, p.stdout.strip()[-200:], p.stderr[-200:], flush=True)
PY
VARIANTS="pred" bash tools/ab_variants.sh r2q
cat gpurun_out/ab_r2q.txt
