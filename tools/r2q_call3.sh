#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_1gpu_r2q.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_1gpu_r2q.log
tail -6 gpurun_out/pytest_1gpu_r2q.log
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
env = dict(os.environ, VDBM_MIRROR_PROFILE="1")
for exe in ("bench_shim", "bench_shim_plaincopy", "bench_shim"):
    for chunk in ("16384", "65536"):
        env["VDBM_MIRROR_CHUNK"] = chunk
        p = subprocess.run(["tools/build/" + exe, "/tmp/scans.bin", "eager", "4", "14"], capture_output=True, text=True, env=env)
        print(exe, chunk, p.stdout.strip()[-120:]); print("\n".join(p.stderr.strip().splitlines()[-2:]), flush=True)
PY
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_r2q.json 2> gpurun_out/bench_r2q.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2q.json"))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], d["roofline"]["frac_of_step_time"])
print("shim", json.dumps({k: v for k, v in d.get("shim", {}).items() if k not in ("api", "modes")}))
PY
