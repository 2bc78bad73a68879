#!/bin/bash
set -u
mkdir -p gpurun_out
./tests/cpp/build/test_shim_kats > gpurun_out/shim_kats_r2q.log 2>&1; echo "shim_kats rc=$?"; tail -12 gpurun_out/shim_kats_r2q.log
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mirror or dirty" 2>&1 | tail -5
python - <<'PY'
import os, struct, subprocess, sys
import numpy as np
sys.path.insert(0, ".")
from vdb_mapping_b200 import scans
c = scans.CONFIGS[2]
n = 30
with open("/tmp/scans.bin", "wb") as f:
    f.write(struct.pack("7d", c.resolution, c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max, n))
    for k in range(n):
        pts, origin = scans.make_scan(2, k)
        p16 = np.ones((pts.shape[0], 4), dtype=np.float32); p16[:, :3] = pts[:, :3]
        f.write(struct.pack("3d", *origin)); f.write(struct.pack("I", p16.shape[0])); f.write(p16.tobytes())
env = dict(os.environ, VDBM_MIRROR_PROFILE="1")
p = subprocess.run(["tools/build/bench_shim", "/tmp/scans.bin", "eager", "4", "12"], capture_output=True, text=True, env=env)
print(p.stdout.strip()[-300:]); print("\n".join(p.stderr.strip().splitlines()[-6:]), flush=True)
for mode in ("lazy", "sources4", "sources4_shared"):
    p = subprocess.run(["tools/build/bench_shim", "/tmp/scans.bin", mode, "4", "30"], capture_output=True, text=True)
    print(mode, p.stdout.strip()[-250:], p.stderr[-300:], flush=True)
PY
