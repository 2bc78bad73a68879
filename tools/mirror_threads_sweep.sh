python - <<'PY'
import os, struct, subprocess, sys
import numpy as np
sys.path.insert(0, ".")
from vdb_mapping_b200 import scans
c = scans.CONFIGS[2]
n = 14
with open("/tmp/scans.bin", "wb") as f:
    f.write(struct.pack("7d", c.resolution, c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max, n))
    for k in range(n):
        pts, origin = scans.make_scan(2, k)
        p16 = np.ones((pts.shape[0], 4), dtype=np.float32); p16[:, :3] = pts[:, :3]
        f.write(struct.pack("3d", *origin)); f.write(struct.pack("I", p16.shape[0])); f.write(p16.tobytes())
print("cpus", os.cpu_count(), open("/proc/cpuinfo").read().count("processor\t"), [l for l in open("/proc/cpuinfo") if "model name" in l][0].strip())
for thr in ("2", "4", "6", "8", "12", "16", "24", "8"):
    env = dict(os.environ, VDBM_MIRROR_PROFILE="1", VDBM_MIRROR_THREADS=thr)
    p = subprocess.run(["tools/build/bench_shim", "/tmp/scans.bin", "eager", "4", "14"], capture_output=True, text=True, env=env)
    import json
    d = json.loads(p.stdout.strip().splitlines()[-1])
    print("threads", thr, "ms_per_scan", d["ms_per_scan"], p.stderr.strip().splitlines()[-1][-90:], flush=True)
PY
