"""Builds tools/bench_shim.cpp against the shim headers + libvdbm_b200.so, writes cfg2 scans to a file and runs it:
the scan-integration throughput seen through the reference-compatible C++ class API.   python tools/bench_shim.py [n_scans]"""
import os, struct, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vdb_mapping_b200 import scans
from vdb_mapping_b200.build import build_lib

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 55
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 2          # BASELINE config of the scans (2 = OS1-128, 4 = 1M-point merged rig)
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else None   # e.g. lazy,group2,group4
so = build_lib()
out = os.path.join(ROOT, "tools", "build")
os.makedirs(out, exist_ok=True)
exe = os.path.join(out, "bench_shim")
subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "bench_shim.cpp"), "-o", exe,
                so, "-Wl,-rpath," + os.path.dirname(so), "-lpthread"], check=True)
c = scans.CONFIGS[cfg]
path = os.path.join(out, "scans_cfg%d.bin" % cfg)
with open(path, "wb") as f:
    f.write(struct.pack("7d", c.resolution, c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max, n_scans))
    for k in range(n_scans):
        pts, origin = scans.make_scan(cfg, k)
        p16 = np.ones((pts.shape[0], 4), dtype=np.float32); p16[:, :3] = pts[:, :3]
        f.write(struct.pack("3d", *origin)); f.write(struct.pack("I", p16.shape[0])); f.write(p16.tobytes())
runs = [("lazy", n_scans), ("eager", min(n_scans, 14)), ("sources4", n_scans), ("sources4_shared", n_scans)]
if modes:
    runs = [(m, min(n_scans, 14) if m == "eager" else n_scans) for m in modes]
for mode, n in runs:
    print(subprocess.run([exe, path, mode, "5", str(n)], capture_output=True, text=True).stdout.strip())
